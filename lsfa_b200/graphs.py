"""The reference graphs around the path, composed from this package's operators.

Dense convolutions (the embedding net SYM:118-130, the Nq net SYM:94-101) are *library* GEMMs here
(``torch.nn.functional.conv2d`` = cuDNN/cuBLAS), exactly as north_star prescribes ("the embedding 1x1
convs are left as dense GEMMs"); everything between them is this package's CUDA:

    key frame, Fgfa (SYM:468-470,473-474,132-148)  K1  warp x scale            lsfa kernel (1 pass)
                                                    --  embed(conv_feat), embed(warp)   cuDNN
                                                    K2  cosine + softmax + blend         lsfa kernels
    key frame, Nq   (SYM:468-472,94-109)            K1, Nq convs (cuDNN), K2 = logits blend

The exact Fgfa graph needs K1's output in HBM (the embedding net consumes the *warped* feature and has
a 3x3 conv), so it is two lsfa passes + the library convs; the single-pass fused operator is exact for
the shipped non-key path and whenever logits / embeddings are given.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops


def prepare_params(params, conv_dtype=None):
    """Convolution parameters in the storage type / memory format the library convolutions will run in (do this once,
    not per call: converting the 3x3 weights costs as much as the convolution of a small batch)."""
    if conv_dtype is None or conv_dtype == torch.float32:
        return tuple(params)
    cl = torch.channels_last
    return tuple(p if p.dtype == conv_dtype and (p.dim() != 4 or p.is_contiguous(memory_format=cl))
                 else (p.to(conv_dtype).contiguous(memory_format=cl) if p.dim() == 4 else p.to(conv_dtype)) for p in params)


def _lowp_input(feats, conv_dtype):
    """Concat_0 of float32 NCHW features as ONE bf16 channels-last tensor, written by this package's transpose kernel
    straight into its slice of the destination (one pass per feature; no float32 concat, no separate cast)."""
    n = sum(f.shape[0] for f in feats)
    _, c, h, w = feats[0].shape
    buf = torch.empty((n, h, w, c), dtype=conv_dtype, device=feats[0].device)
    at = 0
    for f in feats:
        ops.to_nhwc(f, conv_dtype, out=buf[at:at + f.shape[0]])
        at += f.shape[0]
    return buf.permute(0, 3, 1, 2)          # logical NCHW, channels-last strides: what cuDNN's tensor-core path wants


def embed_net(x: torch.Tensor, w1, b1, w2, b2, w3, b3) -> torch.Tensor:
    """get_embednet SYM:118-130: 1x1 (C->512) + ReLU, 3x3 (512->512, pad 1) + ReLU, 1x1 (512->2048), in whatever
    storage type / memory format ``x`` and the parameters have."""
    x = F.relu(F.conv2d(x, w1, b1))
    x = F.relu(F.conv2d(x, w2, b2, padding=1))
    return F.conv2d(x, w3, b3)


def nq_net(x: torch.Tensor, w1, b1, w2, b2, w3, b3) -> torch.Tensor:
    """Nq_net convs SYM:97-101: 3x3 (C->256, pad 1) + ReLU, 1x1 (256->16) + ReLU, 1x1 (16->1)."""
    x = F.relu(F.conv2d(x, w1, b1, padding=1))
    x = F.relu(F.conv2d(x, w2, b2))
    return F.conv2d(x, w3, b3)


def pack_embed_params(embed_params):
    """(w1,b1,w2,b2,w3,b3) as MXNet holds them -> the tensor-core kernels' form (bf16 K-major weights, f32 biases)."""
    w1, b1, w2, b2, w3, b3 = embed_params
    return (ops.pack_conv_weight(w1), b1.contiguous(), ops.pack_conv_weight(w2), b2.contiguous(), ops.pack_conv_weight(w3),
            b3.contiguous())


def pack_nq_params(nq_params):
    w1, b1, w2, b2, w3, b3 = nq_params
    return (ops.pack_conv_weight(w1), b1.contiguous(), w2.contiguous(), b2.contiguous(), w3.contiguous(), b3.contiguous())


def _tc_supported(c_in, *c_outs):
    return c_in % 64 == 0 and all(c % 256 == 0 for c in c_outs)


def key_frame_fgfa(feat_key_old, flow, scale_map, conv_feat, embed_params, is_first_frame=None,
                   flow_kind="flow", conv_dtype=None, packed=None, **kw) -> torch.Tensor:
    """get_key_test_symbol with add_Fgfa_net (SYM:468-470,473-474,477), exact two-phase form.

    ``conv_dtype=torch.bfloat16`` runs the library convolutions on cuDNN's bf16 channels-last (tensor-core) path:
    measured on 16 key frames of 1024x38x63 they take 3.15 ms in fp32/TF32 (191 TFLOP/s) and 1.01 ms in bf16
    (599 TFLOP/s) - 89 % of this graph in fp32.  The features enter through this package's transpose kernel and the
    2048-channel embeddings stay bf16 channels-last for the cosine kernel (no float32 copy of them is ever made).
    fp32 is the default: it is the reference's precision."""
    warp = ops.warp_scale_aggregate(feat_key_old, flow, scale_map=scale_map, flow_kind=flow_kind, **kw)      # K1
    n = conv_feat.shape[0]
    if conv_dtype is None or conv_dtype == torch.float32:
        emb = embed_net(torch.cat([conv_feat, warp], dim=0), *embed_params)                                # SYM:133-134
        emb_cur, emb_warp = emb[:n].contiguous(), emb[n:].contiguous()                                     # SYM:135
        logits = ops.cosine_logits(emb_warp, emb_cur)                                                      # SYM:137-139
    elif conv_dtype == "tc":
        # this package's tcgen05 implicit-GEMM convolutions; compute_weight's reductions are em_conv3's epilogue, so the
        # 2048-channel embeddings are never written (packed = pack_embed_params(embed_params), once per parameter set)
        x = _lowp_input([conv_feat, warp], torch.bfloat16).permute(0, 2, 3, 1)
        logits = ops.embed_cosine_logits(x, packed if packed is not None else pack_embed_params(embed_params))
    else:
        emb = embed_net(_lowp_input([conv_feat, warp], conv_dtype), *prepare_params(embed_params, conv_dtype))
        emb = emb.permute(0, 2, 3, 1)                                    # (2N,H,W,E) view of the channels-last result
        if not emb.is_contiguous():
            emb = emb.contiguous()
        logits = ops.cosine_logits(emb[n:], emb[:n], layout="nhwc_bf16")
    # K2: blend the already warped+scaled feature with conv_feat: an identity warp is NOT needed -
    # the fused op takes the warped feature as `key` with zero flow only in the oracle; here we use
    # the logits blend on top of K1's output via the op-by-op tail (one pass over 3F).
    return ops.blend_logits(warp, conv_feat, logits, bypass=is_first_frame)


def key_frame_nq(feat_key_old, flow, scale_map, conv_feat, nq_params, is_first_frame=None,
                 flow_kind="flow", conv_dtype=None, packed=None, **kw) -> torch.Tensor:
    """get_key_test_symbol with add_Nq_net (shipped, SYM:468-472,477).  ``conv_dtype``: see key_frame_fgfa."""
    warp = ops.warp_scale_aggregate(feat_key_old, flow, scale_map=scale_map, flow_kind=flow_kind, **kw)      # K1
    n = conv_feat.shape[0]
    if conv_dtype == "tc":
        x = _lowp_input([warp, conv_feat], torch.bfloat16).permute(0, 2, 3, 1)
        logits = ops.nq_logits(x, packed if packed is not None else pack_nq_params(nq_params))             # SYM:95-101, one launch
        return ops.blend_logits(warp, conv_feat, logits, bypass=is_first_frame)
    if conv_dtype is None or conv_dtype == torch.float32:
        q = nq_net(torch.cat([warp, conv_feat], dim=0), *nq_params)                                        # SYM:95-101
    else:
        q = nq_net(_lowp_input([warp, conv_feat], conv_dtype), *prepare_params(nq_params, conv_dtype)).float()
    logits = torch.cat([q[:n], q[n:]], dim=1).contiguous()                                                 # (N,2,H,W): [warp, conv]
    return ops.blend_logits(warp, conv_feat, logits, bypass=is_first_frame)


# ------------------------------------------------------------------------------------------------------------------
# The non-key step with every graph switch of the reference (SYM:57-67, 209-272, 326-328, 570-586), in bf16 channels-last:
# convolutions on this package's tensor-core kernels (ops.conv_bf16_nhwc), the warp / add through the fused operator, the
# two tiny SE-block layers (1x1 convolutions on a (N,C,1,1) pooled vector) and the channel concat in torch.  The shipped
# configuration is fuse_type='add', res_fuse='add', rnet_num_conv=0: that one is ops.cur_frame_path + ONE convolution.
# BatchNorm variants (res_diff_bn, small_net_bn_before_fuse) are not built: they need running statistics the path does not own.
# ------------------------------------------------------------------------------------------------------------------
def _nhwc_bf16(x):
    """(N,C,H,W) float32 -> (N,H,W,C) bf16 through this package's transpose kernel (C padded to a multiple of 64 with zeros
    when needed: the 3-channel residual of res_diff_ada)."""
    n, c, h, w = x.shape
    if c % 64:
        pad = torch.zeros((n, 64 - c % 64, h, w), dtype=x.dtype, device=x.device)
        x = torch.cat([x, pad], dim=1).contiguous()
    return ops.to_nhwc(x.contiguous(), torch.bfloat16)


def _conv(x_nhwc, wb, relu=False):
    """mx.sym.Convolution (+ReLU) on tensor cores; wb = (weight (Cout,Cin,k,k) f32, bias (Cout,) f32) as MXNet holds them."""
    w, b = wb
    cin = x_nhwc.shape[3]
    if w.shape[1] != cin:                                   # input channels were zero-padded to 64
        w = torch.cat([w, torch.zeros((w.shape[0], cin - w.shape[1]) + tuple(w.shape[2:]), dtype=w.dtype, device=w.device)], dim=1)
    return ops.conv_bf16_nhwc(x_nhwc, ops.pack_conv_weight(w.contiguous()), b.contiguous(), relu=relu)


def _se(cat_nhwc, p1, p2):
    """global average pool -> 1x1 conv + ReLU -> 1x1 conv + sigmoid (SYM:255-259, 266-270): (N,1,1,C') float32."""
    s = cat_nhwc.float().mean(dim=(1, 2))                                    # (N, C)
    s = torch.relu(s @ p1[0].reshape(p1[0].shape[0], -1).t() + p1[1])
    return torch.sigmoid(s @ p2[0].reshape(p2[0].shape[0], -1).t() + p2[1])[:, None, None, :]


def cur_frame_step(feat_key, motion_vector, res_diff, rnet_convs, cur_small, small_params, fuse_type="add", res_fuse="add",
                   fuse_downsample=None, flow_kind="flow", im_scale=1.0) -> torch.Tensor:
    """get_cur_test_symbol's tail (SYM:570-586) for every `small_net_fuse_type` (SYM:229-272), `rnet_num_conv`
    (SYM:57-67) and `fuse_type` (SYM:323-328).  Inputs NCHW float32 as the reference holds them: feat_key (N,1024,H,W),
    res_diff (N,3,H,W), cur_small (N,256,H,W); returns (N,H,W,1024) bf16 channels-last."""
    key = _nhwc_bf16(feat_key)
    x = _nhwc_bf16(res_diff)
    for wb in rnet_convs[:-1]:
        x = _conv(x, wb, relu=True)
    rd = _conv(x, rnet_convs[-1])                                            # rnet_conv{num_conv}: 1x1 -> 1024
    small = _nhwc_bf16(cur_small)
    kw = dict(flow_kind=flow_kind, im_scale=im_scale, layout="nhwc_bf16")
    if res_fuse == "add" and fuse_type in ("add", "addv2"):
        # out = cur + (warp + res_diff): one fused pass - the convolved residual enters as the `cur` of an 'add', then the
        # small-net branch is added the same way
        if fuse_type == "add":
            cur = _conv(small, small_params["fuse_reduce_add"])
        else:
            cur = _conv(_conv(small, small_params["fuse_reduce_add_conv1"], relu=True), small_params["fuse_reduce_add_conv2"])
        cur = (cur.float() + rd.float()).to(torch.bfloat16)
        return ops.warp_scale_aggregate(key, motion_vector, cur=cur, weight_mode="add", **kw)
    if res_fuse == "add":
        wp = ops.warp_scale_aggregate(key, motion_vector, cur=rd, weight_mode="add", **kw)
    else:
        wp = ops.warp_scale_aggregate(key, motion_vector, **kw)
        wp = _conv(torch.cat([wp, rd], dim=3).contiguous(), fuse_downsample)
    if fuse_type in ("add", "addv2"):
        if fuse_type == "add":
            cur = _conv(small, small_params["fuse_reduce_add"])
        else:
            cur = _conv(_conv(small, small_params["fuse_reduce_add_conv1"], relu=True), small_params["fuse_reduce_add_conv2"])
        return (cur.float() + wp.float()).to(torch.bfloat16)
    if fuse_type in ("concat", "concatv1"):
        c1 = _conv(small, small_params["fuse_reduce_c1"])
        c2 = _conv(wp, small_params["fuse_reduce_c2"])
        cat = _conv(torch.cat([c2, c1], dim=3).contiguous(), small_params["fuse_reduce"], relu=(fuse_type == "concatv1"))
        if fuse_type == "concat":
            return cat
        return (cat.float() * _se(cat, small_params["s_feat_conv1"], small_params["s_feat_conv2"]) + cat.float()).to(torch.bfloat16)
    if fuse_type == "concatv2":
        c = _conv(small, small_params["fuse_reduce_c1"])
        sfeat = _se(torch.cat([wp, c], dim=3), small_params["s_feat_conv1"], small_params["s_feat_conv2"])
        return (c.float() * sfeat + wp.float()).to(torch.bfloat16)
    raise ValueError("unknown small_net_fuse_type %r" % (fuse_type,))
