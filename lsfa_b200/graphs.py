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
