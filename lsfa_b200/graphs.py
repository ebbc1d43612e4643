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


def _lowp(x, params, conv_dtype):
    """Library convolutions in a lower storage type (bf16, channels-last: the tensor-core path of cuDNN).  Measured on
    16 key frames of 1024x38x63: the embedding convs are 3.15 ms in fp32/TF32 (191 TFLOP/s) and 1.01 ms in bf16
    (599 TFLOP/s) - 89 % of the Fgfa key-frame graph either way (profiles/README.md).  fp32 stays the default: it is the
    reference's precision."""
    if conv_dtype is None or conv_dtype == torch.float32:
        return x, params
    cl = torch.channels_last
    x = x.to(conv_dtype).contiguous(memory_format=cl)
    params = tuple(p.to(conv_dtype).contiguous(memory_format=cl) if p.dim() == 4 else p.to(conv_dtype) for p in params)
    return x, params


def embed_net(x: torch.Tensor, w1, b1, w2, b2, w3, b3, conv_dtype=None) -> torch.Tensor:
    """get_embednet SYM:118-130: 1x1 (C->512) + ReLU, 3x3 (512->512, pad 1) + ReLU, 1x1 (512->2048).
    Returns float32 NCHW whatever ``conv_dtype`` the library convolutions ran in."""
    x, (w1, b1, w2, b2, w3, b3) = _lowp(x, (w1, b1, w2, b2, w3, b3), conv_dtype)
    x = F.relu(F.conv2d(x, w1, b1))
    x = F.relu(F.conv2d(x, w2, b2, padding=1))
    return F.conv2d(x, w3, b3).float().contiguous()


def nq_net(x: torch.Tensor, w1, b1, w2, b2, w3, b3, conv_dtype=None) -> torch.Tensor:
    """Nq_net convs SYM:97-101: 3x3 (C->256, pad 1) + ReLU, 1x1 (256->16) + ReLU, 1x1 (16->1)."""
    x, (w1, b1, w2, b2, w3, b3) = _lowp(x, (w1, b1, w2, b2, w3, b3), conv_dtype)
    x = F.relu(F.conv2d(x, w1, b1, padding=1))
    x = F.relu(F.conv2d(x, w2, b2))
    return F.conv2d(x, w3, b3).float().contiguous()


def key_frame_fgfa(feat_key_old, flow, scale_map, conv_feat, embed_params, is_first_frame=None,
                   flow_kind="flow", conv_dtype=None, **kw) -> torch.Tensor:
    """get_key_test_symbol with add_Fgfa_net (SYM:468-470,473-474,477), exact two-phase form."""
    warp = ops.warp_scale_aggregate(feat_key_old, flow, scale_map=scale_map, flow_kind=flow_kind, **kw)      # K1
    emb = embed_net(torch.cat([conv_feat, warp], dim=0), *embed_params, conv_dtype=conv_dtype)             # SYM:133-134
    n = conv_feat.shape[0]
    emb_cur, emb_warp = emb[:n].contiguous(), emb[n:].contiguous()                                         # SYM:135
    logits = ops.cosine_logits(emb_warp, emb_cur)                                                          # SYM:137-139
    # K2: blend the already warped+scaled feature with conv_feat: an identity warp is NOT needed -
    # the fused op takes the warped feature as `key` with zero flow only in the oracle; here we use
    # the logits blend on top of K1's output via the op-by-op tail (one pass over 3F).
    return ops.blend_logits(warp, conv_feat, logits, bypass=is_first_frame)


def key_frame_nq(feat_key_old, flow, scale_map, conv_feat, nq_params, is_first_frame=None,
                 flow_kind="flow", conv_dtype=None, **kw) -> torch.Tensor:
    """get_key_test_symbol with add_Nq_net (shipped, SYM:468-472,477)."""
    warp = ops.warp_scale_aggregate(feat_key_old, flow, scale_map=scale_map, flow_kind=flow_kind, **kw)      # K1
    n = conv_feat.shape[0]
    q = nq_net(torch.cat([warp, conv_feat], dim=0), *nq_params, conv_dtype=conv_dtype)                     # SYM:95-101
    logits = torch.cat([q[:n], q[n:]], dim=1).contiguous()                                                 # (N,2,H,W): [warp, conv]
    return ops.blend_logits(warp, conv_feat, logits, bypass=is_first_frame)
