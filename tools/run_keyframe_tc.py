#!/usr/bin/env python
"""A short run of the tensor-core key-frame networks for ncu: embed net + fused cosine (3 GEMM launches + finalise) and
the Nq net (1 launch) on 16 key frames at 1024x38x63."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsfa_b200 import graphs, ops  # noqa: E402

dev = torch.device("cuda", 0)
N, C, H, W, E = 16, 1024, 38, 63, 2048
g = torch.Generator(device=dev).manual_seed(6)
rn = lambda *sh: 0.01 * torch.randn(sh, device=dev, generator=g)  # noqa: E731
emb = (rn(512, C, 1, 1), rn(512), rn(512, 512, 3, 3), rn(512), rn(E, 512, 1, 1), rn(E))
nq = (rn(256, C, 3, 3), rn(256), rn(16, 256, 1, 1), rn(16), rn(1, 16, 1, 1), rn(1))
x = torch.randn((2 * N, H, W, C), device=dev, generator=g).clamp_(min=0).to(torch.bfloat16)
pe, pq = graphs.pack_embed_params(emb), graphs.pack_nq_params(nq)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    lg = ops.embed_cosine_logits(x, pe)
    lq = ops.nq_logits(x, pq)
torch.cuda.synchronize()
print("ok", float(lg.sum()), float(lq.sum()))
