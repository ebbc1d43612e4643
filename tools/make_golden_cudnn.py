#!/usr/bin/env python
"""Mint gpurun_out/cudnn_sampler.npz ON THE GPU BOX: outputs of cuDNN's spatial-transformer sampler
(cudnnSpatialTfSamplerForward/Backward through torch.cudnn_grid_sampler - the routine the reference's GPU build
dispatches mx.sym.BilinearSampler to) on seeded inputs.  Copy the file to tests/golden/ and commit it: the CPU
suite then checks the oracle's a7+a8 restatement (forward and backward) against what cuDNN computed, on every box."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lsfa_oracle as O  # noqa: E402  (input synthesis + GridGenerator restatement only)

dev = torch.device("cuda", 0)
out = {}
rng = np.random.default_rng(7)
cases = [("blocks", 2, 8, 38, 63), ("half", 1, 8, 17, 23), ("outside", 2, 6, 12, 20), ("subpixel", 1, 4, 68, 120)]
for i, (kind, N, C, H, W) in enumerate(cases):
    data = O.synth_features(rng, (N, C, H, W))
    if kind == "blocks":
        flow = np.ascontiguousarray(O.mv_pool(O.synth_raw_mv(rng, N, 16 * H, 16 * W, 96)))
    elif kind == "half":
        flow = (rng.integers(-6, 7, size=(N, 2, H, W)) * 0.5).astype(np.float32)
    elif kind == "outside":
        flow = (rng.standard_normal((N, 2, H, W)) * np.array([W, H], np.float32).reshape(1, 2, 1, 1) * 0.6).astype(np.float32)
    else:
        flow = (rng.standard_normal((N, 2, H, W)) * 2.5).astype(np.float32)
    grid = O.grid_generator_warp(flow)
    og = rng.standard_normal((N, C, H, W), dtype=np.float32)
    d = torch.from_numpy(data).to(dev).requires_grad_(True)
    g = torch.from_numpy(grid).to(dev).requires_grad_(True)
    y = torch.cudnn_grid_sampler(d, g.permute(0, 2, 3, 1).contiguous())
    y.backward(torch.from_numpy(og).to(dev))
    out["c%d_kind" % i] = np.array(kind)
    out["c%d_data" % i] = data
    out["c%d_flow" % i] = flow
    out["c%d_og" % i] = og
    out["c%d_out" % i] = y.detach().cpu().numpy()
    out["c%d_gdata" % i] = d.grad.cpu().numpy()
    out["c%d_ggrid" % i] = g.grad.cpu().numpy()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
path = os.path.join(ROOT, "gpurun_out", "cudnn_sampler.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes; cudnn", torch.backends.cudnn.version())
