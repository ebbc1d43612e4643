#!/usr/bin/env python
"""Mint golden vectors for the backward of GridGenerator(warp) + BilinearSampler from an implementation that is
independent of this repo's oracle: torch autograd through
``grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True)`` on the CPU (the reference ships no
tests or vectors: SURVEY.md section 4; MXNet itself is not available offline).  Inputs and gradients are stored.
Output: tests/golden/backward_small.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lsfa_oracle as O  # noqa: E402


def main():
    rng = np.random.default_rng(20261017)
    N, C, H, W = 3, 6, 14, 18
    key = rng.standard_normal((N, C, H, W), dtype=np.float32)
    og = rng.standard_normal((N, C, H, W), dtype=np.float32)
    flow = (rng.standard_normal((N, 2, H, W)) * 2.5).astype(np.float32)
    flow[0, :, :2] = 40.0                      # rows sampling outside the plane
    flow[1] = np.float32(0.37)                 # uniform sub-cell motion
    grid = O.grid_generator_warp(flow)         # a7 forward (float32 add/div/sub): input of the sampler's backward
    tk = torch.tensor(key, requires_grad=True)
    tg = torch.tensor(grid, requires_grad=True)
    out = torch.nn.functional.grid_sample(tk, tg.permute(0, 2, 3, 1), mode="bilinear", padding_mode="zeros",
                                          align_corners=True)
    out.backward(torch.tensor(og))
    grad_key, grad_grid = tk.grad.numpy(), tg.grad.numpy()
    half = np.array([(W - 1) / 2.0, (H - 1) / 2.0], np.float32).reshape(1, 2, 1, 1)
    dst = os.path.join(ROOT, "tests", "golden", "backward_small.npz")
    np.savez_compressed(dst, key=key, flow=flow, grid=grid, out_grad=og, out=out.detach().numpy(), grad_key=grad_key,
                        grad_grid=grad_grid, grad_flow=(grad_grid / half).astype(np.float32))
    print("wrote", dst, os.path.getsize(dst))


if __name__ == "__main__":
    main()
