#!/usr/bin/env python
"""Mint tests/golden/fused_backward_small.npz: gradients of the fused operator
    out = wc*cur + ww*(grid_sample(key, grid(flow)) * scale + rnet(res)),  (ww,wc) = softmax(logits), bypass -> cur
from an implementation independent of this repo's oracle: torch autograd in float64 on the CPU
(grid_sample(bilinear, zeros, align_corners=True) = a7+a8; SYM:306-338 is what get_train_symbol differentiates)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rng = np.random.default_rng(20261018)
    N, C, H, W = 3, 5, 9, 11
    f = lambda *s: rng.standard_normal(s)  # noqa: E731
    key, cur, sm = np.maximum(f(N, C, H, W), 0), np.maximum(f(N, C, H, W), 0), 1 + 0.1 * f(N, C, H, W)
    flow = 1.7 * f(N, 2, H, W)
    flow[0, :, :2] = 30.0
    logits, res, rw, rb, og = f(N, 2, H, W), 30 * f(N, 3, H, W), 0.05 * f(C, 3), 0.05 * f(C), f(N, C, H, W)
    byp = np.array([0, 1, 0], np.uint8)
    out = {}
    for mode in ("logits", "add"):
        t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in
             dict(key=key, cur=cur, scale=sm, flow=flow, logits=logits, res=res, rnet_w=rw, rnet_b=rb).items()}
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
        gx = (t["flow"][:, 0] + xs) / ((W - 1) / 2.0) - 1
        gy = (t["flow"][:, 1] + ys) / ((H - 1) / 2.0) - 1
        wp = torch.nn.functional.grid_sample(t["key"], torch.stack([gx, gy], -1), mode="bilinear", padding_mode="zeros", align_corners=True)
        if mode == "logits":
            src0 = wp * t["scale"]
            w = torch.softmax(t["logits"], dim=1)
            o = w[:, 0:1] * src0 + w[:, 1:2] * t["cur"]
        else:
            src0 = wp + torch.einsum("cj,njhw->nchw", t["rnet_w"], t["res"]) + t["rnet_b"].reshape(1, C, 1, 1)
            o = src0 + t["cur"]
        live = torch.tensor((byp == 0).astype(np.float64)).reshape(N, 1, 1, 1)
        o = live * o + (1 - live) * t["cur"]
        o.backward(torch.tensor(og))
        for k, v in t.items():
            if v.grad is not None:
                out["%s_grad_%s" % (mode, k)] = v.grad.numpy().astype(np.float32)
        out["%s_out" % mode] = o.detach().numpy().astype(np.float32)
    for k, v in dict(key=key, cur=cur, scale=sm, flow=flow, logits=logits, res=res, rnet_w=rw, rnet_b=rb, out_grad=og).items():
        out[k] = v.astype(np.float32)
    out["bypass"] = byp
    dst = os.path.join(ROOT, "tests", "golden", "fused_backward_small.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst))


if __name__ == "__main__":
    main()
