#!/usr/bin/env python
"""Bring-up diagnostics for csrc/conv_gemm_tc.cu (run on the GPU box): graded cases from one k-step of a 1x1
convolution to the full networks, each printing error statistics and, on mismatch, WHERE the output differs
(rows = pixels, columns = output channels), so that a wrong descriptor / swizzle / coordinate shows its pattern."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsfa_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def case(name, NB, H, W, Cin, Cout, k, relu=False, ident=False, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn((NB, H, W, Cin), device=dev, generator=g).to(torch.bfloat16)
    if ident:
        w = torch.zeros((Cout, Cin, k, k), device=dev)
        for co in range(Cout):
            w[co, co % Cin, k // 2, k // 2] = 1.0 + (co // Cin)
    else:
        w = (torch.randn((Cout, Cin, k, k), device=dev, generator=g) / (Cin * k * k) ** 0.5).to(torch.bfloat16).float()
    b = 0.1 * torch.randn((Cout,), device=dev, generator=g)
    want = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w, b, padding=k // 2)
    if relu:
        want = want.clamp_min(0)
    want = want.permute(0, 2, 3, 1)
    try:
        got = ops.conv_bf16_nhwc(x, ops.pack_conv_weight(w), b, relu=relu).float()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print("%-44s EXCEPTION %r" % (name, e), flush=True)
        return False
    err = (got - want).abs()
    tol = 2.0 ** -7 * want.abs() + 2.0 ** -8 * 1e-1 * want.abs().max()
    bad = err > tol
    nbad = int(bad.sum())
    print("%-44s max err %.3g (scale %.3g)  bad %d / %d" % (name, err.max().item(), want.abs().max().item(), nbad, bad.numel()), flush=True)
    if nbad:
        b4 = bad.reshape(NB, H * W, Cout)
        print("   bad per image:", b4.sum((1, 2)).tolist())
        rows = b4.any(2).sum(1).tolist()
        cols = b4.any(1).sum(1).tolist()
        print("   pixels with any bad channel per image:", rows, " channels with any bad pixel per image:", cols)
        i0 = b4[0].nonzero()[:8].tolist()
        print("   first bad (pixel, channel) in image 0:", i0)
        for (p, c) in i0[:4]:
            print("      got %.4f want %.4f" % (got.reshape(NB, H * W, Cout)[0, p, c].item(), want.reshape(NB, H * W, Cout)[0, p, c].item()))
    return nbad == 0


ok = True
ok &= case("1x1 K=64 one tile identity", 2, 2, 64, 64, 256, 1, ident=True)
ok &= case("1x1 K=64 one tile random", 2, 2, 64, 64, 256, 1)
ok &= case("1x1 K=128 one tile", 2, 2, 64, 128, 256, 1)
ok &= case("1x1 K=512 4 tiles, 2 chunks", 2, 4, 128, 512, 512, 1)
ok &= case("1x1 ragged 38x63", 2, 38, 63, 64, 256, 1)
ok &= case("3x3 K=64 one tile identity (centre tap)", 2, 2, 64, 64, 256, 3, ident=True)
ok &= case("3x3 K=64 one tile", 2, 2, 64, 64, 256, 3)
ok &= case("3x3 38x63 Cin=128", 2, 38, 63, 128, 256, 3, relu=True)
ok &= case("3x3 4 images 17x23", 4, 17, 23, 64, 512, 3)
ok &= case("1x1 many items (persistent loop)", 8, 38, 63, 64, 512, 1, relu=True)
print("ALL OK" if ok else "SOME FAILED")
