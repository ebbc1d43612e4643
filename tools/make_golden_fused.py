#!/usr/bin/env python
"""Mint the small golden vectors of the fused operator from the CPU oracle (the reference
ships none: SURVEY.md section 4).  Inputs AND oracle outputs are stored so the GPU tests do
not depend on the RNG.  Output: tests/golden/fused_small.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lsfa_oracle as O  # noqa: E402
from tests._util import make_case, oracle_fused  # noqa: E402


def main():
    d = make_case(424242, 3, 16, 10, 12, E=24, max_px=48, raw="ragged", with_bypass=True)
    x0, y0, wx, wy = O.sampler_coords(O.grid_generator_warp(d["flow"]), 10, 12)
    d.update(x0=x0, y0=y0, wx=wx, wy=wy, grid=O.grid_generator_warp(d["flow"]))
    for name, mode in (("none", O.W_NONE), ("add", O.W_ADD), ("mean", O.W_MEAN), ("logits", O.W_LOGITS),
                       ("cosine", O.W_COSINE)):
        d["out_" + name] = oracle_fused(d, mode)
    out = os.path.join(ROOT, "tests", "golden", "fused_small.npz")
    np.savez_compressed(out, **d)
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()
