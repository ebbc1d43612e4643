"""Mint tests/golden/coviar_ref.npz: outputs of the reference's OWN create_and_load_mv_residual
(coviar_data_loader.c:71-177, compiled from /root/reference by `make -C oracle ref`) on seeded
per-frame motion-vector lists.  Run in the authoring container (the GPU box has no /root/reference);
the fixture lets every box check the oracle restatement and the GPU kernels against what the
reference itself computed.

    python tools/make_golden_coviar_ref.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lsfa_oracle as O      # noqa: E402  (input synthesis only)
from oracle import ref_coviar as R       # noqa: E402

CASES = [(1, 32, 48, 0), (3, 48, 80, 6), (5, 40, 56, 6), (11, 96, 160, 8), (4, 45, 77, 4)]


def main():
    R.build(force=True)
    out = {}
    rng = np.random.default_rng(2024)
    for i, (T, h, w, extra) in enumerate(CASES):
        mvs, counts = O.synth_mv_lists(rng, T, h, w, extra=extra)
        if i == 2:
            counts[1] = 0                                  # a P-frame without vectors
        iframe = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        cur = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out["c%d_shape" % i] = np.array([T, h, w], np.int32)
        out["c%d_mvs" % i] = mvs.astype(np.int16)          # all values fit (FFmpeg's own fields are int16/uint8)
        out["c%d_counts" % i] = counts.astype(np.int32)
        out["c%d_iframe" % i] = iframe
        out["c%d_cur" % i] = cur
        out["c%d_mv" % i] = R.mv_accumulate(mvs, counts, h, w).astype(np.int16)
        out["c%d_res" % i] = R.residual(iframe, cur, mvs, counts).astype(np.int16)
        out["c%d_single" % i] = R.mv_single(mvs[0][:int(counts[0])], h, w).astype(np.int16)
    path = os.path.join(ROOT, "tests", "golden", "coviar_ref.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(CASES), "cases")


if __name__ == "__main__":
    main()
