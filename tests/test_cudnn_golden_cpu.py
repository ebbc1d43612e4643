"""The oracle's a7+a8 restatement (forward AND backward) against outputs of cuDNN's spatial-transformer sampler -
the routine the reference's GPU build dispatches mx.sym.BilinearSampler to (SYM:307,321,469,572,679) - recorded on a
B200 by tools/make_golden_cudnn.py (torch.cudnn_grid_sampler = cudnnSpatialTfSamplerForward/Backward, cuDNN 9.22).
Runs on every box: this is what pins the oracle for the rows whose arithmetic lives in MXNet/cuDNN."""
import os

import numpy as np
import pytest

from oracle import c_port as P
from oracle import lsfa_oracle as O
from tests._util import assert_close_f32

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cudnn_sampler.npz")


def _cases():
    g = np.load(GOLD)
    i = 0
    while "c%d_data" % i in g:
        yield {k: g["c%d_%s" % (i, k)] for k in ("kind", "data", "flow", "og", "out", "gdata", "ggrid")}
        i += 1


@pytest.mark.skipif(not os.path.exists(GOLD), reason="fixture not minted yet")
def test_oracle_forward_and_backward_equal_cudnn():
    P.build()
    n = 0
    for c in _cases():
        grid = O.grid_generator_warp(c["flow"])
        scale = np.abs(c["data"]).max()
        assert_close_f32(O.bilinear_sampler(c["data"], grid), c["out"], scale=scale, what="NumPy oracle vs cuDNN (%s)" % c["kind"])
        assert_close_f32(P.bilinear_sampler(c["data"], P.grid_generator_warp(c["flow"])), c["out"], scale=scale,
                         what="C port vs cuDNN (%s)" % c["kind"])
        gd, gg = O.bilinear_sampler_backward(c["data"], grid, c["og"])
        # cuDNN sums with float atomics in an unspecified order: gate = the forward's, with the absolute part scaled to the gradient
        assert np.abs(gd - c["gdata"]).max() <= 4e-6 * np.abs(c["gdata"]).max()
        assert np.abs(gg - c["ggrid"]).max() <= 2e-5 * np.abs(c["ggrid"]).max()
        n += 1
    assert n >= 4
