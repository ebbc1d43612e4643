"""The C port of the reference CPU path (oracle/lsfa_oracle.c, used as cpu_baseline) agrees
bit for bit with the NumPy oracle it is checked against."""
import numpy as np

from oracle import c_port as P
from oracle import lsfa_oracle as O
from tests._util import make_case, oracle_fused


def test_cport_ops_bit_exact():
    P.build()
    rng = np.random.default_rng(0)
    mv = O.synth_raw_mv(rng, 2, 600 - 8, 1000 - 11, 96)
    for mode in (0, 1):
        assert np.array_equal(P.mv_pool(mv, 0.78125, mode), O.mv_pool(mv, 0.78125, mode))
    flow = O.mv_pool(mv)
    flow[0, :, 0, :3] = [[1000.0, -1000.0, 0.5]] * 2
    grid = O.grid_generator_warp(flow)
    assert np.array_equal(P.grid_generator_warp(flow), grid)
    key = O.synth_features(rng, (2, 8, 38, 63))
    assert np.array_equal(P.bilinear_sampler(key, grid), O.bilinear_sampler(key, grid))
    g2 = (rng.random((2, 2, 11, 13)) * 2.4 - 1.2).astype(np.float32)
    assert np.array_equal(P.bilinear_sampler(key, g2), O.bilinear_sampler(key, g2))


def test_cport_chain_matches_fused_oracle():
    d = make_case(12, 2, 16, 38, 63)
    out, _ = P.chain_nq(d["mv"], d["key"], d["scale_map"], d["cur"], d["logits"])
    want = oracle_fused(d, O.W_LOGITS)
    # expf of glibc vs numpy may differ in the last ulp of the softmax weights
    assert np.abs(out - want).max() <= 1e-6 * max(np.abs(d["key"]).max(), np.abs(d["cur"]).max())
    assert P.num_threads() >= 1
