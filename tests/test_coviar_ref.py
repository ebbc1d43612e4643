"""Pin of SURVEY 8f rank 3 (coviar MV accumulation + residual) against the REFERENCE ITSELF.

`oracle/_ref/libcoviar_ref.so` is the reference's own external/data_loader_py2/coviar_data_loader.c
compiled where it lies (`make -C oracle ref`; shims for the absent FFmpeg headers only, real
CPython/NumPy headers).  `tests/golden/coviar_ref.npz` holds outputs of its
`create_and_load_mv_residual` (:71-177) minted by tools/make_golden_coviar_ref.py, so boxes without
/root/reference (the GPU box) check against what the reference computed, not against a restatement.

CPU: the C restatement (oracle/lsfa_oracle.c) and the NumPy oracle == the reference, live and via the fixture.
GPU: lsfa_mv_accumulate_i32 / lsfa_coviar_residual_u8 == the reference's outputs, bit for bit.
"""
import os

import numpy as np
import pytest

from oracle import c_port as P
from oracle import lsfa_oracle as O
from oracle import ref_coviar as R

GOLD = os.path.join(os.path.dirname(__file__), "golden", "coviar_ref.npz")


def _cases():
    g = np.load(GOLD)
    i = 0
    while "c%d_shape" % i in g:
        T, h, w = (int(v) for v in g["c%d_shape" % i])
        yield dict(T=T, h=h, w=w, mvs=g["c%d_mvs" % i].astype(np.int32), counts=g["c%d_counts" % i],
                   iframe=g["c%d_iframe" % i], cur=g["c%d_cur" % i], mv=g["c%d_mv" % i].astype(np.int32),
                   res=g["c%d_res" % i].astype(np.int32))
        i += 1


def test_oracle_restatements_equal_the_reference_fixture():
    P.build()
    n = 0
    for c in _cases():
        got = P.mv_accumulate(c["mvs"], c["counts"], c["h"], c["w"])
        assert np.array_equal(got, c["mv"])
        assert np.array_equal(P.coviar_residual(c["iframe"], c["cur"], got), c["res"])
        assert np.array_equal(O.coviar_residual(c["iframe"], c["cur"], got), c["res"])
        if c["h"] * c["w"] * c["T"] <= 48 * 80 * 3:           # the literal Python loop: small cases only
            assert np.array_equal(O.coviar_accumulate(c["mvs"], c["counts"], c["h"], c["w"]), c["mv"])
        n += 1
    assert n >= 5


@pytest.mark.skipif(not R.available(), reason="neither /root/reference nor a prebuilt oracle/_ref")
def test_reference_library_live_against_fixture_and_oracle():
    """The compiled reference reproduces its own fixture, and agrees with the C restatement on fresh seeds
    (ragged counts, an empty P-frame, vectors leaving the frame, 8x8 vectors overlapping 16x16 ones)."""
    P.build()
    for c in _cases():
        assert np.array_equal(R.mv_accumulate(c["mvs"], c["counts"], c["h"], c["w"]), c["mv"])
        assert np.array_equal(R.residual(c["iframe"], c["cur"], c["mvs"], c["counts"]), c["res"])
    rng = np.random.default_rng(77)
    for (T, h, w) in [(2, 32, 32), (6, 64, 96), (11, 80, 112)]:
        mvs, counts = O.synth_mv_lists(rng, T, h, w, extra=8, max_disp=40)
        counts[T // 2] = max(0, counts[T // 2] - 5)
        ref = R.mv_accumulate(mvs, counts, h, w)
        assert np.array_equal(P.mv_accumulate(mvs, counts, h, w), ref)
        iframe = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        cur = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(P.coviar_residual(iframe, cur, ref), R.residual(iframe, cur, mvs, counts))


@pytest.mark.skipif(not R.available(), reason="neither /root/reference nor a prebuilt oracle/_ref")
def test_reference_library_on_adversarial_lists():
    """Unaligned blocks of odd sizes, static vectors on top of moving ones, sources / destinations outside the frame: the
    C restatement still equals the compiled reference (these lists are what the GPU back-trace's slow path is tested on)."""
    P.build()
    rng = np.random.default_rng(98)
    T, h, w, M = 5, 53, 71, 60
    mvs = np.zeros((T, M, 6), np.int32)
    for t in range(T):
        for i in range(M):
            bw, bh = (int(v) for v in rng.choice([3, 4, 5, 8, 15, 16, 17], 2))
            dx, dy = int(rng.integers(-6, w + 6)), int(rng.integers(-6, h + 6))
            ox, oy = (0, 0) if rng.random() < 0.25 else (int(v) for v in rng.integers(-20, 21, 2))
            mvs[t, i] = (bw, bh, dx + ox, dy + oy, dx, dy)
    counts = np.full(T, M, np.int32)
    counts[2] = 0
    assert np.array_equal(P.mv_accumulate(mvs, counts, h, w), R.mv_accumulate(mvs, counts, h, w))


@pytest.mark.gpu
def test_gpu_mv_accumulate_and_residual_equal_the_reference(cuda):
    import torch
    from lsfa_b200 import ops
    for c in _cases():
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)  # noqa: E731
        field = ops.mv_accumulate(t(c["mvs"][None]), t(c["counts"][None]), c["h"], c["w"])
        res = ops.coviar_residual(t(c["iframe"][None]), t(c["cur"][None]), field)
        torch.cuda.synchronize()
        assert np.array_equal(field[0].cpu().numpy(), c["mv"]), "accumulated MV differs from the reference"
        assert np.array_equal(res[0].cpu().numpy(), c["res"]), "residual differs from the reference"


@pytest.mark.gpu
@pytest.mark.skipif(not R.available(), reason="no prebuilt oracle/_ref on this box")
def test_gpu_against_the_live_reference_library_at_720p(cuda):
    """Full-size GOP (T = 11 P-frames, 720p) against the compiled reference running on the host."""
    import torch
    from lsfa_b200 import ops
    rng = np.random.default_rng(5)
    T, h, w = 11, 720, 1280
    mvs, counts = O.synth_mv_lists(rng, T, h, w, extra=16, max_disp=32)
    want = R.mv_accumulate(mvs, counts, h, w)
    for algo in ("trace", "field"):
        got = ops.mv_accumulate(torch.from_numpy(mvs[None]).to(cuda), torch.from_numpy(counts[None]).to(cuda), h, w, algo=algo)
        torch.cuda.synchronize()
        assert np.array_equal(got[0].cpu().numpy(), want), algo
