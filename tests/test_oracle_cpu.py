"""CPU suite part 1: pin the oracle.

 * against fixtures produced by the REFERENCE's own Python code (lib/utils/image.py imported
   unmodified by tools/make_golden_from_reference.py) - rows a2-a6;
 * against cv2.resize, the third-party routine the reference calls (image.py:204-205,221-222);
 * against torch.grid_sample(align_corners=True), an independent implementation of a7+a8
   (MXNet itself cannot be built offline: those rows stay "parity unpinned");
 * known-answer cases the survey lists (8c i-ix);
 * the committed golden vectors of the fused op are reproducible from the oracle.
"""
import os

import numpy as np
import pytest

from oracle import lsfa_oracle as O
from tests._util import make_case, oracle_fused

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gref():
    return np.load(os.path.join(GOLD, "ref_transform_mv_res.npz"))


def test_transform_mv_res_matches_reference_fixture(gref):
    for i in range(int(gref["n_cases"])):
        s = float(gref["scale_%d" % i])
        mv, res = O.transform_mv_res(gref["mv_in_%d" % i], gref["res_in_%d" % i], s)
        want_mv, want_res = gref["mv_out_%d" % i], gref["res_out_%d" % i]
        assert mv.shape == want_mv.shape and res.shape == want_res.shape
        assert np.array_equal(mv, want_mv), "MV case %d (scale %g) not bit-exact" % (i, s)
        if s == 1.0 or s == 2.0:
            assert np.array_equal(res, want_res), "residual case %d" % i
        else:
            # cv2 >= 4 resizes 3-channel float images through a different (SIMD/IPP) routine
            # than 2-channel ones; our float32 transcription is the 2-channel routine.
            assert np.abs(res - want_res).max() <= 2e-5 * 64


def test_residual_mean_aliasing_matches_reference_fixture(gref):
    mv, res = O.transform_mv_res(gref["mv_in_m"], gref["res_in_m"], 1.0, gref["means_m"],
                                 float(gref["pscale_m"]))
    assert np.array_equal(mv, gref["mv_out_m"])
    assert np.array_equal(res, gref["res_out_m"])
    # the documented consequence of image.py:217-218 with zero means: (ch2, ch1, ch2)
    r = np.arange(16 * 16 * 3, dtype=np.float64).reshape(16, 16, 3)
    a = O.res_colour_mean_inplace(r.copy())
    assert np.array_equal(a[..., 0], r[..., 2]) and np.array_equal(a[..., 1], r[..., 1])
    assert np.array_equal(a[..., 2], r[..., 2])


def test_im_scale_rule_matches_reference_fixture(gref):
    for (h, w), s in zip(gref["resize_shapes"], gref["resize_scales"]):
        assert O.im_scale_for(int(h), int(w)) == float(s)


def test_pool_is_centre2x2_of_cv2_resize():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for shape in [(608, 1008, 2), (96, 160, 3), (16, 32, 2)]:
        p = rng.standard_normal(shape) * 37.0
        if shape[2] == 3:
            # cv2 sums 3-channel float64 rows in another order (last-ulp); the reference only ever
            # feeds float32-valued data here (image.py:205 -> :213), for which any order is exact
            p = p.astype(np.float32).astype(np.float64)
        want = cv2.resize(p, None, None, fx=1 / 16.0, fy=1 / 16.0, interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(O.pool_stride16(p, O.POOL_CENTRE2X2), want.reshape(shape[0] // 16, shape[1] // 16, -1))
    ints = rng.integers(-40, 41, size=(64, 64, 2)).astype(np.float64)
    assert not np.array_equal(O.pool_stride16(ints, O.POOL_CENTRE2X2), O.pool_stride16(ints, O.POOL_AVG16))
    assert np.allclose(O.pool_stride16(ints, O.POOL_AVG16), ints.reshape(4, 16, 4, 16, 2).mean(axis=(1, 3)))


def test_stage1_resize_matches_cv2_two_channel():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    for (h, w, s) in [(720, 1280, 0.78125), (1080, 1920, 1000 / 1920.0), (300, 500, 2.0), (96, 160, 1.0)]:
        mv = rng.integers(-40, 41, size=(h, w, 2)).astype(np.float32)
        want = cv2.resize(mv, None, None, fx=s, fy=s, interpolation=cv2.INTER_LINEAR)
        got = O.resize_linear_f32(mv, s)
        assert got.shape == want.shape
        assert np.array_equal(got, want), (h, w, s)
    for (h, w, s) in [(480, 640, 1.25), (360, 480, 600 / 360.0)]:      # up-scaling: <= 1 ulp on a handful
        mv = rng.integers(-40, 41, size=(h, w, 2)).astype(np.float32)
        want = cv2.resize(mv, None, None, fx=s, fy=s, interpolation=cv2.INTER_LINEAR)
        got = O.resize_linear_f32(mv, s)
        assert np.abs(got - want).max() <= 4e-6 and (got != want).mean() < 1e-4


def test_mv_sign_and_flip():
    mv = np.arange(2 * 3 * 2, dtype=np.int32).reshape(2, 3, 2)
    out = O.mv_sign_flip(mv, flipped=False)
    assert np.array_equal(out, -mv.astype(np.float32))
    f = O.mv_sign_flip(mv, flipped=True)
    assert np.array_equal(f[:, :, 0], mv[:, ::-1, 0].astype(np.float32))      # -(-x) reversed
    assert np.array_equal(f[:, :, 1], -mv[:, ::-1, 1].astype(np.float32))


def test_warp_matches_torch_grid_sample():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(0)
    key = O.synth_features(rng, (2, 32, 38, 63))
    flow = O.mv_pool(O.synth_raw_mv(rng, 2, 600, 1000, 96))
    grid = O.grid_generator_warp(flow)
    out = O.bilinear_sampler(key, grid)
    ref = torch.nn.functional.grid_sample(torch.from_numpy(key), torch.from_numpy(grid).permute(0, 2, 3, 1),
                                          mode="bilinear", padding_mode="zeros", align_corners=True).numpy()
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(key).max()


def test_grid_round_trip_floor_flips_documented():
    """integer flows land on k-eps after the fp32 round trip."""
    z = np.zeros((1, 2, 38, 63), np.float32)
    x0, y0, wx, wy = O.sampler_coords(O.grid_generator_warp(z), 38, 63)
    xs, ys = np.meshgrid(np.arange(63), np.arange(38))
    fx, fy = (x0[0] != xs).mean(), (y0[0] != ys).mean()
    assert 0.03 < fx < 0.10 and 0.10 < fy < 0.25
    key = O.synth_features(np.random.default_rng(0), (1, 8, 38, 63))
    assert np.abs(O.warp(key, z) - key).max() <= 4e-6 * np.abs(key).max()


def test_known_answers_oracle():
    rng = np.random.default_rng(0)
    N, C, H, W = 1, 4, 10, 12
    key = O.synth_features(rng, (N, C, H, W))
    cur = O.synth_features(rng, (N, C, H, W))
    z = np.zeros((N, 2, H, W), np.float32)
    f = z.copy(); f[:, 0] = 0.5
    want = np.zeros_like(key); want[..., :W - 1] = 0.5 * (key[..., :W - 1] + key[..., 1:]); want[..., W - 1] = 0.5 * key[..., W - 1]
    assert np.abs(O.warp(key, f) - want).max() <= 4e-6 * key.max()
    f = z.copy(); f[:, 0] = 500
    assert np.abs(O.warp(key, f)).max() == 0
    f = z.copy(); f[:, 1] = -0.25          # y_r in (-1,0): only the bottom taps are inside
    got = O.warp(key, f)
    assert np.allclose(got[:, :, 0], 0.75 * key[:, :, 0], rtol=1e-5, atol=1e-6)
    lg = np.full((N, 2, H, W), 0.3, np.float32)
    ones = np.ones_like(key)
    a = O.warp_scale_aggregate(key, z, cur=cur, scale_map=ones, weight_mode=O.W_LOGITS, logits=lg)
    b = O.warp_scale_aggregate(key, z, cur=cur, weight_mode=O.W_MEAN)
    assert np.abs(a - b).max() <= 1e-6 * max(key.max(), cur.max())
    e = np.zeros((N, 8, H, W), np.float32)
    assert np.all(O.cosine_weight(e, e) == 0)          # eps path
    e2 = rng.standard_normal((N, 8, H, W)).astype(np.float32)
    assert np.allclose(O.cosine_weight(e2, e2), 1.0, atol=1e-6)
    w1, w2 = O.softmax_pair(O.cosine_weight(e2, -e2), O.cosine_weight(e2, e2))
    assert np.allclose(w1, 1 / (1 + np.e ** 2), atol=1e-6) and np.allclose(w1 + w2, 1, atol=1e-6)
    byp = np.array([1], np.uint8)
    assert np.array_equal(O.warp_scale_aggregate(key, z, cur=cur, weight_mode=O.W_MEAN, bypass=byp), cur)


def test_reference_graph_compositions():
    d = make_case(3, 2, 8, 10, 12, E=16, with_res=True)
    nonkey = O.cur_frame_path(d["key"], d["flow"], d["res"], d["rnet_w"], d["rnet_b"], d["cur"])
    manual = d["cur"] + (O.warp(d["key"], d["flow"]) + O.rnet_conv0(d["res"], d["rnet_w"], d["rnet_b"]))
    assert np.array_equal(nonkey, manual)
    first = np.array([1, 0], np.uint8)
    nq = O.key_frame_path_nq(d["key"], d["flow"], d["scale_map"], d["cur"], d["logits"], first)
    assert np.array_equal(nq[0], d["cur"][0]) and not np.array_equal(nq[1], d["cur"][1])
    fg = O.key_frame_path_fgfa(d["key"], d["flow"], d["scale_map"], d["cur"], d["emb_warp"], d["emb_cur"], first)
    assert fg.shape == nq.shape
    bp = O.batch_path(d["key"][:1], d["flow"], d["scale_map"])
    assert np.array_equal(bp[1], O.warp(d["key"][:1], d["flow"][1:2])[0] * d["scale_map"][1])


def test_fused_golden_reproducible():
    g = np.load(os.path.join(GOLD, "fused_small.npz"))
    d = {k: g[k] for k in g.files}
    for name, mode in (("none", O.W_NONE), ("add", O.W_ADD), ("mean", O.W_MEAN), ("logits", O.W_LOGITS),
                       ("cosine", O.W_COSINE)):
        assert np.array_equal(oracle_fused(d, mode), d["out_" + name]), name
    x0, y0, wx, wy = O.sampler_coords(O.grid_generator_warp(d["flow"]), 10, 12)
    assert np.array_equal(x0, d["x0"]) and np.array_equal(y0, d["y0"])
    assert np.array_equal(O.mv_pool(d["mv"]), d["flow"])


def test_algorithmic_bytes_match_survey():
    assert O.algorithmic_bytes_per_frame(1024, 38, 63, 4, "V0") == 19_630_800
    assert O.algorithmic_bytes_per_frame(1024, 38, 63, 4, "V1") == 29_539_192
    assert O.algorithmic_bytes_per_frame(1024, 38, 63, 4, "V2") == 39_319_056
    assert O.algorithmic_bytes_per_frame(1024, 38, 63, 2, "V2") == 19_707_408
    assert O.algorithmic_bytes_per_frame(1024, 38, 63, 4, "V3") == 78_542_352
    assert O.algorithmic_bytes_per_frame(1024, 68, 120, 4, "V2") == 134_019_840


def test_oracle_properties_linearity_and_partition_of_unity():
    """Size-independent properties of a7+a8 the GPU tests rely on at full size: the warp is linear in the key feature,
    an interior constant plane stays constant (the four tap weights sum to 1), and zero padding only ever shrinks it."""
    rng = np.random.default_rng(12)
    N, C, H, W = 2, 5, 14, 17
    k1, k2 = O.synth_features(rng, (N, C, H, W)), O.synth_features(rng, (N, C, H, W))
    flow = rng.uniform(-3, 3, size=(N, 2, H, W)).astype(np.float32)
    a, b = np.float32(0.75), np.float32(-1.5)
    lhs = O.warp((a * k1 + b * k2).astype(np.float32), flow)
    rhs = a * O.warp(k1, flow) + b * O.warp(k2, flow)
    assert np.abs(lhs - rhs).max() <= 1e-5 * max(np.abs(k1).max(), np.abs(k2).max())
    ones = np.ones((N, C, H, W), np.float32)
    w = O.warp(ones, flow)
    assert w.max() <= 1.0 + 1e-6 and w.min() >= 0.0
    x0, y0, _, _ = O.sampler_coords(O.grid_generator_warp(flow), H, W)
    inside = (x0 >= 0) & (x0 + 1 <= W - 1) & (y0 >= 0) & (y0 + 1 <= H - 1)
    assert np.abs(w[:, 0][inside] - 1.0).max() <= 1e-6


def test_residual_front_end_matches_cv2_end_to_end():
    """a1..a5 for the residual (image.py:52,59,205,207-222): the oracle's own float32 resize against cv2's, through the
    whole chain, for the scales the reference's resize() produces; flip commutes with the chain as image.py:59 applies it."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(21)
    for (h, w), s in (((96, 160), 1.0), ((48, 80), 2.0), ((90, 160), 0.78125), ((72, 96), 1.25)):
        res = rng.integers(-64, 65, size=(h, w, 3)).astype(np.float32)
        mv = np.zeros((h, w, 2), np.float32)
        ours = O.transform_mv_res(mv, res, s, (1.5, -2.0, 3.25), 0.5)[1]
        ref = O.transform_mv_res(mv, res, s, (1.5, -2.0, 3.25), 0.5, use_cv2=True)[1]
        assert ours.shape == ref.shape
        if s in (1.0, 2.0):
            assert np.array_equal(ours, ref)
        else:
            assert np.abs(ours - ref).max() <= 2e-5 * 64        # cv2 >= 4: another SIMD routine for 3-channel images
        flipped = O.transform_mv_res(mv, res[:, ::-1], s)[1]
        assert flipped.shape == ours.shape


def test_centre_rows_are_exactly_what_the_stride16_reduction_reads():
    """The host path sends only rows 16k+7 / 16k+8 of an MV field (lsfa_mv_centre_rows_h2d): zeroing every other row
    must not change the oracle's pooled flow, for heights with and without a lone last row."""
    rng = np.random.default_rng(5)
    for h in (600, 599, 24, 23, 9, 8):
        mv = rng.integers(-64, 65, size=(1, h, 50, 2), dtype=np.int32)
        keep = np.zeros_like(mv)
        rows = [r for r in range(h) if r % 16 in (7, 8)]
        keep[:, rows] = mv[:, rows]
        assert np.array_equal(O.mv_pool(mv), O.mv_pool(keep)), h
