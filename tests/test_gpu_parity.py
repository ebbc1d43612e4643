"""GPU parity tests proper: every call goes through the C ABI (lsfa_b200.ops -> ctypes ->
liblsfa_b200.so) and is compared with the CPU oracle on the same seeded inputs.

Gates: MV pooling and sampler indices bit-exact; fp32 features
|a-b| <= 1e-5|b| + 1e-6 max|data|; bf16 variant rtol 2^-8 against the fp32 oracle evaluated
on bf16-rounded inputs."""
import numpy as np
import pytest
import torch

from oracle import lsfa_oracle as O
from tests._util import assert_close_bf16, assert_close_f32, make_case, oracle_fused

pytestmark = pytest.mark.gpu


def dev(x, cuda):
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


def host(t):
    torch.cuda.synchronize()
    return t.detach().float().cpu().numpy()


@pytest.fixture(scope="module")
def ops(cuda):
    from lsfa_b200 import ops as _ops
    return _ops


# ------------------------------------------------------------------------------------------
# integer / indexing math: bit-exact
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("h,w", [(608, 1008), (600, 1000), (75, 131), (16, 16), (7, 9), (1080, 1920)])
@pytest.mark.parametrize("mode", ["centre2x2", "avg16"])
@pytest.mark.parametrize("scale", [1.0, 0.78125, 1.0 / 0.6])
def test_mv_pool_bit_exact(ops, cuda, h, w, mode, scale):
    rng = np.random.default_rng(h * 31 + w)
    n = 2
    mv = rng.integers(-96, 97, size=(n, h, w, 2), dtype=np.int32)
    want = O.mv_pool(mv, scale, O.POOL_CENTRE2X2 if mode == "centre2x2" else O.POOL_AVG16)
    got = host(ops.mv_pool(dev(mv, cuda), scale, mode))
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    mvf = (mv.astype(np.float32) * np.float32(0.37)).astype(np.float32)   # non-integer field
    want = O.mv_pool(mvf, scale, O.POOL_CENTRE2X2 if mode == "centre2x2" else O.POOL_AVG16)
    got = host(ops.mv_pool(dev(mvf, cuda), scale, mode))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("h,w", [(600, 1000), (75, 131), (64, 64)])
@pytest.mark.parametrize("means,ps", [((0.0, 0.0, 0.0), 1.0), ((103.06, 115.90, 123.15), 0.5)])
def test_res_pool_bit_exact(ops, cuda, h, w, means, ps):
    rng = np.random.default_rng(7)
    res = rng.integers(-64, 65, size=(2, h, w, 3), dtype=np.int32)
    for mode, om in (("centre2x2", O.POOL_CENTRE2X2), ("avg16", O.POOL_AVG16)):
        want = O.res_pool(res, means, ps, om)
        got = host(ops.res_pool(dev(res, cuda), means, ps, mode))
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), mode


def test_mv_pool_matches_reference_golden(ops, cuda, golden_ref):
    """Fixtures produced by the reference's own transform_mv_res (tools/make_golden_from_reference.py)."""
    g = golden_ref
    for i in range(int(g["n_cases"])):
        s = float(g["scale_%d" % i])
        if s != 1.0:
            continue   # stage-1 resize is covered by test_mv_prepare_*
        got_mv, got_res = ops.transform_mv_res(dev(g["mv_in_%d" % i][None], cuda),
                                               dev(g["res_in_%d" % i][None], cuda), s)
        assert np.array_equal(host(got_mv), g["mv_out_%d" % i].astype(np.float32))
        assert np.array_equal(host(got_res), g["res_out_%d" % i].astype(np.float32))
    got_mv, got_res = ops.transform_mv_res(dev(g["mv_in_m"][None], cuda), dev(g["res_in_m"][None], cuda),
                                           1.0, tuple(g["means_m"]), float(g["pscale_m"]))
    assert np.array_equal(host(got_mv), g["mv_out_m"].astype(np.float32))
    assert np.array_equal(host(got_res), g["res_out_m"].astype(np.float32))


@pytest.fixture(scope="module")
def golden_ref():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_transform_mv_res.npz"))


@pytest.mark.parametrize("h,w,scale", [(90, 160, 0.78125), (72, 96, 1.25), (48, 80, 2.0), (96, 160, 1.0),
                                       (117, 203, 600.0 / 117.0 / 4.0)])
@pytest.mark.parametrize("flip", [False, True])
def test_mv_prepare_matches_oracle(ops, cuda, h, w, scale, flip):
    rng = np.random.default_rng(3)
    raw = rng.integers(-48, 49, size=(2, h, w, 2), dtype=np.int32)
    got = host(ops.mv_prepare(dev(raw, cuda), scale, negate=True, flipped=flip))
    want = np.stack([O.resize_linear_f32(O.mv_sign_flip(r, flip), scale) for r in raw])
    assert got.shape == want.shape
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("h,w,scale", [(90, 160, 0.78125), (72, 96, 1.25), (48, 80, 2.0), (96, 160, 1.0), (75, 131, 1.0),
                                       (117, 203, 600.0 / 117.0 / 4.0)])
@pytest.mark.parametrize("flip", [False, True])
def test_res_coviar_pool_matches_oracle(ops, cuda, h, w, scale, flip):
    """image.py:52,59,205,207-222 for the residual in one launch, bit-exact against the oracle's whole chain."""
    rng = np.random.default_rng(5)
    raw = rng.integers(-64, 65, size=(2, h, w, 3), dtype=np.int32)
    means, ps = (3.5, -2.25, 10.0), 0.5
    for mode, om in (("centre2x2", O.POOL_CENTRE2X2), ("avg16", O.POOL_AVG16)):
        for mm, pp in (((0.0, 0.0, 0.0), 1.0), (means, ps)):
            got = host(ops.res_coviar_pool(dev(raw, cuda), scale, flipped=flip, pixel_means=mm, pixel_scale=pp, mode=mode))
            want = np.concatenate([O.to_f32(O.transform_mv_res(np.zeros((h, w, 2), np.float32),
                                                                r[:, ::-1] if flip else r, scale, mm, pp, om)[1]) for r in raw])
            assert got.shape == want.shape
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (mode, mm)


def test_res_coviar_pool_matches_reference_golden(ops, cuda, golden_ref):
    """Against the reference's own transform_mv_res outputs: bit-exact where cv2 4.x resizes 3-channel images with
    the routine the oracle transcribes (scale 1, 2), 2e-5 relative otherwise (see tests/test_oracle_cpu.py)."""
    g = golden_ref
    for i in range(int(g["n_cases"])):
        s = float(g["scale_%d" % i])
        rin = g["res_in_%d" % i]
        assert np.array_equal(rin, np.rint(rin)), "fixture residuals are integer-valued"
        got = host(ops.res_coviar_pool(dev(rin.astype(np.int32)[None], cuda), s))
        want = g["res_out_%d" % i].astype(np.float32)
        if s in (1.0, 2.0):
            assert np.array_equal(got, want), "case %d" % i
        else:
            assert np.abs(got - want).max() <= 2e-5 * 64, "case %d" % i
    rin = g["res_in_m"]
    got = host(ops.res_coviar_pool(dev(rin.astype(np.int32)[None], cuda), 1.0, pixel_means=tuple(g["means_m"]),
                                   pixel_scale=float(g["pscale_m"])))
    assert np.array_equal(got, g["res_out_m"].astype(np.float32))


@pytest.mark.parametrize("N,H,W", [(2, 38, 63), (1, 68, 120), (3, 7, 9), (1, 1, 5), (1, 5, 1)])
def test_grid_generator_bit_exact(ops, cuda, N, H, W):
    rng = np.random.default_rng(11)
    flow = (rng.standard_normal((N, 2, H, W)) * 3).astype(np.float32)
    flow[:, :, ::2, ::3] = np.round(flow[:, :, ::2, ::3])     # integer flows: the floor-flip cases
    flow[0, :, 0, 0] = 0.0
    with np.errstate(all="ignore"):
        want = O.grid_generator_warp(flow)
    got = host(ops.GridGenerator(dev(flow, cuda), transform_type="warp"))
    nan = np.isnan(want)            # H or W == 1 divides by (dim-1)/2 == 0: NaN payloads are not compared
    assert np.array_equal(np.isnan(got), nan)
    assert np.array_equal(got.view(np.uint32)[~nan], want.view(np.uint32)[~nan])


@pytest.mark.parametrize("N,H,W,Hi,Wi", [(2, 38, 63, 38, 63), (1, 68, 120, 68, 120), (2, 9, 7, 12, 5)])
def test_sampler_indices_bit_exact(ops, cuda, N, H, W, Hi, Wi):
    rng = np.random.default_rng(5)
    flow = (rng.standard_normal((N, 2, H, W)) * 4).astype(np.float32)
    flow[:, :, ::2] = np.round(flow[:, :, ::2])
    flow[0, :, :2] = 0.0
    flow[0, 0, 3] = 1000.0      # far out of bounds
    flow[0, 1, 3] = -1000.0
    flow[-1, :, -1] = 0.5       # half-cell
    if (H, W) == (Hi, Wi):
        grid = O.grid_generator_warp(flow)
        x0, y0, wx, wy = O.sampler_coords(grid, Hi, Wi)
        gx0, gy0, gwx, gwy = ops.sampler_coords(dev(flow, cuda), (Hi, Wi), is_grid=False)
        assert np.array_equal(host(gx0).astype(np.int32), x0)
        assert np.array_equal(host(gy0).astype(np.int32), y0)
        assert np.array_equal(host(gwx).view(np.uint32), wx.view(np.uint32))
        assert np.array_equal(host(gwy).view(np.uint32), wy.view(np.uint32))
    grid = (rng.random((N, 2, H, W)) * 2.4 - 1.2).astype(np.float32)
    x0, y0, wx, wy = O.sampler_coords(grid, Hi, Wi)
    gx0, gy0, gwx, gwy = ops.sampler_coords(dev(grid, cuda), (Hi, Wi), is_grid=True)
    assert np.array_equal(gx0.cpu().numpy(), x0) and np.array_equal(gy0.cpu().numpy(), y0)
    assert np.array_equal(host(gwx).view(np.uint32), wx.view(np.uint32))
    assert np.array_equal(host(gwy).view(np.uint32), wy.view(np.uint32))


# ------------------------------------------------------------------------------------------
# a8 BilinearSampler on its own (plane-resident and generic kernels, req write/add)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,C,Hi,Wi,Ho,Wo", [(2, 64, 38, 63, 38, 63), (1, 6, 68, 120, 68, 120),
                                             (2, 5, 9, 7, 11, 13), (1, 3, 4, 4, 2, 2), (2, 8, 37, 63, 37, 63)])
def test_bilinear_sampler(ops, cuda, N, C, Hi, Wi, Ho, Wo):
    rng = np.random.default_rng(C)
    data = O.synth_features(rng, (N, C, Hi, Wi))
    grid = (rng.random((N, 2, Ho, Wo)) * 2.3 - 1.15).astype(np.float32)   # partly out of bounds
    want = O.bilinear_sampler(data, grid)
    got = ops.BilinearSampler(dev(data, cuda), dev(grid, cuda))
    assert_close_f32(host(got), want, scale=np.abs(data).max(), what="write")
    acc = ops.BilinearSampler(dev(data, cuda), dev(grid, cuda), out=got, req="add")
    assert_close_f32(host(acc), 2 * want, scale=2 * np.abs(data).max(), what="add")


def test_bilinear_sampler_matches_torch_grid_sample(ops, cuda):
    """Second, independent implementation: align_corners=True, zeros padding."""
    rng = np.random.default_rng(0)
    data = O.synth_features(rng, (2, 32, 38, 63))
    flow = O.mv_pool(O.synth_raw_mv(rng, 2, 600, 1000, 96))
    grid = ops.GridGenerator(dev(flow, cuda))
    got = host(ops.BilinearSampler(dev(data, cuda), grid))
    ref = torch.nn.functional.grid_sample(torch.from_numpy(data), grid.cpu().permute(0, 2, 3, 1),
                                          mode="bilinear", padding_mode="zeros", align_corners=True)
    assert np.abs(got - ref.numpy()).max() <= 2e-6 * np.abs(data).max()


# ------------------------------------------------------------------------------------------
# the fused operator: every weight mode x layout x kernel
# ------------------------------------------------------------------------------------------
MODES = [("none", O.W_NONE), ("add", O.W_ADD), ("mean", O.W_MEAN), ("logits", O.W_LOGITS), ("cosine", O.W_COSINE)]
SHAPES = [(3, 64, 38, 63), (2, 8, 68, 120), (4, 16, 7, 9), (2, 32, 37, 63), (1, 8, 69, 67)]


def run_fused(ops, cuda, d, mode_name, layout, use_scale=True, use_res=False, flow_kind="raw",
              force_generic=False, req="write", out=None, workspace=None):
    bf16 = layout == "nhwc_bf16"
    nhwc = layout != "nchw"

    def feat(x):
        t = dev(x, cuda)
        if not nhwc:
            return t
        return ops.to_nhwc(t, torch.bfloat16 if bf16 else torch.float32)

    kw = dict(weight_mode=mode_name, layout=layout, force_generic=force_generic, req=req)
    if mode_name != "none":
        kw["cur"] = feat(d["cur"])
    if use_scale:
        kw["scale_map"] = feat(d["scale_map"])
    if use_res:
        kw.update(res=dev(d["res"], cuda), rnet_w=dev(d["rnet_w"], cuda), rnet_b=dev(d["rnet_b"], cuda))
    if mode_name == "logits":
        kw["logits"] = dev(d["logits"], cuda)
    if mode_name == "cosine":
        kw.update(emb_warp=feat(d["emb_warp"]), emb_cur=feat(d["emb_cur"]))
    if "bypass" in d and mode_name != "none":
        kw["bypass"] = dev(d["bypass"], cuda)
    if "key_index" in d:
        kw["key_index"] = dev(d["key_index"], cuda)
    if flow_kind == "raw":
        flow = dev(d["mv"], cuda)
    elif flow_kind == "flow":
        flow = dev(d["flow"], cuda)
    else:
        flow = ops.GridGenerator(dev(d["flow"], cuda))
    if out is not None:
        kw["out"] = out
    if workspace is not None:
        kw["workspace"] = workspace
    res = ops.warp_scale_aggregate(feat(d["key"]), flow, flow_kind=flow_kind, **kw)
    return ops.to_nchw(res) if nhwc else res


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mode_name,mode", MODES)
# kernel choice for NCHW: 0 auto (all-TMA if it can), 1 generic gather, 2 plane-resident LDG/STG
@pytest.mark.parametrize("layout,generic", [("nchw", 0), ("nchw", 1), ("nchw", 2), ("nhwc_f32", 0)])
def test_fused_f32(ops, cuda, shape, mode_name, mode, layout, generic):
    N, C, H, W = shape
    d = make_case(hash((shape, mode)) % 1000, N, C, H, W, E=32 if mode == O.W_COSINE else 0,
                  with_bypass=(mode != O.W_NONE and N >= 3))
    want = oracle_fused(d, mode)
    got = host(run_fused(ops, cuda, d, mode_name, layout, force_generic=generic))
    scale = max(np.abs(d["key"]).max(), np.abs(d["cur"]).max())
    assert_close_f32(got, want, scale=scale, what="%s %s %s" % (shape, mode_name, layout))


@pytest.mark.parametrize("shape", SHAPES[:3])
@pytest.mark.parametrize("mode_name,mode", MODES)
def test_fused_bf16_nhwc(ops, cuda, shape, mode_name, mode):
    N, C, H, W = shape
    d = make_case(17 + mode, N, C, H, W, E=32 if mode == O.W_COSINE else 0, with_bypass=N >= 3 and mode != O.W_NONE)
    for k in ("key", "cur", "scale_map", "emb_warp", "emb_cur"):
        if k in d:
            d[k] = O.bf16_round(d[k])           # oracle sees exactly what the kernel reads
    want = oracle_fused(d, mode)
    got = host(run_fused(ops, cuda, d, mode_name, "nhwc_bf16"))
    assert_close_bf16(got, want, what="%s %s" % (shape, mode_name))


@pytest.mark.parametrize("layout,generic", [("nchw", 0), ("nchw", 1), ("nchw", 2), ("nhwc_f32", 0), ("nhwc_bf16", 0)])
def test_cur_frame_path_as_shipped(ops, cuda, layout, generic):
    """get_cur_test_symbol (SYM:570-586): warp(MV) + rnet_conv0(res) + small-net feature, no scale."""
    d = make_case(101, 3, 64, 38, 63, with_res=True, raw="ragged")
    if layout == "nhwc_bf16":
        for k in ("key", "cur"):
            d[k] = O.bf16_round(d[k])
    want = O.cur_frame_path(d["key"], d["flow"], d["res"], d["rnet_w"], d["rnet_b"], d["cur"])
    got = host(run_fused(ops, cuda, d, "add", layout, use_scale=False, use_res=True, force_generic=generic))
    if layout == "nhwc_bf16":
        assert_close_bf16(got, want, what=layout)
    else:
        assert_close_f32(got, want, scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()), what=layout)


@pytest.mark.parametrize("flow_kind", ["raw", "flow", "grid"])
def test_flow_sources_agree(ops, cuda, flow_kind):
    """raw MV pooled in-kernel == pre-pooled flow == GridGenerator output fed as a grid: identical bits."""
    d = make_case(5, 2, 16, 38, 63)
    base = host(run_fused(ops, cuda, d, "logits", "nchw", flow_kind="raw"))
    got = host(run_fused(ops, cuda, d, "logits", "nchw", flow_kind=flow_kind))
    assert np.array_equal(base.view(np.uint32), got.view(np.uint32))


@pytest.mark.parametrize("variant", ["warp", "scale", "scale_cur", "res_cur"])
@pytest.mark.parametrize("shape", [(5, 64, 38, 63), (3, 16, 60, 60), (2, 8, 16, 24), (1, 4, 38, 63), (3, 8, 68, 120), (2, 4, 100, 132)])
def test_all_tma_kernel_every_variant(ops, cuda, variant, shape):
    """force_generic=3 pins the warp-specialised all-TMA kernel; it must serve these shapes and agree
    with the oracle (bypass frames included: their cur is carried HBM -> smem -> HBM by TMA alone)."""
    N, C, H, W = shape
    d = make_case(7 + N, N, C, H, W, with_res=(variant == "res_cur"), with_bypass=(variant in ("scale_cur", "res_cur") and N >= 3))
    scale = max(np.abs(d["key"]).max(), np.abs(d["cur"]).max())
    if variant == "warp":
        want = oracle_fused(d, O.W_NONE, use_scale=False)
        got = run_fused(ops, cuda, d, "none", "nchw", use_scale=False, force_generic=3)
    elif variant == "scale":
        want = oracle_fused(d, O.W_NONE)
        got = run_fused(ops, cuda, d, "none", "nchw", force_generic=3)
    elif variant == "scale_cur":
        want = oracle_fused(d, O.W_LOGITS)
        got = run_fused(ops, cuda, d, "logits", "nchw", force_generic=3)
    else:
        want = oracle_fused(d, O.W_ADD, use_scale=False, use_res=True)
        got = run_fused(ops, cuda, d, "add", "nchw", use_scale=False, use_res=True, force_generic=3)
    assert_close_f32(host(got), want, scale=scale, what="tma %s %s" % (variant, shape))


@pytest.mark.parametrize("variant", ["warp", "scale", "scale_cur", "res_cur"])
@pytest.mark.parametrize("shape", [(3, 8, 68, 120), (5, 6, 60, 60), (2, 4, 40, 100)])
def test_cluster_multicast_kernel(ops, cuda, variant, shape):
    """force_generic=4 pins the 2-CTA cluster kernel (multicast key load, one pixel part per CTA); it must give
    the same bits as the single-CTA all-TMA kernel and pass the oracle gate, bypass frames included."""
    N, C, H, W = shape
    d = make_case(70 + N, N, C, H, W, max_px=64, with_res=(variant == "res_cur"),
                  with_bypass=(variant in ("scale_cur", "res_cur")))
    args = {"warp": ("none", dict(use_scale=False)), "scale": ("none", {}), "scale_cur": ("logits", {}),
            "res_cur": ("add", dict(use_scale=False, use_res=True))}[variant]
    mode = {"none": O.W_NONE, "logits": O.W_LOGITS, "add": O.W_ADD}[args[0]]
    want = oracle_fused(d, mode, **args[1])
    got4 = run_fused(ops, cuda, d, args[0], "nchw", force_generic=4, **args[1])
    got3 = run_fused(ops, cuda, d, args[0], "nchw", force_generic=3, **args[1])
    assert torch.equal(got4, got3)
    assert_close_f32(host(got4), want, scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()), what="cluster %s" % variant)
    # static split (no scratch) and repeated launches
    t = lambda k: dev(d[k], cuda)  # noqa: E731
    if variant == "scale_cur":
        kw = dict(flow_kind="raw", cur=t("cur"), scale_map=t("scale_map"), weight_mode="logits", logits=t("logits"),
                  bypass=t("bypass"), force_generic=4)
        st = ops.warp_scale_aggregate(t("key"), t("mv"), workspace=False, **kw)
        assert torch.equal(st, got4)
        prep = ops.PreparedAggregate(t("key"), t("mv"), **kw)
        for _ in range(5):
            prep.run()
        assert torch.equal(prep.out, got4)


def test_all_tma_static_and_dynamic_split_agree(ops, cuda):
    """With a workspace the all-TMA kernel claims work dynamically (per-frame queues); without one it
    uses a static contiguous split.  Same bits either way, and repeated launches reuse the scratch."""
    d = make_case(91, 9, 64, 38, 63, with_bypass=True)
    t = lambda k: dev(d[k], cuda)  # noqa: E731
    kw = dict(flow_kind="raw", cur=t("cur"), scale_map=t("scale_map"), weight_mode="logits", logits=t("logits"),
              bypass=t("bypass"), force_generic=3)
    dyn = ops.PreparedAggregate(t("key"), t("mv"), **kw)
    a = host(dyn.run()).copy()
    b = host(dyn.run()).copy()
    st = host(ops.warp_scale_aggregate(t("key"), t("mv"), workspace=False, **kw))
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(a.view(np.uint32), st.view(np.uint32))
    assert_close_f32(a, oracle_fused(d, O.W_LOGITS), scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()), what="dyn")


@pytest.mark.parametrize("src_hw,scale", [((360, 480), 600 / 360.0), ((480, 854), 1.0), ((720, 1280), 0.78125), ((203, 317), 1.25)])
@pytest.mark.parametrize("flip", [False, True])
def test_coviar_front_end_folded_into_the_kernel(ops, cuda, src_hw, scale, flip):
    """flow_kind='coviar' = image.py:53-60 (sign, flip) + :204 (cv2 resize by im_scale) + :207-228 inside the
    fused op.  Bit-identical to running mv_prepare and mv_pool first, and inside the fp32 gate vs the oracle."""
    h0, w0 = src_hw
    rng = np.random.default_rng(h0)
    raw = O.synth_raw_mv(rng, 2, h0, w0, 24)                       # what coviar returns (before the sign flip)
    stage1 = np.stack([O.resize_linear_f32(O.mv_sign_flip(r, flip), scale) for r in raw])
    flow = O.mv_pool(stage1, scale)
    N, _, H, W = flow.shape
    C = 16
    key = O.synth_features(rng, (N, C, H, W)); cur = O.synth_features(rng, (N, C, H, W))
    sm = O.synth_scale_map(rng, (N, C, H, W)); lg = rng.standard_normal((N, 2, H, W), dtype=np.float32)
    want = O.warp_scale_aggregate(key, flow, cur=cur, scale_map=sm, weight_mode=O.W_LOGITS, logits=lg)
    kw = dict(cur=dev(cur, cuda), scale_map=dev(sm, cuda), weight_mode="logits", logits=dev(lg, cuda))
    two_step = ops.warp_scale_aggregate(dev(key, cuda), ops.mv_prepare(dev(raw, cuda), scale, True, flip),
                                        flow_kind="raw", im_scale=scale, **kw)
    for fg in (0, 1, 2):
        got = ops.warp_scale_aggregate(dev(key, cuda), dev(raw, cuda), flow_kind="coviar", im_scale=scale,
                                       negate=True, flipped=flip, force_generic=fg, **kw)
        assert torch.equal(got, two_step), "kernel choice %d" % fg
    assert_close_f32(host(got), want, scale=max(np.abs(key).max(), np.abs(cur).max()), what="coviar")
    nh = lambda a: ops.to_nhwc(dev(a, cuda))  # noqa: E731
    got_nhwc = ops.to_nchw(ops.warp_scale_aggregate(nh(key), dev(raw, cuda), flow_kind="coviar", im_scale=scale, negate=True,
                                                    flipped=flip, layout="nhwc_f32", cur=nh(cur), scale_map=nh(sm),
                                                    weight_mode="logits", logits=dev(lg, cuda)))
    assert torch.equal(got_nhwc, two_step)


def test_shared_key_feature_tile_as(ops, cuda):
    """get_batch_test_symbol (SYM:675-680): one key feature, many frames (tile_as -> key_index)."""
    d = make_case(9, 5, 32, 38, 63, shared_key=True)
    want = O.batch_path(d["key"], d["flow"], d["scale_map"])
    for layout in ("nchw", "nhwc_f32"):
        got = host(run_fused(ops, cuda, d, "none", layout))
        assert_close_f32(got, want, scale=np.abs(d["key"]).max(), what=layout)


def test_req_add_and_null(ops, cuda):
    d = make_case(21, 2, 16, 38, 63)
    want = oracle_fused(d, O.W_LOGITS)
    for layout, generic in (("nchw", 0), ("nchw", 1), ("nchw", 2), ("nhwc_f32", 0)):
        first = run_fused(ops, cuda, d, "logits", layout, force_generic=generic)
        if layout == "nchw":
            acc = run_fused(ops, cuda, d, "logits", layout, force_generic=generic, req="add", out=first)
            assert_close_f32(host(acc), 2 * want, scale=2 * np.abs(want).max(), what="add " + layout)
            sentinel = torch.full_like(first, 7.0)
            run_fused(ops, cuda, d, "logits", layout, req="null", out=sentinel)
            assert float(host(sentinel).min()) == 7.0 == float(host(sentinel).max())


@pytest.mark.parametrize("variant", ["warp", "scale", "scale_cur", "mean", "res_cur"])
@pytest.mark.parametrize("layout", ["nhwc_f32", "nhwc_bf16"])
@pytest.mark.parametrize("shape", [(5, 64, 38, 63), (2, 1024, 38, 63), (3, 16, 17, 23), (2, 200, 9, 7), (1, 8, 1, 5), (2, 40, 68, 120)])
def test_nhwc_all_tma_kernel_every_variant(ops, cuda, variant, layout, shape):
    """force_generic=3 pins the channels-last all-TMA kernel (taps, scale and cur moved by bulk copies, one bulk store
    per pixel group); it must serve these shapes, agree with the oracle, and agree BIT FOR BIT with the LDG/STG tile
    kernel (force_generic=1) - same expression chain - with the claim counter (workspace) and without (static stride).
    Shapes cover partial 32-pixel batches, partial pixel groups, channel runs that are not a multiple of 512 bytes,
    bypass frames and ragged raw MV images."""
    N, C, H, W = shape
    bf16 = layout == "nhwc_bf16"
    d = make_case(11 + N + C, N, C, H, W, with_res=(variant == "res_cur"), raw="ragged" if H > 1 else True,
                  with_bypass=(variant in ("scale_cur", "mean", "res_cur") and N >= 3))
    if bf16:
        for k in ("key", "cur", "scale_map"):
            d[k] = O.bf16_round(d[k])
    call = {"warp": ("none", O.W_NONE, dict(use_scale=False)), "scale": ("none", O.W_NONE, {}),
            "scale_cur": ("logits", O.W_LOGITS, {}), "mean": ("mean", O.W_MEAN, {}),
            "res_cur": ("add", O.W_ADD, dict(use_scale=False, use_res=True))}[variant]
    want = oracle_fused(d, call[1], **call[2])
    got = host(run_fused(ops, cuda, d, call[0], layout, force_generic=3, **call[2]))
    if bf16:
        assert_close_bf16(got, want, what="nhwc tma %s %s" % (variant, shape))
    else:
        assert_close_f32(got, want, scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()), what="nhwc tma %s %s" % (variant, shape))
    ldg = host(run_fused(ops, cuda, d, call[0], layout, force_generic=1, **call[2]))
    assert np.array_equal(got.view(np.uint32), ldg.view(np.uint32)), "all-TMA and LDG/STG kernels differ"
    static = host(run_fused(ops, cuda, d, call[0], layout, force_generic=3, workspace=False, **call[2]))
    assert np.array_equal(got.view(np.uint32), static.view(np.uint32)), "claimed and static batch orders differ"


@pytest.mark.parametrize("variant", ["warp", "scale", "scale_cur", "mean", "res_cur"])
@pytest.mark.parametrize("layout", ["nhwc_f32", "nhwc_bf16"])
@pytest.mark.parametrize("shape", [(5, 128, 38, 63), (2, 1024, 38, 63), (3, 128, 17, 23), (2, 256, 9, 7), (1, 128, 1, 5), (2, 128, 68, 120)])
def test_nhwc_window_kernel_every_variant(ops, cuda, variant, layout, shape):
    """force_generic=5 pins the window-resident channels-last kernel (one tensor-map copy of a 12x12 key window per 8x8
    output tile and 256-byte channel chunk, zero padding by the copy engine's out-of-bounds fill).  It must agree with the
    oracle and BIT FOR BIT with the LDG/STG tile kernel, with the claim counter and with the static tile stride.  Shapes
    cover planes smaller than a tile, tiles overhanging the plane on both sides, bypass frames, ragged raw MV images."""
    N, C, H, W = shape
    bf16 = layout == "nhwc_bf16"
    d = make_case(21 + N + C, N, C, H, W, with_res=(variant == "res_cur"), raw="ragged" if H > 1 else True,
                  with_bypass=(variant in ("scale_cur", "mean", "res_cur") and N >= 3))
    if bf16:
        for k in ("key", "cur", "scale_map"):
            d[k] = O.bf16_round(d[k])
    call = {"warp": ("none", O.W_NONE, dict(use_scale=False)), "scale": ("none", O.W_NONE, {}),
            "scale_cur": ("logits", O.W_LOGITS, {}), "mean": ("mean", O.W_MEAN, {}),
            "res_cur": ("add", O.W_ADD, dict(use_scale=False, use_res=True))}[variant]
    want = oracle_fused(d, call[1], **call[2])
    got = host(run_fused(ops, cuda, d, call[0], layout, force_generic=5, **call[2]))
    if bf16:
        assert_close_bf16(got, want, what="nhwc window %s %s" % (variant, shape))
    else:
        assert_close_f32(got, want, scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()), what="nhwc window %s %s" % (variant, shape))
    ldg = host(run_fused(ops, cuda, d, call[0], layout, force_generic=1, **call[2]))
    assert np.array_equal(got.view(np.uint32), ldg.view(np.uint32)), "window and LDG/STG kernels differ"
    static = host(run_fused(ops, cuda, d, call[0], layout, force_generic=5, workspace=False, **call[2]))
    assert np.array_equal(got.view(np.uint32), static.view(np.uint32)), "claimed and static tile orders differ"


@pytest.mark.parametrize("max_px", [96, 400, 4000])
def test_nhwc_window_kernel_large_motion_falls_back_to_per_pixel_boxes(ops, cuda, max_px):
    """Motion larger than the window slack (taps of one 8x8 tile spread over more than 12 key pixels, or leaving the plane
    altogether): the tile is served as gather items, each pixel with its own 2x2 box - same bits as the tile kernel - and
    the shared-key (tile_as) form on top."""
    d = make_case(500 + max_px, 4, 256, 38, 63, max_px=max_px, with_bypass=True, shared_key=True)
    for layout in ("nhwc_bf16", "nhwc_f32"):
        dd = dict(d)
        if layout == "nhwc_bf16":
            for k in ("key", "cur", "scale_map"):
                dd[k] = O.bf16_round(d[k])
        got = host(run_fused(ops, cuda, dd, "logits", layout, force_generic=5))
        ldg = host(run_fused(ops, cuda, dd, "logits", layout, force_generic=1))
        assert np.array_equal(got.view(np.uint32), ldg.view(np.uint32)), (layout, max_px)


def test_nhwc_window_kernel_key_plane_of_another_size(ops, cuda):
    """BilinearSampler with a data plane that is not the output's size (Hi x Wi != Ho x Wo) through the window kernel."""
    rng = np.random.default_rng(5)
    N, C, Hk, Wk, H, W = 2, 128, 21, 40, 13, 18
    key = O.synth_features(rng, (N, C, Hk, Wk))
    grid = rng.uniform(-1.2, 1.2, size=(N, 2, H, W)).astype(np.float32)
    want = O.bilinear_sampler(key, grid)
    kt = ops.to_nhwc(dev(key, cuda), torch.float32)
    outs = [host(ops.to_nchw(ops.warp_scale_aggregate(kt, dev(grid, cuda), flow_kind="grid", layout="nhwc_f32", force_generic=fg)))
            for fg in (5, 1)]
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    assert_close_f32(outs[0], want, scale=np.abs(key).max(), what="window kernel, other plane size")


def test_nhwc_all_tma_kernel_full_size_and_shared_key(ops, cuda):
    """Config-3 shape (1024 x 38 x 63 bf16, raw int32 MVs) with more batches than one wave of CTAs, and the tile_as
    form (one key feature shared by every frame through key_index)."""
    d = make_case(77, 6, 1024, 38, 63, with_bypass=True, shared_key=True)
    for k in ("key", "cur", "scale_map"):
        d[k] = O.bf16_round(d[k])
    want = oracle_fused(d, O.W_LOGITS)
    got = host(run_fused(ops, cuda, d, "logits", "nhwc_bf16", force_generic=3))
    assert_close_bf16(got, want, what="nhwc tma full size")
    ldg = host(run_fused(ops, cuda, d, "logits", "nhwc_bf16", force_generic=1))
    assert np.array_equal(got.view(np.uint32), ldg.view(np.uint32))


def test_nhwc_all_tma_kernel_random_shapes_match_tile_kernel(ops, cuda):
    """Seeded sweep over odd shapes (channel runs from 16 B to 8 KB, planes from 1 to ~700 pixels, 1-7 frames): the
    all-TMA channels-last kernel and the LDG/STG tile kernel must agree bit for bit, three launches each (a hand-shake
    error between record warp, producers and consumer groups shows up as a hang or as a wrong pixel group)."""
    rng = np.random.default_rng(2024)
    for it in range(24):
        bf16 = bool(it & 1)
        lanes = 8 if bf16 else 4
        C = int(rng.choice([lanes, 2 * lanes, 24, 40, 72, 136, 256, 520, 1024, 2048]))
        C = (C + lanes - 1) // lanes * lanes
        H, W = int(rng.integers(1, 27)), int(rng.integers(1, 27))
        N = int(rng.integers(1, 8))
        variant = ["warp", "scale", "scale_cur", "res_cur"][it % 4]
        layout = "nhwc_bf16" if bf16 else "nhwc_f32"
        d = make_case(300 + it, N, C, H, W, with_res=(variant == "res_cur"), with_bypass=(variant in ("scale_cur", "res_cur") and N >= 3))
        call = {"warp": ("none", dict(use_scale=False)), "scale": ("none", {}), "scale_cur": ("logits", {}),
                "res_cur": ("add", dict(use_scale=False, use_res=True))}[variant]
        ref = host(run_fused(ops, cuda, d, call[0], layout, force_generic=1, **call[1]))
        for rep in range(3):
            got = host(run_fused(ops, cuda, d, call[0], layout, force_generic=3, workspace=False if rep == 2 else None, **call[1]))
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (it, N, C, H, W, variant, layout, rep)


@pytest.mark.parametrize("layout", ["nhwc_f32", "nhwc_bf16"])
def test_nhwc_all_tma_kernel_key_plane_of_another_size(ops, cuda, layout):
    """BilinearSampler semantics in channels-last: the key plane (12x5) and the sampling grid (9x7) differ in size and the
    grid reaches outside [-1,1] (zero padding); all-TMA kernel = tile kernel bit for bit, both = oracle."""
    rng = np.random.default_rng(9)
    N, C, Hk, Wk, H, W = 3, 48, 12, 5, 9, 7
    key = O.synth_features(rng, (N, C, Hk, Wk))
    grid = rng.uniform(-1.3, 1.3, size=(N, 2, H, W)).astype(np.float32)
    bf16 = layout == "nhwc_bf16"
    if bf16:
        key = O.bf16_round(key)
    want = O.bilinear_sampler(key, grid)
    kt = ops.to_nhwc(dev(key, cuda), torch.bfloat16 if bf16 else torch.float32)
    outs = [host(ops.to_nchw(ops.warp_scale_aggregate(kt, dev(grid, cuda), flow_kind="grid", layout=layout, force_generic=fg)))
            for fg in (3, 1)]
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    if bf16:
        assert_close_bf16(outs[0], want, what="nhwc tma sampler")
    else:
        assert_close_f32(outs[0], want, scale=np.abs(key).max(), what="nhwc tma sampler")


def test_plane_generic_nhwc_identical_bits(ops, cuda):
    """The three f32 kernels evaluate the same fmaf chain: results must agree bit for bit."""
    d = make_case(33, 3, 64, 38, 63, with_bypass=True)
    a = host(run_fused(ops, cuda, d, "logits", "nchw", force_generic=3))      # all-TMA
    b = host(run_fused(ops, cuda, d, "logits", "nchw", force_generic=1))      # generic gather
    c = host(run_fused(ops, cuda, d, "logits", "nhwc_f32"))
    e = host(run_fused(ops, cuda, d, "logits", "nchw", force_generic=2))      # plane-resident LDG/STG
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert np.array_equal(a.view(np.uint32), c.view(np.uint32))
    assert np.array_equal(a.view(np.uint32), e.view(np.uint32))


def test_unfused_chain_matches_fused(ops, cuda):
    d = make_case(44, 2, 32, 38, 63)
    want = oracle_fused(d, O.W_LOGITS)
    got = host(ops.unfused_chain(dev(d["key"], cuda), dev(d["flow"], cuda), dev(d["scale_map"], cuda),
                                 dev(d["cur"], cuda), dev(d["logits"], cuda)))
    assert_close_f32(got, want, scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()), what="unfused")


def test_cosine_logits_op(ops, cuda):
    rng = np.random.default_rng(2)
    ew = rng.standard_normal((2, 2048, 9, 7), dtype=np.float32)
    ec = rng.standard_normal((2, 2048, 9, 7), dtype=np.float32)
    ec[:, :, 0, 0] = 0
    want = np.concatenate([O.cosine_weight(ew, ec), O.cosine_weight(ec, ec)], axis=1)
    got = host(ops.cosine_logits(dev(ew, cuda), dev(ec, cuda), "nchw", workspace=False))       # LDG kernel
    assert np.abs(got - want).max() < 2e-6
    ew2, ec2 = np.ascontiguousarray(ew[:, :, :, :6]), np.ascontiguousarray(ec[:, :, :, :6])    # 9x6: even plane -> all-TMA pre-pass
    want2 = np.concatenate([O.cosine_weight(ew2, ec2), O.cosine_weight(ec2, ec2)], axis=1)
    got2 = host(ops.cosine_logits(dev(ew2, cuda), dev(ec2, cuda), "nchw"))
    assert np.abs(got2 - want2).max() < 2e-6
    assert np.array_equal(got2, host(ops.cosine_logits(dev(ew2, cuda), dev(ec2, cuda), "nchw")))   # deterministic
    got = host(ops.cosine_logits(ops.to_nhwc(dev(ew, cuda)), ops.to_nhwc(dev(ec, cuda)), "nhwc_f32"))
    assert np.abs(got - want).max() < 2e-6
    assert got[0, 0, 0, 0] == 0.0 and got[0, 1, 0, 0] == 0.0     # eps path: 0/sqrt(1e-10)


def test_choose_feat(ops, cuda):
    rng = np.random.default_rng(8)
    a = rng.standard_normal((4, 8, 5, 6), dtype=np.float32)
    b = rng.standard_normal((4, 8, 5, 6), dtype=np.float32)
    flag = np.array([1, 0, 0, 1], dtype=np.uint8)
    got = host(ops.ChooseFeat(dev(a, cuda), dev(b, cuda), dev(flag, cuda)))
    assert np.array_equal(got, O.choose_feat(a, b, flag))


# ------------------------------------------------------------------------------------------
# known-answer cases (SURVEY.md 8c i-ix)
# ------------------------------------------------------------------------------------------
def test_known_answers(ops, cuda):
    rng = np.random.default_rng(0)
    N, C, H, W = 2, 16, 38, 63
    key = O.synth_features(rng, (N, C, H, W))
    cur = O.synth_features(rng, (N, C, H, W))
    ones = np.ones((N, C, H, W), np.float32)
    K, Cu = dev(key, cuda), dev(cur, cuda)
    # (i) zero flow: out == data up to the floor-flip epsilon
    z = np.zeros((N, 2, H, W), np.float32)
    got = host(ops.warp_scale_aggregate(K, dev(z, cuda)))
    assert np.abs(got - key).max() <= 4e-6 * np.abs(key).max()
    # (ii) integer flow = pure shift with zero padding
    f = z.copy(); f[:, 0] = 2.0; f[:, 1] = -1.0
    got = host(ops.warp_scale_aggregate(K, dev(f, cuda)))
    want = np.zeros_like(key); want[:, :, 1:, :W - 2] = key[:, :, :H - 1, 2:]
    assert np.abs(got - want).max() <= 4e-6 * np.abs(key).max()
    # (iii) half-cell flow: exact 2-tap mean in x
    f = z.copy(); f[:, 0] = 0.5
    got = host(ops.warp_scale_aggregate(K, dev(f, cuda)))
    want = np.zeros_like(key); want[..., :W - 1] = 0.5 * (key[..., :W - 1] + key[..., 1:]); want[..., W - 1] = 0.5 * key[..., W - 1]
    assert np.abs(got - want).max() <= 4e-6 * np.abs(key).max()
    # (iv) fully out of bounds -> zeros
    f = z.copy(); f[:, 0] = 500.0
    assert float(np.abs(host(ops.warp_scale_aggregate(K, dev(f, cuda)))).max()) == 0.0
    # (v)+(vi) scale == 1, equal logits -> plain average
    lg = np.zeros((N, 2, H, W), np.float32) + 0.3
    f = (rng.standard_normal((N, 2, H, W)) * 2).astype(np.float32)
    a = host(ops.warp_scale_aggregate(K, dev(f, cuda), cur=Cu, scale_map=dev(ones, cuda), weight_mode="logits", logits=dev(lg, cuda)))
    b = host(ops.warp_scale_aggregate(K, dev(f, cuda), cur=Cu, weight_mode="mean"))
    assert np.abs(a - b).max() <= 1e-6 * max(np.abs(key).max(), np.abs(cur).max())
    # (viii) bypass -> cur, bit for bit
    byp = np.array([1, 0], np.uint8)
    a = host(ops.warp_scale_aggregate(K, dev(f, cuda), cur=Cu, weight_mode="mean", bypass=dev(byp, cuda)))
    assert np.array_equal(a[0], cur[0]) and not np.array_equal(a[1], cur[1])
    # (ix) centre-2x2 vs avg-16 pooling differ on a non-constant field
    mv = rng.integers(-40, 41, size=(1, 64, 64, 2), dtype=np.int32)
    p0 = host(ops.mv_pool(dev(mv, cuda), 1.0, "centre2x2")); p1 = host(ops.mv_pool(dev(mv, cuda), 1.0, "avg16"))
    assert not np.array_equal(p0, p1)


def test_fused_golden_fixture(ops, cuda):
    """Committed golden vectors (tests/golden/fused_small.npz, made by tools/make_golden_fused.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fused_small.npz"))
    d = {k: g[k] for k in g.files}
    for mode_name, mode in MODES:
        want = d["out_" + mode_name]
        for layout in ("nchw", "nhwc_f32"):
            got = host(run_fused(ops, cuda, d, mode_name, layout))
            assert_close_f32(got, want, scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()),
                             what="golden %s %s" % (mode_name, layout))


# ------------------------------------------------------------------------------------------
# BASELINE.json sizes: oracle on 2 frames + size-independent properties on the full batch
# ------------------------------------------------------------------------------------------
def test_full_size_config2_against_oracle(ops, cuda):
    d = make_case(2026, 2, 1024, 38, 63, raw=True)
    want = oracle_fused(d, O.W_LOGITS)
    got = host(run_fused(ops, cuda, d, "logits", "nchw"))
    assert_close_f32(got, want, scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()), what="1024x38x63 nchw")
    got = host(run_fused(ops, cuda, d, "logits", "nhwc_f32"))
    assert_close_f32(got, want, scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()), what="1024x38x63 nhwc")


def test_full_size_config4_against_oracle(ops, cuda):
    d = make_case(4, 1, 1024, 68, 120, raw=True, max_px=96)
    want = oracle_fused(d, O.W_LOGITS)
    got = host(run_fused(ops, cuda, d, "logits", "nchw"))
    assert_close_f32(got, want, scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()), what="1024x68x120")


def test_full_batch_properties(ops, cuda):
    """Config 2 at full batch (64 x 1024 x 38 x 63): linearity in the key feature, bypass identity,
    and agreement of the plane-resident kernel with the generic one - no oracle needed."""
    N, C, H, W = 64, 1024, 38, 63
    g = torch.Generator(device=cuda).manual_seed(1)
    key1 = torch.rand((N, C, H, W), device=cuda, generator=g)
    key2 = torch.rand((N, C, H, W), device=cuda, generator=g)
    cur = torch.rand((N, C, H, W), device=cuda, generator=g)
    mv = dev(O.synth_raw_mv(np.random.default_rng(1), N, 600, 1000, 96), cuda)
    logits = torch.randn((N, 2, H, W), device=cuda, generator=g)
    w1 = ops.warp_scale_aggregate(key1, mv, flow_kind="raw")
    w2 = ops.warp_scale_aggregate(key2, mv, flow_kind="raw")
    w12 = ops.warp_scale_aggregate(key1 + key2, mv, flow_kind="raw")
    assert float((w12 - (w1 + w2)).abs().max()) <= 1e-5
    full = ops.warp_scale_aggregate(key1, mv, flow_kind="raw", cur=cur, weight_mode="logits", logits=logits)
    gen = ops.warp_scale_aggregate(key1, mv, flow_kind="raw", cur=cur, weight_mode="logits", logits=logits,
                                   force_generic=1)
    assert torch.equal(full, gen)
    pl = ops.warp_scale_aggregate(key1, mv, flow_kind="raw", cur=cur, weight_mode="logits", logits=logits,
                                  force_generic=2)
    assert torch.equal(full, pl)
    del gen, pl
    # convexity: softmax weights sum to 1 -> out between min and max of the two sources
    lo = torch.minimum(w1, cur) - 1e-5
    hi = torch.maximum(w1, cur) + 1e-5
    assert bool(((full >= lo) & (full <= hi)).all())
    byp = torch.ones(N, dtype=torch.uint8, device=cuda)
    same = ops.warp_scale_aggregate(key1, mv, flow_kind="raw", cur=cur, weight_mode="logits", logits=logits, bypass=byp)
    assert torch.equal(same, cur)


def test_errors_are_loud(ops, cuda):
    from lsfa_b200 import LsfaError
    key = torch.zeros((1, 8, 4, 4), device=cuda)
    flow = torch.zeros((1, 2, 4, 4), device=cuda)
    with pytest.raises(LsfaError):
        ops.warp_scale_aggregate(key, flow, weight_mode="logits", cur=key)      # logits missing
    with pytest.raises(ValueError):
        ops.warp_scale_aggregate(key.cpu(), flow)                               # no CPU path
    with pytest.raises(NotImplementedError):
        ops.GridGenerator(flow, transform_type="affine")
    with pytest.raises(LsfaError):
        ops.warp_scale_aggregate(torch.zeros((1, 4, 4, 6), device=cuda, dtype=torch.bfloat16), flow,
                                 layout="nhwc_bf16")                            # C % 8 != 0


def test_torch_library_ops_dispatch_to_cuda(ops, cuda):
    """torch.ops.lsfa.* (the custom-op harness) run the same kernels as lsfa_b200.ops."""
    import lsfa_b200.torch_ops  # noqa: F401
    d = make_case(77, 2, 16, 38, 63)
    t = lambda k: dev(d[k], cuda)  # noqa: E731
    grid = torch.ops.lsfa.grid_generator_warp(t("flow"))
    warped = torch.ops.lsfa.bilinear_sampler(t("key"), grid)
    assert_close_f32(host(warped), O.warp(d["key"], d["flow"]), scale=np.abs(d["key"]).max(), what="torch.ops warp")
    assert np.array_equal(host(torch.ops.lsfa.mv_pool(t("mv"), 1.0, 0)), d["flow"])
    out = torch.ops.lsfa.warp_scale_aggregate(t("key"), t("mv"), t("cur"), t("scale_map"), t("logits"), None, 3, 2, 1.0, 0)
    assert_close_f32(host(out), oracle_fused(d, O.W_LOGITS), scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()),
                     what="torch.ops fused")


def test_host_aggregator_pipeline(ops, cuda):
    """The host-buffer front end (pinned in/out, chunked 3-stream pipeline) gives the same result."""
    from lsfa_b200.host import HostAggregator
    N, C, H, W = 7, 32, 38, 63
    d = make_case(55, N, C, H, W)
    host_in = {k: torch.from_numpy(d[k]).pin_memory() for k in ("key", "cur", "scale_map", "mv", "logits")}
    out_host = torch.empty((N, C, H, W), dtype=torch.float32).pin_memory()
    agg = HostAggregator(N, C, H, W, d["mv"].shape[1:3], cuda, chunk=2, depth=2)
    for _ in range(3):      # slots are reused across calls
        agg(host_in, out_host)
    agg.synchronize()
    assert_close_f32(out_host.numpy(), oracle_fused(d, O.W_LOGITS),
                     scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()), what="host pipeline")


@pytest.mark.parametrize("h", [600, 608, 599, 24, 23, 9, 8, 7, 3])
def test_mv_centre_rows_h2d_copies_exactly_what_the_reduction_reads(ops, cuda, h):
    """lsfa_mv_centre_rows_h2d moves rows 16k+7 and 16k+8 (those below h) and nothing else; pooling the partial device
    image gives the bits of pooling the whole field."""
    import ctypes
    from lsfa_b200 import _cabi as A
    w, N = 50, 3
    rng = np.random.default_rng(h)
    mv = rng.integers(-64, 65, size=(N, h, w, 2), dtype=np.int32)
    hmv = torch.from_numpy(mv).pin_memory()
    dmv = torch.full((N, h, w, 2), 12345, dtype=torch.int32, device=cuda)
    nbytes = ctypes.c_size_t(0)
    A.check(A.load().lsfa_mv_centre_rows_h2d(hmv.data_ptr(), dmv.data_ptr(), N, h, w, ctypes.cast(ctypes.byref(nbytes), ctypes.c_void_p),
                                             torch.cuda.current_stream().cuda_stream))
    got = host(dmv).astype(np.int32)
    rows = [r for r in range(h) if r % 16 in (7, 8)]
    assert nbytes.value == N * len(rows) * w * 8
    want = np.full_like(mv, 12345)
    want[:, rows] = mv[:, rows]
    assert np.array_equal(got, want)
    if rows:
        a = host(ops.mv_pool(dmv))
        b = host(ops.mv_pool(dev(mv, cuda)))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_host_aggregator_ragged_mv(ops, cuda):
    """Host pipeline with an MV field whose height is not a multiple of 16 (the last block keeps row 16k+7 only)."""
    from lsfa_b200.host import HostAggregator
    N, C, H, W = 3, 16, 12, 9
    d = make_case(56, N, C, H, W, raw="ragged")
    host_in = {k: torch.from_numpy(d[k]).pin_memory() for k in ("key", "cur", "scale_map", "mv", "logits")}
    out_host = torch.empty((N, C, H, W), dtype=torch.float32).pin_memory()
    agg = HostAggregator(N, C, H, W, d["mv"].shape[1:3], cuda, chunk=2, depth=2)
    agg(host_in, out_host)
    agg.synchronize()
    assert_close_f32(out_host.numpy(), oracle_fused(d, O.W_LOGITS),
                     scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()), what="host pipeline, ragged mv")
    bi, _ = agg.bytes_per_call()
    assert bi == N * (3 * C * H * W * 4 + agg.mv_rows_copied() * d["mv"].shape[2] * 8 + 2 * H * W * 4)


def test_exact_two_phase_key_frame_graphs(ops, cuda):
    """get_key_test_symbol end to end (SYM:468-477): K1 warp x scale (lsfa) -> embedding / Nq convolutions
    (library GEMMs: cuDNN) -> K2 cosine + softmax blend (lsfa), against the oracle with NumPy convolutions."""
    from lsfa_b200 import graphs
    rng = np.random.default_rng(4)
    N, C, H, W = 2, 16, 12, 14
    d = make_case(4, N, C, H, W)
    first = np.array([0, 1], np.uint8)
    mk = lambda *shape: (0.2 * rng.standard_normal(shape)).astype(np.float32)  # noqa: E731
    emb = (mk(8, C, 1, 1), mk(8), mk(8, 8, 3, 3), mk(8), mk(32, 8, 1, 1), mk(32))
    nq = (mk(8, C, 3, 3), mk(8), mk(4, 8, 1, 1), mk(4), mk(1, 4, 1, 1), mk(1))
    t = lambda a: dev(a, cuda)  # noqa: E731
    # TF32 would break the 1e-5 gate on the convolutions: pin cuDNN to fp32 for this check
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        got_f = graphs.key_frame_fgfa(t(d["key"]), t(d["mv"]), t(d["scale_map"]), t(d["cur"]), [t(a) for a in emb],
                                      is_first_frame=t(first), flow_kind="raw")
        got_q = graphs.key_frame_nq(t(d["key"]), t(d["mv"]), t(d["scale_map"]), t(d["cur"]), [t(a) for a in nq],
                                    is_first_frame=t(first), flow_kind="raw")
    finally:
        torch.backends.cudnn.allow_tf32 = old
    want_f = O.key_frame_fgfa_full(d["key"], d["flow"], d["scale_map"], d["cur"], emb, first)
    want_q = O.key_frame_nq_full(d["key"], d["flow"], d["scale_map"], d["cur"], nq, first)
    scale = max(np.abs(d["key"]).max(), np.abs(d["cur"]).max())
    # the convolutions are library code with their own summation order: 1e-4 relative on the blend weights
    assert np.abs(host(got_f) - want_f).max() <= 2e-4 * scale
    assert np.abs(host(got_q) - want_q).max() <= 2e-4 * scale
    assert np.array_equal(host(got_f)[1], d["cur"][1]) and np.array_equal(host(got_q)[1], d["cur"][1])
    # the same graphs with the library convolutions in bf16 channels-last (cuDNN's tensor-core path, 3x faster at full
    # size): only the blend weights move, by the bf16 rounding of the embeddings / quality logits
    got_fb = graphs.key_frame_fgfa(t(d["key"]), t(d["mv"]), t(d["scale_map"]), t(d["cur"]), [t(a) for a in emb],
                                   is_first_frame=t(first), flow_kind="raw", conv_dtype=torch.bfloat16)
    got_qb = graphs.key_frame_nq(t(d["key"]), t(d["mv"]), t(d["scale_map"]), t(d["cur"]), [t(a) for a in nq],
                                 is_first_frame=t(first), flow_kind="raw", conv_dtype=torch.bfloat16)
    assert np.abs(host(got_fb) - want_f).max() <= 1e-2 * scale
    assert np.abs(host(got_qb) - want_q).max() <= 1e-2 * scale


def test_stream_scheduler_batches_non_key_frames_of_many_streams(ops, cuda):
    """The batched driver (one key feature per stream, key_index per frame) equals the reference's
    frame-by-frame loop: every non-key frame of every stream is warped from ITS stream's key feature."""
    from lsfa_b200.driver import StreamScheduler
    rng = np.random.default_rng(12)
    seg_lens = [14, 26, 13]
    C, H, W = 16, 10, 12
    sch = StreamScheduler(seg_lens, C, (H, W), cuda)
    keys = {s: O.synth_features(rng, (1, C, H, W)) for s in sch.stream_ids}
    for s in sch.stream_ids:
        sch.set_key_feature(s, dev(keys[s], cuda))
    rw = (0.01 * rng.standard_normal((C, 3))).astype(np.float32); rb = (0.01 * rng.standard_normal(C)).astype(np.float32)
    seen = 0
    for sid, fid, slot in sch.batches(16):
        B = len(sid)
        mv = O.synth_raw_mv(rng, B, 16 * H, 16 * W, 32)
        res = O.res_pool(rng.integers(-64, 65, size=(B, 16 * H, 16 * W, 3), dtype=np.int32))
        cur = O.synth_features(rng, (B, C, H, W))
        got = host(sch.run_non_key_batch(slot, dev(mv, cuda), dev(cur, cuda), res=dev(res, cuda), rnet_w=dev(rw, cuda),
                                         rnet_b=dev(rb, cuda)))
        for i in range(B):   # the reference's per-frame call (tester.py:251-253 -> SYM:570-586)
            want = O.cur_frame_path(keys[int(sid[i])], O.mv_pool(mv[i:i + 1]), res[i:i + 1], rw, rb, cur[i:i + 1])
            assert_close_f32(got[i:i + 1], want, scale=max(np.abs(cur).max(), 5.0), what="stream %d frame %d" % (sid[i], fid[i]))
        seen += B
    assert seen == sum(int((sch.flags[s] == 2).sum()) for s in sch.stream_ids)


def test_config1_single_frame_full_size(ops, cuda):
    """BASELINE.json configs[0]: one non-key frame, 1x1024x38x63 fp32 key feature + a 600x1000 MV field,
    GridGenerator(warp) + BilinearSampler as two drop-in operators."""
    rng = np.random.default_rng(1)
    key = O.synth_features(rng, (1, 1024, 38, 63))
    mv = O.synth_raw_mv(rng, 1, 600, 1000, 96)
    flow = ops.mv_pool(dev(mv, cuda))
    assert np.array_equal(host(flow), O.mv_pool(mv))
    grid = ops.GridGenerator(flow, transform_type="warp")
    out = ops.BilinearSampler(dev(key, cuda), grid)
    want_grid = O.grid_generator_warp(O.mv_pool(mv))
    assert np.array_equal(host(grid).view(np.uint32), want_grid.view(np.uint32))
    assert_close_f32(host(out), O.bilinear_sampler(key, want_grid), scale=np.abs(key).max(), what="config 1")


# ------------------------------------------------------------------------------------------
# backward of a7 / a8
# ------------------------------------------------------------------------------------------
def _bwd_case(seed, N, C, Hi, Wi, Ho, Wo, spread=1.15):
    rng = np.random.default_rng(seed)
    data = rng.standard_normal((N, C, Hi, Wi), dtype=np.float32)
    grid = ((rng.random((N, 2, Ho, Wo), dtype=np.float32) * 2 - 1) * np.float32(spread)).astype(np.float32)
    og = rng.standard_normal((N, C, Ho, Wo), dtype=np.float32)
    return data, grid, og


@pytest.mark.parametrize("kernel", ["scatter", "gather"])
@pytest.mark.parametrize("shape", [(2, 8, 38, 63, 38, 63), (3, 6, 12, 20, 9, 14), (1, 4, 20, 24, 30, 36), (2, 3, 7, 8, 5, 8),
                                   (1, 2, 60, 72, 60, 72)])
def test_bilinear_sampler_backward(ops, cuda, kernel, shape):
    N, C, Hi, Wi, Ho, Wo = shape
    data, grid, og = _bwd_case(sum(shape), *shape)
    want_d, want_g = O.bilinear_sampler_backward(data, grid, og)
    if kernel == "gather" and ((C % 2 or (Hi * Wi) % 2 or (Ho * Wo) % 2) and ((Hi * Wi) % 4 or (Ho * Wo) % 4)):
        with pytest.raises(Exception):
            ops.BilinearSampler_backward(dev(data, cuda), dev(grid, cuda), dev(og, cuda), kernel=kernel)
        return
    gd, gg = ops.BilinearSampler_backward(dev(data, cuda), dev(grid, cuda), dev(og, cuda), kernel=kernel)
    assert_close_f32(host(gd), want_d, scale=np.abs(want_d).max(), what="grad_data " + kernel)
    assert_close_f32(host(gg), want_g, scale=np.abs(want_g).max(), what="grad_grid " + kernel)
    # one gradient at a time (SYM:320-321 needs grad_data only), and kAddTo on both
    gd1, none = ops.BilinearSampler_backward(dev(data, cuda), dev(grid, cuda), dev(og, cuda), req_grid="null", kernel=kernel)
    assert none is None
    assert_close_f32(host(gd1), want_d, scale=np.abs(want_d).max(), what="grad_data only")
    none, gg1 = ops.BilinearSampler_backward(dev(data, cuda), dev(grid, cuda), dev(og, cuda), req_data="null", kernel=kernel)
    assert none is None
    assert_close_f32(host(gg1), want_g, scale=np.abs(want_g).max(), what="grad_grid only")
    base_d = np.full_like(want_d, 0.5)
    base_g = np.full_like(want_g, -0.25)
    gd2, gg2 = ops.BilinearSampler_backward(dev(data, cuda), dev(grid, cuda), dev(og, cuda), grad_data=dev(base_d, cuda),
                                            grad_grid=dev(base_g, cuda), req_data="add", req_grid="add", kernel=kernel)
    assert_close_f32(host(gd2), want_d + base_d, scale=np.abs(want_d).max(), what="grad_data kAddTo")
    assert_close_f32(host(gg2), want_g + base_g, scale=np.abs(want_g).max(), what="grad_grid kAddTo")


def test_gather_backward_is_deterministic_and_matches_scatter(ops, cuda):
    """Lists that fit their slots (here: a sub-cell flow per frame, 4 entries per input pixel) are summed in a
    fixed order: bit-identical from run to run, unlike the atomic scatter."""
    data, _, og = _bwd_case(5, 4, 16, 38, 63, 38, 63)
    flow = np.zeros((4, 2, 38, 63), np.float32)
    # (an exactly integer flow is the one case that overflows the slots: the fp32 grid round trip lands a hair
    # below some integers, and a 3x3 neighbourhood then reaches the pixel with weights of 1e-6 .. 1e-13)
    for n, (fx, fy) in enumerate(((0.37, -0.21), (-1.6, 0.4), (2.25, 3.5), (0.5, 0.5))):
        flow[n, 0], flow[n, 1] = fx, fy
    a, fa = ops.warp_backward(dev(data, cuda), dev(flow, cuda), dev(og, cuda), kernel="gather")
    b, fb = ops.warp_backward(dev(data, cuda), dev(flow, cuda), dev(og, cuda), kernel="gather")
    assert np.array_equal(host(a).view(np.uint32), host(b).view(np.uint32))
    c, fc = ops.warp_backward(dev(data, cuda), dev(flow, cuda), dev(og, cuda), kernel="scatter")
    assert_close_f32(host(a), host(c), scale=np.abs(host(c)).max(), what="gather vs scatter")
    assert_close_f32(host(fa), host(fc), scale=np.abs(host(fc)).max(), what="gather vs scatter grad_flow")
    want_k, want_f = O.warp_backward(data, flow, og)
    assert_close_f32(host(a), want_k, scale=np.abs(want_k).max(), what="gather vs oracle")


def test_gather_backward_overflowing_lists(ops, cuda):
    """Every output pixel samples the same spot: one input pixel's list has H*W entries (far beyond the
    L slots); the overflow fix-up must deliver all of them."""
    rng = np.random.default_rng(6)
    N, C, H, W = 2, 4, 38, 63
    data = rng.standard_normal((N, C, H, W), dtype=np.float32)
    og = rng.standard_normal((N, C, H, W), dtype=np.float32)
    grid = np.zeros((N, 2, H, W), np.float32)
    grid[:, 0] = 0.013
    grid[:, 1] = -0.2
    want_d, want_g = O.bilinear_sampler_backward(data, grid, og)
    gd, gg = ops.BilinearSampler_backward(dev(data, cuda), dev(grid, cuda), dev(og, cuda), kernel="gather")
    # ~2400 terms per sum: the gate scales with the number of addends
    assert np.abs(host(gd) - want_d).max() <= 2e-5 * np.abs(want_d).max()
    assert_close_f32(host(gg), want_g, scale=np.abs(want_g).max(), what="grad_grid")


def test_gather_backward_is_deterministic_when_lists_overflow(ops, cuda):
    """Lists longer than their slots are rebuilt in (p, tap) order by the pre-pass and their tail is added by one thread
    per channel in that order: bit-identical from run to run for ANY sampling map - every pixel sampling one spot (a list
    of H*W entries), integer flows (3x3 neighbourhoods with weights of 1e-6..1e-13), and +-6-cell random motion."""
    rng = np.random.default_rng(61)
    N, C, H, W = 3, 8, 38, 63
    data = rng.standard_normal((N, C, H, W), dtype=np.float32)
    og = rng.standard_normal((N, C, H, W), dtype=np.float32)
    grid = np.zeros((N, 2, H, W), np.float32)
    grid[0, 0], grid[0, 1] = 0.013, -0.2                                   # one spot
    grid[1] = rng.uniform(-0.3, 0.3, size=(2, H, W)).astype(np.float32)    # everything lands in a third of the plane
    grid[2] = rng.uniform(-1.1, 1.1, size=(2, H, W)).astype(np.float32)
    runs = [ops.BilinearSampler_backward(dev(data, cuda), dev(grid, cuda), dev(og, cuda), kernel="gather") for _ in range(3)]
    for gd, gg in runs[1:]:
        assert np.array_equal(host(gd).view(np.uint32), host(runs[0][0]).view(np.uint32)), "grad_data differs between runs"
    want_d, _ = O.bilinear_sampler_backward(data, grid, og)
    assert np.abs(host(runs[0][0]) - want_d).max() <= 2e-5 * np.abs(want_d).max()
    flow = np.round(rng.uniform(-6, 6, size=(N, 2, H, W))).astype(np.float32)   # integer flows
    a = [ops.warp_backward(dev(data, cuda), dev(flow, cuda), dev(og, cuda), kernel="gather")[0] for _ in range(3)]
    assert np.array_equal(host(a[0]).view(np.uint32), host(a[1]).view(np.uint32))
    assert np.array_equal(host(a[0]).view(np.uint32), host(a[2]).view(np.uint32))
    want_k, _ = O.warp_backward(data, flow, og)
    assert_close_f32(host(a[0]), want_k, scale=np.abs(want_k).max(), what="integer flows, gather vs oracle")


def test_recorded_graph_through_the_c_abi_replays_a_frame(ops, cuda):
    """lsfa_graph_begin / _end / _launch: the three drop-in operators and the fused op of ONE frame (BASELINE configs[0])
    recorded on a side stream, replayed with new input contents in the same buffers; results = the eager calls, bit for bit.
    The one-launch cooperative form of the fused kernel is part of the recording."""
    d = make_case(31, 1, 64, 38, 63)
    key, mv, cur, sm, lg = (dev(d[k], cuda) for k in ("key", "mv", "cur", "scale_map", "logits"))
    side = torch.cuda.Stream()
    st = side.cuda_stream
    grid = torch.empty((1, 2, 38, 63), device=cuda)
    flow = torch.empty((1, 2, 38, 63), device=cuda)
    out_s = torch.empty_like(key)
    with torch.cuda.stream(side):
        p = ops.PreparedAggregate(key, mv, flow_kind="raw", cur=cur, scale_map=sm, weight_mode="logits", logits=lg)
        for _ in range(2):                                  # warm-up outside the recording (attribute opt-ins happen once)
            p.run(st)
            ops.BilinearSampler(key, ops.GridGenerator(ops.mv_pool(mv, out=flow), out=grid), out=out_s)
        side.synchronize()
        g = ops.RecordedGraph(st)
        with g:
            p.run(st)
            ops.BilinearSampler(key, ops.GridGenerator(ops.mv_pool(mv, out=flow), out=grid), out=out_s)
        for rep in range(3):
            d2 = make_case(40 + rep, 1, 64, 38, 63)
            for t, k in ((key, "key"), (mv, "mv"), (cur, "cur"), (sm, "scale_map"), (lg, "logits")):
                t.copy_(dev(d2[k], cuda))
            g.launch()
            side.synchronize()
            got_f, got_s = p.out.clone(), out_s.clone()
            want_f = ops.warp_scale_aggregate(key, mv, flow_kind="raw", cur=cur, scale_map=sm, weight_mode="logits", logits=lg)
            want_s = ops.BilinearSampler(key, ops.GridGenerator(ops.mv_pool(mv)))
            side.synchronize()
            assert torch.equal(got_f, want_f) and torch.equal(got_s, want_s), rep
            assert_close_f32(host(got_f), oracle_fused(d2, O.W_LOGITS), scale=max(np.abs(d2["key"]).max(), np.abs(d2["cur"]).max()), what="replay %d" % rep)
        g.close()


def test_warp_backward_and_grid_generator_backward(ops, cuda):
    d = make_case(21, 3, 8, 38, 63, max_px=96)
    rng = np.random.default_rng(22)
    og = rng.standard_normal(d["cur"].shape, dtype=np.float32)
    flow = d["flow"].copy()
    flow[0, :, 0, :4] = [[1000.0, -1000.0, 0.5, -0.5]] * 2          # out of the plane / half-cell flows
    want_k, want_f = O.warp_backward(d["key"], flow, og)
    for kernel in ("gather", "scatter"):
        gk, gf = ops.warp_backward(dev(d["key"], cuda), dev(flow, cuda), dev(og, cuda), kernel=kernel)
        assert_close_f32(host(gk), want_k, scale=np.abs(want_k).max(), what="grad_key " + kernel)
        assert_close_f32(host(gf), want_f, scale=np.abs(want_f).max(), what="grad_flow " + kernel)
    # the same through the two drop-in operators
    grid = ops.GridGenerator(dev(flow, cuda))
    gk2, gg = ops.BilinearSampler_backward(dev(d["key"], cuda), grid, dev(og, cuda))
    gf2 = ops.GridGenerator_backward(gg)
    assert_close_f32(host(gk2), want_k, scale=np.abs(want_k).max(), what="two-op grad_key")
    assert_close_f32(host(gf2), want_f, scale=np.abs(want_f).max(), what="two-op grad_flow")
    g = rng.standard_normal((2, 2, 38, 63), dtype=np.float32)
    assert np.array_equal(host(ops.GridGenerator_backward(dev(g, cuda))), O.grid_generator_warp_backward(g))


def test_backward_full_size_adjoint_property(ops, cuda):
    """BASELINE size (1024x38x63): <og, forward(data)> == <grad_data, data> - the backward is the exact
    transpose of the forward operator, checked without an oracle pass over the full tensors."""
    rng = np.random.default_rng(8)
    N, C, H, W = 4, 1024, 38, 63
    data = torch.from_numpy(rng.standard_normal((N, C, H, W), dtype=np.float32)).to(cuda)
    og = torch.from_numpy(rng.standard_normal((N, C, H, W), dtype=np.float32)).to(cuda)
    mv = O.synth_raw_mv(rng, N, 600, 1000, 96)
    flow = ops.mv_pool(dev(mv, cuda))
    out = ops.warp_scale_aggregate(data, flow)
    gk, gf = ops.warp_backward(data, flow, og)
    lhs = float((og.double() * out.double()).sum())
    rhs = float((gk.double() * data.double()).sum())
    assert abs(lhs - rhs) <= 1e-6 * (og.double().abs() * out.double().abs()).sum().item()
    assert torch.isfinite(gf).all()


def test_torch_autograd_through_the_custom_ops(ops, cuda):
    from lsfa_b200 import torch_ops as T
    data, grid, og = _bwd_case(9, 2, 4, 38, 63, 38, 63)
    flow = (np.random.default_rng(10).standard_normal((2, 2, 38, 63)) * 2).astype(np.float32)
    td = dev(data, cuda).requires_grad_(True)
    tf = dev(flow, cuda).requires_grad_(True)
    out = T.bilinear_sampler(td, T.grid_generator_warp(tf))
    out.backward(dev(og, cuda))
    want_k, want_f = O.warp_backward(data, flow, og)
    assert_close_f32(host(td.grad), want_k, scale=np.abs(want_k).max(), what="autograd grad_data")
    assert_close_f32(host(tf.grad), want_f, scale=np.abs(want_f).max(), what="autograd grad_flow")


@pytest.mark.parametrize("shape", [(5, 8, 68, 120), (5, 4, 100, 132), (6, 6, 60, 80)])
def test_row_trimmed_key_loads_any_motion(ops, cuda, shape):
    """Planes cut into pixel parts load only the key rows each part's taps read (row ranges from the record
    pre-pass).  Whatever the motion - huge flows, everything out of the plane, a still frame, bypass frames,
    a shared key - the result must be bit-identical to the untrimmed kernel and pass the oracle gate."""
    N, C, H, W = shape
    d = make_case(90 + N, N, C, H, W, with_bypass=True, shared_key=(N == 4))
    rng = np.random.default_rng(N)
    flow = (rng.standard_normal((N, 2, H, W)) * 25).astype(np.float32)      # taps all over the plane
    flow[0] = 0.0                                                          # still frame: each part needs its own rows only
    if N > 2:
        flow[2] = 500.0                                                    # everything out of the plane
        flow[2, :, H // 2:] = rng.uniform(-3, 3, size=(2, H - H // 2, W))   # ... except the second part: local motion
    d["flow"] = flow
    want = oracle_fused(d, O.W_LOGITS)
    got = run_fused(ops, cuda, d, "logits", "nchw", flow_kind="flow", force_generic=3)
    assert_close_f32(host(got), want, scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()), what="trimmed %s" % (shape,))
    full = run_fused(ops, cuda, d, "logits", "nchw", flow_kind="flow", force_generic=2)
    assert torch.equal(got, full)


@pytest.mark.parametrize("variant", ["warp", "scale_cur", "res_cur"])
@pytest.mark.parametrize("shape", [(10, 8, 38, 63), (10, 4, 68, 120)])
def test_cooperative_one_launch_form_is_bit_identical(ops, cuda, variant, shape):
    """Batches of up to 8 (frame, pixel part) pairs run as ONE cooperative launch (the kernel's own consumers build the
    sampling records before a grid-wide barrier; BASELINE configs[0] is this mode); larger batches use the record pre-pass.
    The same frames through both forms must agree bit for bit, and the launch accounting must say 1 vs 2."""
    N, C, H, W = shape
    d = make_case(700 + C, N, C, H, W, with_res=True, with_bypass=True)
    kw = {"warp": dict(mode="none", use_scale=False), "scale_cur": dict(mode="logits", use_scale=True),
          "res_cur": dict(mode="add", use_scale=False, use_res=True)}[variant]
    big = run_fused(ops, cuda, d, kw["mode"], "nchw", use_scale=kw["use_scale"], use_res=kw.get("use_res", False), force_generic=3)
    for lo in (0, 3, 7):
        hi = lo + 3 if H * W <= 4320 else lo + 2        # <= 8 virtual frames
        sub = {k: (v[lo:hi] if isinstance(v, np.ndarray) and v.shape[:1] == (N,) else v) for k, v in d.items()}
        small = run_fused(ops, cuda, sub, kw["mode"], "nchw", use_scale=kw["use_scale"], use_res=kw.get("use_res", False), force_generic=3)
        assert torch.equal(small, big[lo:hi]), (variant, lo)


def test_backward_golden_fixture_from_torch_autograd(ops, cuda):
    """Committed vectors minted from an independent implementation (tools/make_golden_backward.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "backward_small.npz"))
    for kernel in ("gather", "scatter"):
        gk, gf = ops.warp_backward(dev(g["key"], cuda), dev(g["flow"], cuda), dev(g["out_grad"], cuda), kernel=kernel)
        assert np.abs(host(gk) - g["grad_key"]).max() <= 4e-6 * np.abs(g["grad_key"]).max(), kernel
        assert np.abs(host(gf) - g["grad_flow"]).max() <= 4e-6 * np.abs(g["grad_flow"]).max(), kernel
    gk, gg = ops.BilinearSampler_backward(dev(g["key"], cuda), dev(g["grid"], cuda), dev(g["out_grad"], cuda))
    assert np.abs(host(gg) - g["grad_grid"]).max() <= 4e-6 * np.abs(g["grad_grid"]).max()
    out = ops.BilinearSampler(dev(g["key"], cuda), dev(g["grid"], cuda))
    assert np.abs(host(out) - g["out"]).max() <= 4e-6 * np.abs(g["out"]).max()


@pytest.mark.parametrize("N,E,H,W", [(3, 2048, 38, 63), (1, 512, 38, 63), (5, 6, 10, 12), (2, 64, 60, 72)])
def test_cosine_prepass_all_tma_vs_ldg_and_oracle(ops, cuda, N, E, H, W):
    """Fgfa weights in NCHW (SYM:111-116,132-148): with the full workspace the cosine pre-pass is the all-TMA kernel
    (static split, one partial per CTA and frame, slots added in order); it must agree with the LDG pre-pass
    (force_generic=1 pins it), be bit-identical from run to run, and pass the oracle gate.  N=1 spreads one frame
    over every CTA (many partial slots)."""
    d = make_case(300 + N, N, 8, H, W, E=E)
    want = oracle_fused(d, O.W_COSINE)
    scale = max(np.abs(d["key"]).max(), np.abs(d["cur"]).max())
    got = run_fused(ops, cuda, d, "cosine", "nchw")
    again = run_fused(ops, cuda, d, "cosine", "nchw")
    ldg = run_fused(ops, cuda, d, "cosine", "nchw", force_generic=1)
    assert torch.equal(got, again)
    assert_close_f32(host(got), want, scale=scale, what="cosine all-TMA pre-pass")
    assert_close_f32(host(ldg), want, scale=scale, what="cosine LDG pre-pass")
    assert ops.num_launches(key=dev(d["key"], cuda), flow=dev(d["flow"], cuda), cur=dev(d["cur"], cuda),
                            scale_map=dev(d["scale_map"], cuda), weight_mode="cosine", emb_warp=dev(d["emb_warp"], cuda),
                            emb_cur=dev(d["emb_cur"], cuda)) == 3     # cosine partials + finalize + fused (N <= 8: the
    # records are built inside the cooperative launch, no pre-pass kernel)
