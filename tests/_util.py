"""Shared helpers for the parity tests: seeded synthetic cases and the
parity gates.  The oracle is imported here and ONLY by tests / smoke / bench baseline."""
import numpy as np

from oracle import lsfa_oracle as O

RTOL_F32 = 1e-5          # north_star: features within 1e-5 relative error in fp32
ATOL_F32_FRAC = 1e-6     # ... + 1e-6 * max|data| (post-ReLU zeros)
RTOL_BF16 = 2.0 ** -8    # bf16 variant: stated tolerance
ATOL_BF16_FRAC = 2.0 ** -8


def assert_close_f32(got, want, scale=None, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = float(np.max(np.abs(want))) if scale is None else float(scale)
    tol = RTOL_F32 * np.abs(want) + ATOL_F32_FRAC * max(scale, 1e-30)
    err = np.abs(got - want)
    bad = err > tol
    assert not bad.any(), "%s: %d/%d outside |a-b| <= 1e-5|b| + 1e-6*%.3g; worst abs err %.3g" % (
        what, int(bad.sum()), bad.size, scale, float(err.max()))


def assert_close_bf16(got, want, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    scale = float(np.max(np.abs(want)))
    tol = RTOL_BF16 * np.abs(want) + ATOL_BF16_FRAC * max(scale, 1e-30)
    err = np.abs(got - want)
    bad = err > tol
    assert not bad.any(), "%s: %d/%d outside bf16 gate; worst abs err %.3g (scale %.3g)" % (
        what, int(bad.sum()), bad.size, float(err.max()), scale)


def make_case(seed, N, C, H, W, E=0, max_px=32, raw=True, with_res=False, with_bypass=False,
              shared_key=False):
    """One seeded synthetic batch.  H,W are feature dims; the raw MV image is (16H-r, 16W-r')
    (ragged when raw='ragged') so zero padding is exercised."""
    rng = np.random.default_rng(seed)
    d = {}
    nk = 1 if shared_key else N
    d["key"] = O.synth_features(rng, (nk, C, H, W))
    d["cur"] = O.synth_features(rng, (N, C, H, W))
    d["scale_map"] = O.synth_scale_map(rng, (N, C, H, W))
    mh, mw = 16 * H, 16 * W
    if raw == "ragged":
        mh, mw = 16 * H - 8, 16 * W - 11   # (.. -8) keeps the centre row 7 and drops row 8
    d["mv"] = O.synth_raw_mv(rng, N, mh, mw, max_px)
    d["flow"] = O.mv_pool(d["mv"], 1.0)
    d["logits"] = rng.standard_normal((N, 2, H, W), dtype=np.float32)
    if E:
        d["emb_warp"] = rng.standard_normal((N, E, H, W), dtype=np.float32)
        d["emb_cur"] = rng.standard_normal((N, E, H, W), dtype=np.float32)
        d["emb_cur"][:, :, 0, 0] = 0.0        # all-zero embedding pixel: eps path
    if with_res:
        d["res_raw"] = rng.integers(-64, 65, size=(N, mh, mw, 3), dtype=np.int32)
        d["res"] = O.res_pool(d["res_raw"])
        d["rnet_w"] = (0.01 * rng.standard_normal((C, 3))).astype(np.float32)
        d["rnet_b"] = (0.01 * rng.standard_normal((C,))).astype(np.float32)
    if with_bypass:
        byp = np.zeros(N, dtype=np.uint8)
        byp[::3] = 1
        d["bypass"] = byp
    if shared_key:
        d["key_index"] = np.zeros(N, dtype=np.int32)
    return d


def oracle_fused(d, mode, use_scale=True, use_res=False):
    return O.warp_scale_aggregate(
        d["key"], d["flow"], cur=d["cur"] if mode != O.W_NONE else None,
        scale_map=d["scale_map"] if use_scale else None,
        res=d.get("res") if use_res else None, rnet_w=d.get("rnet_w"), rnet_b=d.get("rnet_b"),
        weight_mode=mode, logits=d["logits"], emb_warp=d.get("emb_warp"), emb_cur=d.get("emb_cur"),
        bypass=d.get("bypass"), key_index=d.get("key_index"))
