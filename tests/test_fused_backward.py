"""Backward of the fused operator's tails (get_train_symbol SYM:306-338): x scale_map, softmax / add /
mean blend, + rnet_conv0(res), ChooseFeat bypass - on top of the a7/a8 backward.

CPU: the NumPy oracle's chain rule against torch float64 autograd (tests/golden/fused_backward_small.npz, minted by
     tools/make_golden_fused_backward.py - independent of the oracle).
GPU: lsfa_warp_scale_aggregate_backward_f32_nchw against both, the adjoint identity at full size, determinism, and
     torch autograd through the registered custom op."""
import os

import numpy as np
import pytest

from oracle import lsfa_oracle as O
from tests._util import make_case

GOLD = os.path.join(os.path.dirname(__file__), "golden", "fused_backward_small.npz")


def _close(got, want, what, rtol=1e-5, atol_frac=4e-6):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    tol = rtol * np.abs(want) + atol_frac * max(np.abs(want).max(), 1e-30)
    err = np.abs(got - want)
    assert (err <= tol).all(), "%s: worst abs err %.3g at scale %.3g (%d outside)" % (what, err.max(), np.abs(want).max(), int((err > tol).sum()))


def _golden_kwargs(g, mode):
    kw = dict(cur=g["cur"], bypass=g["bypass"])
    if mode == "logits":
        kw.update(scale_map=g["scale"], weight_mode=O.W_LOGITS, logits=g["logits"])
    else:
        kw.update(res=g["res"], rnet_w=g["rnet_w"], rnet_b=g["rnet_b"], weight_mode=O.W_ADD)
    return kw


def test_oracle_backward_equals_torch_autograd_fixture():
    g = np.load(GOLD)
    for mode, names in (("logits", ("key", "flow", "scale", "cur", "logits")), ("add", ("key", "flow", "cur", "res", "rnet_w", "rnet_b"))):
        got = O.warp_scale_aggregate_backward(g["out_grad"], g["key"], g["flow"], **_golden_kwargs(g, mode))
        fwd = O.warp_scale_aggregate(g["key"], g["flow"], **_golden_kwargs(g, mode))
        _close(fwd, g["%s_out" % mode], "forward (%s)" % mode)
        for nme in names:
            _close(got[nme], g["%s_grad_%s" % (mode, nme)], "%s grad_%s" % (mode, nme))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["logits", "add"])
def test_gpu_backward_equals_torch_autograd_fixture(cuda, mode):
    import torch
    from lsfa_b200 import ops
    g = np.load(GOLD)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)  # noqa: E731
    kw = dict(cur=t(g["cur"]), bypass=t(g["bypass"]), flow_kind="flow")
    if mode == "logits":
        kw.update(scale_map=t(g["scale"]), weight_mode="logits", logits=t(g["logits"]))
        names = ("key", "flow", "scale", "cur", "logits")
    else:
        kw.update(res=t(g["res"]), rnet_w=t(g["rnet_w"]), rnet_b=t(g["rnet_b"]), weight_mode="add")
        names = ("key", "flow", "cur", "res", "rnet_w", "rnet_b")
    got = ops.warp_scale_aggregate_backward(t(g["out_grad"]), t(g["key"]), t(g["flow"]), **kw)
    torch.cuda.synchronize()
    for nme in names:
        _close(got[nme].cpu().numpy(), g["%s_grad_%s" % (mode, nme)], "%s grad_%s" % (mode, nme), rtol=2e-5, atol_frac=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 64, 38, 63), (2, 40, 17, 23), (1, 8, 68, 120)])
@pytest.mark.parametrize("variant", ["logits_scale", "mean_scale", "add_res", "none"])
def test_gpu_backward_matches_oracle(cuda, shape, variant):
    import torch
    from lsfa_b200 import ops
    N, C, H, W = shape
    d = make_case(900 + C, N, C, H, W, with_res=True, with_bypass=True)
    rng = np.random.default_rng(C)
    og = rng.standard_normal((N, C, H, W), dtype=np.float32)
    flow = np.ascontiguousarray(d["flow"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)  # noqa: E731
    if variant == "logits_scale":
        okw = dict(cur=d["cur"], scale_map=d["scale_map"], weight_mode=O.W_LOGITS, logits=d["logits"], bypass=d["bypass"])
        gkw = dict(cur=t(d["cur"]), scale_map=t(d["scale_map"]), weight_mode="logits", logits=t(d["logits"]), bypass=t(d["bypass"]))
    elif variant == "mean_scale":
        okw = dict(cur=d["cur"], scale_map=d["scale_map"], weight_mode=O.W_MEAN)
        gkw = dict(cur=t(d["cur"]), scale_map=t(d["scale_map"]), weight_mode="mean")
    elif variant == "add_res":
        okw = dict(cur=d["cur"], res=d["res"], rnet_w=d["rnet_w"], rnet_b=d["rnet_b"], weight_mode=O.W_ADD, bypass=d["bypass"])
        gkw = dict(cur=t(d["cur"]), res=t(d["res"]), rnet_w=t(d["rnet_w"]), rnet_b=t(d["rnet_b"]), weight_mode="add", bypass=t(d["bypass"]))
    else:
        okw = dict(weight_mode=O.W_NONE)
        gkw = dict(weight_mode="none")
    want = O.warp_scale_aggregate_backward(og, d["key"], flow, **okw)
    got = ops.warp_scale_aggregate_backward(t(og), t(d["key"]), t(flow), flow_kind="flow", **gkw)
    again = ops.warp_scale_aggregate_backward(t(og), t(d["key"]), t(flow), flow_kind="flow", **gkw)
    torch.cuda.synchronize()
    assert set(got) == set(want), (sorted(got), sorted(want))
    for nme in want:
        # sums of C (logits, res) or N*H*W (rnet) products: the absolute part scales with the sum's own magnitude
        _close(got[nme].cpu().numpy(), want[nme], "%s grad_%s" % (variant, nme), rtol=2e-5, atol_frac=2e-5)
        if nme not in ("key", "flow"):   # those two come from the a7/a8 backward: d/d(flow) combines per-CTA partial sums
            # atomically, and d/d(key) is summed in a fixed order only where the gather kernel serves the shape (planes
            # whose slices are 16-byte addressable; tests/test_gpu_parity.py checks its determinism) - the scatter
            # fallback of the other shapes adds atomically
            assert torch.equal(got[nme], again[nme]), "grad_%s is not deterministic" % nme


@pytest.mark.gpu
def test_gpu_backward_raw_mv_gives_key_gradient_only(cuda):
    """SYM:319-321: on the motion-vector path the MV is data - grad_key only; the pooled flow is rebuilt for the sampler."""
    import torch
    from lsfa_b200 import ops
    N, C, H, W = 2, 16, 12, 9
    d = make_case(33, N, C, H, W)
    og = np.random.default_rng(1).standard_normal((N, C, H, W), dtype=np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)  # noqa: E731
    got = ops.warp_scale_aggregate_backward(t(og), t(d["key"]), t(d["mv"]), flow_kind="raw", cur=t(d["cur"]), weight_mode="add")
    want = O.warp_scale_aggregate_backward(og, d["key"], d["flow"], cur=d["cur"], weight_mode=O.W_ADD)
    assert set(got) == {"key", "cur"}
    _close(got["key"].cpu().numpy(), want["key"], "grad_key (raw MV)", rtol=2e-5, atol_frac=1e-5)
    assert np.array_equal(got["cur"].cpu().numpy(), og)


@pytest.mark.gpu
def test_gpu_backward_full_size_adjoint_identity(cuda):
    """<og, d out> == <grad_x, dx> for every input, at 1024x38x63 (the oracle's float64 chain would take minutes):
    the backward is the transpose of THIS library's forward."""
    import torch
    from lsfa_b200 import ops
    N, C, H, W = 2, 1024, 38, 63
    g = torch.Generator(device=cuda).manual_seed(0)
    rn = lambda *s: torch.randn(s, device=cuda, generator=g)  # noqa: E731
    key, cur, sm, flow, lg, og = rn(N, C, H, W).clamp_(min=0), rn(N, C, H, W).clamp_(min=0), 1 + 0.1 * rn(N, C, H, W), 2.0 * rn(N, 2, H, W), rn(N, 2, H, W), rn(N, C, H, W)
    kw = dict(flow_kind="flow", cur=cur, scale_map=sm, weight_mode="logits", logits=lg)
    gr = ops.warp_scale_aggregate_backward(og, key, flow, **kw)
    f0 = ops.warp_scale_aggregate(key, flow, **kw).double()
    eps = 1e-2
    for nme, x in (("key", key), ("cur", cur), ("scale", sm)):          # the op is LINEAR in each of these: exact identity
        dx = rn(*x.shape)
        kw2 = dict(kw)
        if nme == "key":
            f1 = ops.warp_scale_aggregate(key + eps * dx, flow, **kw2)
        else:
            kw2["cur" if nme == "cur" else "scale_map"] = x + eps * dx
            f1 = ops.warp_scale_aggregate(key, flow, **kw2)
        lhs = ((f1.double() - f0) * og.double()).sum().item() / eps
        rhs = (gr[nme].double() * dx.double()).sum().item()
        assert abs(lhs - rhs) <= 2e-3 * max(abs(lhs), abs(rhs), 1.0), (nme, lhs, rhs)


@pytest.mark.gpu
def test_torch_autograd_through_the_fused_custom_op(cuda):
    import torch
    import lsfa_b200.torch_ops  # noqa: F401  (registers lsfa::*)
    g = np.load(GOLD)
    t = lambda a, rg=True: torch.from_numpy(np.ascontiguousarray(a)).to(cuda).requires_grad_(rg)  # noqa: E731
    key, flow, cur, sm, lg = t(g["key"]), t(g["flow"]), t(g["cur"]), t(g["scale"]), t(g["logits"])
    byp = torch.from_numpy(g["bypass"]).to(cuda)
    out = torch.ops.lsfa.warp_scale_aggregate(key, flow, cur, sm, lg, byp, 3, 0, 1.0, 0)
    out.backward(torch.from_numpy(g["out_grad"]).to(cuda))
    for nme, x in (("key", key), ("flow", flow), ("cur", cur), ("scale", sm), ("logits", lg)):
        _close(x.grad.cpu().numpy(), g["logits_grad_%s" % nme], "autograd grad_%s" % nme, rtol=2e-5, atol_frac=1e-5)
