"""Pin of the MXNet rows a7/a8 (and their backward) against what the reference's GPU build actually RAN.

The reference runs on MXNet built with cuDNN (README.md:24-50), where `mx.sym.BilinearSampler`
dispatches to cuDNN's spatial-transformer sampler (`cudnnSpatialTfSamplerForward` /
`...Backward`, MXNet src/operator/cudnn_bilinear_sampler-inl.h; call sites SYM:307,321,469,572,679).
MXNet itself is not vendored, but that cuDNN routine is on the box inside torch:
`torch.cudnn_grid_sampler(input, grid)` is a direct binding of it.  These tests compare the
drop-in operators and the fused warp with it on identical inputs, at the reference's full sizes,
inside the fp32 gate of BASELINE.md section 4 (|a-b| <= 1e-5|b| + 1e-6 max|data|).

What cuDNN cannot pin: the grid itself (a7 is three mshadow expressions, no cuDNN call) - it is
fed to both sides here - and floor-flip ties: when a sampling coordinate is within one ulp of an
integer the two implementations may pick neighbouring cells; bilinear interpolation is continuous
there, so the VALUES still agree inside the gate (integer and half-cell flows are in the test set
for exactly this reason).
"""
import numpy as np
import pytest
import torch

from oracle import lsfa_oracle as O
from tests._util import ATOL_F32_FRAC, RTOL_F32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda):
    from lsfa_b200 import ops as _ops
    if not torch.backends.cudnn.is_available():
        pytest.skip("cuDNN not available in this torch build")
    return _ops


def cudnn_sampler(data, grid):
    """cudnnSpatialTfSamplerForward: data (N,C,H,W), grid (N,2,Ho,Wo) in MXNet's layout."""
    return torch.cudnn_grid_sampler(data, grid.permute(0, 2, 3, 1).contiguous())


def gate(got, want, scale, what):
    got = got.double()
    want = want.double()
    err = (got - want).abs()
    tol = RTOL_F32 * want.abs() + ATOL_F32_FRAC * scale
    bad = err > tol
    assert not bool(bad.any()), "%s: %d/%d outside |a-b| <= 1e-5|b| + 1e-6*%.3g; worst abs err %.3g" % (
        what, int(bad.sum()), bad.numel(), scale, float(err.max()))
    return float(err.max())


_SEED = {"blocks": 1, "integer": 2, "half": 3, "outside": 4, "subpixel": 5}


def flows(rng, N, H, W, kind):
    if kind == "blocks":         # SURVEY 8d synthetic macroblock motion, accumulated-GOP magnitude
        return np.ascontiguousarray(O.mv_pool(O.synth_raw_mv(rng, N, 16 * H, 16 * W, 96)))
    if kind == "integer":
        return rng.integers(-3, 4, size=(N, 2, H, W)).astype(np.float32)
    if kind == "half":
        return (rng.integers(-6, 7, size=(N, 2, H, W)) * 0.5).astype(np.float32)
    if kind == "outside":        # most taps leave the plane on at least one side
        return (rng.standard_normal((N, 2, H, W)) * np.array([W, H], np.float32).reshape(1, 2, 1, 1) * 0.6).astype(np.float32)
    if kind == "subpixel":
        return (rng.standard_normal((N, 2, H, W)) * 2.5).astype(np.float32)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["blocks", "integer", "half", "outside", "subpixel"])
@pytest.mark.parametrize("N,C,H,W", [(2, 1024, 38, 63), (1, 1024, 68, 120), (3, 40, 17, 23)])
def test_bilinear_sampler_equals_cudnn_spatial_tf_sampler(ops, cuda, kind, N, C, H, W):
    rng = np.random.default_rng(_SEED[kind] * 1000 + H)
    data = torch.from_numpy(O.synth_features(rng, (N, C, H, W))).to(cuda)
    flow = torch.from_numpy(flows(rng, N, H, W, kind)).to(cuda)
    grid = ops.GridGenerator(flow)
    want = cudnn_sampler(data, grid)
    scale = float(data.abs().max())
    gate(ops.BilinearSampler(data, grid), want, scale, "BilinearSampler vs cuDNN (%s)" % kind)
    # the fused warp (flow in, grid generated in-kernel), every kernel of the NCHW path, and channels-last
    for generic in (0, 1, 2):
        got = ops.warp_scale_aggregate(data, flow, flow_kind="flow", weight_mode="none", force_generic=generic)
        gate(got, want, scale, "fused warp NCHW kernel %d vs cuDNN (%s)" % (generic, kind))
    got = ops.to_nchw(ops.warp_scale_aggregate(ops.to_nhwc(data), flow, flow_kind="flow", weight_mode="none", layout="nhwc_f32"))
    gate(got, want, scale, "fused warp NHWC vs cuDNN (%s)" % kind)


def test_key_plane_of_another_size_equals_cudnn(ops, cuda):
    """Hi x Wi != Ho x Wo (the sampler's general form)."""
    rng = np.random.default_rng(3)
    data = torch.from_numpy(O.synth_features(rng, (2, 16, 24, 31))).to(cuda)
    grid = torch.from_numpy((rng.random((2, 2, 38, 63)) * 2.3 - 1.15).astype(np.float32)).to(cuda)
    gate(ops.BilinearSampler(data, grid), cudnn_sampler(data, grid), float(data.abs().max()), "Hi,Wi != Ho,Wo")


@pytest.mark.parametrize("kind", ["blocks", "half", "subpixel"])
@pytest.mark.parametrize("N,C,H,W", [(2, 256, 38, 63), (1, 64, 68, 120)])
def test_backward_equals_cudnn_spatial_tf_sampler_backward(ops, cuda, kind, N, C, H, W):
    """cudnnSpatialTfSamplerBackward (what MXNet's GPU BilinearSampler backward calls) through torch autograd.
    cuDNN accumulates with float atomics in an unspecified order, so the gate on the gradients is the forward's
    relative part plus an absolute part scaled by the number of summed terms (<= 4 taps x a few output pixels per
    input pixel for grad_data; C terms per pixel for grad_grid)."""
    rng = np.random.default_rng(_SEED[kind] * 1000 + W)
    data = torch.from_numpy(O.synth_features(rng, (N, C, H, W))).to(cuda).requires_grad_(True)
    flow = torch.from_numpy(flows(rng, N, H, W, kind)).to(cuda)
    grid = ops.GridGenerator(flow).requires_grad_(True)
    og = torch.from_numpy(rng.standard_normal((N, C, H, W), dtype=np.float32)).to(cuda)
    out = cudnn_sampler(data, grid)
    out.backward(og)
    for kernel in (("gather", "scatter") if H * W <= 4320 else ("auto", "scatter")):   # gather: planes of <= 4,320 pixels
        gd, gg = ops.BilinearSampler_backward(data.detach(), grid.detach(), og, kernel=kernel)
        s_d = float(data.grad.abs().max())
        s_g = float(grid.grad.abs().max())
        err_d = (gd.double() - data.grad.double()).abs()
        err_g = (gg.double() - grid.grad.double()).abs()
        assert bool((err_d <= 1e-5 * data.grad.abs().double() + 4e-6 * s_d).all()), \
            "grad_data (%s) vs cuDNN: worst %.3g of scale %.3g" % (kernel, float(err_d.max()), s_d)
        # grad_grid sums C products of O(1) terms: relative to the sum's own scale
        assert bool((err_g <= 1e-4 * grid.grad.abs().double() + 2e-5 * s_g).all()), \
            "grad_grid (%s) vs cuDNN: worst %.3g of scale %.3g" % (kernel, float(err_g.max()), s_g)


def test_oracle_restatement_equals_cudnn_on_the_golden_fixture(ops, cuda):
    """Closes the loop for the CPU side: the NumPy oracle's a7+a8 (the thing every other parity test uses as
    truth) against cuDNN on the same inputs."""
    rng = np.random.default_rng(11)
    data = O.synth_features(rng, (2, 32, 38, 63))
    flow = flows(rng, 2, 38, 63, "blocks")
    grid = O.grid_generator_warp(flow)
    want = cudnn_sampler(torch.from_numpy(data).to(cuda), torch.from_numpy(grid).to(cuda)).cpu()
    gate(torch.from_numpy(O.bilinear_sampler(data, grid)), want, float(np.abs(data).max()), "oracle vs cuDNN")
