"""Backward oracle: the NumPy restatement, the C port of MXNet's sequential
BilinearSamplerBackward loop and an independent implementation (torch autograd of grid_sample with
align_corners=True) agree; plus known-answer cases minted here (the reference ships none)."""
import numpy as np
import torch

from oracle import c_port as P
from oracle import lsfa_oracle as O
from tests._util import assert_close_f32


def _case(seed, N=2, C=5, Hi=7, Wi=9, Ho=6, Wo=8, spread=1.3):
    rng = np.random.default_rng(seed)
    data = rng.standard_normal((N, C, Hi, Wi), dtype=np.float32)
    grid = ((rng.random((N, 2, Ho, Wo), dtype=np.float32) * 2 - 1) * np.float32(spread)).astype(np.float32)
    og = rng.standard_normal((N, C, Ho, Wo), dtype=np.float32)
    return data, grid, og


def test_backward_oracle_vs_cport_and_torch_autograd():
    P.build()
    for seed, shape in ((0, {}), (1, dict(Hi=38, Wi=63, Ho=38, Wo=63, C=4)), (2, dict(Hi=5, Wi=4, Ho=9, Wo=11, spread=2.5))):
        data, grid, og = _case(seed, **shape)
        gd, gg = O.bilinear_sampler_backward(data, grid, og)
        gd_c, gg_c = P.bilinear_sampler_backward(data, grid, og)
        assert_close_f32(gd_c, gd, what="C port grad_data")
        assert_close_f32(gg_c, gg, scale=np.abs(gg).max(), what="C port grad_grid")
        td = torch.tensor(data, requires_grad=True)
        tg = torch.tensor(grid, requires_grad=True)
        out = torch.nn.functional.grid_sample(td, tg.permute(0, 2, 3, 1), mode="bilinear", padding_mode="zeros",
                                              align_corners=True)
        out.backward(torch.tensor(og))
        assert np.abs(gd - td.grad.numpy()).max() <= 2e-6 * np.abs(gd).max()
        assert np.abs(gg - tg.grad.numpy()).max() <= 2e-6 * np.abs(gg).max()


def test_backward_is_the_adjoint_of_the_forward():
    """<og, BilinearSampler(data, grid)> == <grad_data, data> for every og: the scatter is the exact transpose."""
    data, grid, og = _case(3, N=1, C=3, Hi=38, Wi=63, Ho=38, Wo=63)
    out = O.bilinear_sampler(data, grid)
    gd, _ = O.bilinear_sampler_backward(data, grid, og)
    lhs = float(np.sum(og.astype(np.float64) * out))
    rhs = float(np.sum(gd.astype(np.float64) * data))
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs) + 1e-4


def test_backward_known_answers():
    # zero flow: the map is the identity, grad_data == out_grad; the grid gradient is the forward difference
    # of the data towards +x / +y (top-left weight 1 -> d/dx = v01 - v00), zero-padded at the far edge
    rng = np.random.default_rng(4)
    N, C, H, W = 1, 2, 6, 7
    data = rng.standard_normal((N, C, H, W), dtype=np.float32)
    og = rng.standard_normal((N, C, H, W), dtype=np.float32)
    flow = np.zeros((N, 2, H, W), np.float32)
    gk, gf = O.warp_backward(data, flow, og)
    # the fp32 grid round trip may land a hair below an integer (SURVEY 8c i): compare with a tolerance
    assert np.abs(gk - og).max() < 1e-4
    right = np.concatenate([data[..., 1:], np.zeros_like(data[..., :1])], -1)
    down = np.concatenate([data[..., 1:, :], np.zeros_like(data[..., :1, :])], -2)
    x0, y0, wx, wy = O.sampler_coords(O.grid_generator_warp(flow), H, W)
    exact = (x0[0] == np.arange(W)[None]) & (y0[0] == np.arange(H)[:, None])   # pixels whose floor did not flip
    want_x = np.sum(og * (right - data), axis=1)
    want_y = np.sum(og * (down - data), axis=1)
    assert np.abs((gf[:, 0] - want_x)[:, exact]).max() < 1e-3
    assert np.abs((gf[:, 1] - want_y)[:, exact]).max() < 1e-3
    # everything outside the plane: both gradients vanish
    far = np.full((N, 2, H, W), 1000.0, np.float32)
    gk, gf = O.warp_backward(data, far, og)
    assert not gk.any() and not gf.any()
    # GridGenerator backward is a division by the half extents
    g = rng.standard_normal((2, 2, 38, 63), dtype=np.float32)
    out = O.grid_generator_warp_backward(g)
    assert np.array_equal(out[:, 0], g[:, 0] / np.float32(31.0)) and np.array_equal(out[:, 1], g[:, 1] / np.float32(18.5))
    assert np.array_equal(P.grid_generator_warp_backward(g), out)


def test_backward_oracle_vs_committed_torch_autograd_fixture():
    """tests/golden/backward_small.npz was minted by tools/make_golden_backward.py from torch autograd of
    grid_sample(align_corners=True) - an implementation independent of this oracle."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "backward_small.npz"))
    gk, gg = O.bilinear_sampler_backward(g["key"], g["grid"], g["out_grad"])
    assert np.abs(gk - g["grad_key"]).max() <= 2e-6 * np.abs(g["grad_key"]).max()
    assert np.abs(gg - g["grad_grid"]).max() <= 2e-6 * np.abs(g["grad_grid"]).max()
    gk2, gf = O.warp_backward(g["key"], g["flow"], g["out_grad"])
    assert np.array_equal(gk2, gk)
    assert np.abs(gf - g["grad_flow"]).max() <= 2e-6 * np.abs(g["grad_flow"]).max()
    assert np.abs(O.warp(g["key"], g["flow"]) - g["out"]).max() <= 2e-6 * np.abs(g["out"]).max()
