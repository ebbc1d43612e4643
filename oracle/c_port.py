"""ctypes loader for oracle/lsfa_oracle.c (TEST INFRASTRUCTURE - see the header of that file).
Used by tests, by bench.py's cpu_baseline / --impl reference legs and by smoke()."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liblsfa_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "lsfa_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B" if force else "-s", "all"], check=True,
                       capture_output=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.lsfa_ref_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads():
    return load().lsfa_ref_num_threads()


def use_all_cores():
    """Use every core this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    load().lsfa_ref_set_num_threads(int(n))
    return num_threads()


def mv_pool(mv, im_scale=1.0, mode=0):
    mv = np.ascontiguousarray(mv, dtype=np.int32)
    N, h, w, _ = mv.shape
    flow = np.empty((N, 2, (h + 15) // 16, (w + 15) // 16), np.float32)
    load().lsfa_ref_mv_pool_i32(_p(mv), _p(flow), N, h, w, C.c_double(im_scale), mode)
    return flow


def grid_generator_warp(flow):
    flow = np.ascontiguousarray(flow, dtype=np.float32)
    N, _, H, W = flow.shape
    grid = np.empty_like(flow)
    load().lsfa_ref_grid_generator_warp(_p(flow), _p(grid), N, H, W)
    return grid


def bilinear_sampler(data, grid):
    data = np.ascontiguousarray(data, dtype=np.float32)
    grid = np.ascontiguousarray(grid, dtype=np.float32)
    N, Cc, Hi, Wi = data.shape
    Ho, Wo = grid.shape[2:]
    out = np.empty((N, Cc, Ho, Wo), np.float32)
    load().lsfa_ref_bilinear_sampler(_p(data), _p(grid), _p(out), N, Cc, Hi, Wi, Ho, Wo)
    return out


def chain_nq(mv, key, scale_map, cur, logits, im_scale=1.0, tmp=None, out=None):
    """The key-frame Nq tail op by op (what MXNet CPU executes): returns (out, tmp)."""
    mv = np.ascontiguousarray(mv, dtype=np.int32)
    N, Cc, H, W = key.shape
    if tmp is None:
        tmp = np.empty(5 * key.size, np.float32)
    if out is None:
        out = np.empty_like(key)
    load().lsfa_ref_chain_nq(_p(mv), mv.shape[1], mv.shape[2], C.c_double(im_scale), _p(key), _p(scale_map),
                             _p(cur), _p(logits), _p(out), _p(tmp), N, Cc, H, W)
    return out, tmp


def mv_accumulate(mvs, counts, height, width):
    """coviar_data_loader.c:71-139 (accumulate=1) for one GOP: mvs (T,M,6) int32, counts (T,) -> (height,width,2)."""
    mvs = np.ascontiguousarray(mvs, dtype=np.int32)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    T, M, _ = mvs.shape
    a = np.empty(height * width * 2, np.int32)
    b = np.empty_like(a)
    out = np.empty((height, width, 2), np.int32)
    load().lsfa_ref_mv_accumulate(_p(mvs), _p(counts), T, M, height, width, _p(a), _p(b), _p(out))
    return out


def coviar_residual(iframe, cur, mv):
    iframe = np.ascontiguousarray(iframe, dtype=np.uint8)
    cur = np.ascontiguousarray(cur, dtype=np.uint8)
    mv = np.ascontiguousarray(mv, dtype=np.int32)
    h, w, _ = cur.shape
    res = np.empty((h, w, 3), np.int32)
    load().lsfa_ref_coviar_residual(_p(iframe), _p(cur), _p(mv), _p(res), h, w)
    return res


def bilinear_sampler_backward(data, grid, out_grad, grad_data=None, grad_grid=None):
    """MXNet's CPU BilinearSamplerBackward (sequential float32 accumulation).  grad_data / grad_grid
    given = kAddTo (accumulated into, in place); None = kWriteTo (zeroed first)."""
    data = np.ascontiguousarray(data, dtype=np.float32)
    grid = np.ascontiguousarray(grid, dtype=np.float32)
    og = np.ascontiguousarray(out_grad, dtype=np.float32)
    N, Cc, Hi, Wi = data.shape
    Ho, Wo = grid.shape[2:]
    gd = np.zeros_like(data) if grad_data is None else grad_data
    gg = np.zeros_like(grid) if grad_grid is None else grad_grid
    load().lsfa_ref_bilinear_sampler_backward(_p(data), _p(grid), _p(og), _p(gd), _p(gg), N, Cc, Hi, Wi, Ho, Wo)
    return gd, gg


def grid_generator_warp_backward(grad_grid):
    g = np.ascontiguousarray(grad_grid, dtype=np.float32)
    N, _, H, W = g.shape
    out = np.empty_like(g)
    load().lsfa_ref_grid_generator_warp_backward(_p(g), _p(out), N, H, W)
    return out
