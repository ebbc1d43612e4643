"""ctypes driver for oracle/_ref/libcoviar_ref.so = the reference's own
external/data_loader_py2/coviar_data_loader.c compiled where it lies (TEST INFRASTRUCTURE; recipe:
`make -C oracle ref`, shims: oracle/ref_shims/README.md).

Only tests/ (and tools/make_golden_coviar_ref.py, which mints tests/golden/coviar_ref.npz for the
GPU box, where /root/reference does not exist) import this.  The function called is the
reference's `create_and_load_mv_residual` (coviar_data_loader.c:71-177), unmodified; this module
restates only the few lines of `decode_video` that drive it per decoded frame
(coviar_data_loader.c:283-352: array creation, the identity initialisation of accu_src /
accu_src_old, one call per frame that carries side data, cur_pos counting from the I-frame).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libcoviar_ref.so")
REF_SRC = "/root/reference/external/data_loader_py2/coviar_data_loader.c"
MV, RESIDUAL = 1, 2          # coviar_data_loader.c:21-22
_lib = None


class AVMotionVector(C.Structure):      # oracle/ref_shims/libavutil/motion_vector.h (FFmpeg public ABI)
    _fields_ = [("source", C.c_int32), ("w", C.c_uint8), ("h", C.c_uint8), ("src_x", C.c_int16), ("src_y", C.c_int16),
                ("dst_x", C.c_int16), ("dst_y", C.c_int16), ("flags", C.c_uint64)]


class AVFrameSideData(C.Structure):     # oracle/ref_shims/libavcodec/avcodec.h
    _fields_ = [("type", C.c_int), ("data", C.c_void_p), ("size", C.c_int), ("metadata", C.c_void_p), ("buf", C.c_void_p)]


def available() -> bool:
    """True when the reference library is built or can be built (the GPU box has neither)."""
    return os.path.exists(LIB) or os.path.exists(REF_SRC)


def build(force=False):
    if not os.path.exists(REF_SRC):
        if os.path.exists(LIB):
            return LIB
        raise FileNotFoundError("reference source absent and no prebuilt oracle/_ref/libcoviar_ref.so")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(REF_SRC):
        subprocess.run(["make", "-C", HERE, "-s"] + (["-B"] if force else []) + ["ref"], check=True, capture_output=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.PyDLL(LIB)             # arguments are PyArrayObject*: keep the GIL
        f = _lib.create_and_load_mv_residual
        f.restype = None
        f.argtypes = [C.POINTER(AVFrameSideData), C.py_object, C.py_object, C.py_object, C.c_int, C.c_int, C.c_int,
                      C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    return _lib


def _side_data(vecs):
    """vecs (n,6) int32 {w,h,src_x,src_y,dst_x,dst_y} -> (AVFrameSideData, keep-alive array)."""
    n = len(vecs)
    arr = (AVMotionVector * max(n, 1))()
    for i in range(n):
        w, h, sx, sy, dx, dy = (int(v) for v in vecs[i])
        arr[i] = AVMotionVector(-1, w, h, sx, sy, dx, dy, 0)
    sd = AVFrameSideData(8, C.cast(arr, C.c_void_p), n * C.sizeof(AVMotionVector), None, None)
    return sd, arr


def _identity(height, width):
    """coviar_data_loader.c:318-328: accu[x][y] = (x, y), in the reference's [x][y][2] int layout."""
    a = np.empty((width, height, 2), np.intc)
    a[..., 0] = np.arange(width, dtype=np.intc)[:, None]
    a[..., 1] = np.arange(height, dtype=np.intc)[None, :]
    return a


def mv_accumulate(mvs, counts, height, width):
    """coviar.load(..., representation=MV, accumulate=True) at GOP position T (= len(counts)): one call of the
    reference function per P-frame; returns the (height,width,2) int32 array the reference returns."""
    lib = load()
    mvs = np.asarray(mvs, np.int32)
    T = len(counts)
    mv_arr = np.zeros((height, width, 2), np.int32)            # :297-303 PyArray_ZEROS
    bgr_arr = np.zeros((2, height, width, 3), np.uint8)        # :289-295
    res_arr = np.zeros((1,), np.int32)                         # not touched for representation == MV
    accu_old = _identity(height, width)
    accu = accu_old.copy()                                     # :329 memcpy
    for t in range(T):
        sd, keep = _side_data(mvs[t][:int(counts[t])])
        lib.create_and_load_mv_residual(C.byref(sd), bgr_arr, mv_arr, res_arr, t + 1, 1, MV,
                                        accu.ctypes.data, accu_old.ctypes.data, width, height, T)
        del keep
    return mv_arr


def residual(iframe, cur, mvs, counts):
    """coviar.load(..., representation=RESIDUAL, accumulate=True): res = cur - iframe[accumulated source]."""
    lib = load()
    mvs = np.asarray(mvs, np.int32)
    height, width, _ = cur.shape
    T = len(counts)
    mv_arr = np.zeros((height, width, 3), np.int32)            # :305-312 (the reference allocates both with 3 channels)
    res_arr = np.zeros((height, width, 3), np.int32)
    bgr_arr = np.ascontiguousarray(np.stack([iframe, cur]).astype(np.uint8))    # [0] = I-frame, [1] = target (:58-64)
    accu_old = _identity(height, width)
    accu = accu_old.copy()
    for t in range(T):
        sd, keep = _side_data(mvs[t][:int(counts[t])])
        lib.create_and_load_mv_residual(C.byref(sd), bgr_arr, mv_arr, res_arr, t + 1, 1, RESIDUAL,
                                        accu.ctypes.data, accu_old.ctypes.data, width, height, T)
        del keep
    return res_arr


def mv_single(vecs, height, width):
    """accumulate=False, representation=MV: the per-frame block splat (coviar_data_loader.c:116-119)."""
    lib = load()
    mv_arr = np.zeros((height, width, 2), np.int32)
    dummy = np.zeros((1,), np.int32)
    bgr_arr = np.zeros((2, height, width, 3), np.uint8)
    sd, keep = _side_data(np.asarray(vecs, np.int32))
    lib.create_and_load_mv_residual(C.byref(sd), bgr_arr, mv_arr, dummy, 1, 0, MV, None, None, width, height, 1)
    del keep
    return mv_arr
