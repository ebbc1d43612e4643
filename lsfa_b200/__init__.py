"""lsfa_b200 - B200-native (sm_100a) implementation of ONE hot path of hustvl/LSFA: the
non-key-frame feature propagation + aggregation step of dff_rfcn (motion vectors ->
stride-16 flow -> warp grid -> bilinear sample of the key-frame feature -> x scale map ->
aggregation with the current-frame feature).

Layout of the package:
  csrc/      hand-written CUDA kernels + the C ABI declared in include/lsfa_ops.h
  _build.py  nvcc recipe (in-tree shared library, no JIT)
  _cabi.py   ctypes loader (no state, no fallback)
  ops.py     the reference's operator surface (GridGenerator, BilinearSampler, ...)
  streams.py stream sharding / key-frame schedule (host logic of the multi-GPU form)
"""
from ._cabi import LsfaError, LsfaLibraryError  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):  # lazy: importing the package must not import torch
    if name in ("ops", "streams"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
