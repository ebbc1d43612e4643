"""In-tree build of the C-ABI shared library (nvcc, sm_100a only).

``python -m lsfa_b200._build`` or ``lsfa_b200._build.build()``.  The resulting
``lsfa_b200/lib/liblsfa_b200.so`` is git-ignored but travels to the GPU box with the
snapshot; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "liblsfa_b200.so")
SOURCES = ["cabi.cu", "aggregate_nchw.cu", "aggregate_nhwc.cu", "aggregate_nhwc_win.cu", "prep_ops.cu", "coviar_accumulate.cu", "sampler_backward.cu", "cosine_nchw.cu", "conv_gemm_tc.cu", "host_pipeline.cu", "aggregate_backward.cu",
           "plane_var0.cu", "plane_var3.cu",
           "tma_var1.cu", "tma_var2.cu", "tma_var3.cu", "tma_var4.cu",
           "tma2_var1.cu", "tma2_var2.cu", "tma2_var3.cu", "tma2_var4.cu"]
HEADERS = [os.path.join(CSRC, "lsfa_device.cuh"), os.path.join(CSRC, "aggregate_nchw_plane.cuh"),
           os.path.join(CSRC, "plane_variant_impl.inc"), os.path.join(CSRC, "aggregate_nchw_tma.cuh"),
           os.path.join(CSRC, "tma_variant_impl.inc"), os.path.join(CSRC, "aggregate_nchw_tma2.cuh"),
           os.path.join(CSRC, "aggregate_nhwc_tma.cuh"), os.path.join(CSRC, "aggregate_nhwc_win.cuh"),
           os.path.join(CSRC, "tma2_variant_impl.inc"), os.path.join(CSRC, "conv_gemm_tc.h"), os.path.join(CSRC, "aggregate_backward.h"),
           os.path.join(os.path.dirname(PKG_DIR), "include", "lsfa_ops.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true",            # value path may contract; the index path uses *_rn intrinsics
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas", "-v",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; this package only builds for sm_100a with CUDA >= 12.8")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(nvcc: str, src: str, obj: str):
    extra = os.environ.get("LSFA_NVCC_EXTRA", "").split()     # experiment knobs, e.g. -DLSFA_LD_POLICY=1
    cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", "-o", obj, src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    return proc.returncode, " ".join(cmd) + "\n" + proc.stdout + proc.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu to an object (in parallel: one nvcc per translation unit) and link."""
    if not force and not is_stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = find_nvcc()
    jobs = [(os.path.join(CSRC, s), os.path.join(obj_dir, s[:-3] + ".o")) for s in SOURCES]
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda j: _compile_one(nvcc, *j), jobs))
    log = "".join(r[1] for r in results)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + [j[1] for j in jobs]
    rc = max(r[0] for r in results)
    if rc == 0:
        proc = subprocess.run(link, capture_output=True, text=True)
        log += " ".join(link) + "\n" + proc.stdout + proc.stderr
        rc = proc.returncode
    with open(os.path.join(LIB_DIR, "build.log"), "w") as f:
        f.write(log)
    if rc != 0:
        raise RuntimeError("nvcc failed:\n" + log[-8000:])
    if verbose:
        print(log)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
