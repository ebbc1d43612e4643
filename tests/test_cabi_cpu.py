"""CPU suite part 2: the C-ABI library builds, loads without a GPU, exports every symbol
include/lsfa_ops.h declares, validates arguments before touching CUDA, and the Python struct
mirrors the C one.  No compute call is made here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lsfa_ops.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"LSFA_API\s+[\w\s\*]+?\b(lsfa_\w+)\s*\(", src)))


def test_header_declares_the_path():
    names = declared_symbols()
    for must in ("lsfa_grid_generator_warp_f32", "lsfa_bilinear_sampler_f32", "lsfa_mv_pool_i32",
                 "lsfa_warp_scale_aggregate_f32_nchw", "lsfa_warp_scale_aggregate_bf16_nhwc",
                 "lsfa_last_error", "lsfa_version"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    from lsfa_b200 import _build, _cabi
    out = subprocess.run(["nm", "-D", "--defined-only", _build.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (lsfa_\w+)", out))
    declared = declared_symbols()
    assert declared, "header parse failed"
    missing = [s for s in declared if s not in exported]
    assert not missing, "declared but not exported: %s" % missing
    assert sorted(_cabi.PROTOTYPES) == declared, "python prototypes and header disagree"
    assert lib.lsfa_version() == _cabi.ABI_VERSION


def test_python_struct_matches_c_layout(lib, tmp_path):
    from lsfa_b200 import _cabi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lsfa_ops.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(LsfaAggArgs), offsetof(LsfaAggArgs,key),'
                   'offsetof(LsfaAggArgs,im_scale), offsetof(LsfaAggArgs,out), offsetof(LsfaAggArgs,force_generic));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    S = _cabi.LsfaAggArgs
    assert got == [ctypes.sizeof(S), S.key.offset, S.im_scale.offset, S.out.offset, S.force_generic.offset]


def test_argument_validation_needs_no_gpu(lib):
    from lsfa_b200 import _cabi as A
    a = A.new_args()
    assert lib.lsfa_warp_scale_aggregate(a, None) == A.E_SHAPE
    assert b"non-positive" in lib.lsfa_last_error()
    a.struct_bytes = 8
    assert lib.lsfa_warp_scale_aggregate(a, None) == A.E_BADARG
    assert b"struct_bytes" in lib.lsfa_last_error()
    assert lib.lsfa_warp_scale_aggregate(None, None) == A.E_BADARG
    a = A.new_args(N=1, C=8, H=4, W=4, layout=A.LAYOUT_NCHW_F32, req=A.REQ_WRITE)
    assert lib.lsfa_warp_scale_aggregate(a, None) == A.E_BADARG            # key/flow/out NULL
    a = A.new_args(N=1, C=6, H=4, W=4, layout=A.LAYOUT_NHWC_BF16, req=A.REQ_WRITE, key=16, flow=16, out=16)
    assert lib.lsfa_warp_scale_aggregate(a, None) == A.E_ALIGN             # C % 8
    a = A.new_args(N=1, C=8, H=4, W=4, req=A.REQ_WRITE, key=16, flow=16, out=16, flow_kind=A.FLOW_RAW_I32,
                   mv_h=100, mv_w=64, im_scale=1.0)
    assert lib.lsfa_warp_scale_aggregate(a, None) == A.E_SHAPE             # 100x64 pools to 7x4
    a = A.new_args(N=1, C=8, H=4, W=4, req=A.REQ_WRITE, key=16, flow=16, out=16, weight_mode=A.W_COSINE,
                   cur=16, emb_warp=16, emb_cur=16, E=32)
    # logits + partial sums of the all-TMA cosine pre-pass (16 channel pairs -> 16 CTAs -> 16 slots x 3 sums x 16 px)
    # + claim counters / row ranges + records
    assert lib.lsfa_warp_scale_aggregate_workspace_bytes(a) == 1 * 2 * 4 * 4 * 4 + 16 * 3 * 16 * 4 + 16 + 16 * 32
    assert lib.lsfa_warp_scale_aggregate_num_launches(a) == 2
    assert lib.lsfa_warp_scale_aggregate(a, None) == A.E_BADARG            # workspace missing
    a.layout = A.LAYOUT_NHWC_F32
    assert lib.lsfa_warp_scale_aggregate_workspace_bytes(a) == 64          # channels-last: one claim counter
    assert lib.lsfa_warp_scale_aggregate_f32_nchw(a, None) == A.E_BADARG   # wrong layout for the suffixed entry
    a = A.new_args(N=1, C=8, H=4, W=4, layout=A.LAYOUT_NHWC_F32, req=A.REQ_WRITE, key=16, flow=16, out=16, force_generic=2)
    assert lib.lsfa_warp_scale_aggregate(a, None) == A.E_BADARG            # kernel 2 does not exist for channels-last
    assert lib.lsfa_res_coviar_pool_i32(None, 16, 1, 32, 32, 32, 32, 1.0, 0, None, 1.0, 0, None) == A.E_BADARG
    assert lib.lsfa_res_coviar_pool_i32(16, 16, 1, 720, 1280, 562, 999, 0.78125, 0, None, 1.0, 0, None) == A.E_SHAPE
    assert b"resizes to 562x1000" in lib.lsfa_last_error()
    assert lib.lsfa_mv_pool_i32(None, None, 1, 16, 16, 1.0, 0, None) == A.E_BADARG
    assert lib.lsfa_mv_pool_i32(16, 16, 1, 16, 16, 1.0, 7, None) == A.E_BADARG
    assert lib.lsfa_grid_generator_warp_f32(16, 16, 0, 4, 4, None) == A.E_SHAPE
    assert lib.lsfa_bilinear_sampler_f32(16, 16, 16, 1, 1, 0, 4, 4, 4, 1, None) == A.E_SHAPE
    assert lib.lsfa_bilinear_sampler_f32(16, 16, 16, 1, 1, 4, 4, 4, 4, 9, None) == A.E_BADARG   # bad req
    a = A.new_args(N=1, C=8, H=38, W=63, req=A.REQ_WRITE, key=16, flow=16, out=16, flow_kind=A.FLOW_COVIAR_I32,
                   mv_src_h=720, mv_src_w=1280, mv_h=562, mv_w=999, im_scale=0.78125)
    assert lib.lsfa_warp_scale_aggregate(a, None) == A.E_SHAPE             # 1280*0.78125 rounds to 1000, not 999
    a = A.new_args(N=1, C=4, H=4096, W=4096, req=A.REQ_WRITE, key=16, flow=16, out=16)
    assert lib.lsfa_warp_scale_aggregate(a, None) == A.E_SHAPE             # planes of 2^24 pixels are refused, not truncated
    a = A.new_args(N=1, C=4, H=4, W=4, req=A.REQ_NULL, key=16, flow=16, out=16)
    assert lib.lsfa_warp_scale_aggregate(a, None) == A.OK                  # kNullOp: validated, nothing launched
    assert lib.lsfa_warp_scale_aggregate_num_launches(a) == 0
    assert lib.lsfa_unfused_chain_num_launches() == 9


def test_no_cpu_fallback_in_product():
    """The product path must fail loudly without its CUDA library and must not import the oracle."""
    import importlib
    from lsfa_b200 import _cabi
    for mod in ("ops.py", "_cabi.py", "streams.py", "__init__.py", "_build.py"):
        text = open(os.path.join(ROOT, "lsfa_b200", mod)).read()
        assert "oracle" not in text.replace("no CPU", ""), "%s mentions the oracle" % mod
    saved, saved_env = _cabi._lib, os.environ.get("LSFA_B200_LIB")
    try:
        _cabi._lib = None
        os.environ["LSFA_B200_LIB"] = "/nonexistent/liblsfa_b200.so"
        with pytest.raises(_cabi.LsfaLibraryError):
            _cabi.load()
    finally:
        _cabi._lib = saved
        if saved_env is None:
            os.environ.pop("LSFA_B200_LIB", None)
        else:
            os.environ["LSFA_B200_LIB"] = saved_env


def test_ops_reject_cpu_tensors(lib):
    torch = pytest.importorskip("torch")
    from lsfa_b200 import ops
    with pytest.raises(ValueError, match="CUDA"):
        ops.GridGenerator(torch.zeros(1, 2, 4, 4))
    with pytest.raises(ValueError, match="CUDA"):
        ops.mv_pool(torch.zeros(1, 16, 16, 2, dtype=torch.int32))
