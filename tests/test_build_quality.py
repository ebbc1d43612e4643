"""Build hygiene that has already cost performance once: the register-critical streaming kernels must not spill
(a 108-byte spill in agg_nchw_tma_kernel<2,5,ResCur> took the shared-key stream sweep from 218k to 167k frames/s), and
the tensor-core translation unit must really contain tcgen05 / TMA-tensor instructions."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG = os.path.join(ROOT, "lsfa_b200", "lib", "build.log")


def _kernel_info():
    text = open(LOG).read()
    out = {}
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\s*\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
                         r"(\d+) bytes spill loads\s*\n.*?Used (\d+) registers", text):
        out[m.group(1)] = dict(stack=int(m.group(2)), spill_st=int(m.group(3)), spill_ld=int(m.group(4)), regs=int(m.group(5)))
    return out


def test_headline_streaming_kernels_do_not_spill(lib):
    from lsfa_b200 import _build
    if not os.path.exists(LOG):
        _build.build(force=True)
    info = _kernel_info()
    assert info, "no ptxas -v records in build.log"
    # agg_nchw_tma_kernel<K=2, PPT=5, VAR> for the four compile-time variants (38x63 planes) and the tensor-core kernels
    wanted = [n for n in info if re.search(r"agg_nchw_tma_kernelILi2ELi5ELi[1-4]E", n) or "conv_gemm_tc_kernel" in n
              or "conv_gemm_tc2_kernel" in n
              or "agg_nhwc_tma_kernel" in n or "cosine_partials_tma_kernel" in n]
    assert len(wanted) >= 11, wanted
    bad = {n: info[n] for n in wanted if info[n]["spill_st"] or info[n]["spill_ld"]}
    assert not bad, "register spills in headline kernels: %s" % bad


def test_tensor_core_unit_contains_tcgen05_and_tma_tensor_sass(lib):
    obj = os.path.join(ROOT, "lsfa_b200", "lib", "obj", "conv_gemm_tc.o")
    if not os.path.exists(obj):
        pytest.skip("object file not kept")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "SYNCS"):
        assert mnemonic in sass, "%s missing from conv_gemm_tc.o" % mnemonic
    # the CTA-pair form: tcgen05.mma.cta_group::2, multicast tcgen05.commit, cta_group::2 tensor copies
    for mnemonic in ("UTCHMMA.2CTA", "UTCBAR.2CTA.MULTICAST", "UTMALDG.4D.2CTA"):
        assert mnemonic in sass, "%s missing from conv_gemm_tc.o" % mnemonic


def test_window_kernel_unit_contains_tensor_map_copies(lib):
    """The window-resident channels-last kernel moves its key windows with tensor-map TMA (UTMALDG), not bulk copies."""
    obj = os.path.join(ROOT, "lsfa_b200", "lib", "obj", "aggregate_nhwc_win.o")
    if not os.path.exists(obj):
        pytest.skip("object file not kept")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass and "SYNCS" in sass
    assert sass.count("Function :") >= 8          # 2 storage types x 4 compile-time variants
