"""The reference-facing HOST path behind the C ABI (lsfa_host_aggregate_f32_nchw, csrc/host_pipeline.cu):
`_load_data` + forward + `asnumpy()` of the reference (core/DataParallelExecutorGroup.py:24-39, core/tester.py:138-145)
with the key feature kept on the device across a GOP (core/tester.py:246-252)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import lsfa_oracle as O
from tests._util import assert_close_f32, make_case, oracle_fused


def _pin(a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()


def test_host_args_struct_and_validation_need_no_gpu(lib, tmp_path):
    import os
    import subprocess
    from lsfa_b200 import _cabi as A
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lsfa_ops.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(LsfaHostAggArgs), offsetof(LsfaHostAggArgs,im_scale),'
                   'offsetof(LsfaHostAggArgs,key), offsetof(LsfaHostAggArgs,staging), offsetof(LsfaHostAggArgs,stream_out));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    S = A.LsfaHostAggArgs
    assert got == [ctypes.sizeof(S), S.im_scale.offset, S.key.offset, S.staging.offset, S.stream_out.offset]

    def args(**kw):
        a = S()
        a.struct_bytes = ctypes.sizeof(S)
        a.N, a.C, a.H, a.W, a.mv_h, a.mv_w, a.im_scale = 4, 8, 2, 3, 32, 48, 1.0
        a.weight_mode, a.chunk, a.depth = A.W_LOGITS, 2, 2
        a.key = a.cur = a.mv = a.logits = a.out = 256
        for k, v in kw.items():
            setattr(a, k, v)
        return a
    assert lib.lsfa_host_aggregate_f32_nchw(None) == A.E_BADARG
    assert lib.lsfa_host_aggregate_f32_nchw(args(struct_bytes=4)) == A.E_BADARG
    assert lib.lsfa_host_aggregate_f32_nchw(args(H=5)) == A.E_SHAPE                  # H != ceil(mv_h/16)
    assert lib.lsfa_host_aggregate_f32_nchw(args(weight_mode=A.W_COSINE)) == A.E_BADARG
    assert b"NONE, ADD, MEAN and LOGITS" in lib.lsfa_last_error()
    assert lib.lsfa_host_aggregate_f32_nchw(args(logits=None)) == A.E_BADARG
    assert lib.lsfa_host_aggregate_f32_nchw(args(depth=1)) == A.E_BADARG
    assert lib.lsfa_host_aggregate_f32_nchw(args(im_scale=0.0)) == A.E_BADARG
    assert lib.lsfa_host_aggregate_f32_nchw(args()) == A.E_BADARG                    # no staging
    assert b"staging too small" in lib.lsfa_last_error()
    idx = (ctypes.c_int32 * 4)(0, 1, 2, 9)
    bad = args(key_index=ctypes.addressof(idx), key_table=256, num_slots=3)
    assert lib.lsfa_host_aggregate_f32_nchw(bad) == A.E_BADARG                       # slot 9 of a 3-slot table
    assert b"key_index[3]=9" in lib.lsfa_last_error()
    # staging: depth x (key + cur + out chunks, mv image, logits, workspace), every part rounded to 256 bytes
    F = 2 * 8 * 2 * 3 * 4
    up = lambda v: (v + 255) // 256 * 256  # noqa: E731
    ws = lib.lsfa_warp_scale_aggregate_workspace_bytes(A.new_args(N=2, C=8, H=2, W=3, layout=A.LAYOUT_NCHW_F32))
    assert lib.lsfa_host_aggregate_staging_bytes(args()) == 2 * (3 * up(F) + up(2 * 32 * 48 * 8) + up(2 * 2 * 2 * 3 * 4) + up(ws))
    bi, bo = ctypes.c_size_t(0), ctypes.c_size_t(0)
    assert lib.lsfa_host_aggregate_bytes(args(), ctypes.byref(bi), ctypes.byref(bo)) == A.OK
    assert bo.value == 4 * 8 * 2 * 3 * 4 and bi.value == 4 * (2 * 8 * 2 * 3 * 4 + 4 * 48 * 8 + 2 * 2 * 3 * 4)


@pytest.mark.gpu
@pytest.mark.parametrize("im_scale", [1.0, 0.78125, 1.25])
@pytest.mark.parametrize("mode,use_scale", [("logits", True), ("add", False), ("mean", True), ("none", True)])
def test_host_path_matches_oracle_with_im_scale(cuda, im_scale, mode, use_scale):
    """im_scale reaches the kernel (image.py:224: flow = pooled * im_scale / 16): the MV image is at network scale,
    as transform_mv_res leaves it after its stage-1 resize."""
    from lsfa_b200.host import HostAggregator
    N, C, H, W = 5, 16, 12, 9
    d = make_case(70, N, C, H, W)
    flow = O.mv_pool(d["mv"], im_scale)
    wm = {"logits": O.W_LOGITS, "add": O.W_ADD, "mean": O.W_MEAN, "none": O.W_NONE}[mode]
    want = O.warp_scale_aggregate(d["key"], flow, cur=d["cur"] if mode != "none" else None,
                                  scale_map=d["scale_map"] if use_scale else None, weight_mode=wm, logits=d["logits"])
    host_in = {k: _pin(d[k]) for k in ("key", "cur", "scale_map", "mv", "logits")}
    out_host = torch.empty((N, C, H, W), dtype=torch.float32).pin_memory()
    agg = HostAggregator(N, C, H, W, d["mv"].shape[1:3], cuda, chunk=2, depth=2, weight_mode=mode, use_scale=use_scale,
                         im_scale=im_scale)
    for _ in range(2):
        agg(host_in, out_host)
    agg.synchronize()
    assert_close_f32(out_host.numpy(), want, scale=max(np.abs(d["key"]).max(), np.abs(d["cur"]).max()), what="host path")
    assert agg.h2d_bytes == 2 * agg.bytes_per_call()[0]


@pytest.mark.gpu
def test_host_path_rejects_what_it_cannot_serve(cuda):
    from lsfa_b200.host import HostAggregator
    with pytest.raises(ValueError, match="weight_mode"):
        HostAggregator(2, 8, 2, 3, (32, 48), cuda, weight_mode="cosine")
    agg = HostAggregator(2, 8, 2, 3, (32, 48), cuda, weight_mode="logits")
    d = make_case(1, 2, 8, 2, 3)
    out_host = torch.empty((2, 8, 2, 3), dtype=torch.float32).pin_memory()
    with pytest.raises(KeyError, match="logits"):
        agg({k: _pin(d[k]) for k in ("key", "cur", "scale_map", "mv")}, out_host)
    with pytest.raises(ValueError, match="pinned"):
        agg({k: torch.from_numpy(np.ascontiguousarray(d[k])) for k in ("key", "cur", "scale_map", "mv", "logits")}, out_host)


@pytest.mark.gpu
def test_host_path_gop_contract_key_uploaded_once_per_gop(cuda):
    """core/tester.py:246-252: the key feature stays on the device; a batch carries the key frames that arrived with it
    (uploaded into their table slots first) and every non-key frame names the slot it samples."""
    from lsfa_b200.host import HostAggregator
    S, C, H, W = 3, 16, 12, 9                       # three streams, one key slot each
    rng = np.random.default_rng(3)
    keys_a = O.synth_features(rng, (S, C, H, W))
    keys_b = O.synth_features(rng, (1, C, H, W))     # stream 1 gets a new key frame in the second batch
    N = 6
    agg = HostAggregator(N, C, H, W, (16 * H, 16 * W), cuda, chunk=4, depth=2, weight_mode="logits", use_scale=True, num_slots=S)
    scale = 0.0
    table = keys_a.copy()
    for batch in range(2):
        d = make_case(80 + batch, N, C, H, W)
        kidx = np.array([0, 1, 2, 2, 1, 0], np.int32)
        host_in = {k: _pin(d[k]) for k in ("cur", "scale_map", "mv", "logits")}
        host_in["key_index"] = _pin(kidx)
        if batch == 0:
            host_in["new_keys"], host_in["key_slot"] = _pin(keys_a), _pin(np.arange(S, dtype=np.int32))
        else:
            host_in["new_keys"], host_in["key_slot"] = _pin(keys_b), _pin(np.array([1], np.int32))
            table[1] = keys_b[0]
        out_host = torch.empty((N, C, H, W), dtype=torch.float32).pin_memory()
        agg(host_in, out_host)
        agg.synchronize()
        want = O.warp_scale_aggregate(table, d["flow"], cur=d["cur"], scale_map=d["scale_map"], weight_mode=O.W_LOGITS,
                                      logits=d["logits"], key_index=kidx)
        scale = max(np.abs(table).max(), np.abs(d["cur"]).max())
        assert_close_f32(out_host.numpy(), want, scale=scale, what="GOP batch %d" % batch)
        bi, bo = agg.bytes_per_call(host_in)
        F = C * H * W * 4
        assert bo == N * F
        assert bi == N * (2 * F + agg.mv_rows_copied() * 16 * W * 8 + 2 * H * W * 4 + 4) + host_in["new_keys"].shape[0] * F
