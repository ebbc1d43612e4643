"""Host-buffer front end of the fused operator: the call a reference-side user makes.

In the reference every forward goes host -> device -> host: ``_load_data`` copies the batch to
the GPU (core/DataParallelExecutorGroup.py:24-39,350), the executor runs, ``asnumpy()`` brings
the result back (core/tester.py:138-145).  ``HostAggregator`` is that contract for the
non-key-frame path with the copies made explicit and overlapped: the batch is cut into chunks
of frames, and three CUDA streams pipeline  H2D(chunk i+1) | kernel(chunk i) | D2H(chunk i-1)
over double-buffered device staging.  Inputs and the output are pinned host tensors.

Only torch memory/stream/event plumbing lives here; the arithmetic is the C-ABI fused kernel.
"""
from __future__ import annotations

import ctypes
from typing import Dict

import torch

from . import _cabi as A


class HostAggregator:
    """The fused operator from pinned HOST buffers to a pinned host buffer, pipelined: a thin holder of the caller-owned
    resources (device staging, three streams, optionally the device key table) around ONE C-ABI call per batch,
    ``lsfa_host_aggregate_f32_nchw`` (include/lsfa_ops.h; lsfa_b200/csrc/host_pipeline.cu) - the entry point a
    reference-side binding would call in place of ``_load_data`` + forward + ``asnumpy()``.

    host inputs (pinned, NCHW float32): scale_map (optional), cur (N,C,H,W); mv (N,h,w,2) int32 at network scale;
    logits (N,2,H,W) when weight_mode='logits'.  Output: pinned (N,C,H,W) float32.
    Keys: private (``host['key']`` (N,C,H,W), one per frame, uploaded with the frame) or - ``num_slots > 0`` - the GOP
    contract of the reference (core/tester.py:246-252): a device key table of num_slots features;
    ``host['key_index']`` (N,) int32 says which slot each frame samples, and ``host['new_keys']`` (K,C,H,W) +
    ``host['key_slot']`` (K,) int32 are the key frames that arrived with this batch (uploaded once, first).
    """

    MODES = {"none": A.W_NONE, "add": A.W_ADD, "mean": A.W_MEAN, "logits": A.W_LOGITS}

    def __init__(self, N: int, C: int, H: int, W: int, mv_hw, device, chunk: int = 8, depth: int = 2,
                 weight_mode: str = "logits", use_scale: bool = True, im_scale: float = 1.0, num_slots: int = 0):
        if weight_mode not in self.MODES:
            raise ValueError("weight_mode must be one of %s (the host path has no embedding inputs), got %r"
                             % (sorted(self.MODES), weight_mode))
        self.N, self.C, self.H, self.W = N, C, H, W
        self.mv_h, self.mv_w = mv_hw
        self.device = torch.device(device)
        self.chunk = max(1, min(chunk, N))
        self.depth = depth
        self.weight_mode = weight_mode
        self.use_scale = use_scale
        self.im_scale = float(im_scale)
        self.num_slots = int(num_slots)
        self._lib = A.load()
        with torch.cuda.device(self.device):
            self.s_in = torch.cuda.Stream(self.device)
            self.s_run = torch.cuda.Stream(self.device)
            self.s_out = torch.cuda.Stream(self.device)
            self.key_table = (torch.empty((self.num_slots, C, H, W), dtype=torch.float32, device=self.device)
                              if self.num_slots else None)
            probe = self._args(None, None, sizing=True)
            need = self._lib.lsfa_host_aggregate_staging_bytes(probe)
            if need == 0:
                A.check(self._lib.lsfa_host_aggregate_f32_nchw(probe))      # raises with the library's message
            # torch.empty: nothing here is read before the pipeline has written it (no fill kernel on another stream)
            self.staging = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        self._staging_off = (-self.staging.data_ptr()) % 256
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.launches = 0
        self.calls = 0

    # -- argument block ----------------------------------------------------------------------------------------
    def _args(self, host, out_host, sizing=False):
        a = A.LsfaHostAggArgs()
        a.struct_bytes = ctypes.sizeof(A.LsfaHostAggArgs)
        a.N, a.C, a.H, a.W = self.N, self.C, self.H, self.W
        a.mv_h, a.mv_w, a.im_scale = self.mv_h, self.mv_w, self.im_scale
        a.weight_mode = self.MODES[self.weight_mode]
        a.chunk, a.depth = self.chunk, self.depth
        a.stream_in, a.stream_run, a.stream_out = self.s_in.cuda_stream, self.s_run.cuda_stream, self.s_out.cuda_stream
        gop = self.num_slots > 0
        if gop:
            a.key_table, a.num_slots = self.key_table.data_ptr(), self.num_slots
        if sizing:      # placeholders: the size query looks at which inputs exist, never at their contents
            z = ctypes.c_int32 * self.N
            self._zero_index = z()
            a.key, a.mv, a.out = 256, 256, 256
            a.key_index = ctypes.addressof(self._zero_index) if gop else None
            a.scale_map = 256 if self.use_scale else None
            a.cur = 256 if self.weight_mode != "none" else None
            a.logits = 256 if self.weight_mode == "logits" else None
            return a
        need = ["mv"] + (["scale_map"] if self.use_scale else []) + (["cur"] if self.weight_mode != "none" else []) \
            + (["logits"] if self.weight_mode == "logits" else []) + (["key_index"] if gop else ["key"])
        for k in need:
            if k not in host:
                raise KeyError("host[%r] is required for weight_mode=%r, use_scale=%r, %s keys"
                               % (k, self.weight_mode, self.use_scale, "table" if gop else "private"))
        for k, t in list(host.items()) + ([("out", out_host)] if out_host is not None else []):
            if t.is_cuda or not t.is_contiguous():
                raise ValueError("%s must be a contiguous HOST tensor" % k)
            if not t.is_pinned():
                raise ValueError("%s must be pinned (page-locked) host memory for the copies to be asynchronous" % k)
        a.mv, a.out = host["mv"].data_ptr(), (out_host.data_ptr() if out_host is not None else 256)
        a.scale_map = host["scale_map"].data_ptr() if self.use_scale else None
        a.cur = host["cur"].data_ptr() if self.weight_mode != "none" else None
        a.logits = host["logits"].data_ptr() if self.weight_mode == "logits" else None
        if gop:
            a.key_index = host["key_index"].data_ptr()
            nk = host["new_keys"].shape[0] if "new_keys" in host else 0
            a.num_new_keys = nk
            if nk:
                a.key, a.key_slot = host["new_keys"].data_ptr(), host["key_slot"].data_ptr()
        else:
            a.key = host["key"].data_ptr()
        a.staging = self.staging.data_ptr() + self._staging_off
        a.staging_bytes = self.staging.numel() - self._staging_off
        return a

    def mv_rows_copied(self) -> int:
        """Rows 16k+7 and 16k+8 below mv_h: what image.py:221 (cv2.resize fx=1/16 = the centre 2x2 of each block) reads."""
        return sum(1 for r in range(self.mv_h) if r % 16 in (7, 8))

    def bytes_per_call(self, host=None):
        """(H2D, D2H) bytes of one call, as the library counts them.  GOP mode: pass the host dict (it depends on how
        many key frames arrive with the batch)."""
        if host is None and self.num_slots:
            raise ValueError("GOP mode: pass the host dict")
        a = self._args(None, None, sizing=True) if host is None else self._args(host, None)
        bi, bo = ctypes.c_size_t(0), ctypes.c_size_t(0)
        A.check(self._lib.lsfa_host_aggregate_bytes(a, ctypes.byref(bi), ctypes.byref(bo)))
        return bi.value, bo.value

    def __call__(self, host: Dict[str, torch.Tensor], out_host: torch.Tensor) -> torch.Tensor:
        """Enqueue the whole batch (one C-ABI call); returns out_host (valid after ``self.synchronize()``)."""
        a = self._args(host, out_host)
        with torch.cuda.device(self.device):
            A.check(self._lib.lsfa_host_aggregate_f32_nchw(a))
        bi, bo = ctypes.c_size_t(0), ctypes.c_size_t(0)
        A.check(self._lib.lsfa_host_aggregate_bytes(a, ctypes.byref(bi), ctypes.byref(bo)))
        self.h2d_bytes += bi.value
        self.d2h_bytes += bo.value
        n_chunks = (self.N + self.chunk - 1) // self.chunk
        self.launches += n_chunks * self.launches_per_chunk()
        self.calls += 1
        return out_host

    def launches_per_chunk(self) -> int:
        """Kernels the fused operator enqueues per chunk (record pre-pass + streaming kernel with the staging workspace)."""
        return 2

    def synchronize(self):
        self.s_in.synchronize()
        self.s_run.synchronize()
        self.s_out.synchronize()


def bind_near_gpu(device_index: int) -> dict:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off (sysfs ``local_cpulist`` of its PCI
    function), so that pinned host buffers allocated afterwards are first-touched on that node and the H2D/D2H copies do
    not cross the socket interconnect.  With one process per GPU (the reference's one-thread-per-GPU model,
    tester.py:301-309) and PCIe-bound transfers this is what keeps the per-GPU host bandwidth when several ranks copy at
    once.  Best effort: returns what it did; never raises."""
    import os
    info = {"bound": False}
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bus
        info["pci"] = bus
        with open(base + "/numa_node") as f:
            info["numa_node"] = int(f.read().strip())
        with open(base + "/local_cpulist") as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        info["cpus_local"] = len(cpus)
        if allowed and allowed != os.sched_getaffinity(0):
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
        info["cpus_now"] = len(os.sched_getaffinity(0))
    except Exception as e:       # no sysfs / no NUMA information / not permitted: stay as we are
        info["error"] = repr(e)
    return info
