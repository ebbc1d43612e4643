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

from typing import Dict, Optional

import torch

from . import _cabi as A
from . import ops

_FEATURE_KEYS = ("key", "scale_map", "cur")


class HostAggregator:
    """Key-frame Nq-style aggregation (warp x scale, softmax-logit blend) from host buffers.

    host inputs (pinned, NCHW float32): key, scale_map, cur (N,C,H,W); mv (N,h,w,2) int32;
    logits (N,2,H,W).  Output: pinned (N,C,H,W) float32.
    """

    def __init__(self, N: int, C: int, H: int, W: int, mv_hw, device, chunk: int = 8, depth: int = 2,
                 weight_mode: str = "logits", use_scale: bool = True):
        self.N, self.C, self.H, self.W = N, C, H, W
        self.mv_h, self.mv_w = mv_hw
        self.device = torch.device(device)
        self.chunk = max(1, min(chunk, N))
        self.depth = depth
        self.weight_mode = weight_mode
        self.use_scale = use_scale
        f = (self.chunk, C, H, W)
        self.stage = []
        for _ in range(depth):
            buf = {k: torch.empty(f, dtype=torch.float32, device=self.device) for k in _FEATURE_KEYS}
            buf["out"] = torch.empty(f, dtype=torch.float32, device=self.device)
            # zeroed once: only the rows the stride-16 reduction reads are ever copied in (see __call__)
            buf["mv"] = torch.zeros((self.chunk, self.mv_h, self.mv_w, 2), dtype=torch.int32, device=self.device)
            buf["logits"] = torch.empty((self.chunk, 2, H, W), dtype=torch.float32, device=self.device)
            self.stage.append(buf)
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]      # inputs of slot landed
        self.ev_run = [torch.cuda.Event() for _ in range(depth)]     # kernel of slot done
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]     # output of slot copied out
        self._primed = [False] * depth   # slot has been used (its events recorded) at least once
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.launches = 0
        self._lib = A.load()

    def mv_rows_copied(self) -> int:
        """Rows 16k+7 and 16k+8 below mv_h: what image.py:221 (cv2.resize fx=1/16 = the centre 2x2 of each block) reads."""
        return sum(1 for r in range(self.mv_h) if r % 16 in (7, 8))

    def bytes_per_call(self):
        per_frame_in = (3 if self.use_scale else 2) * self.C * self.H * self.W * 4 \
            + self.mv_rows_copied() * self.mv_w * 2 * 4 + 2 * self.H * self.W * 4
        per_frame_out = self.C * self.H * self.W * 4
        return self.N * per_frame_in, self.N * per_frame_out

    def __call__(self, host: Dict[str, torch.Tensor], out_host: torch.Tensor) -> torch.Tensor:
        """Enqueue the whole batch; returns out_host (valid after ``self.synchronize()``)."""
        n_chunks = (self.N + self.chunk - 1) // self.chunk
        feats = [k for k in _FEATURE_KEYS if (k != "scale_map" or self.use_scale)]
        for i in range(n_chunks):
            slot = i % self.depth
            lo, hi = i * self.chunk, min(self.N, (i + 1) * self.chunk)
            m = hi - lo
            buf = self.stage[slot]
            with torch.cuda.stream(self.s_in):
                if self._primed[slot]:
                    self.s_in.wait_event(self.ev_run[slot])      # previous kernel on this slot has read its inputs
                for k in feats + ["logits"]:
                    buf[k][:m].copy_(host[k][lo:hi], non_blocking=True)
                # motion vectors: the parity-mode reduction reads 2 rows of every 16, so only those cross PCIe
                # (1/8 of the field: 0.6 MB instead of 4.8 MB per 600x1000 frame), into a full-size device image
                A.check(self._lib.lsfa_mv_centre_rows_h2d(host["mv"][lo:hi].data_ptr(), buf["mv"].data_ptr(), m, self.mv_h,
                                                          self.mv_w, None, self.s_in.cuda_stream))
                self.ev_in[slot].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[slot])
                if self._primed[slot]:
                    self.s_run.wait_event(self.ev_out[slot])     # previous output of this slot has left
                ops.warp_scale_aggregate(buf["key"][:m], buf["mv"][:m], flow_kind="raw", cur=buf["cur"][:m],
                                         scale_map=buf["scale_map"][:m] if self.use_scale else None,
                                         weight_mode=self.weight_mode,
                                         logits=buf["logits"][:m] if self.weight_mode == "logits" else None,
                                         out=buf["out"][:m])
                self.launches += 1
                self.ev_run[slot].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_run[slot])
                out_host[lo:hi].copy_(buf["out"][:m], non_blocking=True)
                self.ev_out[slot].record(self.s_out)
            self._primed[slot] = True
        bi, bo = self.bytes_per_call()
        self.h2d_bytes += bi
        self.d2h_bytes += bo
        return out_host

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
