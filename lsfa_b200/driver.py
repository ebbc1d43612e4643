"""Batched multi-stream driver of the non-key-frame path (scheduler half).

What the reference does one frame at a time on the host - ``TestLoader`` emits ``key_frame_flag``
(core/loader.py:113-131), ``pred_eval`` keeps the last key feature and swaps it into the next
non-key batch as ``feat_key`` (core/tester.py:242-253) - becomes: a device-resident table with ONE
key feature per stream (the current GOP's), and one fused launch per batch of non-key frames taken
from many streams, each frame pointing at its stream's slot through ``key_index``.  Streams are
sharded over ranks with ``streams.shard_streams`` (test_rcnn.py:69-75); no collective.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import ops, streams


class StreamScheduler:
    def __init__(self, frame_seg_lens: Sequence[int], channels: int, feat_hw, device, rank: int = 0,
                 world_size: int = 1, interval: int = streams.KEY_FRAME_INTERVAL):
        self.seg_lens = [int(v) for v in frame_seg_lens]
        self.stream_ids = streams.shard_streams(self.seg_lens, world_size)[rank]
        self.slot_of = {s: i for i, s in enumerate(self.stream_ids)}
        self.interval = interval
        self.device = torch.device(device)
        h, w = feat_hw
        # one key feature per local stream: 9.8 MB fp32 each at 1024x38x63 (1024 streams = 10 GB)
        self.key_table = torch.zeros((max(1, len(self.stream_ids)), channels, h, w), dtype=torch.float32,
                                     device=self.device)
        self.flags = {s: streams.key_frame_flags(self.seg_lens[s], interval) for s in self.stream_ids}

    def set_key_feature(self, stream_id: int, feat: torch.Tensor) -> None:
        """After a key frame's forward (tester.py:246-249): remember its (aggregated) feature."""
        self.key_table[self.slot_of[stream_id]].copy_(feat.reshape(self.key_table.shape[1:]), non_blocking=True)

    def batches(self, batch: int):
        """(stream_id, frame_id, key_slot) arrays covering every non-key frame of this rank's streams."""
        return streams.non_key_batches(self.seg_lens, self.stream_ids, batch, self.interval)

    def run_non_key_batch(self, key_slots, mv: torch.Tensor, cur: torch.Tensor, res: Optional[torch.Tensor] = None,
                          rnet_w: Optional[torch.Tensor] = None, rnet_b: Optional[torch.Tensor] = None,
                          flow_kind: str = "raw", im_scale: float = 1.0, out: Optional[torch.Tensor] = None,
                          **kw) -> torch.Tensor:
        """get_cur_test_symbol for a whole batch (SYM:570-586): out[i] = cur[i] + warp(key_table[slot[i]], mv[i])
        [+ rnet(res[i])].  key_slots: (B,) int array of table slots."""
        if isinstance(key_slots, torch.Tensor):        # already on the device: no copy, no sync (batched harness)
            idx = key_slots.to(device=self.device, dtype=torch.int32)
        else:
            idx = torch.as_tensor(np.asarray(key_slots, dtype=np.int32), device=self.device)
        return ops.warp_scale_aggregate(self.key_table, mv, key_index=idx, cur=cur, res=res, rnet_w=rnet_w,
                                        rnet_b=rnet_b, weight_mode="add", flow_kind=flow_kind, im_scale=im_scale,
                                        out=out, **kw)
