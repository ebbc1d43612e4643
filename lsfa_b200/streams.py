"""Host-side plumbing for the multi-GPU form of the path: independent video streams are
sharded over ranks with NO data-path collective.

Mirrors the reference's own scheme:
  * ``shard_streams``   - greedy arg-min bin packing of videos over GPUs by frame count,
                          dff_rfcn/function/test_rcnn.py:69-75;
  * ``key_frame_flags`` - the 0/1/2 key-frame schedule of TestLoader,
                          dff_rfcn/core/loader.py:87-131 (KEY_FRAME_INTERVAL = 12,
                          dff_rfcn/config/config.py:163);
  * ``non_key_batches`` - groups the non-key frames (flag 2) of a rank's streams into
                          fixed-size batches with the key_index that points every frame
                          at its GOP's key feature (what tester.py:251-252 does one
                          frame at a time by swapping ``feat_key`` into the batch).
Pure Python / NumPy; the only torch.distributed use is the scalar gather in
``gather_frame_counts`` (bench bookkeeping, works on gloo and nccl).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

KEY_FRAME_INTERVAL = 12


def shard_streams(frame_seg_lens: Sequence[int], world_size: int) -> List[List[int]]:
    """Stream indices per rank.  Same greedy rule as test_rcnn.py:69-75: each stream, in
    order, goes to the rank with the fewest frames so far (np.argmin: lowest rank on ties)."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    shards: List[List[int]] = [[] for _ in range(world_size)]
    loads = np.zeros(world_size, dtype=np.int64)
    for i, n in enumerate(frame_seg_lens):
        if n < 0:
            raise ValueError("negative frame_seg_len")
        r = int(np.argmin(loads))
        shards[r].append(i)
        loads[r] += int(n)
    return shards


def key_frame_flags(seg_len: int, interval: int = KEY_FRAME_INTERVAL) -> np.ndarray:
    """Per-frame flag of one video: 0 = first key frame, 1 = later key frame (also the last
    frame of the video, loader.py:124-127), 2 = non-key frame."""
    flags = np.full(seg_len, 2, dtype=np.int8)
    key = 0
    for f in range(seg_len):
        if f == key:
            flags[f] = 0 if key == 0 else 1
        elif f + 1 == seg_len:
            flags[f] = 1
        if f + 1 - key == interval:
            key = f + 1
    return flags


def non_key_batches(frame_seg_lens: Sequence[int], stream_ids: Sequence[int], batch: int,
                    interval: int = KEY_FRAME_INTERVAL) -> List[Tuple[np.ndarray, np.ndarray, np.ndarray]]:
    """Batches of (stream_id, frame_id, key_slot) arrays covering every non-key frame of the
    given streams.  key_slot indexes a per-rank table with one key feature per stream (the
    current GOP's), i.e. the ``key_index`` argument of the fused operator."""
    sid, fid, slot = [], [], []
    for s_local, s in enumerate(stream_ids):
        flags = key_frame_flags(int(frame_seg_lens[s]), interval)
        nk = np.nonzero(flags == 2)[0]
        sid.append(np.full(nk.shape, s, dtype=np.int64))
        fid.append(nk.astype(np.int64))
        slot.append(np.full(nk.shape, s_local, dtype=np.int32))
    if not sid:
        return []
    sid, fid, slot = np.concatenate(sid), np.concatenate(fid), np.concatenate(slot)
    return [(sid[i:i + batch], fid[i:i + batch], slot[i:i + batch]) for i in range(0, len(sid), batch)]


def gather_frame_counts(local_frames: int, local_ms: float):
    """(total frames over ranks, max elapsed ms over ranks).  Uses torch.distributed when a
    process group is initialised, otherwise returns the local values."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return int(local_frames), float(local_ms)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(local_frames)], dtype=torch.float64, device=dev)
    m = torch.tensor([float(local_ms)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return int(round(t.item())), float(m.item())
