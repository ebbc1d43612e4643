"""CPU suite part 3: host logic of the multi-GPU form (stream sharding, key-frame schedule)
including a world_size-2 gloo run of the bookkeeping the bench uses."""
import os
import subprocess
import sys

import numpy as np

from lsfa_b200 import streams as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_streams_is_the_reference_greedy_rule():
    lens = [464, 100, 200, 50, 300, 12, 12, 900]
    shards = S.shard_streams(lens, 3)
    # literal restatement of test_rcnn.py:69-75
    loads = np.zeros(3, dtype=np.int64); want = [[] for _ in range(3)]
    for i, n in enumerate(lens):
        g = int(np.argmin(loads)); want[g].append(i); loads[g] += n
    assert shards == want
    assert sorted(sum(shards, [])) == list(range(len(lens)))
    assert S.shard_streams([], 4) == [[], [], [], []]
    assert S.shard_streams([5, 5], 1) == [[0, 1]]


def test_key_frame_flags_follow_testloader():
    f = S.key_frame_flags(30)
    assert f[0] == 0 and f[12] == 1 and f[24] == 1 and f[29] == 1
    assert (f[1:12] == 2).all() and (f[13:24] == 2).all() and (f[25:29] == 2).all()
    assert list(S.key_frame_flags(1)) == [0]
    assert list(S.key_frame_flags(2)) == [0, 1]
    assert (S.key_frame_flags(13) == np.array([0] + [2] * 11 + [1])).all()
    # literal simulation of loader.py:87-131
    for seg in (5, 12, 13, 25, 40):
        cur, key, out = 0, 0, []
        while cur < seg:
            if key == cur:
                out.append(0 if key == 0 else 1)
            elif cur + 1 == seg:
                out.append(1)
            else:
                out.append(2)
            cur += 1
            if cur - key == 12:
                key = cur
        assert list(S.key_frame_flags(seg)) == out


def test_non_key_batches_cover_every_non_key_frame():
    lens = [30, 14, 1, 40]
    batches = S.non_key_batches(lens, [0, 1, 3], 16)
    sid = np.concatenate([b[0] for b in batches]); fid = np.concatenate([b[1] for b in batches])
    slot = np.concatenate([b[2] for b in batches])
    total = sum(int((S.key_frame_flags(lens[s]) == 2).sum()) for s in (0, 1, 3))
    assert len(sid) == total and all(len(b[0]) <= 16 for b in batches)
    assert set(zip(sid.tolist(), slot.tolist())) == {(0, 0), (1, 1), (3, 2)}
    for s, f in zip(sid, fid):
        assert S.key_frame_flags(lens[s])[f] == 2


WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from lsfa_b200 import streams as S
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
lens = [36, 24, 48, 12, 60]
mine = S.shard_streams(lens, world)[rank]
frames = sum(len(b[0]) for b in S.non_key_batches(lens, mine, 8))
total, ms = S.gather_frame_counts(frames, 10.0 + rank)
expect = sum(int((S.key_frame_flags(n) == 2).sum()) for n in lens)
assert total == expect, (total, expect)
assert ms == 10.0 + world - 1
if rank == 0:
    print("OK", total, ms)
dist.destroy_process_group()
"""


def test_world_size_2_gloo_bookkeeping(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    assert "OK" in outs[0][0]


def test_bind_near_gpu_is_best_effort_without_a_gpu():
    """host.bind_near_gpu never raises: without a CUDA device (or without NUMA information in sysfs) it reports why and
    leaves the process affinity alone."""
    import os
    pytest = __import__("pytest")
    pytest.importorskip("torch")
    from lsfa_b200.host import bind_near_gpu
    before = os.sched_getaffinity(0)
    info = bind_near_gpu(0)
    assert isinstance(info, dict) and "bound" in info
    if not info["bound"]:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
