"""world_size-2 gloo test of the N>1 path's host logic (frame sharding + result gather)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from osep_path_planner_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 11
    mine = shard.frames_of_rank(B, rank, world)
    # stand-in for the per-frame result (V, vertices): deterministic function of the frame id
    local = {b: np.full((b % 4 + 1, 3), float(b), np.float32) for b in mine}
    counts, verts = shard.gather_frame_results(local, B, cap=2)      # cap below the largest count: nothing may be truncated
    tmax = shard.max_over_ranks(float(rank + 1))
    if rank == 0:
        q.put((sorted(mine), counts.tolist(), [v.tolist() for v in verts], tmax))
    dist.barrier(); dist.destroy_process_group()


def test_batch_sharding_and_gather_world2():
    from osep_path_planner_b200 import shard
    assert shard.frames_of_rank(10, 0, 1) == list(range(10))
    a, b = shard.frames_of_rank(11, 0, 2), shard.frames_of_rank(11, 1, 2)
    assert sorted(a + b) == list(range(11)) and abs(len(a) - len(b)) <= 1 and a == list(range(0, 6)) and b == list(range(6, 11))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    mine, counts, verts, tmax = q.get(timeout=120)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert mine == list(range(6)) and counts == [b % 4 + 1 for b in range(11)] and tmax == 2.0
    for b in range(11):
        assert np.array_equal(np.asarray(verts[b], np.float32), np.full((b % 4 + 1, 3), float(b), np.float32))
