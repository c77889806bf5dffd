"""world_size-2 gloo test of the tiled-map path's host logic: slab partition, halo exchange (P2P send/recv),
ownership filter and the all-gather of owned vertices.  The per-tile skeleton is a numpy stand-in here (voxel
centroids); the real per-tile kernel parity is a GPU test."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fake_skeleton(tile):
    k = np.floor(tile / 2.0).astype(np.int64)
    _, inv = np.unique(k, axis=0, return_inverse=True)
    out = np.zeros((inv.max() + 1, 3), np.float64); cnt = np.zeros(inv.max() + 1)
    np.add.at(out, inv.reshape(-1), tile); np.add.at(cnt, inv.reshape(-1), 1)
    return (out / cnt[:, None]).astype(np.float32)


def make_map():
    rng = np.random.default_rng(0)
    return (rng.uniform([-30, -5, 0], [30, 5, 20], (6000, 3))).astype(np.float32)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from osep_path_planner_b200 import tiled
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = make_map()
    axis, bounds = tiled.split_axis_and_bounds(P, world)
    own = tiled.slab_of(P, axis, bounds, rank)
    assert tiled.global_axis_and_interval(own) == axis
    res = tiled.tiled_skeleton(own, (bounds[rank], bounds[rank + 1]), axis, 3.0, fake_skeleton)
    q.put((rank, res["tile"], res["local"], [np.asarray(v) for v in res["vertices"]], res["halo_points"]))
    dist.barrier(); dist.destroy_process_group()


def test_halo_exchange_and_vertex_gather_world2():
    sys.path.insert(0, ROOT)
    from osep_path_planner_b200 import tiled
    P = make_map()
    axis, bounds = tiled.split_axis_and_bounds(P, 2)
    assert axis == 0
    slabs = [tiled.slab_of(P, axis, bounds, r) for r in range(2)]
    assert abs(len(slabs[0]) - len(slabs[1])) <= 1 and len(slabs[0]) + len(slabs[1]) == len(P)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    got = {}
    for _ in range(2):
        r, tile, local, allv, nh = q.get(timeout=120)
        got[r] = (tile, local, allv, nh)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    halo = 3.0
    cut = bounds[1]
    # rank 0 receives exactly rank 1's points within `halo` of the cut, and vice versa
    exp0 = np.concatenate([slabs[0], slabs[1][slabs[1][:, 0] < cut + halo]])
    exp1 = np.concatenate([slabs[1], slabs[0][slabs[0][:, 0] >= cut - halo]])
    assert np.array_equal(got[0][0], exp0) and np.array_equal(got[1][0], exp1)
    assert got[0][3] == len(exp0) - len(slabs[0]) > 0 and got[1][3] == len(exp1) - len(slabs[1]) > 0
    # owned vertices: inside the rank's interval only; the gather gives both ranks the same union
    v0 = tiled.owned_filter(fake_skeleton(exp0), 0, bounds[0], bounds[1]); v1 = tiled.owned_filter(fake_skeleton(exp1), 0, bounds[1], bounds[2])
    assert np.array_equal(got[0][1], v0) and np.array_equal(got[1][1], v1)
    for r in range(2):
        assert np.array_equal(got[r][2][0], v0) and np.array_equal(got[r][2][1], v1)
