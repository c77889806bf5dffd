"""gloo tests (world sizes 2 and 3) of the tiled-map path's exchange logic: slab partition, multi-hop halo exchange
(P2P send/recv to every rank whose widened interval reaches into the slab), ownership filter and the all-gather of owned
vertices.  The per-tile skeleton is a numpy stand-in here (voxel centroids); the real per-tile kernel parity and the
device-resident pipeline are GPU tests."""
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


def _worker3(rank, world, port, q, bounds, halo):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from osep_path_planner_b200 import tiled
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = make_map()
    own = tiled.slab_of(P, 0, bounds, rank)
    res = tiled.tiled_skeleton(own, (bounds[rank], bounds[rank + 1]), 0, halo, fake_skeleton)
    q.put((rank, res["tile"], res["local"], [np.asarray(v) for v in res["vertices"]], res["exchange"]))
    dist.barrier(); dist.destroy_process_group()


def test_multi_hop_halo_when_a_slab_is_thinner_than_the_halo():
    """Three slabs, the middle one 4 m thin, halo 10 m: ranks 0 and 2 must receive each other's points THROUGH the middle
    slab (rank +-2), not only their direct neighbour's -- a +-1-only exchange would hand them an incomplete neighbourhood."""
    sys.path.insert(0, ROOT)
    from osep_path_planner_b200 import tiled
    P = make_map()
    bounds = np.array([-np.inf, -2.0, 2.0, np.inf]); halo = 10.0
    slabs = [tiled.slab_of(P, 0, bounds, r) for r in range(3)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + (os.getpid() % 40)
    ps = [ctx.Process(target=_worker3, args=(r, 3, port, q, bounds, halo)) for r in range(3)]
    [p.start() for p in ps]
    got = {}
    for _ in range(3):
        r, tile, local, allv, st = q.get(timeout=180)
        got[r] = (tile, local, allv, st)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    for r in range(3):
        lo, hi = bounds[r], bounds[r + 1]
        parts = [slabs[r]]
        for s in range(3):
            if s != r:
                c = slabs[s][:, 0]
                parts.append(slabs[s][((c >= lo - halo) & (c < lo)) | ((c >= hi) & (c < hi + halo))])
        exp = np.concatenate(parts)
        assert np.array_equal(got[r][0], exp), "tile of rank %d" % r
    assert got[0][3]["peers"] == 2 and got[2][3]["peers"] == 2       # the outer ranks serve both other ranks
    n02 = ((slabs[2][:, 0] >= 2.0) & (slabs[2][:, 0] < 8.0)).sum()
    assert n02 > 0 and len(got[0][0]) == len(slabs[0]) + len(slabs[1]) + n02
    owned = [tiled.owned_filter(fake_skeleton(got[r][0]), 0, bounds[r], bounds[r + 1]) for r in range(3)]
    for r in range(3):
        for s in range(3):
            assert np.array_equal(got[r][2][s], owned[s])


def test_recursive_bisection_partitions_the_map():
    """rcb_boxes: disjoint boxes that cover space, equal point counts, owner == box membership; a map whose points are thin
    along its longest axis (two walls) is cut ACROSS the walls, where few points are within a halo of the cut."""
    sys.path.insert(0, ROOT)
    from osep_path_planner_b200 import tiled
    rng = np.random.default_rng(5)
    walls = [np.c_[rng.uniform(-20, 20, 8000), rng.normal(y, 0.05, 8000), rng.uniform(0, 30, 8000)] for y in (0.0, 60.0)]
    P = np.concatenate(walls).astype(np.float32)
    for world in (2, 3, 4, 8):
        boxes, owner = tiled.rcb_boxes(P, world, 5.0)
        inside = np.stack([np.all((P >= boxes[r, 0]) & (P < boxes[r, 1]), axis=1) for r in range(world)])
        assert np.array_equal(inside.sum(0), np.ones(len(P)))                      # every point in exactly one box
        assert np.array_equal(np.argmax(inside, 0), owner)
        cnt = np.bincount(owner, minlength=world)
        assert cnt.max() - cnt.min() <= world
        dup = sum(int(np.all((P >= boxes[r, 0] - 5.0) & (P < boxes[r, 1] + 5.0), axis=1).sum()) for r in range(world)) - len(P)
        if world == 2:
            assert dup == 0                                                         # the cut separates the two walls: no halo at all
        ax, b = tiled.split_axis_and_bounds(P, world)
        slab_dup = sum(int(((P[:, ax] >= b[r] - 5.0) & (P[:, ax] < b[r + 1] + 5.0)).sum()) for r in range(world)) - len(P)
        assert dup <= slab_dup
