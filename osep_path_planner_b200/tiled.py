"""Tiled skeleton extraction of a large accumulated map (BASELINE.json configs[4]).

The map is split into P slabs of (about) equal point count along its longest axis, one slab per rank /
GPU.  Each tile is an independent "frame" of the hot path (own adaptive leaf, own ~1000-point ROSA) once
it owns its halo, so the only data-path communication is

  1. all_reduce(min/max) of the 6 bbox floats        -> common split axis,
  2. all_gather of the slab intervals [lo_r, hi_r)    -> who borders whom,
  3. HALO EXCHANGE: points within `halo` of a slab face go to the neighbouring rank
     (torch.distributed batch_isend_irecv: NCCL send/recv over NVLink on GPUs, gloo in the CPU tests),
  4. all_gather of the owned vertices (capacity padded + counts), then the seam graph
     (graph_adj + mst over the union, gskel.cpp:201-281) on every rank.

Halo points take part as neighbours (normals, similarity, active sets) but a tile only EMITS the vertices
that fall inside its own interval.  This is a new capability (the reference has no tiling): parity is defined
per tile, against the CPU restatement run on the identical tile + halo input (tests/test_parity_gpu.py).
"""
import numpy as np
import torch
import torch.distributed as dist


def split_axis_and_bounds(points, world):
    """Single-process helper: longest axis and P+1 slab boundaries with equal point counts."""
    ext = points.max(0) - points.min(0)
    axis = int(np.argmax(ext))
    q = np.quantile(points[:, axis].astype(np.float64), np.linspace(0, 1, world + 1))
    q[0], q[-1] = -np.inf, np.inf
    return axis, q


def slab_of(points, axis, bounds, r):
    c = points[:, axis]
    return points[(c >= bounds[r]) & (c < bounds[r + 1])]


def default_halo(points, max_points=1000):
    """10 x the initial leaf estimate (similarity radius = 10 leaf, rosa.cpp:212; leaf0 = cbrt(bbox volume / max_points), rosa.hpp:99-110)."""
    ext = np.maximum(points.max(0) - points.min(0), 1e-6).astype(np.float64)
    return float(10.0 * np.cbrt(ext.prod() / max(1, max_points)))


def _dev(t):
    return t.cuda() if dist.is_initialized() and dist.get_backend() == "nccl" else t


def exchange_halo(own, axis, lo, hi, halo):
    """own: (n,3) float32 numpy array of this rank's slab, interval [lo, hi).  Returns the halo points received
    from the left and right neighbours (numpy).  Ranks are ordered along the axis (rank r borders r-1, r+1)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        z = np.zeros((0, 3), np.float32)
        return z, z
    rank, world = dist.get_rank(), dist.get_world_size()
    c = own[:, axis]
    to_left = np.ascontiguousarray(own[c < lo + halo]) if rank > 0 else np.zeros((0, 3), np.float32)
    to_right = np.ascontiguousarray(own[c >= hi - halo]) if rank < world - 1 else np.zeros((0, 3), np.float32)
    # sizes first (tiny all_gather), then the payloads point to point
    sizes = torch.tensor([len(to_left), len(to_right)], dtype=torch.int64)
    all_sizes = [torch.zeros_like(_dev(sizes)) for _ in range(world)]
    dist.all_gather(all_sizes, _dev(sizes))
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    n_from_left = int(all_sizes[rank - 1][1]) if rank > 0 else 0          # left neighbour's "to_right"
    n_from_right = int(all_sizes[rank + 1][0]) if rank < world - 1 else 0  # right neighbour's "to_left"
    recv_left = _dev(torch.empty(n_from_left, 3, dtype=torch.float32))
    recv_right = _dev(torch.empty(n_from_right, 3, dtype=torch.float32))
    send_left = _dev(torch.from_numpy(to_left)); send_right = _dev(torch.from_numpy(to_right))
    ops = []
    if rank > 0:
        if len(to_left):
            ops.append(dist.P2POp(dist.isend, send_left, rank - 1))
        if n_from_left:
            ops.append(dist.P2POp(dist.irecv, recv_left, rank - 1))
    if rank < world - 1:
        if len(to_right):
            ops.append(dist.P2POp(dist.isend, send_right, rank + 1))
        if n_from_right:
            ops.append(dist.P2POp(dist.irecv, recv_right, rank + 1))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return recv_left.cpu().numpy(), recv_right.cpu().numpy()


def global_axis_and_interval(own):
    """Collectives 1 + 2: common split axis from the global bbox, this rank's interval from its own extent
    (slabs are contiguous along the axis and ordered by rank)."""
    mn = torch.from_numpy(own.min(0).astype(np.float32)); mx = torch.from_numpy(own.max(0).astype(np.float32))
    if dist.is_initialized() and dist.get_world_size() > 1:
        mn, mx = _dev(mn), _dev(mx)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        mn, mx = mn.cpu(), mx.cpu()
    axis = int(torch.argmax(mx - mn))
    return axis


def owned_filter(vertices, axis, lo, hi):
    if len(vertices) == 0:
        return vertices
    c = vertices[:, axis]
    return vertices[(c >= lo) & (c < hi)]


def tiled_skeleton(own, bounds_lo_hi, axis, halo, skeletonize, gather_cap=2048):
    """One rank's part of the tiled pipeline.
    own: this rank's slab (n,3) float32; bounds_lo_hi: (lo, hi) of the slab along `axis`;
    skeletonize: callable (m,3) float32 -> (V,3) float32 vertices (Rosa on the GPU; a stand-in in the CPU tests).
    Returns dict(tile=tile+halo input, local=owned vertices, counts, vertices=[per-rank owned vertices])."""
    from . import shard
    lo, hi = bounds_lo_hi
    left, right = exchange_halo(own, axis, lo, hi, halo)
    tile = np.ascontiguousarray(np.concatenate([own, left, right], axis=0), dtype=np.float32)
    verts = skeletonize(tile)
    mine = np.ascontiguousarray(owned_filter(np.asarray(verts, np.float32).reshape(-1, 3), axis, lo, hi))
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    counts, allv = shard.gather_frame_results({rank: mine}, world, cap=gather_cap)
    return {"tile": tile, "halo_points": len(left) + len(right), "local": mine, "counts": counts, "vertices": allv}
