"""Tiled skeleton extraction of a large accumulated map (BASELINE.json configs[4]).

The map is split into P axis-aligned boxes of (about) equal point count, one per rank / GPU -- slabs along one axis
(split_axis_and_bounds) or, for structure-like maps whose points are thin along every axis somewhere, a recursive
coordinate bisection that cuts where the fewest points lie within a halo width of the cut (rcb_boxes) -- and STAYS ON
THE DEVICE: a rank's slab is a CUDA tensor, the halo points are selected by a kernel of the
library (osep_points_select_range: stable compaction by the axis test), travel rank to rank with NCCL
send/recv straight into the tail of the tile buffer, the tile runs as one frame of the hot path
(osep_rosa_run_device), the owned vertices are masked on the device and all-gathered with NCCL, and the seam
graph (graph_adj + mst over the union, gskel.cpp:201-281) is built from the gathered device buffer.  Only
scalars (intervals, halo widths, counts) pass through the host.

  1. all_reduce(min/max) of the 6 bbox floats                       -> common split axis,
  2. all_gather of (lo_r, hi_r, halo_r)                              -> who needs what,
  3. HALO EXCHANGE: rank s sends rank r every own point within halo_r of r's interval -- to EVERY rank whose
     widened interval reaches into s's slab, not only to r = s +- 1 (equal-count slabs can be thinner than the halo),
  4. all_gather of the owned vertices (capacity padded + counts), then the seam graph on every rank.

halo_r = 10 x the leaf size the adaptive VoxelGrid law converges to on the rank's own points
(osep_rosa_leaf_device; 10 leaf = the similarity radius, rosa.cpp:212) unless the caller fixes it.
Halo points take part as neighbours (normals, similarity, active sets) but a tile only EMITS the vertices that
fall inside its own interval.  This is a new capability (the reference has no tiling): parity is defined per
tile, against the CPU restatement run on the identical tile + halo input (tests/test_parity_gpu.py).

The same functions run on CPU tensors with the gloo backend (tests/test_tiled_cpu.py: the exchange / ownership /
gather logic with world sizes 2 and 3); the selection is then a torch mask -- host-logic test path only.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


# ---- partition helpers (single process, host: used to build the synthetic map and by the tests) ---------------------
def split_axis_and_bounds(points, world, halo=None):
    """Split axis and P+1 slab boundaries with equal point counts.  halo=None: the longest axis (the single-rank / test
    default).  With a halo estimate the axis is the one whose largest TILE (slab + halo on both sides) is smallest: on
    structure-like maps the longest axis can be the one most of the points are thin along (rotor planes, facades), and
    equal-count slabs thinner than the halo make every tile several slabs wide."""
    def bounds_of(axis):
        q = np.quantile(points[:, axis].astype(np.float64), np.linspace(0, 1, world + 1))
        q[0], q[-1] = -np.inf, np.inf
        return q
    ext = points.max(0) - points.min(0)
    if halo is None or world == 1:
        axis = int(np.argmax(ext))
        return axis, bounds_of(axis)
    best = None
    for axis in range(3):
        q = bounds_of(axis)
        c = np.sort(points[:, axis].astype(np.float64))
        sizes = [np.searchsorted(c, q[r + 1] + halo) - np.searchsorted(c, q[r] - halo) for r in range(world)]
        if best is None or max(sizes) < best[0]:
            best = (max(sizes), axis, q)
    return best[1], best[2]


def slab_of(points, axis, bounds, r):
    c = points[:, axis]
    return points[(c >= bounds[r]) & (c < bounds[r + 1])]


def default_halo(points, max_points=1000):
    """10 x the initial leaf estimate (similarity radius = 10 leaf, rosa.cpp:212; leaf0 = cbrt(bbox volume / max_points), rosa.hpp:99-110)."""
    ext = np.maximum(points.max(0) - points.min(0), 1e-6).astype(np.float64)
    return float(10.0 * np.cbrt(ext.prod() / max(1, max_points)))


def owned_filter(vertices, axis, lo, hi):
    if len(vertices) == 0:
        return vertices
    c = vertices[:, axis]
    return vertices[(c >= lo) & (c < hi)]


def rcb_boxes(points, world, halo):
    """Recursive coordinate bisection: `world` boxes (lo[3], hi[3]) of equal point count; every cut is taken along the axis
    with the fewest points within `halo` of the cut plane (the points both sides would have to exchange).  Returns
    (boxes float64 (world, 2, 3), owner index per point)."""
    n = len(points)
    boxes = np.empty((world, 2, 3)); owner = np.zeros(n, np.int32)
    def rec(idx, lo, hi, first, k):
        if k == 1:
            boxes[first, 0] = lo; boxes[first, 1] = hi; owner[idx] = first
            return
        k1 = k // 2
        best = None
        for axis in range(3):
            c = points[idx, axis].astype(np.float64)
            cut = float(np.quantile(c, k1 / k))
            dup = int(((c >= cut - halo) & (c < cut + halo)).sum())
            if best is None or dup < best[0]:
                best = (dup, axis, cut)
        _, axis, cut = best
        left = points[idx, axis] < cut
        hi1 = hi.copy(); hi1[axis] = cut; lo2 = lo.copy(); lo2[axis] = cut
        rec(idx[left], lo, hi1, first, k1); rec(idx[~left], lo2, hi, first + k1, k - k1)
    rec(np.arange(n), np.full(3, -np.inf), np.full(3, np.inf), 0, world)
    return boxes, owner


def slab_box(axis, lo, hi):
    b = np.array([[-np.inf] * 3, [np.inf] * 3]); b[0, axis] = lo; b[1, axis] = hi
    return b


def owned_filter_box(vertices, box):
    if len(vertices) == 0:
        return vertices
    m = np.all((vertices[:, :3] >= box[0]) & (vertices[:, :3] < box[1]), axis=1)
    return vertices[m]


# ---- device-side primitives -------------------------------------------------------------------------------------
def _world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def select_range(t, axis, lo, hi, out=None, out_stride=None):
    """Rows of t (n, 3|4) with lo <= t[:, axis] < hi, in input order.  CUDA tensors: the library's compaction kernel
    (into `out` when given: returns the row count); CPU tensors (gloo tests): a torch mask."""
    lo = float(max(lo, -3.0e38)); hi = float(min(hi, 3.0e38))
    if not t.is_cuda:
        sel = t[(t[:, axis] >= lo) & (t[:, axis] < hi)]
        if out is None:
            return sel
        out[:len(sel), :3] = sel[:, :3]
        if out.shape[1] > 3:
            out[:len(sel), 3] = 1.0
        return len(sel)
    L = _lib.load()
    n = int(t.shape[0])
    alloc = out is None
    if alloc:
        out = torch.empty((max(n, 1), out_stride or 3), dtype=torch.float32, device=t.device)
    cnt = C.c_int(0)
    s = torch.cuda.current_stream(t.device).cuda_stream
    rc = L.osep_points_select_range(t.device.index, C.c_void_p(t.data_ptr()), n, int(t.stride(0)), int(axis), lo, hi,
                                    C.c_void_p(out.data_ptr()), int(out.shape[1]), int(out.shape[0]), C.byref(cnt), C.c_void_p(s))
    if rc != 0:
        raise _lib.OsepError("osep_points_select_range failed (%d): %s" % (rc, _lib.last_error()))
    return out[:cnt.value] if alloc else cnt.value


def select_box(t, lo3, hi3, out=None, out_stride=None):
    """Rows of t (n, 3|4) inside the box lo3 <= p < hi3, in input order (CUDA: osep_points_select_box; CPU tensors: a torch mask)."""
    lo = np.maximum(np.asarray(lo3, np.float64), -3.0e38).astype(np.float32); hi = np.minimum(np.asarray(hi3, np.float64), 3.0e38).astype(np.float32)
    if not t.is_cuda:
        m = torch.ones(t.shape[0], dtype=torch.bool)
        for d in range(3):
            m &= (t[:, d] >= float(lo[d])) & (t[:, d] < float(hi[d]))
        sel = t[m]
        if out is None:
            return sel
        out[:len(sel), :3] = sel[:, :3]
        if out.shape[1] > 3:
            out[:len(sel), 3] = 1.0
        return len(sel)
    L = _lib.load()
    n = int(t.shape[0])
    alloc = out is None
    if alloc:
        out = torch.empty((max(n, 1), out_stride or 3), dtype=torch.float32, device=t.device)
    cnt = C.c_int(0)
    s = torch.cuda.current_stream(t.device).cuda_stream
    rc = L.osep_points_select_box(t.device.index, C.c_void_p(t.data_ptr()), n, int(t.stride(0)), lo.ctypes.data_as(C.c_void_p), hi.ctypes.data_as(C.c_void_p),
                                  C.c_void_p(out.data_ptr()), int(out.shape[1]), int(out.shape[0]), C.byref(cnt), C.c_void_p(s))
    if rc != 0:
        raise _lib.OsepError("osep_points_select_box failed (%d): %s" % (rc, _lib.last_error()))
    return out[:cnt.value] if alloc else cnt.value


def global_axis(own):
    """Collective 1: common split axis from the global bbox."""
    mn = own[:, :3].min(0).values.float(); mx = own[:, :3].max(0).values.float()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return int(torch.argmax(mx - mn))


def global_axis_and_interval(own):
    """numpy front end of global_axis (kept for callers that hold the slab on the host)."""
    return global_axis(torch.from_numpy(np.ascontiguousarray(own, dtype=np.float32)))


def gather_table(box, halo, device, own=None):
    """Collective 2: (lo[3], hi[3], halo, bbox_lo[3], bbox_hi[3]) of every rank, as a float64 numpy (world, 13) array
    (scalars only).  bbox = the extent of the rank's own points: a peer whose widened box misses it gets nothing."""
    rank, world = _world()
    if own is not None and own.shape[0] > 0:
        bb = torch.cat([own[:, :3].min(0).values, own[:, :3].max(0).values]).double().cpu().numpy()
    else:
        bb = np.array([-np.inf] * 3 + [np.inf] * 3)
    row = np.concatenate([np.asarray(box, np.float64).reshape(6), [float(halo)], bb])
    row = np.clip(row, -1.0e300, 1.0e300)                              # +-inf does not survive every collective backend
    mine = torch.tensor(row, dtype=torch.float64, device=device)
    if world == 1:
        return mine.cpu().numpy().reshape(1, 13)
    rows = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(rows, mine)
    return torch.stack(rows).cpu().numpy()


def halo_plan(table, rank):
    """What rank `rank` sends: [(peer, lo3, hi3)] = peer's box widened by peer's halo, clipped to this rank's box and to the
    extent of its points, for every peer whose widened box reaches them (boxes are disjoint, so every own point inside is a
    halo point of the peer; with thin tiles this is more than the direct neighbours: multi-hop)."""
    lo_s, hi_s = np.maximum(table[rank, 0:3], table[rank, 7:10]), np.minimum(table[rank, 3:6], np.nextafter(table[rank, 10:13].astype(np.float32), np.float32(np.inf)).astype(np.float64))
    plan = []
    for r in range(len(table)):
        if r == rank:
            continue
        lo_r, hi_r, h = table[r, 0:3], table[r, 3:6], table[r, 6]
        a = np.maximum(lo_s, lo_r - h); b = np.minimum(hi_s, hi_r + h)
        if np.all(a < b):
            plan.append((r, a, b))
    return plan


def exchange_halo(own, table, tile=None):
    """Collective 3.  own: (n, 3|4) tensor of this rank's points (rows ordered as stored).  Returns (tile, n_halo, stats):
    tile = own rows followed by the halo rows received from the other ranks in rank order, (n + n_halo, 3)."""
    rank, world = _world()
    n = int(own.shape[0])
    dev = own.device
    plan = halo_plan(table, rank)
    send = torch.empty((max(n * max(len(plan), 1), 1), 3), dtype=torch.float32, device=dev)      # a thin tile may go whole to several peers
    counts = torch.zeros(world, dtype=torch.int64)
    segs, off = {}, 0
    for peer, a, b in plan:
        c = select_box(own, a, b, out=send[off:])
        segs[peer] = (off, c); counts[peer] = c; off += c
    if world > 1:
        allc = [torch.zeros(world, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(allc, counts.to(dev))
        recv = torch.stack(allc).cpu().numpy()[:, rank]          # recv[s] = rows rank s sends here
    else:
        recv = np.zeros(1, np.int64)
    n_halo = int(recv.sum())
    if tile is None or tile.shape[0] < n + n_halo:
        tile = torch.empty((n + n_halo, 3), dtype=torch.float32, device=dev)
    tile[:n] = own[:, :3]
    ops, roff = [], n
    for s in range(world):
        if s != rank and recv[s] > 0:
            ops.append(dist.P2POp(dist.irecv, tile[roff:roff + int(recv[s])], s)); roff += int(recv[s])
    for peer, (o, c) in segs.items():
        if c > 0:
            ops.append(dist.P2POp(dist.isend, send[o:o + c], peer))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return tile[:n + n_halo], n_halo, {"sent_rows": int(off), "recv_rows": n_halo, "peers": len(plan)}


def gather_vertices(mine, cap=512):
    """Collective 4: mine = (V, 3|4) owned vertices of this rank.  Returns (counts[world], all (sum V, C) tensor) on every
    rank; the padded buffers are sized from the largest count (all_reduce MAX), so nothing is ever truncated."""
    rank, world = _world()
    V = int(mine.shape[0]); Cc = int(mine.shape[1])
    if world == 1:
        return np.array([V], np.int64), mine
    dev = mine.device
    vmax = torch.tensor([V], dtype=torch.int64, device=dev)
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
    rows = max(int(vmax.item()), cap)
    buf = torch.zeros((rows + 1, Cc), dtype=torch.float32, device=dev)
    buf[:V] = mine; buf[rows, 0] = float(V)                       # the count rides in the last row: one collective
    allb = torch.empty((world * (rows + 1), Cc), dtype=torch.float32, device=dev)
    if dev.type == "cuda":
        dist.all_gather_into_tensor(allb, buf)
    else:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf); allb = torch.cat(parts)
    allb = allb.view(world, rows + 1, Cc)
    counts = allb[:, rows, 0].to(torch.int64).cpu().numpy()
    return counts, torch.cat([allb[r, :int(counts[r])] for r in range(world)])


class TileRunner:
    """One rank's part of the tiled pipeline on its GPU: leaf -> halo -> exchange -> frame -> owned mask -> gather -> seam graph."""

    def __init__(self, device, cfg=None, fuse_dist_th=3.0):
        from .rosa import Rosa, RosaConfig
        self.device = int(device)
        self.cfg = cfg or RosaConfig(pts_dist_lim=1.0e4)
        self._Rosa = Rosa
        self.rosa = None
        self.gskel = None
        self.fuse = fuse_dist_th
        self.tile_buf = None
        self.L = _lib.load()

    def _handle(self, n):
        if self.rosa is None or self.rosa.capacity < n:
            if self.rosa is not None:
                self.rosa.close()
            self.rosa = self._Rosa(self.cfg, device=self.device, capacity_points=int(n * 1.25) + 4096)
            self.rosa.capacity = int(n * 1.25) + 4096
        return self.rosa

    def own_leaf(self, own):
        """converged leaf size of the rank's own points (two-phase halo)"""
        r = self._handle(int(own.shape[0]))
        leaf = C.c_float(0.0); nf = C.c_int(0)
        rc = self.L.osep_rosa_leaf_device(r._h, C.c_void_p(own.data_ptr()), int(own.shape[0]), int(own.stride(0)), C.byref(leaf), C.byref(nf))
        if rc < 0:
            raise _lib.OsepError("osep_rosa_leaf_device failed (%d): %s" % (rc, _lib.last_error()))
        return float(leaf.value)

    def run(self, own, axis=None, lo=None, hi=None, halo=None, box=None):
        """own: CUDA tensor (n, 3|4) of this rank's points; the tile is either the slab [lo, hi) along `axis` or the box
        `box` = (lo[3], hi[3]).  Returns a dict with the owned vertices (CUDA tensor), the per-rank counts, all vertices
        (CUDA tensor) and the seam graph."""
        import time
        ev = lambda: torch.cuda.Event(enable_timing=True)
        e = [ev() for _ in range(5)]
        t0 = time.perf_counter()
        if box is None:
            box = slab_box(axis, lo, hi)
        box = np.asarray(box, np.float64)
        e[0].record()
        if halo is None:
            halo = 10.0 * self.own_leaf(own)
        table = gather_table(box, halo, own.device, own)
        tile, n_halo, st = exchange_halo(own, table, self.tile_buf)
        self.tile_buf = tile if self.tile_buf is None or self.tile_buf.shape[0] < tile.shape[0] else self.tile_buf
        e[1].record()
        r = self._handle(int(tile.shape[0]))
        torch.cuda.current_stream().synchronize()           # the library runs on its own stream
        r.run_device(tile.data_ptr(), int(tile.shape[0]), 3)
        ok = r.finish()
        e[2].record()
        vp = C.c_void_p(); nv = C.c_int(0)
        self.L.osep_rosa_vertices_device(r._h, C.byref(vp), C.byref(nv))
        V = nv.value if ok else 0
        mine = torch.zeros((0, 4), dtype=torch.float32, device=own.device)
        if V:
            out = torch.empty((V, 4), dtype=torch.float32, device=own.device)
            cnt = C.c_int(0)
            blo = np.maximum(box[0], -3.0e38).astype(np.float32); bhi = np.minimum(box[1], 3.0e38).astype(np.float32)
            rc = self.L.osep_points_select_box(self.device, vp, V, 4, blo.ctypes.data_as(C.c_void_p), bhi.ctypes.data_as(C.c_void_p),
                                               C.c_void_p(out.data_ptr()), 4, V, C.byref(cnt), C.c_void_p(torch.cuda.current_stream().cuda_stream))
            if rc != 0:
                raise _lib.OsepError("owned-vertex mask failed (%d): %s" % (rc, _lib.last_error()))
            mine = out[:cnt.value]
        counts, allv = gather_vertices(mine)
        e[3].record()
        graph = {"n_edges": 0}
        if allv.shape[0] > 1:
            from .gskel import GSkel, _view_arrays
            from ._lib import GraphView
            if self.gskel is None or self.gskel.cap < allv.shape[0]:
                self.gskel = GSkel(device=self.device, capacity_vertices=max(1024, int(allv.shape[0]) * 2))
                self.gskel.cap = max(1024, int(allv.shape[0]) * 2)
            v = GraphView()
            torch.cuda.current_stream().synchronize()
            rc = self.L.osep_graph_build_device(self.gskel._h, C.c_void_p(allv.data_ptr()), int(allv.shape[0]), 4, self.fuse, C.byref(v))
            if rc != 0:
                raise _lib.OsepError("osep_graph_build_device failed (%d): %s" % (rc, _lib.last_error()))
            graph = _view_arrays(v)
        e[4].record(); torch.cuda.synchronize()
        ms = [e[i].elapsed_time(e[i + 1]) for i in range(4)]
        return {"ok": ok, "tile_points": int(tile.shape[0]), "halo_points": n_halo, "halo_m": float(halo), "halo_bytes": 12 * n_halo,
                "local": mine, "counts": counts, "vertices": allv, "graph": graph, "exchange": st, "gpu_ms": r.result.gpu_ms,
                "ms": {"leaf_select_exchange": ms[0], "frame": ms[1], "mask_gather": ms[2], "seam_graph": ms[3]},
                "wall_s": time.perf_counter() - t0, "status": r.result.status, "M": r.result.n_downsampled}

    def close(self):
        if self.rosa is not None:
            self.rosa.close(); self.rosa = None
        if self.gskel is not None:
            self.gskel.close(); self.gskel = None


def tiled_skeleton(own, bounds_lo_hi, axis, halo, skeletonize, box=None):
    """Host-logic form of TileRunner.run for CPU tensors / numpy (gloo tests): `skeletonize` is a callable
    (m, 3) float32 numpy -> (V, 3) float32 numpy.  Same exchange, ownership and gather code as the device path."""
    if box is None:
        box = slab_box(axis, bounds_lo_hi[0], bounds_lo_hi[1])
    box = np.asarray(box, np.float64)
    t = torch.from_numpy(np.ascontiguousarray(own, dtype=np.float32))
    table = gather_table(box, halo, t.device, t)
    tile, n_halo, st = exchange_halo(t, table)
    verts = np.asarray(skeletonize(tile.numpy()), np.float32).reshape(-1, 3)
    mine = torch.from_numpy(np.ascontiguousarray(owned_filter_box(verts, box)))
    counts, allv = gather_vertices(mine)
    allv = allv.numpy(); offs = np.concatenate([[0], np.cumsum(counts)])
    return {"tile": tile.numpy(), "halo_points": n_halo, "local": mine.numpy(), "counts": counts,
            "vertices": [allv[offs[r]:offs[r + 1]] for r in range(len(counts))], "exchange": st}
