"""Tiled skeleton extraction of a large accumulated map (BASELINE.json configs[4]).

The map is split into P slabs of (about) equal point count along its longest axis, one slab per rank /
GPU, and STAYS ON THE DEVICE: a rank's slab is a CUDA tensor, the halo points are selected by a kernel of the
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
def split_axis_and_bounds(points, world):
    """Longest axis and P+1 slab boundaries with equal point counts."""
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


def owned_filter(vertices, axis, lo, hi):
    if len(vertices) == 0:
        return vertices
    c = vertices[:, axis]
    return vertices[(c >= lo) & (c < hi)]


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


def global_axis(own):
    """Collective 1: common split axis from the global bbox."""
    mn = own[:, :3].min(0).values.float(); mx = own[:, :3].max(0).values.float()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return int(torch.argmax(mx - mn))


def global_axis_and_interval(own):
    """numpy front end of global_axis (kept for callers that hold the slab on the host)."""
    return global_axis(torch.from_numpy(np.ascontiguousarray(own, dtype=np.float32)))


def gather_table(lo, hi, halo, device):
    """Collective 2: (lo_r, hi_r, halo_r) of every rank, as a float64 numpy (world, 3) array (scalars only)."""
    rank, world = _world()
    mine = torch.tensor([lo, hi, halo], dtype=torch.float64, device=device)
    if world == 1:
        return mine.cpu().numpy().reshape(1, 3)
    rows = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(rows, mine)
    return torch.stack(rows).cpu().numpy()


def halo_plan(table, rank):
    """What rank `rank` sends: [(peer, lo, hi)] = the part of peer's widened interval that reaches into this rank's slab.
    Every peer whose interval lies within its own halo width of this slab gets a segment (multi-hop)."""
    lo_s, hi_s = table[rank, 0], table[rank, 1]
    plan = []
    for r in range(len(table)):
        if r == rank:
            continue
        lo_r, hi_r, halo_r = table[r]
        if r > rank:                      # peer is on the right: it needs [lo_r - halo_r, lo_r)
            a, b = max(lo_s, lo_r - halo_r), min(hi_s, lo_r)
        else:                             # peer is on the left: it needs [hi_r, hi_r + halo_r)
            a, b = max(lo_s, hi_r), min(hi_s, hi_r + halo_r)
        if a < b:
            plan.append((r, a, b))
    return plan


def exchange_halo(own, axis, table, tile=None):
    """Collective 3.  own: (n, 3|4) tensor of this rank's slab (rows ordered as stored).  Returns (tile, n_halo, stats):
    tile = own rows followed by the halo rows received from the other ranks in rank order, (n + n_halo, 3)."""
    rank, world = _world()
    n = int(own.shape[0])
    dev = own.device
    plan = halo_plan(table, rank)
    send = torch.empty((max(n * max(len(plan), 1), 1), 3), dtype=torch.float32, device=dev)      # a thin slab may go whole to several peers
    counts = torch.zeros(world, dtype=torch.int64)
    segs, off = {}, 0
    for peer, a, b in plan:
        c = select_range(own, axis, a, b, out=send[off:])
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

    def run(self, own, axis, lo, hi, halo=None):
        """own: CUDA tensor (n, 3|4) of the slab; [lo, hi) its interval along `axis`.  Returns a dict with the owned vertices
        (CUDA tensor), the per-rank counts, all vertices (CUDA tensor) and the seam graph."""
        import time
        ev = lambda: torch.cuda.Event(enable_timing=True)
        e = [ev() for _ in range(5)]
        t0 = time.perf_counter()
        e[0].record()
        if halo is None:
            halo = 10.0 * self.own_leaf(own)
        table = gather_table(lo, hi, halo, own.device)
        tile, n_halo, st = exchange_halo(own, axis, table, self.tile_buf)
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
            rc = self.L.osep_points_select_range(self.device, vp, V, 4, int(axis), float(max(lo, -3.0e38)), float(min(hi, 3.0e38)),
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


def tiled_skeleton(own, bounds_lo_hi, axis, halo, skeletonize):
    """Host-logic form of TileRunner.run for CPU tensors / numpy (gloo tests): `skeletonize` is a callable
    (m, 3) float32 numpy -> (V, 3) float32 numpy.  Same exchange, ownership and gather code as the device path."""
    lo, hi = bounds_lo_hi
    t = torch.from_numpy(np.ascontiguousarray(own, dtype=np.float32))
    table = gather_table(lo, hi, halo, t.device)
    tile, n_halo, st = exchange_halo(t, axis, table)
    verts = np.asarray(skeletonize(tile.numpy()), np.float32).reshape(-1, 3)
    mine = torch.from_numpy(np.ascontiguousarray(owned_filter(verts, axis, lo, hi)))
    counts, allv = gather_vertices(mine)
    allv = allv.numpy(); offs = np.concatenate([[0], np.cumsum(counts)])
    return {"tile": tile.numpy(), "halo_points": n_halo, "local": mine.numpy(), "counts": counts,
            "vertices": [allv[offs[r]:offs[r + 1]] for r in range(len(counts))], "exchange": st}
