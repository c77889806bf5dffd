"""Host-side mirror of the reference's `Rosa` class (osep_skeleton_decomp/include/rosa.hpp:48-53)
over the C ABI.  Same members, same meaning:

    Rosa(cfg)            explicit Rosa(const RosaConfig&)
    input_cloud()        mutable host buffer the caller fills each tick (rosa_node.cpp:129)
    rosa_run()           bool; prints the reference's stdout lines (rosa.cpp:38-52, RUN_STEP)
    output_vertices()    V x 3 float32, column-major storage like Eigen::MatrixXf

The compute is entirely in libosep_skel.so (CUDA, sm_100a).  There is no CPU path here.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import RosaConfig, RosaResult, GraphView, OsepError, ROSA_STAGES

_INT_TAPS = {"filt_idx", "knn_full", "nvox_hist", "vox_count", "surf_off", "surf_idx", "simi_off", "simi_idx",
             "poor_idx", "active_size", "corresp", "seeds", "levels", "launches", "knn_stats", "graph_replays"}
_U32_TAPS = {"vox_keys"}
_V3_TAPS = {"filtered", "normals_full", "pts", "nrms", "vset", "vset_iters", "pset", "pset_drosa", "skelver_sampled", "skelver"}


def _view_arrays(v):
    def arr(p, n, dt):
        return np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True) if n > 0 and p else np.zeros(0, dt)
    G, E = v.n_vertices, v.n_edges
    out = {
        "n_vertices": G, "n_edges": E,
        "edges": arr(v.edges, 2 * E, np.int32).reshape(-1, 2), "edge_w": arr(v.edge_w, E, np.float32),
        "adj_off": arr(v.adj_off, G + 1, np.int32), "type": arr(v.type, G, np.int32),
        "joints": arr(v.joints, v.n_joints, np.int32), "leafs": arr(v.leafs, v.n_leafs, np.int32),
    }
    out["adj_idx"] = arr(v.adj_idx, int(out["adj_off"][-1]) if G > 0 else 0, np.int32)
    if v.rng_off:
        out["rng_off"] = arr(v.rng_off, G + 1, np.int32)
        out["rng_idx"] = arr(v.rng_idx, int(out["rng_off"][-1]) if G > 0 else 0, np.int32)
    return out


class Rosa:
    def __init__(self, cfg=None, device=0, capacity_points=200000, verbose=False):
        self._L = _lib.load()
        self.cfg = cfg or RosaConfig()
        self._h = C.c_void_p()
        self.verbose = verbose
        rc = self._L.osep_rosa_create(C.byref(self.cfg), device, capacity_points, C.byref(self._h))
        if rc != 0:
            raise OsepError("osep_rosa_create failed (%d): %s" % (rc, _lib.last_error()))
        self.capacity = capacity_points
        self._cloud = np.zeros((0, 4), dtype=np.float32)
        self._vertices = np.zeros((0, 3), dtype=np.float32, order="F")
        self.result = RosaResult()
        self.running = True

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.osep_rosa_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- reference surface -------------------------------------------------------------
    def input_cloud(self, n=None):
        """Mutable (n, 4) float32 host buffer in pcl::PointXYZ layout {x, y, z, 1}."""
        if n is not None and n != self._cloud.shape[0]:
            self._cloud = np.ones((n, 4), dtype=np.float32)
        return self._cloud

    def set_input(self, xyz):
        xyz = np.asarray(xyz, dtype=np.float32)
        buf = self.input_cloud(xyz.shape[0])
        buf[:, :xyz.shape[1]] = xyz
        return buf

    def rosa_run(self):
        cloud = np.ascontiguousarray(self._cloud, dtype=np.float32)
        n = cloud.shape[0]
        if self.verbose:
            print("PointCloud size before downsampling: %d" % n)
        rc = self._L.osep_rosa_run(self._h, cloud.ctypes.data_as(C.c_void_p), n, cloud.shape[1] if n else 4, C.byref(self.result))
        if rc < 0:
            raise OsepError("osep_rosa_run failed (%d): %s" % (rc, _lib.last_error()))
        self._collect_vertices()
        self.running = rc == 0
        if self.verbose:
            self._print_steps(rc)
        return self.running

    def rosa_run_pointcloud2(self, data, n_points, point_step, off_x=0):
        """rosa_run() straight from a sensor_msgs/PointCloud2 data block (bytes / uint8 array of n_points records of
        point_step bytes; float32 x, y, z at byte offsets off_x, off_x + 4, off_x + 8) -- replaces
        pcl::fromROSMsg + rosa_run (rosa_node.cpp:129,136) without the host-side repack."""
        buf = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8).reshape(-1))
        rc = self._L.osep_rosa_run_pointcloud2(self._h, buf.ctypes.data_as(C.c_void_p), n_points, point_step, off_x, off_x + 4, off_x + 8, C.byref(self.result))
        if rc < 0:
            raise OsepError("osep_rosa_run_pointcloud2 failed (%d): %s" % (rc, _lib.last_error()))
        self._collect_vertices()
        self.running = rc == 0
        if self.verbose:
            self._print_steps(rc)
        return self.running

    def output_vertices(self):
        return self._vertices

    def output_vertices_now(self):
        """vertices of the last frame when the C entry points were called directly (bench): re-reads the library's buffer"""
        self._collect_vertices()
        return self._vertices

    # --- device-resident entry points (bench / pipelines) -----------------------------------
    def run_device(self, dev_ptr, n, stride_floats=4):
        rc = self._L.osep_rosa_run_device(self._h, C.c_void_p(dev_ptr), n, stride_floats)
        if rc < 0:
            raise OsepError("osep_rosa_run_device failed (%d): %s" % (rc, _lib.last_error()))

    def finish(self):
        rc = self._L.osep_rosa_finish(self._h, C.byref(self.result))
        if rc < 0:
            raise OsepError("osep_rosa_finish failed (%d): %s" % (rc, _lib.last_error()))
        self._collect_vertices()
        self.running = rc == 0
        return self.running

    @property
    def stream(self):
        return self._L.osep_rosa_stream(self._h)

    def set_debug(self, on=True):
        rc = self._L.osep_rosa_set_debug(self._h, 1 if on else 0)
        if rc != 0:
            raise OsepError("osep_rosa_set_debug failed: %s" % _lib.last_error())

    STAGE_NAMES = ["filter", "leaf_loop", "knn_grid", "knn_normals", "voxel_final", "similarity", "surf_knn_levels", "-",
                   "drosa_orient_smooth", "position_fixup", "dcrosa", "vertex_sampling", "vertex_smooth"]

    def set_profile(self, on=True):
        rc = self._L.osep_rosa_set_profile(self._h, 1 if on else 0)
        if rc != 0:
            raise OsepError("osep_rosa_set_profile failed: %s" % _lib.last_error())

    def stage_ms(self):
        return dict(zip(self.STAGE_NAMES, self.tap("stage_ms").tolist()))

    @property
    def launches(self):
        return int(self.tap("launches")[0])

    def set_lanes(self, lanes):
        rc = self._L.osep_rosa_set_lanes(self._h, lanes)
        if rc != 0:
            raise OsepError("osep_rosa_set_lanes failed (%d): %s" % (rc, _lib.last_error()))

    def run_batch(self, frames, device_ptrs=False, counts=None, stride_floats=4):
        """frames: list of (n_i, 3|4) float32 host arrays, or device pointers with counts."""
        B = len(frames)
        ptrs = (C.c_void_p * B)()
        ns = (C.c_int * B)()
        keep = []
        if device_ptrs:
            for b in range(B):
                ptrs[b] = frames[b]; ns[b] = counts[b]
            stride = stride_floats
        else:
            stride = frames[0].shape[1] if B else 4
            for b in range(B):
                a = np.ascontiguousarray(frames[b], dtype=np.float32)
                assert a.shape[1] == stride
                keep.append(a); ptrs[b] = a.ctypes.data; ns[b] = a.shape[0]
        res = (RosaResult * B)()
        fn = self._L.osep_rosa_run_batch_device if device_ptrs else self._L.osep_rosa_run_batch
        rc = fn(self._h, ptrs, ns, B, stride, res)
        if rc < 0:
            raise OsepError("osep_rosa_run_batch failed (%d): %s" % (rc, _lib.last_error()))
        out = []
        for b in range(B):
            p = C.POINTER(C.c_float)(); nv = C.c_int()
            self._L.osep_rosa_batch_vertices(self._h, b, C.byref(p), C.byref(nv))
            V = nv.value
            v = np.ctypeslib.as_array(p, shape=(3 * V,)).copy().reshape(3, V).T if V else np.zeros((0, 3), np.float32)
            out.append(v)
        return list(res), out

    def set_graph(self, enable=True, fuse_dist_th=3.0):
        """build the skeleton graph inside every frame (osep_rosa_set_graph); graph(fuse_dist_th) then costs no device work"""
        rc = self._L.osep_rosa_set_graph(self._h, 1 if enable else 0, fuse_dist_th)
        if rc != 0:
            raise OsepError("osep_rosa_set_graph failed (%d): %s" % (rc, _lib.last_error()))

    def graph(self, fuse_dist_th=3.0):
        """graph_adj + mst + graph_decomp (gskel.cpp:201-281,539-567) of the last frame's vertices."""
        v = GraphView()
        rc = self._L.osep_rosa_graph(self._h, fuse_dist_th, C.byref(v))
        if rc != 0:
            raise OsepError("osep_rosa_graph failed (%d): %s" % (rc, _lib.last_error()))
        return _view_arrays(v)

    def tap(self, name):
        n = self._L.osep_rosa_get(self._h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        if name in _V3_TAPS:
            out = np.empty((n // 3, 3), dtype=np.float32)
        elif name in _INT_TAPS:
            out = np.empty(n, dtype=np.int32)
        elif name in _U32_TAPS:
            out = np.empty(n, dtype=np.uint32)
        else:
            out = np.empty(n, dtype=np.float32)
        if n:
            self._L.osep_rosa_get(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes)
        return out

    def csr(self, name):
        return self.tap(name + "_off"), self.tap(name + "_idx")

    TIMELINE_NAMES = ["filter", "leaf_pass0", "knn_grid", "knn_scan", "normals", "voxel_final", "voxel_reduce_positions", "surf_knn", "schedule", "schedule_end",
                      "voxel_reduce_normals", "similarity", "drosa", "position_fixup", "dcrosa_vertex_sampling", "tail", "end"]

    def timeline(self):
        """Kernel entry times of the last frame in microseconds after the filter's scan kernel (globaltimer stamps written by
        block 0 of selected kernels in a normal, overlapped run; csrc/osep_common.cuh FrameStamp)."""
        buf = np.zeros(18, dtype=np.uint64)
        self._L.osep_rosa_get(self._h, b"timeline", buf.ctypes.data_as(C.c_void_p), buf.nbytes)
        t = buf.astype(np.int64)
        return {n: float(t[i] - t[0]) / 1e3 for i, n in enumerate(self.TIMELINE_NAMES) if t[i] > 0}

    # --- helpers ----------------------------------------------------------------------------
    def _collect_vertices(self):
        p = C.POINTER(C.c_float)(); nv = C.c_int()
        self._L.osep_rosa_vertices(self._h, C.byref(p), C.byref(nv))
        V = nv.value
        if V:
            flat = np.ctypeslib.as_array(p, shape=(3 * V,)).copy()
            self._vertices = np.asfortranarray(flat.reshape(3, V).T)     # column-major V x 3
        else:
            self._vertices = np.zeros((0, 3), dtype=np.float32, order="F")

    def _print_steps(self, rc):
        for i, name in enumerate(ROSA_STAGES, 1):
            if rc == 0 or i < rc:
                print("[SUCCESS] " + name)
                if i == 1:
                    print("PointCloud size after downsampling: %d" % self.result.n_downsampled)
            else:
                print("[FAILED] " + name)
                return
        print("Number of vertices: %d" % self.result.n_vertices)
        print("[ROSA] Time Elapsed: %d ms" % int(self.result.gpu_ms))
