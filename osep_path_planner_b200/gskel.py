"""Host-side mirror of the reference's `GSkel` class (include/gskel.hpp:106-112) over the C ABI:
GSkel(cfg), input_vertices(), gskel_run(), output_gskel(); plus the graph the reference keeps
internal (adjacency, joints, leafs, branches -- "TODO: Publish edges", gskel_node.cpp:5)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import GSkelConfig, GraphView, OsepError, GSKEL_STAGES
from .rosa import _view_arrays

_INT = {"adj_off", "adj_idx", "branch_off", "branch_idx", "joints", "leafs", "new_idx", "type", "obs", "prelim_frozen"}


class GSkel:
    def __init__(self, cfg=None, device=0, capacity_vertices=8192, verbose=False):
        self._L = _lib.load()
        self.cfg = cfg or GSkelConfig()
        self._h = C.c_void_p()
        self.verbose = verbose
        rc = self._L.osep_gskel_create(C.byref(self.cfg), device, capacity_vertices, C.byref(self._h))
        if rc != 0:
            raise OsepError("osep_gskel_create failed (%d): %s" % (rc, _lib.last_error()))
        self._cands = np.zeros((0, 4), dtype=np.float32)
        self.failed_stage = 0
        self.running = True

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.osep_gskel_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def input_vertices(self, n=None):
        if n is not None and n != self._cands.shape[0]:
            self._cands = np.ones((n, 4), dtype=np.float32)
        return self._cands

    def set_input(self, xyz):
        xyz = np.asarray(xyz, dtype=np.float32)
        buf = self.input_vertices(xyz.shape[0])
        buf[:, :xyz.shape[1]] = xyz
        return buf

    def gskel_run(self):
        c = np.ascontiguousarray(self._cands, dtype=np.float32)
        rc = self._L.osep_gskel_run(self._h, c.ctypes.data_as(C.c_void_p), c.shape[0], c.shape[1] if c.shape[0] else 4)
        if rc < 0:
            raise OsepError("osep_gskel_run failed (%d): %s" % (rc, _lib.last_error()))
        self.failed_stage = rc
        self.running = rc == 0
        if self.verbose:
            for i, name in enumerate(GSKEL_STAGES, 1):
                if rc == 0 or i < rc:
                    print("[SUCCESS] " + name)
                else:
                    print("[FAILED] " + name)
                    break
        return self.running

    def gskel_run_tf(self, translation, rotation_xyzw):
        """tf2::doTransform + fromROSMsg + gskel_run (gskel_node.cpp:137-142): the candidates set with set_input() are in
        the sensor frame; translation / rotation_xyzw are the TransformStamped fields (float64)."""
        c = np.ascontiguousarray(self._cands, dtype=np.float32)
        t = np.ascontiguousarray(translation, dtype=np.float64); q = np.ascontiguousarray(rotation_xyzw, dtype=np.float64)
        rc = self._L.osep_gskel_run_tf(self._h, c.ctypes.data_as(C.c_void_p), c.shape[0], c.shape[1] if c.shape[0] else 4,
                                       t.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p))
        if rc < 0:
            raise OsepError("osep_gskel_run_tf failed (%d): %s" % (rc, _lib.last_error()))
        self.failed_stage = rc
        self.running = rc == 0
        return self.running

    def gskel_run_tf_device(self, rosa, translation, rotation_xyzw):
        """gskel_run_tf with the candidates taken from `rosa`'s DEVICE vertices (the frame that just ran): the transform
        runs on the device (osep_gskel_run_tf_device)."""
        t = np.ascontiguousarray(translation, dtype=np.float64); q = np.ascontiguousarray(rotation_xyzw, dtype=np.float64)
        rc = self._L.osep_gskel_run_tf_device(self._h, rosa._h, t.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p))
        if rc < 0:
            raise OsepError("osep_gskel_run_tf_device failed (%d): %s" % (rc, _lib.last_error()))
        self.failed_stage = rc
        self.running = rc == 0
        return self.running

    def output_gskel(self):
        p = C.POINTER(C.c_float)(); n = C.c_int()
        self._L.osep_gskel_vertices(self._h, C.byref(p), C.byref(n))
        G = n.value
        return np.ctypeslib.as_array(p, shape=(4 * G,)).copy().reshape(G, 4) if G else np.zeros((0, 4), np.float32)

    def graph(self):
        v = GraphView()
        self._L.osep_gskel_graph(self._h, C.byref(v))
        return _view_arrays(v)

    def tap(self, name):
        n = self._L.osep_gskel_get(self._h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.int32) if name in _INT else np.empty((n // 3, 3), dtype=np.float32)
        if n:
            self._L.osep_gskel_get(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes)
        return out

    def build_graph(self, xyz, fuse_dist_th=3.0):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        v = GraphView()
        rc = self._L.osep_graph_build(self._h, xyz.ctypes.data_as(C.c_void_p), xyz.shape[0], xyz.shape[1] if xyz.ndim == 2 else 3,
                                      fuse_dist_th, C.byref(v))
        if rc != 0:
            raise OsepError("osep_graph_build failed (%d): %s" % (rc, _lib.last_error()))
        return _view_arrays(v)


def build_graph(xyz, fuse_dist_th=3.0, device=0):
    g = GSkel(device=device, capacity_vertices=max(16, len(xyz)))
    try:
        return g.build_graph(xyz, fuse_dist_th)
    finally:
        g.close()
