"""ctypes loader for the CPU oracle (oracle/liborc_oracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product package (osep_path_planner_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class RosaConfig(C.Structure):
    # rosa.hpp:20-33, defaults rosa_node.cpp:50-60
    _fields_ = [("max_points", C.c_int), ("min_points", C.c_int), ("pts_dist_lim", C.c_float),
                ("ne_knn", C.c_int), ("nb_knn", C.c_int), ("max_proj_range", C.c_float),
                ("niter_drosa", C.c_int), ("niter_dcrosa", C.c_int), ("niter_smooth", C.c_int),
                ("alpha_recenter", C.c_float), ("radius_smooth", C.c_float)]

    @classmethod
    def default(cls, **kw):
        c = cls(1000, 100, 50.0, 20, 10, 10.0, 6, 5, 3, 0.3, 5.0)
        for k, v in kw.items():
            setattr(c, k, v)
        return c


class RosaOpts(C.Structure):
    _fields_ = [("faithful_normals", C.c_int), ("faithful_order", C.c_int), ("libm_cbrt", C.c_int),
                ("stop_after_stage", C.c_int)]


class GSkelConfig(C.Structure):
    # gskel.hpp:23-33, defaults gskel_node.cpp:62-70
    _fields_ = [("gnd_th", C.c_float), ("fuse_dist_th", C.c_float), ("fuse_conf_th", C.c_float),
                ("lkf_pn", C.c_float), ("lkf_mn", C.c_float), ("max_obs_wo_conf", C.c_int),
                ("niter_smooth_vertex", C.c_int), ("vertex_smooth_coef", C.c_float), ("min_branch_length", C.c_int)]

    @classmethod
    def default(cls, **kw):
        c = cls(60.0, 3.0, 0.1, 1e-4, 0.1, 5, 3, 0.3, 5)
        for k, v in kw.items():
            setattr(c, k, v)
        return c


def build(force=False):
    so = os.path.join(_HERE, "liborc_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_rosa_create.restype = C.c_void_p
        L.orc_rosa_create.argtypes = [C.POINTER(RosaConfig), C.POINTER(RosaOpts)]
        L.orc_rosa_destroy.argtypes = [C.c_void_p]
        L.orc_rosa_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_rosa_run.restype = C.c_int
        L.orc_rosa_leaf_size.argtypes = [C.c_void_p]; L.orc_rosa_leaf_size.restype = C.c_float
        L.orc_rosa_leaf_used.argtypes = [C.c_void_p]; L.orc_rosa_leaf_used.restype = C.c_float
        L.orc_rosa_flags.argtypes = [C.c_void_p]; L.orc_rosa_flags.restype = C.c_uint
        L.orc_rosa_stage_ms.argtypes = [C.c_void_p, C.c_int]; L.orc_rosa_stage_ms.restype = C.c_double
        L.orc_rosa_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long]; L.orc_rosa_get.restype = C.c_long
        L.orc_graph_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float]; L.orc_graph_build.restype = C.c_void_p
        L.orc_graph_destroy.argtypes = [C.c_void_p]
        L.orc_graph_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long]; L.orc_graph_get.restype = C.c_long
        L.orc_gskel_create.argtypes = [C.POINTER(GSkelConfig)]; L.orc_gskel_create.restype = C.c_void_p
        L.orc_gskel_destroy.argtypes = [C.c_void_p]
        L.orc_gskel_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]; L.orc_gskel_run.restype = C.c_int
        L.orc_gskel_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long]; L.orc_gskel_get.restype = C.c_long
        L.orc_eig3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]; L.orc_eig3.restype = C.c_int
        L.orc_cbrtf.argtypes = [C.c_float]; L.orc_cbrtf.restype = C.c_float
        L.orc_pcl_eigen33.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_llt3.argtypes = [C.c_void_p] * 3; L.orc_llt3.restype = C.c_int
        L.orc_ldlt3.argtypes = [C.c_void_p] * 3; L.orc_ldlt3.restype = C.c_int
        L.orc_knn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


_INT_TAPS = {"filt_idx", "knn_full", "nvox_hist", "vox_count", "surf_off", "surf_idx", "simi_off", "simi_idx",
             "poor_idx", "active_size", "corresp", "seeds", "rng_off", "rng_idx", "adj_off", "adj_idx", "joints",
             "leafs", "type", "mst_edges", "new_idx", "branch_off", "branch_idx", "obs", "prelim_frozen"}
_U32_TAPS = {"vox_keys"}
_V3_TAPS = {"filtered", "normals_full", "pts", "nrms", "vset", "vset_iters", "pset", "pset_drosa",
            "skelver_sampled", "skelver", "global", "prelim"}


def _fetch(getter, h, name):
    n = getter(h, name.encode(), None, 0)
    if n < 0:
        raise KeyError(name)
    if name in _V3_TAPS:
        out = np.empty((n, 3), dtype=np.float32)
    elif name in _INT_TAPS:
        out = np.empty(n, dtype=np.int32)
    elif name in _U32_TAPS:
        out = np.empty(n, dtype=np.uint32)
    else:
        out = np.empty(n, dtype=np.float32)
    if n:
        getter(h, name.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes)
    return out


class RosaOracle:
    """Mirror of the reference's Rosa class surface (rosa.hpp:48-53) over the CPU oracle."""

    def __init__(self, cfg=None, faithful_normals=0, faithful_order=0, libm_cbrt=0, stop_after_stage=7):
        self.cfg = cfg or RosaConfig.default()
        self.opts = RosaOpts(faithful_normals, faithful_order, libm_cbrt, stop_after_stage)
        self._h = lib().orc_rosa_create(C.byref(self.cfg), C.byref(self.opts))
        self.failed_stage = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_rosa_destroy(self._h)
            self._h = None

    def run(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        assert xyz.ndim == 2 and xyz.shape[1] in (3, 4)
        self.failed_stage = lib().orc_rosa_run(self._h, xyz.ctypes.data_as(C.c_void_p), xyz.shape[0], xyz.shape[1])
        return self.failed_stage == 0

    def tap(self, name):
        return _fetch(lib().orc_rosa_get, self._h, name)

    def csr(self, name):
        return self.tap(name + "_off"), self.tap(name + "_idx")

    @property
    def leaf_size(self):
        return lib().orc_rosa_leaf_size(self._h)

    @property
    def leaf_used(self):
        return lib().orc_rosa_leaf_used(self._h)

    @property
    def flags(self):
        return lib().orc_rosa_flags(self._h)

    def stage_ms(self):
        return [lib().orc_rosa_stage_ms(self._h, s) for s in range(1, 8)]


class GraphOracle:
    """Stateless graph_adj + mst + graph_decomp (gskel.cpp:201-281,539-567)."""

    def __init__(self, pos, fuse_dist_th=3.0):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        self._h = lib().orc_graph_build(pos.ctypes.data_as(C.c_void_p), pos.shape[0], pos.shape[1], fuse_dist_th)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_graph_destroy(self._h)
            self._h = None

    def tap(self, name):
        return _fetch(lib().orc_graph_get, self._h, name)


class GSkelOracle:
    """Mirror of the reference's GSkel class surface (gskel.hpp:106-112) over the CPU oracle."""

    def __init__(self, cfg=None):
        self.cfg = cfg or GSkelConfig.default()
        self._h = lib().orc_gskel_create(C.byref(self.cfg))
        self.failed_stage = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_gskel_destroy(self._h)
            self._h = None

    def run(self, cands):
        cands = np.ascontiguousarray(cands, dtype=np.float32)
        self.failed_stage = lib().orc_gskel_run(self._h, cands.ctypes.data_as(C.c_void_p), cands.shape[0], cands.shape[1])
        return self.failed_stage == 0

    def tap(self, name):
        return _fetch(lib().orc_gskel_get, self._h, name)


def tf_points(xyz, translation, rotation_xyzw):
    """tf2::doTransform(PointCloud2) (gskel_node.cpp:137-138), float32 arithmetic of tf2_sensor_msgs."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    t = np.ascontiguousarray(translation, dtype=np.float64); q = np.ascontiguousarray(rotation_xyzw, dtype=np.float64)
    out = np.zeros((xyz.shape[0], 3), np.float32)
    L = lib()
    L.orc_tf_points.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_tf_points(xyz.ctypes.data_as(C.c_void_p), xyz.shape[0], xyz.shape[1], t.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def eig3(A):
    A = np.ascontiguousarray(A, dtype=np.float32).reshape(3, 3)
    ev = np.zeros(3, np.float32); Q = np.zeros((3, 3), np.float32)
    ok = lib().orc_eig3(A.ctypes.data_as(C.c_void_p), ev.ctypes.data_as(C.c_void_p), Q.ctypes.data_as(C.c_void_p))
    return bool(ok), ev, Q


def cbrtf(x):
    return float(lib().orc_cbrtf(C.c_float(x)))


def pcl_eigen33(A):
    A = np.ascontiguousarray(A, dtype=np.float32).reshape(3, 3)
    ev = np.zeros(1, np.float32); v = np.zeros(3, np.float32)
    lib().orc_pcl_eigen33(A.ctypes.data_as(C.c_void_p), ev.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
    return float(ev[0]), v


def solve3(A, b, kind="llt"):
    A = np.ascontiguousarray(A, dtype=np.float32).reshape(3, 3)
    b = np.ascontiguousarray(b, dtype=np.float32)
    x = np.zeros(3, np.float32)
    f = lib().orc_llt3 if kind == "llt" else lib().orc_ldlt3
    ok = f(A.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p))
    return bool(ok), x


def knn(xyz, k):
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    out = np.empty((xyz.shape[0], k), dtype=np.int32)
    lib().orc_knn(xyz.ctypes.data_as(C.c_void_p), xyz.shape[0], xyz.shape[1], k, out.ctypes.data_as(C.c_void_p))
    return out
