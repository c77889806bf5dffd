"""ctypes binding of libosep_skel.so (include/osep_skel.h).  No fallback: if the CUDA library
is missing or no B200-class device is usable, every entry point raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OSEP_SKEL_LIB") or os.path.join(_HERE, "libosep_skel.so")     # OSEP_SKEL_LIB: an instrumented build of the same library (make stress)

OSEP_OK = 0
ERR_CUDA, ERR_ARG, ERR_CAPACITY = -1, -2, -3
FLAG_LEAF_NOT_CONVERGED, FLAG_POOR_FIXUP_OOB, FLAG_VOXEL_OVERFLOW, FLAG_NONFINITE_PSET = 1, 2, 4, 8
ROSA_STAGES = ["preprocess", "rosa_init", "similarity_neighbor_extraction", "drosa", "dcrosa", "vertex_sampling", "vertex_smooth"]
GSKEL_STAGES = ["increment_skeleton", "graph_adj", "mst", "vertex_merge", "prune", "smooth_vertex_positions", "extract_branches"]


class RosaConfig(C.Structure):
    """RosaConfig (rosa.hpp:20-33); defaults = ROS parameter defaults (rosa_node.cpp:50-60)."""
    _fields_ = [("max_points", C.c_int32), ("min_points", C.c_int32), ("pts_dist_lim", C.c_float),
                ("ne_knn", C.c_int32), ("nb_knn", C.c_int32), ("max_proj_range", C.c_float),
                ("niter_drosa", C.c_int32), ("niter_dcrosa", C.c_int32), ("niter_smooth", C.c_int32),
                ("alpha_recenter", C.c_float), ("radius_smooth", C.c_float)]

    def __init__(self, **kw):
        super().__init__(1000, 100, 50.0, 20, 10, 10.0, 6, 5, 3, 0.3, 5.0)
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError("unknown RosaConfig field %r" % k)
            setattr(self, k, v)


class GSkelConfig(C.Structure):
    """GSkelConfig (gskel.hpp:23-33); defaults gskel_node.cpp:62-70."""
    _fields_ = [("gnd_th", C.c_float), ("fuse_dist_th", C.c_float), ("fuse_conf_th", C.c_float),
                ("lkf_pn", C.c_float), ("lkf_mn", C.c_float), ("max_obs_wo_conf", C.c_int32),
                ("niter_smooth_vertex", C.c_int32), ("vertex_smooth_coef", C.c_float), ("min_branch_length", C.c_int32)]

    def __init__(self, **kw):
        super().__init__(60.0, 3.0, 0.1, 1e-4, 0.1, 5, 3, 0.3, 5)
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError("unknown GSkelConfig field %r" % k)
            setattr(self, k, v)


class RosaResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("flags", C.c_uint32), ("n_input", C.c_int32), ("n_filtered", C.c_int32),
                ("n_downsampled", C.c_int32), ("n_vertices", C.c_int32), ("n_passes", C.c_int32), ("n_simi", C.c_int32),
                ("leaf_size", C.c_float), ("leaf_used", C.c_float), ("gpu_ms", C.c_float)]


class GraphView(C.Structure):
    _fields_ = [("n_vertices", C.c_int32), ("n_edges", C.c_int32), ("edges", C.POINTER(C.c_int32)), ("edge_w", C.POINTER(C.c_float)),
                ("adj_off", C.POINTER(C.c_int32)), ("adj_idx", C.POINTER(C.c_int32)), ("type", C.POINTER(C.c_int32)),
                ("joints", C.POINTER(C.c_int32)), ("n_joints", C.c_int32), ("leafs", C.POINTER(C.c_int32)), ("n_leafs", C.c_int32),
                ("rng_off", C.POINTER(C.c_int32)), ("rng_idx", C.POINTER(C.c_int32))]


# every symbol include/osep_skel.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "osep_last_error", "osep_abi_version", "osep_device_count", "osep_rosa_default_config", "osep_gskel_default_config",
    "osep_rosa_create", "osep_rosa_destroy", "osep_rosa_run", "osep_rosa_run_pointcloud2", "osep_rosa_run_device", "osep_rosa_finish", "osep_rosa_stream",
    "osep_rosa_vertices", "osep_rosa_get", "osep_rosa_set_debug", "osep_rosa_set_profile", "osep_rosa_run_batch", "osep_rosa_run_batch_device",
    "osep_rosa_batch_vertices", "osep_rosa_set_lanes", "osep_rosa_graph", "osep_rosa_set_graph",
    "osep_gskel_create", "osep_gskel_destroy", "osep_gskel_run", "osep_gskel_run_tf", "osep_gskel_vertices", "osep_gskel_graph", "osep_gskel_get",
    "osep_graph_build", "osep_graph_build_device",
    "osep_points_select_range", "osep_points_select_box", "osep_rosa_vertices_device", "osep_rosa_leaf_device",
    "osep_tf_points_device", "osep_gskel_run_tf_device",
    "osep_map_create", "osep_map_destroy", "osep_map_insert_device", "osep_map_insert", "osep_map_points_device", "osep_map_clear",
]

_LIB = None


class OsepError(RuntimeError):
    pass


def load():
    """Load libosep_skel.so; raises OsepError when it has not been built (no CPU fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise OsepError("libosep_skel.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(make -C osep_path_planner_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, ip, fp = C.c_void_p, C.c_int, C.c_float
    L.osep_last_error.restype = C.c_char_p
    L.osep_abi_version.restype = ip
    L.osep_device_count.restype = ip
    L.osep_rosa_default_config.argtypes = [C.POINTER(RosaConfig)]
    L.osep_gskel_default_config.argtypes = [C.POINTER(GSkelConfig)]
    L.osep_rosa_create.argtypes = [C.POINTER(RosaConfig), ip, ip, C.POINTER(vp)]
    L.osep_rosa_destroy.argtypes = [vp]
    L.osep_rosa_run.argtypes = [vp, vp, ip, ip, C.POINTER(RosaResult)]
    L.osep_rosa_run_device.argtypes = [vp, vp, ip, ip]
    L.osep_rosa_run_pointcloud2.argtypes = [vp, vp, ip, ip, ip, ip, ip, C.POINTER(RosaResult)]
    L.osep_rosa_finish.argtypes = [vp, C.POINTER(RosaResult)]
    L.osep_rosa_stream.argtypes = [vp]; L.osep_rosa_stream.restype = vp
    L.osep_rosa_vertices.argtypes = [vp, C.POINTER(C.POINTER(C.c_float)), C.POINTER(ip)]
    L.osep_rosa_get.argtypes = [vp, C.c_char_p, vp, C.c_long]; L.osep_rosa_get.restype = C.c_long
    L.osep_rosa_set_debug.argtypes = [vp, ip]
    L.osep_rosa_set_profile.argtypes = [vp, ip]
    L.osep_rosa_run_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(ip), ip, ip, C.POINTER(RosaResult)]
    L.osep_rosa_run_batch_device.argtypes = [vp, C.POINTER(vp), C.POINTER(ip), ip, ip, C.POINTER(RosaResult)]
    L.osep_rosa_batch_vertices.argtypes = [vp, ip, C.POINTER(C.POINTER(C.c_float)), C.POINTER(ip)]
    L.osep_rosa_set_lanes.argtypes = [vp, ip]
    L.osep_rosa_graph.argtypes = [vp, fp, C.POINTER(GraphView)]
    L.osep_rosa_set_graph.argtypes = [vp, ip, fp]
    L.osep_gskel_create.argtypes = [C.POINTER(GSkelConfig), ip, ip, C.POINTER(vp)]
    L.osep_gskel_destroy.argtypes = [vp]
    L.osep_gskel_run.argtypes = [vp, vp, ip, ip]
    L.osep_gskel_run_tf.argtypes = [vp, vp, ip, ip, vp, vp]
    L.osep_gskel_vertices.argtypes = [vp, C.POINTER(C.POINTER(C.c_float)), C.POINTER(ip)]
    L.osep_gskel_graph.argtypes = [vp, C.POINTER(GraphView)]
    L.osep_gskel_get.argtypes = [vp, C.c_char_p, vp, C.c_long]; L.osep_gskel_get.restype = C.c_long
    L.osep_graph_build.argtypes = [vp, vp, ip, ip, fp, C.POINTER(GraphView)]
    L.osep_graph_build_device.argtypes = [vp, vp, ip, ip, fp, C.POINTER(GraphView)]
    L.osep_points_select_range.argtypes = [ip, vp, ip, ip, ip, fp, fp, vp, ip, ip, C.POINTER(ip), vp]
    L.osep_points_select_box.argtypes = [ip, vp, ip, ip, vp, vp, vp, ip, ip, C.POINTER(ip), vp]
    L.osep_rosa_vertices_device.argtypes = [vp, C.POINTER(vp), C.POINTER(ip)]
    L.osep_rosa_leaf_device.argtypes = [vp, vp, ip, ip, C.POINTER(fp), C.POINTER(ip)]
    L.osep_tf_points_device.argtypes = [ip, vp, ip, ip, vp, vp, vp, ip, vp]
    L.osep_gskel_run_tf_device.argtypes = [vp, vp, vp, vp]
    L.osep_map_create.argtypes = [ip, ip, ip, fp, C.POINTER(vp)]
    L.osep_map_destroy.argtypes = [vp]
    L.osep_map_insert_device.argtypes = [vp, vp, ip, ip, vp, vp, C.POINTER(ip)]
    L.osep_map_insert.argtypes = [vp, vp, ip, ip, vp, vp, C.POINTER(ip)]
    L.osep_map_points_device.argtypes = [vp, C.POINTER(vp), C.POINTER(ip)]
    L.osep_map_clear.argtypes = [vp]
    _LIB = L
    return L


def last_error():
    return (load().osep_last_error() or b"").decode()
