/* include/osep_skel.h -- C ABI of libosep_skel.so, the B200 (sm_100a) implementation of
 * osep_skeleton_decomp's per-frame skeleton-extraction hot path.
 *
 * The reference has no FFI/plugin layer: its seam is the two C++ classes
 *   Rosa  (osep_skeleton_decomp/include/rosa.hpp:48-53:  ctor, rosa_run(), input_cloud(), output_vertices())
 *   GSkel (osep_skeleton_decomp/include/gskel.hpp:106-112: ctor, gskel_run(), input_vertices(), output_gskel())
 * driven by RosaNode::process_tick (src/rosa_node.cpp:120-140) and GSkelNode::process_tick
 * (src/gskel_node.cpp:114-147).  Every entry point below names the member it replaces.  The
 * C++ shim classes with the reference's member names live in osep_path_planner_b200/host/
 * (rosa.hpp, gskel.hpp) and call only this ABI; INTEGRATION.md shows the node-side binding.
 *
 * Conventions: plain pointers and sizes, POD structs, no exceptions, no PCL/Eigen/ROS/torch
 * types.  Inputs are caller owned; outputs are library owned and valid until the next call on
 * the same handle.  A handle is used from one thread at a time (the reference classes are
 * non-reentrant too: rosa.hpp:74-95 member scratch).  Different handles may run concurrently.
 * There is no CPU fallback: every call fails with OSEP_ERR_CUDA when no sm_100-class device
 * is usable.
 */
#ifndef OSEP_SKEL_H_
#define OSEP_SKEL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSEP_ABI_VERSION 1

/* ---- status codes.  1..7 = index of the failing stage, mirroring RUN_STEP (rosa.hpp:12-18,
 * gskel.hpp:15-21): the reference returns false at the first stage that fails. */
enum {
    OSEP_OK = 0,
    OSEP_STAGE_1 = 1, /* Rosa: preprocess                      | GSkel: increment_skeleton      */
    OSEP_STAGE_2 = 2, /* Rosa: rosa_init                       | GSkel: graph_adj               */
    OSEP_STAGE_3 = 3, /* Rosa: similarity_neighbor_extraction  | GSkel: mst                     */
    OSEP_STAGE_4 = 4, /* Rosa: drosa                           | GSkel: vertex_merge            */
    OSEP_STAGE_5 = 5, /* Rosa: dcrosa                          | GSkel: prune                   */
    OSEP_STAGE_6 = 6, /* Rosa: vertex_sampling                 | GSkel: smooth_vertex_positions */
    OSEP_STAGE_7 = 7, /* Rosa: vertex_smooth                   | GSkel: extract_branches        */
    OSEP_ERR_CUDA = -1,     /* CUDA runtime error / no usable device (see osep_last_error) */
    OSEP_ERR_ARG = -2,      /* bad argument */
    OSEP_ERR_CAPACITY = -3  /* a device buffer sized at create time is too small for this frame */
};

/* ---- result flags (bit set) */
enum {
    OSEP_FLAG_LEAF_NOT_CONVERGED = 1u, /* 8 VoxelGrid passes exhausted (rosa.cpp:125-146): leaf_size != leaf_used */
    OSEP_FLAG_POOR_FIXUP_OOB = 2u,     /* goodCenters[gid] out of bounds = UB in the reference (rosa.cpp:374); pset left unchanged */
    OSEP_FLAG_VOXEL_OVERFLOW = 4u,     /* PCL VoxelGrid int32 guard hit (output = input) */
    OSEP_FLAG_NONFINITE_PSET = 8u      /* vertex_sampling dropped non-finite rows (rosa.cpp:439-448) */
};

/* RosaConfig (rosa.hpp:20-33); defaults are the ROS parameter defaults (rosa_node.cpp:50-60). */
typedef struct osep_rosa_config {
    int32_t max_points;     /* rosa_max_points            1000 */
    int32_t min_points;     /* rosa_min_points             100 */
    float pts_dist_lim;     /* rosa_point_dist_lim         50  */
    int32_t ne_knn;         /* rosa_ne_knn                  20 */
    int32_t nb_knn;         /* rosa_nb_knn                  10 */
    float max_proj_range;   /* rosa_max_projection_range    10 */
    int32_t niter_drosa;    /* rosa_niter_drosa              6 */
    int32_t niter_dcrosa;   /* rosa_niter_dcrosa             5 */
    int32_t niter_smooth;   /* rosa_niter_smooth             3 */
    float alpha_recenter;   /* rosa_recenter_alpha         0.3 */
    float radius_smooth;    /* rosa_smooth_radius            5 */
} osep_rosa_config;

/* GSkelConfig (gskel.hpp:23-33); defaults gskel_node.cpp:62-70. */
typedef struct osep_gskel_config {
    float gnd_th;               /* gskel_gnd_th               60  */
    float fuse_dist_th;         /* gskel_fuse_dist_th          3  */
    float fuse_conf_th;         /* gskel_fuse_conf_th         0.1 */
    float lkf_pn;               /* gskel_lkf_pn              1e-4 */
    float lkf_mn;               /* gskel_lkf_mn               0.1 */
    int32_t max_obs_wo_conf;    /* gskel_max_obs_wo_conf       5  */
    int32_t niter_smooth_vertex;/* gskel_niter_vertex_smooth   3  */
    float vertex_smooth_coef;   /* gskel_vertex_smooth_coef   0.3 */
    int32_t min_branch_length;  /* gskel_min_branch_length     5  */
} osep_gskel_config;

void osep_rosa_default_config(osep_rosa_config* cfg);
void osep_gskel_default_config(osep_gskel_config* cfg);

/* Summary of one frame (what rosa_run prints, rosa.cpp:38,40,48, plus parity scalars). */
typedef struct osep_rosa_result {
    int32_t status;        /* OSEP_OK, failing stage 1..7, or OSEP_ERR_* */
    uint32_t flags;        /* OSEP_FLAG_* */
    int32_t n_input;       /* "PointCloud size before downsampling" */
    int32_t n_filtered;    /* N' after the distance/finite filter (rosa.cpp:67-88) */
    int32_t n_downsampled; /* M = pcd_size_ after preprocess */
    int32_t n_vertices;    /* V = skelver.rows() */
    int32_t n_passes;      /* VoxelGrid passes executed (1..8) */
    int32_t n_simi;        /* sum of similarity-neighbour list lengths */
    float leaf_size;       /* Rosa::leaf_size (rosa.cpp:148) */
    float leaf_used;       /* leaf of the pass that produced the downsampled cloud */
    float gpu_ms;          /* device time of the frame (CUDA events), 0 when not measured */
} osep_rosa_result;

typedef struct osep_rosa osep_rosa;
typedef struct osep_gskel osep_gskel;

const char* osep_last_error(void);
int osep_abi_version(void);
int osep_device_count(void);

/* ---- Rosa ---------------------------------------------------------------------------- */

/* Rosa::Rosa(const RosaConfig&) (rosa.cpp:15-34).  Allocates the device arena once:
 * capacity_points = largest frame the handle accepts. */
int osep_rosa_create(const osep_rosa_config* cfg, int device, int capacity_points, osep_rosa** out);
int osep_rosa_destroy(osep_rosa* h);

/* input_cloud() write + rosa_run() (rosa_node.cpp:129,137): HOST cloud, `stride_floats`
 * floats per point (4 for pcl::PointXYZ's 16-byte layout, 3 for packed xyz).  Copies the
 * cloud to the device, runs the 7 stages, copies the vertices back.  Returns res->status. */
int osep_rosa_run(osep_rosa* h, const float* xyz_host, int n, int stride_floats, osep_rosa_result* res);

/* pcl::fromROSMsg(*msg, rosa_->input_cloud()) + rosa_run() (rosa_node.cpp:129,136) without the host-side repack of
 * the message: the raw sensor_msgs/PointCloud2 data block (n_points = width * height records of point_step bytes,
 * little endian) is copied to the device as it is and the filter kernel reads x, y, z in place.
 * Requires float32 fields x, y, z at the consecutive byte offsets off_x, off_x + 4, off_x + 8 -- the layout of every
 * PointXYZ* message -- and off_x % 4 == 0, point_step % 4 == 0; anything else returns OSEP_ERR_ARG.  Like
 * fromROSMsg, only x, y, z are used (the 4th float of pcl::PointXYZ is 1.0f). */
int osep_rosa_run_pointcloud2(osep_rosa* h, const void* data_host, int n_points, int point_step, int off_x, int off_y, int off_z,
                              osep_rosa_result* res);

/* Same frame, DEVICE-resident input; enqueues on the handle's stream and returns without a
 * host synchronisation.  Follow with osep_rosa_finish() to wait and read the summary. */
int osep_rosa_run_device(osep_rosa* h, const float* xyz_dev, int n, int stride_floats);
int osep_rosa_finish(osep_rosa* h, osep_rosa_result* res);
/* cudaStream_t the handle enqueues on (for event timing by the caller). */
void* osep_rosa_stream(osep_rosa* h);

/* output_vertices() (rosa.hpp:53): V x 3 float32, COLUMN-MAJOR like Eigen::MatrixXf
 * (x[V], y[V], z[V]); host pointer owned by the handle, valid until the next run. */
int osep_rosa_vertices(osep_rosa* h, const float** data, int* n_vertices);

/* Debug taps for parity tests.  Copies the named intermediate to `dst`
 * (host) when it fits in cap_bytes; returns the element count, or -1 for an unknown name.
 * float3 rows count 3 per row.  Names: "filtered" "filt_idx" "normals_full" "knn_full"
 * "leaf_hist" "nvox_hist" "vox_keys" "vox_count" "pts" "nrms" "surf_off" "surf_idx"
 * "simi_off" "simi_idx" "vset" "vset_iters" "vvar" "pset" "pset_drosa" "poor_idx"
 * "active_size" "corresp" "seeds" "skelver_sampled" "skelver" "stage_ms". */
long osep_rosa_get(osep_rosa* h, const char* name, void* dst, long cap_bytes);
/* Enable per-stage taps that cost extra copies (vset_iters, knn_full, skelver_sampled ...). */
int osep_rosa_set_debug(osep_rosa* h, int enable);
/* Record a CUDA event after every stage group; "stage_ms" then returns 13 floats: device ms of
 * filter, leaf loop, kNN grid build, kNN+normals, final voxel pass, similarity, surf kNN+levels,
 * (unused), drosa orientation loop, position+fix-up, dcrosa, vertex_sampling, vertex_smooth.
 * "launches" returns the number of kernels this handle has launched since create. */
int osep_rosa_set_profile(osep_rosa* h, int enable);

/* Batched independent frames (BASELINE.json configs[3]): B host clouds, processed on an
 * internal pool of `lanes` streams of this handle's device.  results[b] receives each
 * frame's summary; vertices of frame b are fetched with osep_rosa_batch_vertices. */
int osep_rosa_run_batch(osep_rosa* h, const float* const* xyz_host, const int* n, int n_frames,
                        int stride_floats, osep_rosa_result* results);
int osep_rosa_run_batch_device(osep_rosa* h, const float* const* xyz_dev, const int* n, int n_frames,
                               int stride_floats, osep_rosa_result* results);
int osep_rosa_batch_vertices(osep_rosa* h, int frame, const float** data, int* n_vertices);
int osep_rosa_set_lanes(osep_rosa* h, int lanes);

/* ---- stateless skeleton graph over a vertex set: graph_adj + mst + graph_decomp
 * (gskel.cpp:201-281, 539-567) on the device.  xyz: host, `stride_floats` per vertex.
 * Outputs (host, library owned): MST edge list (u,v) in Kruskal order, CSR adjacency,
 * type per vertex (1 = leaf, 0 otherwise -- the reference's fall-through, gskel.cpp:554-562),
 * joints, leafs. */
typedef struct osep_graph_view {
    int32_t n_vertices;
    int32_t n_edges;
    const int32_t* edges;     /* 2 * n_edges */
    const float* edge_w;      /* n_edges */
    const int32_t* adj_off;   /* n_vertices + 1 */
    const int32_t* adj_idx;   /* 2 * n_edges, push order of gskel.cpp:270-275 */
    const int32_t* type;      /* n_vertices */
    const int32_t* joints; int32_t n_joints;
    const int32_t* leafs; int32_t n_leafs;
    const int32_t* rng_off;   /* graph_adj adjacency before the MST (duplicates kept) */
    const int32_t* rng_idx;
} osep_graph_view;
int osep_rosa_graph(osep_rosa* h, float fuse_dist_th, osep_graph_view* out);  /* graph of the last frame's vertices */
/* enable != 0: graph_adj + mst (gskel.cpp:201-281) run INSIDE every frame, behind vertex_smooth on the frame's stream (one
 * launch sequence / CUDA graph from point cloud to skeleton graph); osep_rosa_graph(h, same fuse_dist_th, ...) then only
 * assembles the host views.  Off by default: the reference's Rosa class stops at the vertices (rosa.hpp:53). */
int osep_rosa_set_graph(osep_rosa* h, int enable, float fuse_dist_th);

/* ---- GSkel --------------------------------------------------------------------------- */
int osep_gskel_create(const osep_gskel_config* cfg, int device, int capacity_vertices, osep_gskel** out);
int osep_gskel_destroy(osep_gskel* h);
/* input_vertices() write + gskel_run() (gskel_node.cpp:139-141).  cand_xyz: host, already in
 * the global frame (the TF transform, gskel_node.cpp:124-138, stays with the caller). */
int osep_gskel_run(osep_gskel* h, const float* cand_xyz, int n, int stride_floats);
/* tf2::doTransform(*msg, gmsg, T) + fromROSMsg + gskel_run() (gskel_node.cpp:137-142) in one call:
 * cand_xyz are Rosa's vertices in the SENSOR frame; translation[3] / rotation_xyzw[4] are the float64 fields of the
 * geometry_msgs/TransformStamped returned by lookupTransform(global_frame, sensor_frame).  The transform is applied
 * with tf2_sensor_msgs' arithmetic (float32 Translation3f * Quaternionf, then linear * p + translation). */
int osep_gskel_run_tf(osep_gskel* h, const float* cand_xyz, int n, int stride_floats, const double* translation, const double* rotation_xyzw);
/* output_gskel() (gskel.hpp:112): G points, 16-byte {x,y,z,1} like pcl::PointXYZ. */
int osep_gskel_vertices(osep_gskel* h, const float** xyz4, int* n_vertices);
int osep_gskel_graph(osep_gskel* h, osep_graph_view* out);
long osep_gskel_get(osep_gskel* h, const char* name, void* dst, long cap_bytes);
/* graph of an arbitrary host vertex set on the device (stateless; used by tests and the tiled-map path) */
int osep_graph_build(osep_gskel* h, const float* xyz, int n, int stride_floats, float fuse_dist_th, osep_graph_view* out);


/* ---- spatially tiled maps (BASELINE.json configs[4]).  The reference has no tiling: these entry points are what
 * a tiled caller (osep_path_planner_b200/tiled.py; one process per GPU) needs so that the map, the halos and the vertices
 * never leave the device between the NCCL exchanges. */

/* Stable compaction on the device: the records of pts_dev (n records of stride_floats floats) whose coordinate `axis`
 * (0, 1, 2) lies in [lo, hi) are written in input order to out_dev as records of out_stride floats (3, or 4 with w = 1).
 * Used for the halo selection (points within `halo` of a slab face) and the owned-vertex mask (vertices inside the tile's own
 * interval).  Runs on `stream` (a cudaStream_t, may be NULL) of `device`; synchronises the stream and stores the number of
 * matches in *count_host.  More than out_cap matches -> OSEP_ERR_CAPACITY (the first out_cap are written, the count is exact). */
int osep_points_select_range(int device, const float* pts_dev, int n, int stride_floats, int axis, float lo, float hi,
                             float* out_dev, int out_stride, int out_cap, int* count_host, void* stream);
/* The same compaction with an axis-aligned box test on all three coordinates: lo[d] <= p[d] < hi[d] (pass -/+3e38 for an open
 * side).  Tiles of a recursive-bisection partition are boxes; their halo is the box widened by the halo width. */
int osep_points_select_box(int device, const float* pts_dev, int n, int stride_floats, const float* lo3, const float* hi3,
                           float* out_dev, int out_stride, int out_cap, int* count_host, void* stream);
/* Device copy of output_vertices() (rosa.hpp:53): V records {x, y, z, 1} of the last frame, valid until the next run. */
int osep_rosa_vertices_device(osep_rosa* h, const float** xyzw_dev, int* n_vertices);
/* filter + leaf-size control law only (rosa.cpp:67-95, 118-148) on a DEVICE cloud: the converged leaf of a tile's own points
 * sizes its halo (10 x leaf = the similarity radius, rosa.cpp:212) before the exchange.  Returns OSEP_OK / OSEP_STAGE_1. */
int osep_rosa_leaf_device(osep_rosa* h, const float* xyz_dev, int n, int stride_floats, float* leaf_size, int* n_filtered);
/* osep_graph_build over DEVICE vertices (the all-gathered vertices of the tiles: the seam graph) */
int osep_graph_build_device(osep_gskel* h, const float* xyz_dev, int n, int stride_floats, float fuse_dist_th, osep_graph_view* out);

/* ---- the stages either side of the hot path, kept on the device -------------------------------- */

/* tf2::doTransform(*msg, gmsg, T) (gskel_node.cpp:137-138) as a kernel: out = R(rotation_xyzw) * p + translation for n DEVICE
 * records, in tf2_sensor_msgs' float32 arithmetic (the same sequence as osep_gskel_run_tf).  NULL translation/rotation = identity. */
int osep_tf_points_device(int device, const float* in_dev, int n, int stride_floats, const double* translation, const double* rotation_xyzw,
                          float* out_dev, int out_stride, void* stream);
/* osep_gskel_run_tf with the candidates taken from the Rosa handle's DEVICE vertices (the frame that just ran): transform
 * on the device, one small device-to-host copy of the transformed candidates for the stateful GSkel stages -- the sensor-frame
 * vertices are never repacked or transformed on the host.  `rosa` must live on the same device as `h`. */
int osep_gskel_run_tf_device(osep_gskel* h, osep_rosa* rosa, const double* translation, const double* rotation_xyzw);

/* Accumulated map (the input of the tiled configuration, BASELINE.json configs[4]; the reference has no counterpart: its
 * RosaNode consumes one PointCloud2 per tick, rosa_node.cpp:120-140).  A voxel hash on the device: every frame is transformed
 * into the map frame (translation / rotation as in osep_tf_points_device) and inserted; a voxel of side voxel_size keeps the
 * FIRST point that ever fell into it (lowest frame, then lowest index within the frame).  The kept points form an append-only
 * dense DEVICE array in exactly that order, whatever the thread interleaving (tests/pyref/map_ref.py restates it on the CPU). */
typedef struct osep_map osep_map;
int osep_map_create(int device, int capacity_points, int max_frame_points, float voxel_size, osep_map** out);
int osep_map_destroy(osep_map* m);
int osep_map_insert_device(osep_map* m, const float* xyz_dev, int n, int stride_floats, const double* translation, const double* rotation_xyzw, int* n_new);
int osep_map_insert(osep_map* m, const float* xyz_host, int n, int stride_floats, const double* translation, const double* rotation_xyzw, int* n_new);
/* the dense point array: n_points records {x, y, z, 1}, DEVICE pointer, valid until the map is destroyed (it only grows) */
int osep_map_points_device(osep_map* m, const float** xyzw_dev, int* n_points);
int osep_map_clear(osep_map* m);

#ifdef __cplusplus
}
#endif
#endif /* OSEP_SKEL_H_ */
