// oracle/oracle_rosa.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's local-skeleton path,
//   /root/reference/osep_skeleton_decomp/src/rosa.cpp (whole file) and
//   include/rosa.hpp:99-128,
// plus the PCL 1.12.1 / FLANN 1.9.1 / Eigen 3.4.0 semantics behind its calls
// (those libraries are NOT vendored under /root/reference
// and the reference has no tests or golden vectors => PARITY UNPINNED).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference leg may use this code; the product (libosep_skel.so) never does.
//
// Two modes (RosaOracleOpts):
//   canonical (default) -- every order the reference leaves unspecified, or that a
//     parallel machine cannot follow, is fixed to a documented canonical one
//     (DESIGN.md "canonical choices" C1..C6).  The CUDA path is bit-compared to this.
//   faithful            -- reference traversal orders (DFS visit order for the
//     active-set sums), PCL's trigonometric eigen33 for normals and libm cbrtf.
//     Used as the CPU baseline and to measure how far canonical drifts.
#pragma once
#include "oracle_math.hpp"
#include "oracle_kdtree.hpp"
#include <vector>
#include <cstdio>

namespace orc {

struct RosaConfig {          // rosa.hpp:20-33, defaults rosa_node.cpp:50-60
    int max_points = 1000;
    int min_points = 100;
    float pts_dist_lim = 50.0f;
    int ne_knn = 20;
    int nb_knn = 10;
    float max_proj_range = 10.0f;
    int niter_drosa = 6;
    int niter_dcrosa = 5;
    int niter_smooth = 3;
    float alpha_recenter = 0.3f;
    float radius_smooth = 5.0f;
};

struct RosaOracleOpts {
    int faithful_normals = 0;   // 1: pcl::eigen33 (libm trig) instead of the trig-free solver
    int faithful_order = 0;     // 1: DFS-order sequential sums over active sets
    int libm_cbrt = 0;          // 1: std::cbrt instead of canon_cbrtf
    int stop_after_stage = 7;   // run stages 1..k only (taps of later stages stay empty)
};

enum : unsigned {
    FLAG_LEAF_NOT_CONVERGED = 1u,    // 8 VoxelGrid passes exhausted (rosa.cpp:125-146 quirk)
    FLAG_POOR_FIXUP_OOB = 2u,        // goodCenters[gid] out of bounds: reference UB (rosa.cpp:374)
    FLAG_VOXEL_OVERFLOW = 4u,        // PCL int32 guard hit: VoxelGrid output = input
    FLAG_NONFINITE_PSET = 8u,        // vertex_sampling compacted non-finite rows (rosa.cpp:439-448 index bug)
};

struct RosaOracle {
    RosaConfig cfg;
    RosaOracleOpts opt;

    // ---- state mirroring Rosa members (rosa.hpp:74-95)
    std::vector<V3> orig;                    // RD.orig_ (filtered in place)
    std::vector<V3> pts, nrms;               // RD.pts_, RD.nrms_
    std::vector<std::vector<int>> surf_nbs, simi_nbs;
    std::vector<V3> pset, vset;
    std::vector<float> vvar;
    std::vector<V3> skelver;                 // RD.skelver rows
    float leaf_size = 0.f;
    size_t pcd_size = 0;
    // ---- taps
    std::vector<int> filt_idx;               // original index of every kept point
    std::vector<V3> normals_full;            // per filtered point
    std::vector<int> knn_full;               // N' x k_eff
    int knn_k = 0;
    std::vector<float> leaf_hist; std::vector<int> nvox_hist;
    float leaf_used = 0.f;                   // leaf of the VoxelGrid pass that produced pts
    std::vector<unsigned> vox_keys;          // key of every output voxel (ascending)
    std::vector<int> vox_count;
    std::vector<V3> vset_iters;              // vset after each drosa iteration (niter x M)
    std::vector<V3> pset_drosa;              // pset after drosa
    std::vector<int> poor_idx;
    std::vector<int> active_size_last;       // |active| per point in the position step
    std::vector<int> corresp;
    std::vector<int> seeds;                  // seed point index of every vertex
    std::vector<V3> skelver_sampled;         // after vertex_sampling, before vertex_smooth
    unsigned flags = 0;
    int failed_stage = 0;                    // 0 = ok, else 1..7 (RUN_STEP order, rosa.cpp:39-46)
    double stage_ms[8] = {0};

    bool run(const float* xyz, int n, int stride_floats);

    bool preprocess(); bool rosa_init(); bool similarity_neighbor_extraction();
    bool drosa(); bool dcrosa(); bool vertex_sampling(); bool vertex_smooth();

    // helpers (rosa.cpp:625-841)
    float similarity_metric(const V3& p1, const V3& v1, const V3& p2, const V3& v2, float range_r, float scale = 5.0f) const;
    void compute_active_samples(int seed, const V3& p, const V3& n, std::vector<int>& out);
    V3 compute_symmetrynormal(const std::vector<int>& idxs);
    float symmnormal_variance(const V3& symm, const std::vector<int>& idxs);
    V3 cov_eigs_from_normals(const std::vector<int>& idxs);
    V3 closest_projection_point(const std::vector<int>& idxs);
    bool local_line_fit(const std::vector<int>& nb, V3& mean_out, V3& dir_out, float& conf_out, int min_nb);

    // voxel grid (PCL VoxelGrid<PointNormal>::applyFilter); returns voxel count
    int voxel_grid(float leaf, bool emit, std::vector<V3>& out_p, std::vector<V3>& out_n);

private:
    KdTree kd_;
    std::vector<char> state_;
    std::vector<int> stack_, order_;
    template <int K, class F> void reduce_active(const std::vector<int>& idxs, F contrib, float out[K]);
};

}  // namespace orc
