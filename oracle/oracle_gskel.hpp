// oracle/oracle_gskel.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's global incremental skeleton graph,
//   /root/reference/osep_skeleton_decomp/src/gskel.cpp (whole file),
//   include/gskel.hpp:35-104 (Edge, UnionFind, Vertex, GSkelData) and
//   include/lkf_vertex_fuse.hpp:6-29 (VertexLKF).
// PARITY UNPINNED (no reference tests / golden vectors; FLANN and Eigen restated).
// Canonical choices: C1 neighbour order (d2, index); C6 Kruskal edge order
// (w, u, v) -- the reference's std::sort on w alone leaves ties unordered
// (gskel.hpp:38-40, gskel.cpp:266).
#pragma once
#include "oracle_math.hpp"
#include "oracle_kdtree.hpp"
#include <vector>
#include <set>

namespace orc {

struct GSkelConfig {          // gskel.hpp:23-33, defaults gskel_node.cpp:62-70
    float gnd_th = 60.0f;
    float fuse_dist_th = 3.0f;
    float fuse_conf_th = 0.1f;
    float lkf_pn = 1e-4f;
    float lkf_mn = 0.1f;
    int max_obs_wo_conf = 5;
    int niter_smooth_vertex = 3;
    float vertex_smooth_coef = 0.3f;
    int min_branch_length = 5;
};

struct GVertex {              // gskel.hpp:70-86 (+ VertexLKF state)
    V3 position{0, 0, 0};
    V3 kx{0, 0, 0};
    float kP[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int obs_count = 0, unconf_check = 0, type = 0, smooth_iters = 0;
    bool just_approved = false, frozen = false, conf_check = false, marked_for_deletion = false, updated = false;
};

struct GEdge { int u, v; float w; };

// Stateless part (G3-G5): graph_adj + mst + graph_decomp over a vertex set.
struct GraphResult {
    std::vector<std::vector<int>> adj_rng;   // after graph_adj (duplicates kept, push order)
    std::vector<GEdge> mst_edges;            // accepted edges in Kruskal order
    std::vector<std::vector<int>> adj;       // MST adjacency (push order)
    std::vector<int> joints, leafs, type;
};
void graph_adj(const std::vector<V3>& pos, float fuse_dist_th, std::vector<std::vector<int>>& adj);
void mst_from_adj(const std::vector<V3>& pos, const std::vector<std::vector<int>>& adj_in, GraphResult& out);
void graph_decomp(const std::vector<std::vector<int>>& adj, std::vector<int>& joints, std::vector<int>& leafs, std::vector<int>& type);
GraphResult build_graph(const std::vector<V3>& pos, float fuse_dist_th);

struct GSkelOracle {
    GSkelConfig cfg;
    std::vector<V3> new_cands;
    std::vector<GVertex> prelim_vers, global_vers;
    std::vector<V3> global_cloud;
    std::vector<int> new_vers_indxs, joints, leafs;
    std::vector<std::vector<int>> branches, global_adj;
    size_t gskel_size = 0;
    int failed_stage = 0;      // 0 ok, else 1..7 (gskel.cpp:24-32 order)

    bool run(const float* xyz, int n, int stride);
    bool increment_skeleton(); bool graph_adj_stage(); bool mst_stage(); bool vertex_merge();
    bool prune(); bool smooth_vertex_positions(); bool extract_branches();
    void decomp(); void merge_into(int keep, int del); bool size_assert() const;
private:
    KdTree kd_;
};

}  // namespace orc
