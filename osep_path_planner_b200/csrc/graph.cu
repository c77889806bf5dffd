// csrc/graph.cu -- skeleton-graph construction (GSkel, gskel.cpp) for sm_100a.
//
// Device (data-parallel, O(G^2) distance work + edge build):
//   G3 graph_adj   kNN(10) + relative-neighbourhood test          (gskel.cpp:201-246)
//   G4 mst         stream-compacted edge list (u<v), sort by (w,u,v), Kruskal with the
//                  reference's UnionFind                          (gskel.cpp:248-281, gskel.hpp:35-68)
// Host (sequential, stateful, G ~ 10..10^3 -- stays with the class state like the reference):
//   G1 gskel_run, G2 increment_skeleton + VertexLKF, G5 graph_decomp, G6 vertex_merge,
//   G7 prune, G8 smooth_vertex_positions, G9 extract_branches, G10 size_assert.
// The host side keeps the exact push orders of the reference's std::vector adjacency lists,
// because vertex_merge / extract_branches iterate them in order.
#include "tail.cuh"
#include "rosa_impl.cuh"
#include <vector>
#include <set>
#include <string>
#include <algorithm>
#include <cstring>

namespace osep {

// brute-force kNN(10) by (d2, index) + RNG test restricted to the kNN list (tail.cuh: graph_adj_warp).
// G: vertex count known on the host, or (st != nullptr) read from the frame state of the Rosa frame that is still running on the
// same stream: the graph of a frame is then built inside the frame's launch sequence / CUDA graph, without a host round trip
__device__ __forceinline__ int graph_count(const FrameState* st, int G) { return st ? (st->status == 0 ? st->V : 0) : G; }

constexpr int GADJ_THREADS = 32;
__global__ void __launch_bounds__(GADJ_THREADS)
k_graph_adj(const float4* __restrict__ pos, int G_arg, const FrameState* __restrict__ st, float max_dist_th, int* __restrict__ acc, int* __restrict__ acc_cnt) {
    graph_adj_thread(pos, graph_count(st, G_arg), blockIdx.x * GADJ_THREADS + threadIdx.x, gridDim.x * GADJ_THREADS, max_dist_th, acc, acc_cnt);
}

// Single block (tail.cuh: graph_mst_block).  Keys and the component arrays live in shared memory when they fit.
constexpr int GMST_SKEYS = 4096;
constexpr int GMST_SWORK = 3072;
__global__ void __launch_bounds__(1024)
k_graph_mst(const float4* __restrict__ pos, int G_arg, const FrameState* __restrict__ st, const int* __restrict__ acc, const int* __restrict__ acc_cnt,
            unsigned long long* gkeys, int keycap, int* gwork, int* __restrict__ mst_uv, float* __restrict__ mst_w, GraphDevState* gs) {
    __shared__ unsigned long long skeys[GMST_SKEYS];
    __shared__ int swork[GMST_SWORK];
    graph_mst_block(pos, graph_count(st, G_arg), acc, acc_cnt, skeys, GMST_SKEYS, gkeys, keycap, swork, GMST_SWORK, gwork, mst_uv, mst_w, gs, nullptr);
}

__global__ void __launch_bounds__(1024) k_export(const FrameState* __restrict__ st, const float4* __restrict__ skel, int Mq, const GraphDevState* __restrict__ gs,
                                                 const int* __restrict__ acc_cnt, const int* __restrict__ acc, const int* __restrict__ mst_uv, const float* __restrict__ mst_w,
                                                 unsigned char* __restrict__ out) {
    export_block(st, skel, Mq, gs, acc_cnt, acc, mst_uv, mst_w, out);
}

// The tail of a frame in ONE single-CTA launch: R22 vertex_smooth -> G3 graph_adj -> G4 mst -> export (tail.cuh), with the vertex
// positions, neighbour lists, distance rows, edge keys and union-find in the shared-memory arena whenever they fit (a frame has
// V ~ 10^2 vertices; larger ones fall back to the global scratch arrays piece by piece -- same results).
struct TailArgs {
    FrameState* st;
    float4* skel; float4* skel2; float4* vtmp; int* vs_off; int* vs_idx; int* tmp_j; float* tmp_d; int cap;
    float radius, alpha; int niter;
    int with_graph; float max_dist_th;
    int* acc; int* acc_cnt; unsigned long long* gkeys; int keycap; int* gwork; int* mst_uv; float* mst_w; GraphDevState* gs;
    int Mq; unsigned char* out; unsigned arena_bytes;
    long long* dbg;        // OSEP_FUSED_DETAIL=1: globaltimer stamps at the phase borders (tools/tail_dbg.py)
};
constexpr int TAIL_THREADS = 512;
constexpr unsigned TAIL_ARENA = 96 * 1024;
__global__ void __launch_bounds__(TAIL_THREADS, 1) k_frame_tail(const __grid_constant__ TailArgs a) {
    extern __shared__ uint4 tail_smem[];
    unsigned char* arena = reinterpret_cast<unsigned char*>(tail_smem);
    auto stamp = [&](int k) { if (a.dbg && threadIdx.x == 0) { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); a.dbg[k] = t; } };
    stamp(0);
    frame_stamp(a.st, TS_TAIL);
    const int status0 = a.st->status, V0 = a.st->V;
    __syncthreads();
    const float4* pos = a.skel;
    if (status0 == 0) {
        if (V0 == 0) { if (threadIdx.x == 0) a.st->status = OSEP_STAGE_7; }
        else {
            const float4* p = vertex_smooth_block(a.st, a.skel, a.skel2, a.vtmp, a.vs_off, a.vs_idx, a.tmp_j, a.tmp_d, a.cap, a.radius, a.alpha, a.niter, arena, a.arena_bytes, a.dbg);
            if (p) pos = p;
        }
    }
    __syncthreads();
    stamp(1);
    if (a.with_graph) {
        const int G = a.st->status == 0 ? a.st->V : 0;
        size_t used = (pos != a.skel) ? smooth_pos_bytes(G) : 0;                 // the smoothed positions stay where they are
        graph_adj_thread(pos, G, threadIdx.x, TAIL_THREADS, a.max_dist_th, a.acc, a.acc_cnt);
        __syncthreads();
        stamp(2);
        // component arrays first (3 G ints), the rest of the arena holds the edge keys
        int* swork = reinterpret_cast<int*>(arena + used);
        const size_t wbytes = tail_align16((size_t)3 * G * 4);
        const bool w_in = used + wbytes <= a.arena_bytes;
        const size_t kbase = used + (w_in ? wbytes : 0);
        unsigned long long* skeys = reinterpret_cast<unsigned long long*>(arena + kbase);
        const int skeys_cap = kbase < a.arena_bytes ? (int)((a.arena_bytes - kbase) / 8) : 0;
        graph_mst_block(pos, G, a.acc, a.acc_cnt, skeys, skeys_cap, a.gkeys, a.keycap, swork, w_in ? 3 * G : 0, a.gwork, a.mst_uv, a.mst_w, a.gs, a.dbg ? a.dbg + 4 : nullptr);
        __syncthreads();
    }
    stamp(3);
    export_block(a.st, a.skel, a.Mq, a.with_graph ? a.gs : nullptr, a.acc_cnt, a.acc, a.mst_uv, a.mst_w, a.out);
    __syncthreads();
    stamp(7);
    if (threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); reinterpret_cast<FrameState*>(a.out)->tstamp[TS_END] = t; }   // (the state was exported above: the last stamp goes into the exported copy)
}

struct GraphHost {
    int cap = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    float4* pos = nullptr; int* acc = nullptr; int* acc_cnt = nullptr; unsigned long long* keys = nullptr; int keycap = 0;
    int* parent = nullptr /* 3 cap ints: component label, parent, best edge */; int* mst_uv = nullptr; float* mst_w = nullptr; GraphDevState* gs = nullptr;
    // host results
    std::vector<int> h_acc, h_acc_cnt, edges, adj_off, adj_idx, type, joints, leafs, rng_off, rng_idx;
    std::vector<float> edge_w;
    std::vector<std::vector<int>> adj;      // MST adjacency in push order
    std::vector<std::vector<int>> rng;      // graph_adj adjacency in push order (duplicates kept)
    int G = 0;
    // pinned mirrors of the in-frame graph (k_frame_tail): the first `mcap` vertices' lists travel with the frame
    // (they point into the owning Rosa handle's pinned export block)
    int mcap = 0; GraphDevState* m_gs = nullptr; int* m_acc = nullptr; int* m_acc_cnt = nullptr; int* m_uv = nullptr; float* m_w = nullptr;
};

static void graph_free(GraphHost* g) {
    if (!g) return;
    void* p[] = {g->pos, g->acc, g->acc_cnt, g->keys, g->parent, g->mst_uv, g->mst_w, g->gs};
    for (void* q : p) if (q) cudaFree(q);
    if (g->own_stream && g->stream) cudaStreamDestroy(g->stream);
    delete g;
}

static GraphHost* graph_alloc(int cap, int device, cudaStream_t stream) {
    GraphHost* g = new GraphHost();
    g->cap = cap; g->device = device; g->keycap = cap * GACC;
    if (stream) g->stream = stream; else { cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking); g->own_stream = true; }
    bool ok = cudaMalloc(&g->pos, sizeof(float4) * cap) == cudaSuccess && cudaMalloc(&g->acc, sizeof(int) * cap * GACC) == cudaSuccess &&
              cudaMalloc(&g->acc_cnt, sizeof(int) * cap) == cudaSuccess;
    int P2 = 1; while (P2 < g->keycap) P2 <<= 1;
    ok = ok && cudaMalloc(&g->keys, sizeof(unsigned long long) * P2) == cudaSuccess && cudaMalloc(&g->parent, sizeof(int) * 3 * cap) == cudaSuccess &&
         cudaMalloc(&g->mst_uv, sizeof(int) * 2 * cap) == cudaSuccess && cudaMalloc(&g->mst_w, sizeof(float) * cap) == cudaSuccess &&
         cudaMalloc(&g->gs, sizeof(GraphDevState)) == cudaSuccess;
    if (!ok) { graph_free(g); return nullptr; }
    return g;
}

// graph_decomp (gskel.cpp:539-567) incl. the case-2 fall-through: type 1 for leaves, 0 otherwise
static void decomp_host(const std::vector<std::vector<int>>& adj, std::vector<int>& joints, std::vector<int>& leafs, std::vector<int>& type) {
    joints.clear(); leafs.clear(); type.assign(adj.size(), 0);
    for (int i = 0; i < (int)adj.size(); ++i) {
        const int degree = (int)adj[i].size();
        if (degree == 1) { leafs.push_back(i); type[i] = 1; }
        else if (degree > 2) joints.push_back(i);
    }
}
static void to_csr(const std::vector<std::vector<int>>& a, std::vector<int>& off, std::vector<int>& idx) {
    off.assign(a.size() + 1, 0); idx.clear();
    for (size_t i = 0; i < a.size(); ++i) { off[i + 1] = off[i] + (int)a[i].size(); idx.insert(idx.end(), a[i].begin(), a[i].end()); }
}

// host half: adjacency lists in the reference's push order, graph_decomp, CSR views
static void graph_assemble(GraphHost* g, int G, int n_mst, const int* acc, const int* acc_cnt, const int* uv, const float* w) {
    g->G = G;
    g->adj.assign(G, {}); g->rng.assign(G, {});
    g->edges.assign(uv, uv + (size_t)2 * n_mst); g->edge_w.assign(w, w + n_mst);
    for (int i = 0; i < G; ++i) for (int t = 0; t < acc_cnt[i]; ++t) { const int nb = acc[(size_t)i * GACC + t]; g->rng[i].push_back(nb); g->rng[nb].push_back(i); }
    for (int e = 0; e < n_mst; ++e) { const int u = uv[2 * e], v = uv[2 * e + 1]; g->adj[u].push_back(v); g->adj[v].push_back(u); }
    decomp_host(g->adj, g->joints, g->leafs, g->type);
    to_csr(g->adj, g->adj_off, g->adj_idx); to_csr(g->rng, g->rng_off, g->rng_idx);
}

// pos_dev: G float4 on the device (already resident).  Runs G3+G4 on the device, G5 on the host.
static int graph_run(GraphHost* g, const float4* pos_dev, int G, float fuse_dist_th) {
    if (G == 0) { graph_assemble(g, 0, 0, nullptr, nullptr, nullptr, nullptr); return OSEP_OK; }
    if (G > g->cap || G > 65535) return OSEP_ERR_CAPACITY;
    const float max_dist_th = 2.5f * fuse_dist_th;
    k_graph_adj<<<div_up(G, GADJ_THREADS), GADJ_THREADS, 0, g->stream>>>(pos_dev, G, nullptr, max_dist_th, g->acc, g->acc_cnt);
    k_graph_mst<<<1, 1024, 0, g->stream>>>(pos_dev, G, nullptr, g->acc, g->acc_cnt, g->keys, g->keycap, g->parent, g->mst_uv, g->mst_w, g->gs);
    GraphDevState gs;
    if (cudaMemcpyAsync(&gs, g->gs, sizeof(gs), cudaMemcpyDeviceToHost, g->stream) != cudaSuccess) return OSEP_ERR_CUDA;
    g->h_acc.resize((size_t)G * GACC); g->h_acc_cnt.resize(G);
    cudaMemcpyAsync(g->h_acc.data(), g->acc, sizeof(int) * G * GACC, cudaMemcpyDeviceToHost, g->stream);
    cudaMemcpyAsync(g->h_acc_cnt.data(), g->acc_cnt, sizeof(int) * G, cudaMemcpyDeviceToHost, g->stream);
    if (cudaStreamSynchronize(g->stream) != cudaSuccess) return OSEP_ERR_CUDA;
    if (gs.overflow) return OSEP_ERR_CAPACITY;
    std::vector<int> uv((size_t)2 * gs.n_mst); std::vector<float> w(gs.n_mst);
    if (gs.n_mst) {
        cudaMemcpy(uv.data(), g->mst_uv, sizeof(int) * 2 * gs.n_mst, cudaMemcpyDeviceToHost);
        cudaMemcpy(w.data(), g->mst_w, sizeof(float) * gs.n_mst, cudaMemcpyDeviceToHost);
    }
    graph_assemble(g, G, gs.n_mst, g->h_acc.data(), g->h_acc_cnt.data(), uv.data(), w.data());
    return OSEP_OK;
}

// after the frame's stream has been synchronised: host half from the mirrors (graphs larger than the mirrors are fetched on demand)
static int graph_collect_frame(GraphHost* g) {
    const GraphDevState gs = *g->m_gs;
    if (gs.overflow) return OSEP_ERR_CAPACITY;
    const int G = gs.pad;
    if (G <= g->mcap) { graph_assemble(g, G, gs.n_mst, g->m_acc, g->m_acc_cnt, g->m_uv, g->m_w); return OSEP_OK; }
    g->h_acc.resize((size_t)G * GACC); g->h_acc_cnt.resize(G);
    std::vector<int> uv((size_t)2 * gs.n_mst); std::vector<float> w(gs.n_mst);
    if (cudaMemcpy(g->h_acc.data(), g->acc, sizeof(int) * G * GACC, cudaMemcpyDeviceToHost) != cudaSuccess) return OSEP_ERR_CUDA;
    cudaMemcpy(g->h_acc_cnt.data(), g->acc_cnt, sizeof(int) * G, cudaMemcpyDeviceToHost);
    if (gs.n_mst) { cudaMemcpy(uv.data(), g->mst_uv, sizeof(int) * 2 * gs.n_mst, cudaMemcpyDeviceToHost); cudaMemcpy(w.data(), g->mst_w, sizeof(float) * gs.n_mst, cudaMemcpyDeviceToHost); }
    graph_assemble(g, G, gs.n_mst, g->h_acc.data(), g->h_acc_cnt.data(), uv.data(), w.data());
    return OSEP_OK;
}

static void graph_view(GraphHost* g, osep_graph_view* out) {
    out->n_vertices = g->G; out->n_edges = (int)g->edge_w.size();
    out->edges = g->edges.data(); out->edge_w = g->edge_w.data();
    out->adj_off = g->adj_off.data(); out->adj_idx = g->adj_idx.data(); out->type = g->type.data();
    out->joints = g->joints.data(); out->n_joints = (int)g->joints.size();
    out->leafs = g->leafs.data(); out->n_leafs = (int)g->leafs.size();
    out->rng_off = g->rng_off.data(); out->rng_idx = g->rng_idx.data();
}

// ------------------------------------------------------------------------------------------
// Stateful GSkel (host side of the class; mirrors GSkelData, gskel.hpp:88-104)
struct GVertex {
    f3 position{0, 0, 0};
    f3 kx{0, 0, 0};
    float kP[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int obs_count = 0, unconf_check = 0, type = 0, smooth_iters = 0;
    bool just_approved = false, frozen = false, conf_check = false, marked_for_deletion = false, updated = false;
};

struct GSkelImpl {
    osep_gskel_config cfg;
    GraphHost* g = nullptr;
    std::vector<f3> new_cands;
    std::vector<GVertex> prelim_vers, global_vers;
    std::vector<f3> global_cloud;
    std::vector<int> new_vers_indxs, joints, leafs;
    std::vector<std::vector<int>> branches, global_adj;
    size_t gskel_size = 0;
    int failed_stage = 0;
    std::vector<float> out_xyz4;
    std::vector<float4> upload;
    std::vector<float> tf_host;        // osep_gskel_run_tf_device: transformed candidates (device -> host, 16 bytes per vertex)
    // views
    std::vector<int> v_adj_off, v_adj_idx, v_type, v_edges; std::vector<float> v_w;

    bool size_assert() const { return global_vers.size() == global_cloud.size() && global_cloud.size() == global_adj.size() && global_adj.size() == gskel_size; }
    void decomp() {
        std::vector<int> type; decomp_host(global_adj, joints, leafs, type);
        for (size_t i = 0; i < global_vers.size() && i < type.size(); ++i) global_vers[i].type = type[i];
    }
};

// VertexLKF::update (lkf_vertex_fuse.hpp:16-28), Q = pn*I, R = mn*I (gskel.hpp:135-136)
static void lkf_update(GVertex& g, const f3& z, float pn, float mn) {
    float Pp[9], S[9], Si[9], K[9], IK[9];
    for (int i = 0; i < 9; ++i) { const float dg = (i % 4 == 0) ? 1.0f : 0.0f; Pp[i] = g.kP[i] + dg * pn; }
    for (int i = 0; i < 9; ++i) { const float dg = (i % 4 == 0) ? 1.0f : 0.0f; S[i] = Pp[i] + dg * mn; }
    mat3_inv(S, Si);
    mat3_mul(Pp, Si, K);
    const f3 inn = sub3(z, g.kx);
    const f3 Kd{sum3(K[0] * inn.x, K[1] * inn.y, K[2] * inn.z), sum3(K[3] * inn.x, K[4] * inn.y, K[5] * inn.z), sum3(K[6] * inn.x, K[7] * inn.y, K[8] * inn.z)};
    g.kx = add3(g.kx, Kd);
    for (int i = 0; i < 9; ++i) { const float dg = (i % 4 == 0) ? 1.0f : 0.0f; IK[i] = dg - K[i]; }
    mat3_mul(IK, Pp, g.kP);
}

// increment_skeleton (gskel.cpp:40-199).  The prelim radius search is a brute-force scan in
// canonical (d2, index) order: only the nearest non-frozen hit and two flags are consumed.
static bool gs_increment(GSkelImpl& S) {
    if (S.new_cands.empty()) return false;
    for (auto& v : S.prelim_vers) v.just_approved = false;
    const size_t n_prelim0 = S.prelim_vers.size();          // the kd-tree is a snapshot taken at tick start
    std::vector<f3> prelim_cloud(n_prelim0);
    for (size_t i = 0; i < n_prelim0; ++i) prelim_cloud[i] = S.prelim_vers[i].position;
    const float fuse_r2 = S.cfg.fuse_dist_th * S.cfg.fuse_dist_th;
    std::vector<f3> spawned;
    std::vector<std::pair<float, int>> hits;
    for (const f3& ver : S.new_cands) {
        if (!fin3(ver)) continue;
        int chosen = -1; float chosen_d2 = FLT_MAX; bool any_in_radius = false, frozen_in_radius = false;
        if (!prelim_cloud.empty()) {
            hits.clear();
            for (size_t i = 0; i < n_prelim0; ++i) { if (!fin3(prelim_cloud[i])) continue; const float d = dist2(ver, prelim_cloud[i]); if (d < fuse_r2) hits.push_back({d, (int)i}); }
            std::sort(hits.begin(), hits.end());
            if (!hits.empty()) {
                any_in_radius = true;
                for (auto& hd : hits) {
                    if (S.prelim_vers[hd.second].frozen) { frozen_in_radius = true; continue; }
                    if (hd.first < chosen_d2) { chosen_d2 = hd.first; chosen = hd.second; }
                }
            }
        }
        bool near_spawned = false;
        if (chosen >= 0 && chosen_d2 <= fuse_r2) {
            GVertex& g = S.prelim_vers[chosen];
            lkf_update(g, ver, S.cfg.lkf_pn, S.cfg.lkf_mn);
            if (!fin3(g.kx)) { g.marked_for_deletion = true; continue; }
            g.position = g.kx;
            g.obs_count++;
            const float trace = sum3(g.kP[0], g.kP[4], g.kP[8]);
            if (!g.conf_check && trace < S.cfg.fuse_conf_th) { g.conf_check = true; g.just_approved = true; g.frozen = true; g.unconf_check = 0; }
            else g.unconf_check++;
            if (g.unconf_check > S.cfg.max_obs_wo_conf) g.marked_for_deletion = true;
        } else if (any_in_radius || frozen_in_radius) {
            continue;
        } else {
            for (const f3& s : spawned) if (sqn3(sub3(s, ver)) <= fuse_r2) { near_spawned = true; break; }
            if (!near_spawned) {
                GVertex nv;
                nv.kx = ver; for (int i = 0; i < 9; ++i) nv.kP[i] = (i % 4 == 0) ? 1.0f : 0.0f;
                nv.position = ver; nv.obs_count++; nv.smooth_iters = S.cfg.niter_smooth_vertex;
                S.prelim_vers.push_back(nv);
                spawned.push_back(ver);
            }
        }
    }
    { std::vector<GVertex> k; for (auto& v : S.prelim_vers) if (!v.marked_for_deletion) k.push_back(v); S.prelim_vers.swap(k); }
    S.new_vers_indxs.clear();
    for (auto& v : S.prelim_vers) {
        if (v.just_approved && v.position.z > S.cfg.gnd_th) {
            S.global_vers.push_back(v);
            S.global_cloud.push_back(v.position);
            S.new_vers_indxs.push_back((int)S.global_vers.size() - 1);
            v.just_approved = false; v.updated = true;
        }
    }
    { std::vector<GVertex> kept; std::vector<f3> clean;
      for (auto& v : S.global_vers) { if (!fin3(v.position)) continue; kept.push_back(v); clean.push_back(v.position); }
      S.global_vers.swap(kept); S.global_cloud.swap(clean); S.gskel_size = S.global_vers.size(); }
    return !S.new_vers_indxs.empty();
}

// graph_adj + mst on the device (gskel.cpp:201-281)
static int gs_graph(GSkelImpl& S, bool& adj_ok, bool& mst_ok) {
    adj_ok = mst_ok = false;
    if (S.global_cloud.empty()) return OSEP_OK;
    const int G = (int)S.gskel_size;
    S.upload.resize(G);
    for (int i = 0; i < G; ++i) S.upload[i] = make_float4(S.global_cloud[i].x, S.global_cloud[i].y, S.global_cloud[i].z, 1.0f);
    if (G > S.g->cap) {
        // the global skeleton outgrew the device arena: grow it (the state machine above has already accepted the tick, so
        // failing here would leave prelim / global state ahead of the graph for good).  u, v travel in 16 bits: 65535 is final.
        if (G > 65535) return OSEP_ERR_CAPACITY;
        GraphHost* bigger = graph_alloc(std::min(65535, std::max(2 * S.g->cap, G + 1024)), S.g->device, nullptr);
        if (!bigger) return OSEP_ERR_CUDA;
        graph_free(S.g); S.g = bigger;
    }
    if (cudaMemcpyAsync(S.g->pos, S.upload.data(), sizeof(float4) * G, cudaMemcpyHostToDevice, S.g->stream) != cudaSuccess) return OSEP_ERR_CUDA;
    const int rc = graph_run(S.g, S.g->pos, G, S.cfg.fuse_dist_th);
    if (rc) return rc;
    S.global_adj = S.g->rng;                       // graph_adj result (gskel.cpp:244)
    adj_ok = S.size_assert();
    if (!adj_ok) return OSEP_OK;
    if (S.gskel_size == 0 || S.global_adj.empty()) return OSEP_OK;
    S.global_adj = S.g->adj;                       // MST adjacency (gskel.cpp:277)
    S.decomp();
    mst_ok = S.size_assert();
    return OSEP_OK;
}

// merge_into (gskel.cpp:569-589)
static void gs_merge_into(GSkelImpl& S, int keep, int del) {
    GVertex& Vi = S.global_vers[keep]; GVertex& Vj = S.global_vers[del];
    const int tot = Vi.obs_count + Vj.obs_count;
    Vi.position = div3(add3(mul3(Vi.position, (float)Vi.obs_count), mul3(Vj.position, (float)Vj.obs_count)), (float)tot);
    Vi.obs_count = tot;
    for (int nb : S.global_adj[del]) {
        if (nb == keep) continue;
        auto& kn = S.global_adj[keep];
        if (std::find(kn.begin(), kn.end(), nb) == kn.end()) kn.push_back(nb);
        auto& nn = S.global_adj[nb];
        std::replace(nn.begin(), nn.end(), del, keep);
    }
}

// vertex_merge (gskel.cpp:283-353)
static bool gs_vertex_merge(GSkelImpl& S) {
    const int N_new = (int)S.new_vers_indxs.size();
    if (S.gskel_size == 0 || N_new == 0) return false;
    if ((float)N_new / (float)S.gskel_size > 0.5) return false;
    auto is_joint = [&](int idx) { return std::find(S.joints.begin(), S.joints.end(), idx) != S.joints.end(); };
    std::set<int> to_delete;
    for (int new_id : S.new_vers_indxs) {
        if (new_id < 0 || new_id >= (int)S.gskel_size) continue;
        if (to_delete.count(new_id)) continue;
        const std::vector<int> nbrs = S.global_adj[new_id];
        for (int nb_id : nbrs) {
            if (nb_id < 0 || nb_id >= (int)S.gskel_size) continue;
            if (new_id == nb_id || to_delete.count(nb_id)) continue;
            bool do_merge = false;
            if (is_joint(new_id) && is_joint(nb_id)) { do_merge = true; S.joints.erase(std::remove(S.joints.begin(), S.joints.end(), nb_id), S.joints.end()); }
            const float dist = norm3(sub3(S.global_vers[new_id].position, S.global_vers[nb_id].position));
            if (!do_merge && dist < 0.5f * S.cfg.fuse_dist_th) do_merge = true;
            if (!do_merge) continue;
            gs_merge_into(S, nb_id, new_id);
            to_delete.insert(new_id);
            break;
        }
    }
    if (to_delete.empty()) return true;
    for (auto it = to_delete.rbegin(); it != to_delete.rend(); ++it) {
        const int del = *it;
        if (del < 0 || del > (int)S.global_vers.size()) continue;
        S.global_vers.erase(S.global_vers.begin() + del);
        S.global_cloud.erase(S.global_cloud.begin() + del);
        S.global_adj.erase(S.global_adj.begin() + del);
        for (auto& nbrs : S.global_adj) {
            nbrs.erase(std::remove(nbrs.begin(), nbrs.end(), del), nbrs.end());
            for (auto& v : nbrs) if (v > del) --v;
            std::sort(nbrs.begin(), nbrs.end());
            nbrs.erase(std::unique(nbrs.begin(), nbrs.end()), nbrs.end());
        }
        S.new_vers_indxs.erase(std::remove(S.new_vers_indxs.begin(), S.new_vers_indxs.end(), del), S.new_vers_indxs.end());
        for (auto& id : S.new_vers_indxs) if (id > del) --id;
    }
    S.gskel_size = S.global_vers.size();
    S.decomp();
    return S.size_assert();
}

// prune (gskel.cpp:355-394)
static bool gs_prune(GSkelImpl& S) {
    const int N_new = (int)S.new_vers_indxs.size();
    if (S.gskel_size == 0 || N_new == 0) return false;
    if ((float)N_new / (float)S.gskel_size > 0.5) return false;
    std::set<int> joint_set(S.joints.begin(), S.joints.end()), new_set(S.new_vers_indxs.begin(), S.new_vers_indxs.end());
    std::vector<int> to_delete;
    for (int leaf : S.leafs) {
        if (!new_set.count(leaf)) continue;
        if (S.global_adj[leaf].size() != 1) continue;
        if (joint_set.count(S.global_adj[leaf][0])) to_delete.push_back(leaf);
    }
    std::sort(to_delete.rbegin(), to_delete.rend());
    for (int idx : to_delete) {
        S.global_vers.erase(S.global_vers.begin() + idx);
        S.global_cloud.erase(S.global_cloud.begin() + idx);
        S.global_adj.erase(S.global_adj.begin() + idx);
        for (auto& nbrs : S.global_adj) { nbrs.erase(std::remove(nbrs.begin(), nbrs.end(), idx), nbrs.end()); for (auto& v : nbrs) if (v > idx) --v; }
        S.new_vers_indxs.erase(std::remove(S.new_vers_indxs.begin(), S.new_vers_indxs.end(), idx), S.new_vers_indxs.end());
        for (auto& id : S.new_vers_indxs) if (id > idx) --id;
    }
    S.gskel_size = S.global_vers.size();
    S.decomp();
    return true;
}

// smooth_vertex_positions (gskel.cpp:396-434)
static bool gs_smooth(GSkelImpl& S) {
    if (S.gskel_size == 0 || S.global_adj.empty()) return false;
    std::vector<f3> new_pos(S.gskel_size, f3{0, 0, 0});
    for (size_t i = 0; i < S.gskel_size; ++i) {
        GVertex& v = S.global_vers[i];
        if (v.smooth_iters > 0) {
            f3 avg{0, 0, 0};
            for (int j : S.global_adj[i]) avg = add3(avg, S.global_vers[j].position);
            avg = div3(avg, (float)S.global_adj[i].size());
            new_pos[i] = add3(mul3(v.position, 1.0f - S.cfg.vertex_smooth_coef), mul3(avg, S.cfg.vertex_smooth_coef));
            --v.smooth_iters;
        } else new_pos[i] = v.position;
    }
    for (size_t i = 0; i < S.gskel_size; ++i) S.global_vers[i].position = new_pos[i];
    S.global_cloud.clear();
    for (const auto& v : S.global_vers) S.global_cloud.push_back(v.position);
    return true;
}

// extract_branches (gskel.cpp:436-536) with its quirks (:448/:485 swap of loop copies, :486 stray ';').
// The branch sort is stable (the reference's std::sort leaves equal lengths unordered).
static bool gs_branches(GSkelImpl& S) {
    if (S.gskel_size == 0) return false;
    S.branches.clear();
    std::set<std::pair<int, int>> visited;
    auto is_endpoint = [&](int v) { return S.global_vers[v].type == 1 || S.global_vers[v].type == 3; };
    auto walk = [&](int endp, int nb) {
        std::vector<int> branch; branch.push_back(endp);
        int prev = endp, curr = nb;
        size_t guard = 0;
        while (guard++ <= S.gskel_size + 1) {
            branch.push_back(curr);
            if (is_endpoint(curr) && curr != endp) break;
            int next = -1;
            for (int nn : S.global_adj[curr]) if (nn != prev) { next = nn; break; }
            if (next == -1) break;
            prev = curr; curr = next;
        }
        for (size_t i = 1; i < branch.size(); ++i) visited.insert({branch[i - 1], branch[i]});
        S.branches.push_back(branch);
    };
    for (int endp0 : S.leafs) {
        int endp = endp0;
        for (int nb0 : S.global_adj[endp0]) { int nb = nb0; if (endp > nb) std::swap(endp, nb); if (visited.count({endp, nb})) continue; walk(endp, nb); }
    }
    for (int endp0 : S.joints) {
        int endp = endp0;
        for (int nb0 : S.global_adj[endp0]) { int nb = nb0; if (endp > nb) std::swap(endp, nb); walk(endp, nb); }
    }
    std::stable_sort(S.branches.begin(), S.branches.end(), [](const std::vector<int>& a, const std::vector<int>& b) { return a.size() < b.size(); });
    const size_t min_bl = (size_t)S.cfg.min_branch_length;
    S.branches.erase(std::remove_if(S.branches.begin(), S.branches.end(), [min_bl](const std::vector<int>& b) { return b.size() < min_bl; }), S.branches.end());
    return true;
}

}  // namespace osep

using namespace osep;

struct osep_gskel { GSkelImpl impl; int device; };
struct osep_rosa;   // defined in api.cu

namespace osep {
// hooks used by api.cu for osep_rosa_graph
GraphHost* graph_create(int cap, int device, cudaStream_t stream) { return graph_alloc(cap, device, stream); }
void graph_destroy(GraphHost* g) { graph_free(g); }
int graph_run_device(GraphHost* g, const float4* pos_dev, int G, float fuse_dist_th, osep_graph_view* out) {
    const int rc = graph_run(g, pos_dev, G, fuse_dist_th);
    if (rc == OSEP_OK && out) graph_view(g, out);
    return rc;
}
// the tail of a frame (vertex_smooth, then -- with g -- graph_adj + mst, then the export block) as one launch; returns the
// bytes the host copy must move
size_t frame_tail_enqueue(osep_rosa_impl* h, GraphHost* g, float fuse_dist_th, cudaStream_t s) {
    RosaDev& d = h->d;
    const osep_rosa_config& c = h->cfg;
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_frame_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAIL_ARENA); attr_set = true; }
    TailArgs a;
    a.st = d.st; a.skel = d.skel; a.skel2 = d.skel2; a.vtmp = d.vtmp; a.vs_off = d.vs_off; a.vs_idx = d.vs_idx; a.tmp_j = d.rv1; a.tmp_d = (float*)d.rk1; a.cap = d.Scap;
    a.radius = c.radius_smooth; a.alpha = c.alpha_recenter; a.niter = c.niter_smooth;
    a.with_graph = g ? 1 : 0; a.max_dist_th = 2.5f * fuse_dist_th;
    a.acc = g ? g->acc : nullptr; a.acc_cnt = g ? g->acc_cnt : nullptr; a.gkeys = g ? g->keys : nullptr; a.keycap = g ? g->keycap : 0; a.gwork = g ? g->parent : nullptr;
    a.mst_uv = g ? g->mst_uv : nullptr; a.mst_w = g ? g->mst_w : nullptr; a.gs = g ? g->gs : nullptr;
    a.Mq = d.Mcap / 4; a.out = d.exp; a.arena_bytes = TAIL_ARENA; a.dbg = h->fused_detail ? d.fused_dbg + 48 : nullptr;
    k_frame_tail<<<1, TAIL_THREADS, TAIL_ARENA, s>>>(a);
    return exp_bytes(d.Mcap / 4, g != nullptr);
}
size_t frame_export_bytes(int Mq) { return exp_bytes(Mq, true); }
void graph_set_mirrors(GraphHost* g, unsigned char* exp_host, int Mq) {
    int* og = reinterpret_cast<int*>(exp_host + exp_off_graph(Mq));
    g->mcap = Mq; g->m_gs = reinterpret_cast<GraphDevState*>(og); g->m_acc_cnt = og + 4; g->m_acc = g->m_acc_cnt + Mq;
    g->m_uv = g->m_acc + (size_t)Mq * GACC; g->m_w = reinterpret_cast<float*>(g->m_uv + 2 * (size_t)Mq);
}
int graph_collect_in_frame(GraphHost* g, osep_graph_view* out) {
    const int rc = graph_collect_frame(g);
    if (rc == OSEP_OK && out) graph_view(g, out);
    return rc;
}
}  // namespace osep

extern "C" {

int osep_gskel_create(const osep_gskel_config* cfg, int device, int capacity_vertices, osep_gskel** out) {
    if (!out || capacity_vertices < 1 || capacity_vertices > 65535) return OSEP_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return OSEP_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return OSEP_ERR_CUDA;
    osep_gskel* h = new osep_gskel();
    osep_gskel_default_config(&h->impl.cfg);
    if (cfg) h->impl.cfg = *cfg;
    h->device = device;
    h->impl.g = graph_alloc(capacity_vertices, device, nullptr);
    if (!h->impl.g) { delete h; return OSEP_ERR_CUDA; }
    *out = h;
    return OSEP_OK;
}
int osep_gskel_destroy(osep_gskel* h) {
    if (!h) return OSEP_ERR_ARG;
    cudaSetDevice(h->device);
    graph_free(h->impl.g);
    delete h;
    return OSEP_OK;
}

// gskel_run (gskel.cpp:21-38): stages in RUN_STEP order; returns 0 or the failing stage 1..7
int osep_gskel_run(osep_gskel* h, const float* cand_xyz, int n, int stride_floats) {
    if (!h || n < 0 || stride_floats < 3 || (n > 0 && !cand_xyz)) return OSEP_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return OSEP_ERR_CUDA;
    GSkelImpl& S = h->impl;
    S.new_cands.resize(n);
    for (int i = 0; i < n; ++i) S.new_cands[i] = {cand_xyz[(size_t)i * stride_floats], cand_xyz[(size_t)i * stride_floats + 1], cand_xyz[(size_t)i * stride_floats + 2]};
    S.failed_stage = 0;
    if (!gs_increment(S)) return S.failed_stage = OSEP_STAGE_1;
    bool adj_ok = false, mst_ok = false;
    const int rc = gs_graph(S, adj_ok, mst_ok);
    if (rc) return rc;
    if (!adj_ok) return S.failed_stage = OSEP_STAGE_2;
    if (!mst_ok) return S.failed_stage = OSEP_STAGE_3;
    if (!gs_vertex_merge(S)) return S.failed_stage = OSEP_STAGE_4;
    if (!gs_prune(S)) return S.failed_stage = OSEP_STAGE_5;
    if (!gs_smooth(S)) return S.failed_stage = OSEP_STAGE_6;
    if (!gs_branches(S)) return S.failed_stage = OSEP_STAGE_7;
    return OSEP_OK;
}

// gskel_node.cpp:137-139: tf2::doTransform(PointCloud2) = Eigen::Translation3f(t) * Eigen::Quaternionf(w,x,y,z) applied to
// every point in float32 (tf2_sensor_msgs.hpp); Quaternion::toRotationMatrix + (linear * p) + translation of Eigen 3.4.
int osep_gskel_run_tf(osep_gskel* h, const float* cand_xyz, int n, int stride_floats, const double* translation, const double* rotation_xyzw) {
    if (!h || n < 0 || stride_floats < 3 || (n > 0 && !cand_xyz) || !translation || !rotation_xyzw) return OSEP_ERR_ARG;
    const float qx = (float)rotation_xyzw[0], qy = (float)rotation_xyzw[1], qz = (float)rotation_xyzw[2], qw = (float)rotation_xyzw[3];
    const float t0 = (float)translation[0], t1 = (float)translation[1], t2 = (float)translation[2];
    const float tx = 2.0f * qx, ty = 2.0f * qy, tz = 2.0f * qz;
    const float twx = tx * qw, twy = ty * qw, twz = tz * qw;
    const float txx = tx * qx, txy = ty * qx, txz = tz * qx;
    const float tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    const float r00 = 1.0f - (tyy + tzz), r01 = txy - twz, r02 = txz + twy;
    const float r10 = txy + twz, r11 = 1.0f - (txx + tzz), r12 = tyz - twx;
    const float r20 = txz - twy, r21 = tyz + twx, r22 = 1.0f - (txx + tyy);
    std::vector<float> g((size_t)n * 3);
    for (int i = 0; i < n; ++i) {
        const float x = cand_xyz[(size_t)i * stride_floats], y = cand_xyz[(size_t)i * stride_floats + 1], z = cand_xyz[(size_t)i * stride_floats + 2];
        g[3 * (size_t)i] = sum3(r00 * x, r01 * y, r02 * z) + t0;
        g[3 * (size_t)i + 1] = sum3(r10 * x, r11 * y, r12 * z) + t1;
        g[3 * (size_t)i + 2] = sum3(r20 * x, r21 * y, r22 * z) + t2;
    }
    return osep_gskel_run(h, g.data(), n, 3);
}

// the Rosa -> TF -> GSkel hand-off without a host-side transform: candidates = the Rosa handle's device vertices
int osep_gskel_run_tf_device(osep_gskel* h, osep_rosa* rosa, const double* translation, const double* rotation_xyzw) {
    if (!h || !rosa || !translation || !rotation_xyzw) return OSEP_ERR_ARG;
    const float* src = nullptr; int n = 0;
    int rc = osep_rosa_vertices_device(rosa, &src, &n);          // waits for the frame
    if (rc != OSEP_OK) return rc;
    if (cudaSetDevice(h->device) != cudaSuccess) return OSEP_ERR_CUDA;
    GraphHost* g = h->impl.g;
    if (n > g->cap) return OSEP_ERR_CAPACITY;
    std::vector<float>& host = h->impl.tf_host;
    host.resize((size_t)n * 4);
    if (n > 0) {
        rc = osep_tf_points_device(h->device, src, n, 4, translation, rotation_xyzw, reinterpret_cast<float*>(g->pos), 4, (void*)g->stream);
        if (rc != OSEP_OK) return rc;
        if (cudaMemcpyAsync(host.data(), g->pos, sizeof(float4) * n, cudaMemcpyDeviceToHost, g->stream) != cudaSuccess || cudaStreamSynchronize(g->stream) != cudaSuccess) return OSEP_ERR_CUDA;
    }
    return osep_gskel_run(h, host.data(), n, 4);
}

int osep_gskel_vertices(osep_gskel* h, const float** xyz4, int* n_vertices) {
    if (!h || !xyz4 || !n_vertices) return OSEP_ERR_ARG;
    GSkelImpl& S = h->impl;
    S.out_xyz4.resize(S.global_cloud.size() * 4);
    for (size_t i = 0; i < S.global_cloud.size(); ++i) { S.out_xyz4[4 * i] = S.global_cloud[i].x; S.out_xyz4[4 * i + 1] = S.global_cloud[i].y; S.out_xyz4[4 * i + 2] = S.global_cloud[i].z; S.out_xyz4[4 * i + 3] = 1.0f; }
    *xyz4 = S.out_xyz4.data(); *n_vertices = (int)S.global_cloud.size();
    return OSEP_OK;
}

int osep_gskel_graph(osep_gskel* h, osep_graph_view* out) {
    if (!h || !out) return OSEP_ERR_ARG;
    GSkelImpl& S = h->impl;
    memset(out, 0, sizeof(*out));
    to_csr(S.global_adj, S.v_adj_off, S.v_adj_idx);
    S.v_type.resize(S.global_vers.size());
    for (size_t i = 0; i < S.global_vers.size(); ++i) S.v_type[i] = S.global_vers[i].type;
    S.v_edges.clear(); S.v_w.clear();
    for (int i = 0; i < (int)S.global_adj.size(); ++i) for (int nb : S.global_adj[i]) if (nb > i) {
        S.v_edges.push_back(i); S.v_edges.push_back(nb);
        S.v_w.push_back(norm3(sub3(S.global_vers[i].position, S.global_vers[nb].position)));
    }
    out->n_vertices = (int)S.global_vers.size(); out->n_edges = (int)S.v_w.size();
    out->edges = S.v_edges.data(); out->edge_w = S.v_w.data();
    out->adj_off = S.v_adj_off.data(); out->adj_idx = S.v_adj_idx.data(); out->type = S.v_type.data();
    out->joints = S.joints.data(); out->n_joints = (int)S.joints.size();
    out->leafs = S.leafs.data(); out->n_leafs = (int)S.leafs.size();
    out->rng_off = nullptr; out->rng_idx = nullptr;
    return OSEP_OK;
}

static long put_ints(const std::vector<int>& v, void* dst, long cap) { if (dst && cap >= (long)(v.size() * 4) && !v.empty()) memcpy(dst, v.data(), v.size() * 4); return (long)v.size(); }

long osep_gskel_get(osep_gskel* h, const char* name, void* dst, long cap) {
    if (!h || !name) return -1;
    GSkelImpl& S = h->impl;
    const std::string nm(name);
    auto put_pos = [&](const std::vector<f3>& p) { if (dst && cap >= (long)(p.size() * 12) && !p.empty()) memcpy(dst, p.data(), p.size() * 12); return (long)p.size() * 3; };
    if (nm == "global") return put_pos(S.global_cloud);
    if (nm == "prelim") { std::vector<f3> p; for (auto& v : S.prelim_vers) p.push_back(v.position); return put_pos(p); }
    if (nm == "prelim_frozen") { std::vector<int> p; for (auto& v : S.prelim_vers) p.push_back(v.frozen ? 1 : 0); return put_ints(p, dst, cap); }
    if (nm == "adj_off" || nm == "adj_idx") { std::vector<int> o, x; to_csr(S.global_adj, o, x); return put_ints(nm == "adj_off" ? o : x, dst, cap); }
    if (nm == "branch_off" || nm == "branch_idx") { std::vector<int> o, x; to_csr(S.branches, o, x); return put_ints(nm == "branch_off" ? o : x, dst, cap); }
    if (nm == "joints") return put_ints(S.joints, dst, cap);
    if (nm == "leafs") return put_ints(S.leafs, dst, cap);
    if (nm == "new_idx") return put_ints(S.new_vers_indxs, dst, cap);
    if (nm == "type") { std::vector<int> t; for (auto& v : S.global_vers) t.push_back(v.type); return put_ints(t, dst, cap); }
    if (nm == "obs") { std::vector<int> t; for (auto& v : S.global_vers) t.push_back(v.obs_count); return put_ints(t, dst, cap); }
    return -1;
}

int osep_graph_build(osep_gskel* h, const float* xyz, int n, int stride_floats, float fuse_dist_th, osep_graph_view* out) {
    if (!h || !out || n < 0 || stride_floats < 3) return OSEP_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return OSEP_ERR_CUDA;
    GraphHost* g = h->impl.g;
    if (n > g->cap) return OSEP_ERR_CAPACITY;
    std::vector<float4>& up = h->impl.upload;
    up.resize(n);
    for (int i = 0; i < n; ++i) up[i] = make_float4(xyz[(size_t)i * stride_floats], xyz[(size_t)i * stride_floats + 1], xyz[(size_t)i * stride_floats + 2], 1.0f);
    if (n && cudaMemcpyAsync(g->pos, up.data(), sizeof(float4) * n, cudaMemcpyHostToDevice, g->stream) != cudaSuccess) return OSEP_ERR_CUDA;
    const int rc = graph_run(g, g->pos, n, fuse_dist_th);
    if (rc == OSEP_OK) graph_view(g, out);
    return rc;
}

int osep_graph_build_device(osep_gskel* h, const float* xyz_dev, int n, int stride_floats, float fuse_dist_th, osep_graph_view* out) {
    if (!h || !out || n < 0 || (stride_floats != 3 && stride_floats != 4)) return OSEP_ERR_ARG;
    if (cudaSetDevice(h->device) != cudaSuccess) return OSEP_ERR_CUDA;
    GraphHost* g = h->impl.g;
    if (n > g->cap) return OSEP_ERR_CAPACITY;
    if (n) {
        cudaError_t e;
        if (stride_floats == 4) e = cudaMemcpyAsync(g->pos, xyz_dev, sizeof(float4) * n, cudaMemcpyDeviceToDevice, g->stream);
        else e = cudaMemcpy2DAsync(g->pos, sizeof(float4), xyz_dev, 12, 12, n, cudaMemcpyDeviceToDevice, g->stream);   // w is not read by the graph kernels
        if (e != cudaSuccess) return OSEP_ERR_CUDA;
    }
    const int rc = graph_run(g, g->pos, n, fuse_dist_th);
    if (rc == OSEP_OK) graph_view(g, out);
    return rc;
}

}  // extern "C"
