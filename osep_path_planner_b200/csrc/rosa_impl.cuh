// csrc/rosa_impl.cuh -- the Rosa handle: device arena + launch entry points of the stages.
#pragma once
#include "osep_common.cuh"
#include <vector>
#include <string>

namespace osep {

struct RosaDev {
    // ---- capacities (fixed at create)
    int Ncap = 0;          // input points
    int Mcap = 0;          // downsampled points
    int Scap = 0;          // similarity CSR entries
    int T = 0;             // hash table slots (power of two >= 2*Ncap)
    int W = 0;             // bitmask words per row = Mcap/32
    int knn = 0, nbk = 0;  // ne_knn, nb_knn

    FrameState* st = nullptr;
    unsigned char* exp = nullptr;      // per-frame result block gathered by k_export (graph.cu)
    FrameParams* prm = nullptr;        // device copy of the per-frame launch parameters (osep_common.cuh)
    // N-scale
    float* in = nullptr;               // staged input (host-buffer path), Ncap * 4 floats
    unsigned char* in_raw = nullptr; size_t in_raw_bytes = 0;   // staged raw PointCloud2 records (osep_rosa_run_pointcloud2), grown on demand
    float4* P = nullptr;               // filtered points (x,y,z,1)
    int* filt_idx = nullptr;
    int* blk_cnt = nullptr;
    unsigned long long* gtab = nullptr;  // kNN grid: (epoch<<32 | cell key)
    int* g_cnt = nullptr; int* g_start = nullptr; int* g_fill = nullptr;
    int* p_slot = nullptr;
    int2* units = nullptr; int* slow_list = nullptr; int* over_list = nullptr; int2* knn_retry = nullptr; int* knn_grow = nullptr;
    int* g_cell = nullptr; int* cell_slots = nullptr; int* pc_cell = nullptr; int2* nbr = nullptr; float* nbr_tau = nullptr; int cell_cap = 0;   // compact cell ids + 27-neighbour table (thread-per-query kNN)   // kNN work units (cell, first query) and slow-path query list
    float4* Pc = nullptr;              // points in cell order, w = filtered index (int bits)
    float4* Nrm = nullptr;             // normals of the filtered cloud
    int* knn_full = nullptr;           // kNN indices of the filtered cloud (Ncap * ne_knn), canonical order
    unsigned long long* vtab = nullptr;  // voxel set: (epoch<<32 | voxel key)
    unsigned* vkeys = nullptr;         // distinct voxel keys, sorted
    int* vid = nullptr;                // voxel id per filtered point
    int* vox_cnt = nullptr; int* vox_off = nullptr; int* vox_fill = nullptr;
    int* blk_hist = nullptr;          // [block of 256 points][voxel] counts / prefixes of the stable partition by voxel (nullptr: arena too large, atomic scatter + per-voxel sort)
    unsigned* rk0 = nullptr; unsigned* rk1 = nullptr; int* rv0 = nullptr; int* rv1 = nullptr;
    // M-scale
    float4* pts = nullptr; float4* nrms = nullptr;
    float4* pset = nullptr; float4* pset2 = nullptr; float4* vset = nullptr; float4* vnew = nullptr;
    float* vvar = nullptr;
    unsigned* simi_bits = nullptr;     // Mcap x W bit matrix of the similarity graph
    float4* nraw = nullptr; float4* vnd = nullptr;   // hoisted per-point normal quantities (drosa.cu)
    float4* vset_it = nullptr; float4* vnew_it = nullptr; int* fused_flags = nullptr; long long* fused_dbg = nullptr; long long* fused_tl = nullptr;   // fused drosa (drosa.cu): per-iteration state + flags
    int* surf = nullptr;               // Mcap * nbk
    int* simi_off = nullptr; int* simi_idx = nullptr;
    int* level = nullptr; int* level_off = nullptr; int* level_pts = nullptr;
    int* good = nullptr; float4* centers = nullptr; int* good_rank = nullptr; int* poor_idx = nullptr;
    int* active_size = nullptr;
    float4* pcloud = nullptr;          // compacted finite pset (vertex_sampling)
    unsigned* adj_bits = nullptr;      // Mcap * W cover adjacency
    int* seeds = nullptr; int* seed_vid = nullptr; int* corresp = nullptr;
    float4* skel = nullptr; float4* skel2 = nullptr; float4* skel_sampled = nullptr; float4* vtmp = nullptr;
    int* vs_off = nullptr; int* vs_idx = nullptr;   // vertex_smooth CSR (reuses Scap sizing)
    float4* vset_iters = nullptr;      // debug tap: niter_drosa * Mcap
};

struct osep_rosa_impl {
    osep_rosa_config cfg;
    int device = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;     // side stream: independent branches of one frame overlap (kNN normals || leaf passes; surf kNN || similarity)
    cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
    cudaStream_t stream3 = nullptr;     // third branch: surf kNN + smoothing schedule need only the downsampled POSITIONS, so they run while the kNN normals branch is still busy
    cudaEvent_t ev_pts = nullptr, ev_join3 = nullptr;
    bool overlap = true;                // disabled while per-stage profiling is on (stage times stay sequential and attributable)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    RosaDev d;
    bool debug = false;
    bool fused = true;                 // drosa as one persistent cooperative kernel (OSEP_DROSA_UNFUSED=1 disables)
    int knn_ctas_per_sm = 0;           // grid of the persistent kNN kernel in CTAs per SM; 0 = by frame size (OSEP_KNN_CTAS)
    float knn_target = 29.0f;          // thread-per-query kNN: points the pilot radius of a cell holds (OSEP_KNN_TARGET); the list window is [k, 40)
    bool knn_warp_path = false;        // OSEP_KNN_WARP=1: the warp-per-cell kernels also for K <= 24 (A/B)
    int knn_ppc = 0;                   // target points per kNN grid cell; 0 = by frame size: 9, or 6 above 400k points (OSEP_KNN_PPC)
    int hot_ctas = 1 << 20;                 // fused drosa: worker CTAs serving the ready queue after the iteration-0 tasks (OSEP_HOT_CTAS)
    int spin_ns = 20;                  // fused drosa: sleep between two polls of an idle worker warp (OSEP_SPIN_NS)
    int sched_lag = 3;                 // fused drosa schedule model: walker steps from G(k-1,j) to the staged O(k,j) (OSEP_DROSA_LAG)
    bool fused_detail = false;         // per-step cycle counters in the chain walkers (OSEP_FUSED_DETAIL=1; costs ~4 % of the stage)
    int fused_grid = 0;                // CTAs of the fused kernel (0 = one per SM)
    int fused_smem = 0;                // dynamic shared memory of the fused kernel in bytes; 0 = by mode (OSEP_FUSED_SMEM_KB)
    bool vs_separate = false;          // vertex_sampling as its own five launches instead of phases of the dcrosa cluster kernel (OSEP_VS_SEPARATE=1; A/B and tests)
    bool dc16_ok = false;              // the device can place a 16-CTA cluster of k_dcrosa_cluster
    int dc_cluster = 0;                // CTAs of the dcrosa cluster: 16 (non-portable size) for single frames when the device can place it, else 8 (OSEP_DC_CLUSTER)
    int simi_side = -1;                // similarity list fill on the side stream next to the fused kernel: 1/0, -1 = only when one frame runs at a time (OSEP_SIMI_SIDE)
    bool profile = false;              // record a CUDA event after every stage group (osep_rosa_set_profile)
    static constexpr int NSTAGE_EV = 16;
    cudaEvent_t stage_ev[NSTAGE_EV] = {};
    long long n_launches = 0;          // kernels launched by this handle since create
    long long n_graph_replays = 0;     // frames that ran as one cudaGraphLaunch
    // CUDA graph of one frame: a frame has no host decisions (osep_common.cuh), so the whole launch sequence of a
    // frame is captured on the second frame and replayed afterwards; n / stride / buffer travel through FrameParams
    struct FrameGraph {
        cudaGraphExec_t exec = nullptr;
        int big = -1, epoch = -1;          // key: size class (launch heuristics) and configuration epoch -- NOT n, stride or the buffer
        int seen_big = -1, seen_epoch = -1;
        long long launches = 0;
    } fg;
    bool use_graphs = true;            // OSEP_NO_GRAPH=1 disables; also cleared when capture is not supported
    int cfg_epoch = 0;                 // bumped by every setter that changes what a frame launches
    // host mirrors
    unsigned char* exp_host = nullptr; // pinned result block: st_host and skel_host point into it
    FrameState* st_host = nullptr;
    FrameParams* prm_host = nullptr;   // pinned; source of the copy node at the head of a frame
    float4* skel_host = nullptr;       // first Mcap/4 vertices
    std::vector<float4> skel_big;      // frames with more vertices than the export block carries
    std::vector<float> vertices_colmajor;
    int last_n = 0;
    // graph outputs (osep_rosa_graph)
    struct GraphHost* graph = nullptr;
    bool auto_graph = false; float graph_fuse = 3.0f;   // osep_rosa_set_graph: graph_adj + mst run inside every frame
    bool graph_in_frame = false;                       // the last enqueued frame carried its graph
    // batch
    std::vector<osep_rosa_impl*> lanes;
    std::vector<std::vector<float>> batch_vertices;
};

// stage-group boundaries for the profile tap ("stage_ms"), in launch order
enum StageMark { SM_FILTER = 0, SM_LEAF, SM_GRID, SM_KNN, SM_VOXEL, SM_SIMI, SM_SURF, SM_DROSA_ORIENT, SM_DROSA_SMOOTH, SM_POSITION,
                 SM_DCROSA, SM_VSAMPLE, SM_VSMOOTH, SM_COUNT };
inline void stage_mark(osep_rosa_impl* h, int which) {
    if (h->profile && h->stage_ev[which]) cudaEventRecord(h->stage_ev[which], h->stream);
}

// stage launchers (all asynchronous on h->stream, no host sync)
void launch_preprocess(osep_rosa_impl* h, int n_class);                               // R3-R11 (n_class: point count the launch heuristics are tuned for)
void launch_mscale(osep_rosa_impl* h);                                                // R12-R22

}  // namespace osep
