// csrc/api.cu -- C ABI of libosep_skel.so (include/osep_skel.h): handle management, the
// host-buffer and device-buffer entry points of Rosa::rosa_run, debug taps, batched frames.
#include "rosa_impl.cuh"
#include <cstdio>
#include <cstring>
#include <cstdarg>
#include <cstdlib>
#include <string>
#include <algorithm>

namespace osep {

void mscale_set_attributes(int Mcap);
bool dcrosa_cluster16_ok();
bool drosa_fused_supported(int sms);
struct GraphHost;
GraphHost* graph_create(int cap, int device, cudaStream_t stream);
void graph_destroy(GraphHost* g);
int graph_run_device(GraphHost* g, const float4* pos_dev, int G, float fuse_dist_th, osep_graph_view* out);
int graph_collect_in_frame(GraphHost* g, osep_graph_view* out);
size_t frame_tail_enqueue(osep_rosa_impl* h, GraphHost* g, float fuse_dist_th, cudaStream_t s);
size_t frame_export_bytes(int Mq);
int select_box_enqueue(const float* in, int n, int stride, const float* lo3, const float* hi3, float* out, int out_stride, int out_cap,
                       int* blk_cnt, int* total_dev, cudaStream_t s);
void launch_leaf_only(osep_rosa_impl* h);
void graph_set_mirrors(GraphHost* g, unsigned char* exp_host, int Mq);


static thread_local std::string g_err;
static void set_err(const char* fmt, ...) {
    char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_err = buf;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return OSEP_ERR_CUDA; } } while (0)

template <class T> static cudaError_t dalloc(T** p, size_t n) { return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)); }

static int rosa_alloc(osep_rosa_impl* h, int capacity_points) {
    RosaDev& d = h->d;
    const osep_rosa_config& c = h->cfg;
    d.Ncap = std::max(capacity_points, 1024);
    d.Mcap = 3072;
    d.W = d.Mcap / 32;
    d.Scap = d.Mcap * 128;
    d.T = 1; while (d.T < 2 * d.Ncap) d.T <<= 1;
    d.knn = c.ne_knn; d.nbk = c.nb_knn;
    const size_t N = d.Ncap, M = d.Mcap, T = d.T;
    const size_t tmpcap = std::max<size_t>(N, d.Scap);
    CK(dalloc(&d.st, 1)); CK(cudaMemset(d.st, 0, sizeof(FrameState)));
    CK(dalloc(&d.prm, 1)); CK(cudaMemset(d.prm, 0, sizeof(FrameParams)));
    CK(dalloc(&d.in, N * 4));
    CK(dalloc(&d.P, N)); CK(dalloc(&d.filt_idx, N)); CK(dalloc(&d.blk_cnt, N / 256 + 2));
    { unsigned* g4; CK(dalloc(&g4, 4 * T)); d.gtab = (unsigned long long*)g4; d.g_cnt = (int*)(g4 + 2 * T); d.g_fill = (int*)(g4 + 3 * T); }
    CK(dalloc(&d.g_start, T)); CK(dalloc(&d.units, N + 8)); CK(dalloc(&d.slow_list, N)); CK(dalloc(&d.over_list, N + 8)); CK(dalloc(&d.knn_retry, N + 8)); CK(dalloc(&d.knn_grow, N + 8));
    d.cell_cap = (int)(N / 3 + 1024); CK(dalloc(&d.g_cell, T)); CK(dalloc(&d.cell_slots, (size_t)d.cell_cap)); CK(dalloc(&d.pc_cell, N)); CK(dalloc(&d.nbr, (size_t)d.cell_cap * 27)); CK(dalloc(&d.nbr_tau, (size_t)d.cell_cap)); CK(dalloc(&d.knn_full, N * (size_t)d.knn)); CK(dalloc(&d.p_slot, N)); CK(dalloc(&d.Pc, N)); CK(dalloc(&d.Nrm, N));
    CK(dalloc(&d.vtab, T)); CK(cudaMemset(d.vtab, 0, T * sizeof(unsigned long long)));
    CK(dalloc(&d.vkeys, M)); CK(dalloc(&d.vid, N)); CK(dalloc(&d.vox_cnt, M + 2)); CK(dalloc(&d.vox_fill, M + 2));
    if ((size_t)(N / 256 + 2) * M * sizeof(int) <= ((size_t)32 << 20)) CK(dalloc(&d.blk_hist, (size_t)(N / 256 + 2) * M));      // frames: stable partition by voxel (nscale.cu); larger arenas keep the atomic scatter + per-voxel sort
    CK(dalloc(&d.rk0, tmpcap)); CK(dalloc(&d.rk1, tmpcap)); CK(dalloc(&d.rv0, tmpcap)); CK(dalloc(&d.rv1, tmpcap));
    CK(dalloc(&d.pts, M)); CK(dalloc(&d.nrms, M)); CK(dalloc(&d.pset, M)); CK(dalloc(&d.pset2, M)); CK(dalloc(&d.vset, M)); CK(dalloc(&d.vnew, M));
    CK(dalloc(&d.vvar, M)); CK(dalloc(&d.simi_bits, M * (size_t)d.W)); CK(dalloc(&d.nraw, M)); CK(dalloc(&d.vnd, M));
    CK(dalloc(&d.vset_it, M * 9)); CK(dalloc(&d.vnew_it, M * 8)); CK(dalloc(&d.fused_flags, M * 16 + 16)); CK(dalloc(&d.fused_dbg, 128)); CK(cudaMemset(d.fused_dbg, 0, 128 * 8));
    { const char* e = getenv("OSEP_KNN_PPC"); if (e && atoi(e) > 0) h->knn_ppc = atoi(e); }
    { const char* e = getenv("OSEP_KNN_TARGET"); if (e && atof(e) > 0) h->knn_target = (float)atof(e); }
    { const char* e = getenv("OSEP_KNN_WARP"); if (e && e[0] == '1') h->knn_warp_path = true; }
    { const char* e = getenv("OSEP_KNN_CTAS"); if (e && atoi(e) > 0) h->knn_ctas_per_sm = atoi(e); }
    { const char* e = getenv("OSEP_DROSA_UNFUSED"); if (e && e[0] == '1') h->fused = false;
      const char* e3 = getenv("OSEP_NO_GRAPH"); if (e3 && e3[0] == '1') h->use_graphs = false;
      const char* e4 = getenv("OSEP_FUSED_DETAIL"); if (e4 && e4[0] == '1') { h->fused_detail = true; CK(dalloc(&d.fused_tl, M * 9 * 16)); CK(cudaMemset(d.fused_tl, 0, M * 9 * 16 * 8)); }
      const char* e6 = getenv("OSEP_FUSED_GRID"); if (e6 && atoi(e6) > 0) h->fused_grid = atoi(e6);
      const char* e8 = getenv("OSEP_FUSED_SMEM_KB"); if (e8 && atoi(e8) > 0) h->fused_smem = atoi(e8) * 1024;
      const char* e11 = getenv("OSEP_VS_SEPARATE"); if (e11 && e11[0] == '1') h->vs_separate = true;
      const char* e10 = getenv("OSEP_DC_CLUSTER"); if (e10 && (atoi(e10) == 8 || atoi(e10) == 16)) h->dc_cluster = atoi(e10);
      const char* e9 = getenv("OSEP_SIMI_SIDE"); if (e9 && (e9[0] == '0' || e9[0] == '1')) h->simi_side = e9[0] - '0';
      const char* e7 = getenv("OSEP_HOT_CTAS"); if (e7 && atoi(e7) > 0) h->hot_ctas = atoi(e7);
      const char* e5 = getenv("OSEP_SPIN_NS"); if (e5 && atoi(e5) >= 0) h->spin_ns = atoi(e5);
      const char* e2 = getenv("OSEP_DROSA_LAG"); if (e2 && atoi(e2) >= 0) h->sched_lag = atoi(e2);
      int coop = 0; cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device); if (!coop) h->fused = false; }
    { int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device)); h->sms = sms > 0 ? sms : 148; } CK(dalloc(&d.surf, M * (size_t)std::max(d.nbk, 1)));
    CK(dalloc(&d.simi_off, M + 2)); CK(dalloc(&d.simi_idx, (size_t)d.Scap));
    CK(dalloc(&d.level, M)); CK(dalloc(&d.level_off, M + 2)); CK(dalloc(&d.level_pts, M));
    CK(dalloc(&d.good, M)); CK(dalloc(&d.centers, M)); CK(dalloc(&d.good_rank, M + 2)); CK(dalloc(&d.poor_idx, M));
    CK(dalloc(&d.active_size, M)); CK(dalloc(&d.pcloud, M)); CK(dalloc(&d.adj_bits, M * (size_t)d.W));
    CK(dalloc(&d.seeds, M)); CK(dalloc(&d.corresp, M));
    CK(dalloc(&d.skel, M)); CK(dalloc(&d.skel2, M)); CK(dalloc(&d.skel_sampled, M)); CK(dalloc(&d.vtmp, M));
    CK(dalloc(&d.vs_off, M + 2)); CK(dalloc(&d.vs_idx, (size_t)d.Scap));
    // one pinned result block per frame: frame state | first Mcap/4 vertices | in-frame graph lists (graph.cu k_export)
    CK(cudaMalloc((void**)&d.exp, frame_export_bytes(d.Mcap / 4)));
    CK(cudaMallocHost((void**)&h->exp_host, frame_export_bytes(d.Mcap / 4)));
    memset(h->exp_host, 0, frame_export_bytes(d.Mcap / 4));
    h->st_host = reinterpret_cast<FrameState*>(h->exp_host);
    h->skel_host = reinterpret_cast<float4*>(h->exp_host + 512);
    CK(cudaMallocHost((void**)&h->prm_host, sizeof(FrameParams)));
    memset(h->st_host, 0, sizeof(FrameState));
    mscale_set_attributes(d.Mcap);
    h->dc16_ok = dcrosa_cluster16_ok();
    if (h->fused && !drosa_fused_supported(h->sms)) h->fused = false;     // falls back to the per-iteration kernels (same results)
    return OSEP_OK;
}

static void rosa_free(osep_rosa_impl* h) {
    RosaDev& d = h->d;
    void* ptrs[] = {d.st, d.prm, d.in, d.in_raw, d.P, d.filt_idx, d.blk_cnt, d.gtab, d.g_start, d.units, d.slow_list, d.over_list, d.knn_retry, d.knn_grow, d.g_cell, d.cell_slots, d.pc_cell, d.nbr, d.nbr_tau, d.p_slot, d.Pc, d.Nrm, d.knn_full, d.vtab, d.vkeys, d.vid,
                    d.vox_cnt, d.vox_fill, d.blk_hist, d.rk0, d.rk1, d.rv0, d.rv1, d.pts, d.nrms, d.pset, d.pset2, d.vset, d.vnew, d.vvar, d.surf,
                    d.simi_off, d.simi_idx, d.simi_bits, d.nraw, d.vnd, d.vset_it, d.vnew_it, d.fused_flags, d.fused_dbg, d.fused_tl, d.level, d.level_off, d.level_pts, d.good, d.centers, d.good_rank, d.poor_idx, d.active_size,
                    d.pcloud, d.adj_bits, d.seeds, d.corresp, d.skel, d.skel2, d.skel_sampled, d.vtmp, d.vs_off, d.vs_idx, d.vset_iters};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->graph) graph_destroy(h->graph);
    if (h->exp_host) cudaFreeHost(h->exp_host);
    if (d.exp) cudaFree(d.exp);
    if (h->prm_host) cudaFreeHost(h->prm_host);
    for (int i = 0; i < osep_rosa_impl::NSTAGE_EV; ++i) if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (int i = 0; i < 2; ++i) { if (h->ev_fork[i]) cudaEventDestroy(h->ev_fork[i]); if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]); }
    if (h->fg.exec) cudaGraphExecDestroy(h->fg.exec);
    if (h->ev_pts) cudaEventDestroy(h->ev_pts);
    if (h->ev_join3) cudaEventDestroy(h->ev_join3);
    if (h->stream3) cudaStreamDestroy(h->stream3);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream) cudaStreamDestroy(h->stream);
}

static int rosa_enqueue(osep_rosa_impl* h, const float* xyz_dev, int n, int stride) {
    RosaDev& d = h->d;
    h->last_n = n;
    h->graph_in_frame = h->auto_graph && h->graph != nullptr;
    // the previous frame of this handle has been collected (one frame in flight per handle), so the pinned mirror is free
    h->prm_host->in = xyz_dev; h->prm_host->n = n; h->prm_host->stride = stride; h->prm_host->pad = 0;
    const int big = n > 400000 ? 1 : 0;              // the only host-side launch heuristics: kNN grid occupancy / CTAs per SM
    cudaEventRecord(h->ev0, h->stream);
    auto enqueue_all = [&]() {
        cudaMemcpyAsync(d.prm, h->prm_host, sizeof(FrameParams), cudaMemcpyHostToDevice, h->stream);    // a copy NODE in the graph: re-read at every replay
        launch_preprocess(h, big ? 400001 : 1);
        launch_mscale(h);
        const bool wg = h->auto_graph && h->graph;
        // R22 vertex_smooth, the skeleton graph (graph_adj + mst) when it is on, and the result block: one single-CTA launch
        const size_t nbytes = frame_tail_enqueue(h, wg ? h->graph : nullptr, h->graph_fuse, h->stream); h->n_launches += 1;
        stage_mark(h, SM_VSMOOTH);
        cudaMemcpyAsync(h->exp_host, d.exp, nbytes, cudaMemcpyDeviceToHost, h->stream);       // the frame's only device-to-host copy (larger results on demand)
    };
    bool done = false;
    if (h->use_graphs && !h->profile && !h->debug) {
        osep_rosa_impl::FrameGraph& g = h->fg;
        if (g.exec && g.big == big && g.epoch == h->cfg_epoch) {
            if (cudaGraphLaunch(g.exec, h->stream) == cudaSuccess) { h->n_launches += g.launches; ++h->n_graph_replays; done = true; }
            else cudaGetLastError();
        } else if (g.seen_big == big && g.seen_epoch == h->cfg_epoch) {
            if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
            const long long l0 = h->n_launches;
            if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                enqueue_all();
                cudaGraph_t graph = nullptr;
                if (cudaStreamEndCapture(h->stream, &graph) == cudaSuccess && graph) {
                    if (cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) g.exec = nullptr;
                    cudaGraphDestroy(graph);
                }
                g.launches = h->n_launches - l0; h->n_launches = l0;
                if (!g.exec) { cudaGetLastError(); h->use_graphs = false; }          // capture unsupported here: plain launches from now on
                else {
                    g.big = big; g.epoch = h->cfg_epoch;
                    if (cudaGraphLaunch(g.exec, h->stream) == cudaSuccess) { h->n_launches += g.launches; ++h->n_graph_replays; done = true; }
                    else cudaGetLastError();
                }
            } else { cudaGetLastError(); h->use_graphs = false; }
        }
        g.seen_big = big; g.seen_epoch = h->cfg_epoch;
    }
    if (!done) enqueue_all();
    cudaEventRecord(h->ev1, h->stream);
    CK(cudaGetLastError());
    return OSEP_OK;
}

static int rosa_collect(osep_rosa_impl* h, osep_rosa_result* res) {
    CK(cudaStreamSynchronize(h->stream));
    const FrameState& s = *h->st_host;
    RosaDev& d = h->d;
    int V = (s.status == 0) ? s.V : 0;
    const float4* sk = h->skel_host;
    if (V > d.Mcap / 4) {                         // more vertices than the export block carries: fetch them all
        h->skel_big.resize(V);
        CK(cudaMemcpy(h->skel_big.data(), d.skel, sizeof(float4) * V, cudaMemcpyDeviceToHost));
        sk = h->skel_big.data();
    }
    h->vertices_colmajor.resize((size_t)3 * V);
    for (int i = 0; i < V; ++i) {
        h->vertices_colmajor[i] = sk[i].x;
        h->vertices_colmajor[(size_t)V + i] = sk[i].y;
        h->vertices_colmajor[(size_t)2 * V + i] = sk[i].z;
    }
    if (res) {
        res->status = s.status; res->flags = s.flags; res->n_input = h->last_n; res->n_filtered = s.n_filt;
        res->n_downsampled = s.M; res->n_vertices = V; res->n_passes = s.n_pass; res->n_simi = s.S;
        res->leaf_size = s.leaf_size; res->leaf_used = s.leaf_used;
        float ms = 0.f; cudaEventElapsedTime(&ms, h->ev0, h->ev1); res->gpu_ms = ms;
    }
    if (s.status < 0) {
        // device-side capacity limits (the reference has none): say which one instead of leaving osep_last_error() stale
        set_err("frame failed on the device with status %d: N'=%d M=%d (arena %d) S=%d (arena %d) passes=%d flags=0x%x%s%s", s.status, s.n_filt, s.M, d.Mcap, s.S,
                d.Scap, s.n_pass, s.flags, s.pass_overflow ? "; VoxelGrid index overflow (PCL would return the input cloud unchanged)" : "",
                s.n_keys > d.Mcap ? "; more voxels than the downsampled-point arena holds" : "");
    }
    return s.status;
}

}  // namespace osep

using namespace osep;

struct osep_rosa { osep_rosa_impl impl; };

extern "C" {

const char* osep_last_error(void) { return g_err.c_str(); }
int osep_abi_version(void) { return OSEP_ABI_VERSION; }
int osep_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

void osep_rosa_default_config(osep_rosa_config* c) {
    if (!c) return;
    c->max_points = 1000; c->min_points = 100; c->pts_dist_lim = 50.0f; c->ne_knn = 20; c->nb_knn = 10; c->max_proj_range = 10.0f;
    c->niter_drosa = 6; c->niter_dcrosa = 5; c->niter_smooth = 3; c->alpha_recenter = 0.3f; c->radius_smooth = 5.0f;
}
void osep_gskel_default_config(osep_gskel_config* c) {
    if (!c) return;
    c->gnd_th = 60.0f; c->fuse_dist_th = 3.0f; c->fuse_conf_th = 0.1f; c->lkf_pn = 1e-4f; c->lkf_mn = 0.1f; c->max_obs_wo_conf = 5;
    c->niter_smooth_vertex = 3; c->vertex_smooth_coef = 0.3f; c->min_branch_length = 5;
}

int osep_rosa_create(const osep_rosa_config* cfg, int device, int capacity_points, osep_rosa** out) {
    if (!out) { set_err("osep_rosa_create: out is null"); return OSEP_ERR_ARG; }
    *out = nullptr;
    osep_rosa_config c; osep_rosa_default_config(&c);
    if (cfg) c = *cfg;
    if (c.ne_knn < 1 || c.ne_knn > KNN_MAX || c.nb_knn < 1 || c.nb_knn > 32) { set_err("ne_knn must be in [1,%d], nb_knn in [1,32]", KNN_MAX); return OSEP_ERR_ARG; }
    if (c.max_points < 1 || c.max_points > 1536) { set_err("max_points must be in [1,1536] (device arena holds 3072 downsampled points)"); return OSEP_ERR_ARG; }
    if (capacity_points < 1) { set_err("capacity_points must be positive"); return OSEP_ERR_ARG; }
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_err("device %d not present (%d CUDA devices)", device, ndev); return OSEP_ERR_CUDA; }
    CK(cudaSetDevice(device));
    osep_rosa* h = new osep_rosa();
    h->impl.cfg = c; h->impl.device = device;
    int rc = OSEP_OK;
    do {
        if (cudaStreamCreateWithFlags(&h->impl.stream, cudaStreamNonBlocking) != cudaSuccess) { set_err("cudaStreamCreate failed"); rc = OSEP_ERR_CUDA; break; }
        cudaEventCreate(&h->impl.ev0); cudaEventCreate(&h->impl.ev1);
        if (cudaStreamCreateWithFlags(&h->impl.stream2, cudaStreamNonBlocking) != cudaSuccess) { set_err("cudaStreamCreate failed"); rc = OSEP_ERR_CUDA; break; }
        for (int i = 0; i < 2; ++i) { cudaEventCreateWithFlags(&h->impl.ev_fork[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&h->impl.ev_join[i], cudaEventDisableTiming); }
        if (cudaStreamCreateWithFlags(&h->impl.stream3, cudaStreamNonBlocking) != cudaSuccess) { set_err("cudaStreamCreate failed"); rc = OSEP_ERR_CUDA; break; }
        cudaEventCreateWithFlags(&h->impl.ev_pts, cudaEventDisableTiming); cudaEventCreateWithFlags(&h->impl.ev_join3, cudaEventDisableTiming);
        rc = rosa_alloc(&h->impl, capacity_points);
    } while (0);
    if (rc != OSEP_OK) { rosa_free(&h->impl); delete h; return rc; }
    *out = h;
    return OSEP_OK;
}

int osep_rosa_destroy(osep_rosa* h) {
    if (!h) return OSEP_ERR_ARG;
    cudaSetDevice(h->impl.device);
    for (osep_rosa_impl* l : h->impl.lanes) { rosa_free(l); delete reinterpret_cast<osep_rosa*>(l); }
    rosa_free(&h->impl);
    delete h;
    return OSEP_OK;
}

void* osep_rosa_stream(osep_rosa* h) { return h ? (void*)h->impl.stream : nullptr; }

int osep_rosa_set_debug(osep_rosa* h, int enable) {
    if (!h) return OSEP_ERR_ARG;
    osep_rosa_impl& I = h->impl;
    CK(cudaSetDevice(I.device));
    I.debug = enable != 0; ++I.cfg_epoch;
    if (I.debug && !I.d.vset_iters) {
        CK(dalloc(&I.d.vset_iters, (size_t)I.d.Mcap * std::max(1, I.cfg.niter_drosa)));
    }
    return OSEP_OK;
}

int osep_rosa_set_profile(osep_rosa* h, int enable) {
    if (!h) return OSEP_ERR_ARG;
    osep_rosa_impl& I = h->impl;
    CK(cudaSetDevice(I.device));
    I.profile = enable != 0; ++I.cfg_epoch;
    I.overlap = !I.profile;
    if (I.profile) for (int i = 0; i < SM_COUNT; ++i) if (!I.stage_ev[i]) CK(cudaEventCreate(&I.stage_ev[i]));
    return OSEP_OK;
}

static int check_frame(osep_rosa_impl& I, int n, int stride) {
    if (n < 0 || stride < 3) { set_err("bad frame: n=%d stride=%d", n, stride); return OSEP_ERR_ARG; }
    if (n > I.d.Ncap) { set_err("frame of %d points exceeds the handle capacity %d", n, I.d.Ncap); return OSEP_ERR_CAPACITY; }
    return OSEP_OK;
}

int osep_rosa_run_device(osep_rosa* h, const float* xyz_dev, int n, int stride_floats) {
    if (!h) return OSEP_ERR_ARG;
    osep_rosa_impl& I = h->impl;
    int rc = check_frame(I, n, stride_floats); if (rc) return rc;
    CK(cudaSetDevice(I.device));
    if (n == 0) {                       // rosa.cpp:63-65: empty input fails preprocess
        CK(cudaStreamSynchronize(I.stream));   // a frame still in flight on this handle writes st_host when it completes
        memset(I.st_host, 0, sizeof(FrameState)); I.st_host->status = OSEP_STAGE_1; I.last_n = 0;
        cudaEventRecord(I.ev0, I.stream); cudaEventRecord(I.ev1, I.stream);
        return OSEP_OK;
    }
    return rosa_enqueue(&I, xyz_dev, n, stride_floats);
}

int osep_rosa_finish(osep_rosa* h, osep_rosa_result* res) {
    if (!h) return OSEP_ERR_ARG;
    CK(cudaSetDevice(h->impl.device));
    return rosa_collect(&h->impl, res);
}

int osep_rosa_run(osep_rosa* h, const float* xyz_host, int n, int stride_floats, osep_rosa_result* res) {
    if (!h) return OSEP_ERR_ARG;
    osep_rosa_impl& I = h->impl;
    int rc = check_frame(I, n, stride_floats);
    if (rc) { if (res) { memset(res, 0, sizeof(*res)); res->status = rc; } return rc; }
    CK(cudaSetDevice(I.device));
    int dev_stride = stride_floats;
    if (n > 0) {
        if (stride_floats > 4) {        // wide records (intensity, ring, time ...): only x, y, z travel, packed into 16-byte records
            CK(cudaMemcpy2DAsync(I.d.in, 16, xyz_host, sizeof(float) * (size_t)stride_floats, 12, (size_t)n, cudaMemcpyHostToDevice, I.stream));
            dev_stride = 4;
        } else {
            CK(cudaMemcpyAsync(I.d.in, xyz_host, sizeof(float) * (size_t)n * stride_floats, cudaMemcpyHostToDevice, I.stream));
        }
    }
    rc = osep_rosa_run_device(h, I.d.in, n, dev_stride);
    if (rc) { if (res) { memset(res, 0, sizeof(*res)); res->status = rc; } return rc; }
    return rosa_collect(&I, res);
}

int osep_rosa_run_pointcloud2(osep_rosa* h, const void* data_host, int n_points, int point_step, int off_x, int off_y, int off_z,
                              osep_rosa_result* res) {
    if (!h) return OSEP_ERR_ARG;
    osep_rosa_impl& I = h->impl;
    auto fail = [&](int rc) { if (res) { memset(res, 0, sizeof(*res)); res->status = rc; } return rc; };
    if (n_points < 0 || point_step < 12 || (point_step & 3) || off_x < 0 || (off_x & 3) || off_y != off_x + 4 || off_z != off_x + 8 || off_z + 4 > point_step) {
        set_err("PointCloud2 layout not supported: point_step=%d offsets x=%d y=%d z=%d (need consecutive 4-byte aligned float32 x,y,z)", point_step, off_x, off_y, off_z);
        return fail(OSEP_ERR_ARG);
    }
    if (n_points > 0 && !data_host) return fail(OSEP_ERR_ARG);
    int rc = check_frame(I, n_points, point_step / 4);
    if (rc) return fail(rc);
    CK(cudaSetDevice(I.device));
    const size_t bytes = (size_t)n_points * point_step;
    if (bytes > I.d.in_raw_bytes) {                      // sized for the widest record seen so far at full capacity
        CK(cudaStreamSynchronize(I.stream));
        if (I.d.in_raw) cudaFree(I.d.in_raw);
        I.d.in_raw = nullptr; I.d.in_raw_bytes = 0;
        const size_t want = (size_t)I.d.Ncap * point_step;
        if (cudaMalloc((void**)&I.d.in_raw, want) != cudaSuccess) { cudaGetLastError(); set_err("cudaMalloc of %zu staging bytes failed", want); return fail(OSEP_ERR_CUDA); }
        I.d.in_raw_bytes = want;
    }
    if (n_points > 0) CK(cudaMemcpyAsync(I.d.in_raw, data_host, bytes, cudaMemcpyHostToDevice, I.stream));
    rc = osep_rosa_run_device(h, reinterpret_cast<const float*>(I.d.in_raw + off_x), n_points, point_step / 4);
    if (rc) return fail(rc);
    return rosa_collect(&I, res);
}

int osep_rosa_vertices(osep_rosa* h, const float** data, int* n_vertices) {
    if (!h || !data || !n_vertices) return OSEP_ERR_ARG;
    *n_vertices = (int)(h->impl.vertices_colmajor.size() / 3);
    *data = h->impl.vertices_colmajor.data();
    return OSEP_OK;
}

// ---- taps
namespace {
long tap_copy(void* dst, long cap, const void* dev, size_t count, size_t elem) {
    if (dst && cap >= (long)(count * elem) && count) {
        if (cudaMemcpy(dst, dev, count * elem, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    }
    return (long)count;
}
long tap_f4_as_f3(void* dst, long cap, const float4* dev, size_t rows) {
    if (dst && cap >= (long)(rows * 12) && rows) {
        std::vector<float4> tmp(rows);
        if (cudaMemcpy(tmp.data(), dev, rows * 16, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
        float* o = (float*)dst;
        for (size_t i = 0; i < rows; ++i) { o[3 * i] = tmp[i].x; o[3 * i + 1] = tmp[i].y; o[3 * i + 2] = tmp[i].z; }
    }
    return (long)rows * 3;
}
}  // namespace

long osep_rosa_get(osep_rosa* h, const char* name, void* dst, long cap) {
    if (!h || !name) return -1;
    osep_rosa_impl& I = h->impl;
    if (cudaSetDevice(I.device) != cudaSuccess) return -2;
    cudaStreamSynchronize(I.stream);
    RosaDev& d = I.d;
    const FrameState& s = *I.st_host;
    const std::string nm(name);
    const size_t N = (size_t)s.n_filt, M = (size_t)s.M, V = (size_t)((s.status == 0 || s.status == OSEP_STAGE_7) ? s.V : 0);
    if (nm == "filtered") return tap_f4_as_f3(dst, cap, d.P, N);
    if (nm == "filt_idx") return tap_copy(dst, cap, d.filt_idx, N, 4);
    if (nm == "normals_full") return tap_f4_as_f3(dst, cap, d.Nrm, N);
    if (nm == "knn_full") { if (!d.knn_full) return 0; return tap_copy(dst, cap, d.knn_full, N * (size_t)std::min<int>(d.knn, (int)N), 4); }
    if (nm == "leaf_hist") { if (dst && cap >= (long)(4 * s.n_pass)) memcpy(dst, s.leaf_hist, 4 * s.n_pass); return s.n_pass; }
    if (nm == "nvox_hist") { if (dst && cap >= (long)(4 * s.n_pass)) memcpy(dst, s.nvox_hist, 4 * s.n_pass); return s.n_pass; }
    if (nm == "vox_keys") return tap_copy(dst, cap, d.vkeys, s.vox_identity ? 0 : M, 4);
    if (nm == "vox_count") {
        if (s.vox_identity) return 0;          // output = input: no voxels were formed (PCL returns before its sort)
        std::vector<int> off(M + 1);
        if (M && cudaMemcpy(off.data(), d.vox_cnt, (M + 1) * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
        if (dst && cap >= (long)(M * 4)) for (size_t i = 0; i < M; ++i) ((int*)dst)[i] = off[i + 1] - off[i];
        return (long)M;
    }
    if (nm == "pts") return tap_f4_as_f3(dst, cap, d.pts, M);
    if (nm == "nrms") return tap_f4_as_f3(dst, cap, d.nrms, M);
    if (nm == "surf_idx") return tap_copy(dst, cap, d.surf, M * (size_t)d.nbk, 4);
    if (nm == "surf_off") { if (dst && cap >= (long)((M + 1) * 4)) for (size_t i = 0; i <= M; ++i) ((int*)dst)[i] = (int)(i * d.nbk); return (long)M + 1; }
    if (nm == "simi_off") return tap_copy(dst, cap, d.simi_off, M + 1, 4);
    if (nm == "simi_idx") return tap_copy(dst, cap, d.simi_idx, (size_t)s.S, 4);
    if (nm == "vset") return tap_f4_as_f3(dst, cap, (I.fused && I.cfg.niter_drosa <= 8) ? d.vset_it + (size_t)I.cfg.niter_drosa * d.Mcap : d.vset, M);   // fused drosa keeps per-iteration rows
    if (nm == "vset_iters") {
        if (!d.vset_iters) return 0;
        const int nit = I.cfg.niter_drosa;
        if (dst && cap >= (long)(M * 12 * nit)) for (int it = 0; it < nit; ++it) tap_f4_as_f3((char*)dst + (size_t)it * M * 12, (long)(M * 12), d.vset_iters + (size_t)it * d.Mcap, M);
        return (long)(M * 3 * nit);
    }
    if (nm == "vvar") return tap_copy(dst, cap, d.vvar, M, 4);
    if (nm == "pset") return tap_f4_as_f3(dst, cap, d.pset, M);
    if (nm == "pset_drosa") { if (!I.debug) return 0; return tap_f4_as_f3(dst, cap, d.centers, M); }
    if (nm == "poor_idx") return tap_copy(dst, cap, d.poor_idx, (size_t)s.n_poor, 4);
    if (nm == "active_size") return tap_copy(dst, cap, d.active_size, M, 4);
    if (nm == "corresp") return tap_copy(dst, cap, d.corresp, (size_t)s.Mc, 4);
    if (nm == "seeds") return tap_copy(dst, cap, d.seeds, V, 4);
    if (nm == "skelver_sampled") return tap_f4_as_f3(dst, cap, d.skel_sampled, V);
    if (nm == "skelver") return tap_f4_as_f3(dst, cap, d.skel, V);
    if (nm == "levels") return tap_copy(dst, cap, d.level, M, 4);
    if (nm == "stage_ms") {      // device ms between consecutive stage marks of the last frame (needs osep_rosa_set_profile)
        if (!I.profile) return 0;
        if (dst && cap >= (long)(4 * SM_COUNT)) {
            cudaEvent_t prev = I.ev0;
            for (int i = 0; i < SM_COUNT; ++i) { float ms = 0.f; cudaEventElapsedTime(&ms, prev, I.stage_ev[i]); ((float*)dst)[i] = ms; prev = I.stage_ev[i]; }
        }
        return SM_COUNT;
    }
    if (nm == "fused_tl") { if (!d.fused_tl) return 0; return tap_copy(dst, cap, d.fused_tl, (size_t)d.Mcap * 9 * 16, 8) * 2; }
    if (nm == "fused_dbg") return tap_copy(dst, cap, d.fused_dbg, 128, 8) * 2;     // 128 int64 = 256 int32 words
    if (nm == "timeline") { if (dst && cap >= sizeof(s.tstamp)) memcpy(dst, s.tstamp, sizeof(s.tstamp)); return (long long)(sizeof(s.tstamp) / 4); }   // globaltimer ns at the entry of selected kernels (FrameStamp)
    if (nm == "knn_stats") { if (dst && cap >= 20) { int* o = (int*)dst; o[0] = s.n_units; o[1] = s.n_slow; o[2] = s.n_over; o[3] = s.n_retry; o[4] = s.n_grow; } return 5; }
    if (nm == "launches") { if (dst && cap >= 4) *(int*)dst = (int)I.n_launches; return 1; }
    if (nm == "graph_replays") { if (dst && cap >= 4) *(int*)dst = (int)I.n_graph_replays; return 1; }
    return -1;
}

// graph_adj + mst + graph_decomp (gskel.cpp:201-281,539-567) over the last frame's vertices,
// which are still resident on the device (no host round trip of the positions).
int osep_rosa_graph(osep_rosa* h, float fuse_dist_th, osep_graph_view* out) {
    if (!h || !out) return OSEP_ERR_ARG;
    osep_rosa_impl& I = h->impl;
    CK(cudaSetDevice(I.device));
    CK(cudaStreamSynchronize(I.stream));
    if (!I.graph) { I.graph = graph_create(I.d.Mcap, I.device, I.stream); if (!I.graph) { set_err("graph arena allocation failed"); return OSEP_ERR_CUDA; } graph_set_mirrors(I.graph, I.exp_host, I.d.Mcap / 4); }
    if (I.graph_in_frame && fuse_dist_th == I.graph_fuse) return graph_collect_in_frame(I.graph, out);    // built inside the frame (osep_rosa_set_graph): host assembly only
    const int V = (I.st_host->status == 0) ? I.st_host->V : 0;
    return graph_run_device(I.graph, I.d.skel, V, fuse_dist_th, out);
}

// Build the skeleton graph (graph_adj + mst, gskel.cpp:201-281) INSIDE every frame: the two kernels and the result copies join the
// frame's launch sequence / CUDA graph, osep_rosa_graph() then only assembles the host views.  enable = 0 restores the default.
int osep_rosa_set_graph(osep_rosa* h, int enable, float fuse_dist_th) {
    if (!h) return OSEP_ERR_ARG;
    osep_rosa_impl& I = h->impl;
    CK(cudaSetDevice(I.device));
    CK(cudaStreamSynchronize(I.stream));
    if (enable && !I.graph) { I.graph = graph_create(I.d.Mcap, I.device, I.stream); if (!I.graph) { set_err("graph arena allocation failed"); return OSEP_ERR_CUDA; } graph_set_mirrors(I.graph, I.exp_host, I.d.Mcap / 4); }
    I.auto_graph = enable != 0; I.graph_fuse = fuse_dist_th; I.graph_in_frame = false; ++I.cfg_epoch;
    return OSEP_OK;
}

// ---- tiled maps
int osep_points_select_box(int device, const float* pts_dev, int n, int stride_floats, const float* lo3, const float* hi3,
                           float* out_dev, int out_stride, int out_cap, int* count_host, void* stream) {
    if (n < 0 || stride_floats < 3 || !lo3 || !hi3 || (out_stride != 3 && out_stride != 4) || out_cap < 0 || !count_host) { set_err("osep_points_select_box: bad argument"); return OSEP_ERR_ARG; }
    *count_host = 0;
    if (n == 0) return OSEP_OK;
    if (!pts_dev || (!out_dev && out_cap > 0)) return OSEP_ERR_ARG;
    CK(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream;
    int* scratch = nullptr;                           // per-block counts + the total (stream-ordered allocation)
    const size_t nb = (size_t)(n + 255) / 256;
    CK(cudaMallocAsync((void**)&scratch, sizeof(int) * (nb + 2), s));
    int rc = select_box_enqueue(pts_dev, n, stride_floats, lo3, hi3, out_dev, out_stride, out_cap, scratch, scratch + nb + 1, s);
    int total = 0;
    if (rc == OSEP_OK && cudaMemcpyAsync(&total, scratch + nb + 1, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = OSEP_ERR_CUDA;
    cudaFreeAsync(scratch, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) rc = OSEP_ERR_CUDA;
    if (rc != OSEP_OK) { set_err("osep_points_select_box: %s", cudaGetErrorString(cudaGetLastError())); return rc; }
    *count_host = total;
    if (total > out_cap) { set_err("osep_points_select_box: %d matches, room for %d", total, out_cap); return OSEP_ERR_CAPACITY; }
    return OSEP_OK;
}

int osep_points_select_range(int device, const float* pts_dev, int n, int stride_floats, int axis, float lo, float hi,
                             float* out_dev, int out_stride, int out_cap, int* count_host, void* stream) {
    if (axis < 0 || axis > 2) { set_err("osep_points_select_range: bad axis"); return OSEP_ERR_ARG; }
    float lo3[3] = {-3.0e38f, -3.0e38f, -3.0e38f}, hi3[3] = {3.0e38f, 3.0e38f, 3.0e38f};
    lo3[axis] = lo; hi3[axis] = hi;
    return osep_points_select_box(device, pts_dev, n, stride_floats, lo3, hi3, out_dev, out_stride, out_cap, count_host, stream);
}

int osep_rosa_vertices_device(osep_rosa* h, const float** xyzw_dev, int* n_vertices) {
    if (!h || !xyzw_dev || !n_vertices) return OSEP_ERR_ARG;
    osep_rosa_impl& I = h->impl;
    CK(cudaSetDevice(I.device));
    CK(cudaStreamSynchronize(I.stream));
    *xyzw_dev = reinterpret_cast<const float*>(I.d.skel);
    *n_vertices = (I.st_host->status == 0) ? I.st_host->V : 0;
    return OSEP_OK;
}

int osep_rosa_leaf_device(osep_rosa* h, const float* xyz_dev, int n, int stride_floats, float* leaf_size, int* n_filtered) {
    if (!h || !leaf_size) return OSEP_ERR_ARG;
    osep_rosa_impl& I = h->impl;
    int rc = check_frame(I, n, stride_floats); if (rc) return rc;
    CK(cudaSetDevice(I.device));
    *leaf_size = 0.0f; if (n_filtered) *n_filtered = 0;
    if (n == 0) return OSEP_STAGE_1;
    CK(cudaStreamSynchronize(I.stream));
    I.prm_host->in = xyz_dev; I.prm_host->n = n; I.prm_host->stride = stride_floats; I.prm_host->pad = 0;
    CK(cudaMemcpyAsync(I.d.prm, I.prm_host, sizeof(FrameParams), cudaMemcpyHostToDevice, I.stream));
    launch_leaf_only(&I);
    CK(cudaMemcpyAsync(I.st_host, I.d.st, sizeof(FrameState), cudaMemcpyDeviceToHost, I.stream));
    CK(cudaStreamSynchronize(I.stream));
    *leaf_size = I.st_host->leaf_size; if (n_filtered) *n_filtered = I.st_host->n_filt;
    const int status = I.st_host->status;
    I.st_host->status = OSEP_STAGE_1; I.st_host->V = 0;      // not a frame: accessors of the last frame's results read "no vertices"
    return status;
}

// ---- batched independent frames: `lanes` sub-handles, each with its own stream and arena
int osep_rosa_set_lanes(osep_rosa* h, int lanes) {
    if (!h || lanes < 1 || lanes > 64) return OSEP_ERR_ARG;
    osep_rosa_impl& I = h->impl;
    CK(cudaSetDevice(I.device));
    while ((int)I.lanes.size() < lanes - 1) {
        osep_rosa* l = nullptr;
        int rc = osep_rosa_create(&I.cfg, I.device, I.d.Ncap, &l);
        if (rc) return rc;
        I.lanes.push_back(&l->impl);
    }
    while ((int)I.lanes.size() > lanes - 1) { osep_rosa_impl* l = I.lanes.back(); I.lanes.pop_back(); rosa_free(l); delete reinterpret_cast<osep_rosa*>(l); }
    // several frames in flight: give each frame's persistent drosa kernel a slice of the SMs so that `lanes` of them are co-resident
    int slice = lanes > 1 ? std::max(I.cfg.niter_drosa + 6, I.sms / lanes) : 0;
    { const char* e = getenv("OSEP_FUSED_SLICE"); if (e && lanes > 1 && atoi(e) > I.cfg.niter_drosa) slice = atoi(e); }   // tuning knob
    I.fused_grid = slice; ++I.cfg_epoch;
    for (osep_rosa_impl* l : I.lanes) { l->fused_grid = slice; ++l->cfg_epoch; }
    return OSEP_OK;
}

static int run_batch(osep_rosa* h, const float* const* xyz, const int* n, int B, int stride, osep_rosa_result* results, bool host) {
    if (!h || !xyz || !n || B < 0) return OSEP_ERR_ARG;
    osep_rosa_impl& I = h->impl;
    CK(cudaSetDevice(I.device));
    if (I.lanes.empty()) { int rc = osep_rosa_set_lanes(h, 16); if (rc) return rc; }   // 12 persistent drosa kernels of 12 CTAs fit 148 SMs; the extra lanes keep N-scale work in flight while drosa kernels queue (measured: 8 / 12 / 16 / 24 lanes = 3.43 / 3.24 / 3.46 / 3.45 k frames/s)
    std::vector<osep_rosa_impl*> L; L.push_back(&I); for (auto* l : I.lanes) L.push_back(l);
    const int nl = (int)L.size();
    I.batch_vertices.assign(B, {});
    int worst = OSEP_OK;
    std::vector<int> inflight(nl, -1);
    std::vector<int> fifo;                        // lanes in enqueue order: the oldest frame in flight is the next one collected
    auto collect = [&](int lane) {
        const int b = inflight[lane];
        if (b < 0) return;
        osep_rosa_result r; memset(&r, 0, sizeof(r));
        const int rc = rosa_collect(L[lane], &r);
        if (rc == OSEP_ERR_CUDA) r.status = OSEP_ERR_CUDA;          // stream error: rosa_collect could not read the frame state
        if (results) results[b] = r;
        I.batch_vertices[b] = L[lane]->vertices_colmajor;
        if (r.status < 0) worst = r.status;
        inflight[lane] = -1;
    };
    auto drain = [&]() { for (int lane : fifo) collect(lane); fifo.clear(); };
    for (int b = 0; b < B; ++b) {
        // next lane: a free one, else BLOCK on the oldest frame in flight (no host polling: a sticky device error surfaces in
        // rosa_collect's cudaStreamSynchronize instead of spinning forever; frames of one batch take similar times)
        int lane = -1;
        for (int k = 0; k < nl && lane < 0; ++k) if (inflight[k] < 0) lane = k;
        if (lane < 0) { lane = fifo.front(); fifo.erase(fifo.begin()); collect(lane); }
        if (worst == OSEP_ERR_CUDA) { drain(); return worst; }
        osep_rosa_impl* l = L[lane];
        int rc = check_frame(*l, n[b], stride);
        if (rc == OSEP_OK && n[b] == 0) rc = OSEP_STAGE_1;
        if (rc != OSEP_OK) { if (results) { memset(&results[b], 0, sizeof(osep_rosa_result)); results[b].status = rc; } if (rc < 0) worst = rc; continue; }
        const float* src = xyz[b];
        // host frames are staged in the lane's own input buffer (H2D on the lane's stream); device-resident frames are read
        // in place -- the lane's CUDA graph takes buffer, n and stride from FrameParams
        if (host) { cudaMemcpyAsync(l->d.in, xyz[b], sizeof(float) * (size_t)n[b] * stride, cudaMemcpyHostToDevice, l->stream); src = l->d.in; }
        rc = rosa_enqueue(l, src, n[b], stride);
        if (rc) { drain(); return rc; }           // frames already in flight are collected (their results[] are valid) before the error is reported
        inflight[lane] = b; fifo.push_back(lane);
    }
    drain();
    return worst;
}

int osep_rosa_run_batch(osep_rosa* h, const float* const* xyz_host, const int* n, int n_frames, int stride_floats, osep_rosa_result* results) {
    if (stride_floats > 4 || stride_floats < 3) { set_err("stride_floats must be 3 or 4"); return OSEP_ERR_ARG; }
    return run_batch(h, xyz_host, n, n_frames, stride_floats, results, true);
}
int osep_rosa_run_batch_device(osep_rosa* h, const float* const* xyz_dev, const int* n, int n_frames, int stride_floats, osep_rosa_result* results) {
    return run_batch(h, xyz_dev, n, n_frames, stride_floats, results, false);
}
int osep_rosa_batch_vertices(osep_rosa* h, int frame, const float** data, int* n_vertices) {
    if (!h || !data || !n_vertices) return OSEP_ERR_ARG;
    osep_rosa_impl& I = h->impl;
    if (frame < 0 || frame >= (int)I.batch_vertices.size()) return OSEP_ERR_ARG;
    *n_vertices = (int)(I.batch_vertices[frame].size() / 3);
    *data = I.batch_vertices[frame].data();
    return OSEP_OK;
}

}  // extern "C"
