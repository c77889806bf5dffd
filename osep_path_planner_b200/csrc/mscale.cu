// csrc/mscale.cu -- M-scale ROSA stages (M ~ 1000 downsampled points) on sm_100a:
//   R12 similarity_neighbor_extraction   (rosa.cpp:201-256, 625-651)
//   R13 surf_nbs kNN(nb_knn)             (rosa.cpp:259-267)
//   R14-R17 drosa orientation loop       (rosa.cpp:270-309, 653-732)
//   R18 position step, R19 poor fix-up   (rosa.cpp:311-376, 734-791)
//   R20 dcrosa + R21 vertex_sampling     (rosa.cpp:382-532, 793-841): ONE thread-block-cluster launch, k_dcrosa_cluster
//   (R22 vertex_smooth, rosa.cpp:534-620, runs in the frame's tail kernel: tail.cuh / graph.cu)
//
// The whole M-scale state (pts, nrms, pset, vset: 4 x 16 KB; similarity CSR ~250 KB) is
// L1/L2 resident; these stages move ~0 algorithmic HBM bytes and are bounded by dependent
// latency, so the design goal is few, warp-cooperative phases:
//   * one WARP per seed for the active-set stages: slab mask by ballot, BFS over bitmasks in
//     shared memory, float sums in the canonical lane-owner order + xor butterfly (C4),
//     3x3 eigen solve in registers;
//   * the in-place (Gauss-Seidel) orientation smoothing keeps the reference's sequential
//     semantics through level scheduling of its dependency DAG;
//   * the greedy radius cover keeps the sequential semantics on one warp over precomputed
//     adjacency bit rows.
#include "rosa_impl.cuh"
#include "tail.cuh"
#include <cooperative_groups.h>
#include <type_traits>

namespace osep {

constexpr unsigned FULL = 0xffffffffu;

template <int K> __device__ __forceinline__ void butterfly(float (&v)[K]) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = v[k] + __shfl_xor_sync(FULL, v[k], off);
    }
}

// ------------------------------------------------------------------------------------------
// R12.  similarity_metric (rosa.cpp:625-651)
__device__ __forceinline__ float similarity_metric(const f3& p1, const f3& v1, const f3& p2, const f3& v2, float range_r) {
    const float scale = 5.0f;
    const f3 d = sub3(p1, p2);
    const float dot12_raw = dot3(v1, v2);
    const float dot12 = smax(0.0f, smin(1.0f, dot12_raw));
    if (dot12 == 0.0f) return 0.0f;
    const float proj1 = dot3(d, v1);
    const float proj2 = -dot3(d, v2);
    const float s1 = scale * proj1, s2 = scale * proj2;
    const float dist1 = norm3(add3(d, mul3(v1, s1))) / range_r;
    const float dist2_ = norm3(sub3(d, mul3(v2, s2))) / range_r;
    float k1, k2;
    if (dist1 > 1.0f) k1 = 0.0f; else { const float t2 = dist1 * dist1, t3 = t2 * dist1; k1 = 2.0f * t3 - 3.0f * t2 + 1.0f; }
    if (k1 == 0.0f) return 0.0f;
    if (dist2_ > 1.0f) k2 = 0.0f; else { const float t2 = dist2_ * dist2_, t3 = t2 * dist2_; k2 = 2.0f * t3 - 3.0f * t2 + 1.0f; }
    if (k2 == 0.0f) return 0.0f;
    const float align2 = dot12 * dot12;
    return smin(k1, k2) * align2;
}

// one warp per point.  FILL = false: count kept neighbours; FILL = true: write them in
// ascending (d2, index) order (radiusSearch returns sorted results; the order is the
// summation order of dcrosa, rosa.cpp:395-397, 804-813).
template <bool FILL>
__device__ __forceinline__ void simi_row(FrameState* st, const float4* __restrict__ pts, const float4* __restrict__ nrms,
                                         int* simi_off, int* __restrict__ simi_idx, int* __restrict__ tmp_j, float* __restrict__ tmp_d,
                                         unsigned* __restrict__ simi_bits, int M, int W, int i, int lane, const float4* pts_scan) {
    const float leaf = st->leaf_size;
    const float radius_r = 10.0f * leaf;
    const float th_sim = 0.1f * leaf;
    const float r2 = radius_r * radius_r;
    const f3 P1 = ld3(pts, i), N1 = ld3(nrms, i);
    const bool n1ok = fin3(N1);
    const f3 v1 = normalize3(N1);
    const int base = FILL ? simi_off[i] : 0;
    int cnt = 0;
    if (n1ok && FILL) {
        // second pass: the accepted set is the bit row the first pass wrote -- no second evaluation of the metric
        for (int j0 = 0; j0 < M; j0 += 32) {
            const int j = j0 + lane;
            const unsigned m = simi_bits[(size_t)i * W + (j0 >> 5)];
            const bool keep = (m >> lane) & 1u;
            if (keep) { const int pos = base + cnt + __popc(m & ((1u << lane) - 1u)); tmp_j[pos] = j; tmp_d[pos] = dist2(P1, ld3(pts, j)); }
            cnt += __popc(m);
        }
    } else if (n1ok) {
        // count pass.  The radius test (cheap, ~1 candidate in 8 passes) and the metric (two square roots, five divisions)
        // are separated: candidates inside the radius are compacted into a 64-entry window and the metric runs on DENSE
        // batches of 32 -- evaluated inside the scan loop it ran once per 32 candidates with ~4 lanes active.  The kept
        // bits are collected in a shared-memory copy of the row (the bit row i of the similarity graph, drosa.cu).
        __shared__ unsigned short s_buf[WARPS_PER_BLOCK][64];
        __shared__ unsigned s_row[WARPS_PER_BLOCK][128];
        unsigned short* buf = s_buf[threadIdx.x >> 5];
        unsigned* row = s_row[threadIdx.x >> 5];
        for (int w = lane; w < W; w += 32) row[w] = 0u;
        __syncwarp();
        const unsigned lt = (1u << lane) - 1u;
        auto batch = [&](int nb) {
            bool keep = false; int j = 0;
            if (lane < nb) {
                j = buf[lane];
                const f3 N2 = ld3(nrms, j);
                if (fin3(N2)) {
                    const f3 P2 = ld3(pts_scan, j);
                    const f3 v2 = normalize3(N2);
                    keep = similarity_metric(P1, v1, P2, v2, radius_r) > th_sim;
                }
            }
            if (keep) atomicOr(&row[j >> 5], 1u << (j & 31));
            cnt += __popc(__ballot_sync(FULL, keep));
        };
        int nbuf = 0;
        for (int j0 = 0; j0 < M; j0 += 32) {
            const int j = j0 + lane;
            const bool in = (j < M && j != i) && (dist2(P1, ld3(pts_scan, j < M ? j : 0)) < r2);
            const unsigned m = __ballot_sync(FULL, in);
            if (in) buf[nbuf + __popc(m & lt)] = (unsigned short)j;
            nbuf += __popc(m);
            __syncwarp();
            if (nbuf >= 32) {
                batch(32);
                const unsigned short mv = buf[32 + lane];       // the window holds < 64 entries: move the rest down
                __syncwarp();
                buf[lane] = mv;
                nbuf -= 32;
                __syncwarp();
            }
        }
        if (nbuf > 0) batch(nbuf);
        __syncwarp();
        for (int w = lane; w < W; w += 32) simi_bits[(size_t)i * W + w] = row[w];      // padding words of the row stay zero
    }
    if (!FILL) {
        if (!n1ok) for (int w = lane; w < W; w += 32) simi_bits[(size_t)i * W + w] = 0u;
        if (lane == 0) simi_off[i] = cnt;
        return;
    }
    __syncwarp();
    for (int e = lane; e < cnt; e += 32) {
        const float d = tmp_d[base + e]; const int j = tmp_j[base + e];
        int rank = 0;
        for (int f = 0; f < cnt; ++f) rank += nb_less(tmp_d[base + f], tmp_j[base + f], d, j) ? 1 : 0;
        simi_idx[base + rank] = j;
    }
}

template <bool FILL>
__global__ void k_simi(FrameState* st, const float4* __restrict__ pts, const float4* __restrict__ nrms,
                       int* simi_off, int* __restrict__ simi_idx, int* __restrict__ tmp_j, float* __restrict__ tmp_d,
                       unsigned* __restrict__ simi_bits, int Scap) {
    if (!FILL) frame_stamp(st, TS_SIMI);
    if (st->status != 0) return;
    const int M = st->M;
    const int W = (((M + 31) >> 5) + 3) & ~3;   // row stride of the bit matrix: padded to 16 bytes (drosa.cu reads rows with 16-byte loads)
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const float4* pts_scan = pts;
    if (!FILL) {
        // count pass: the block's 8 warps scan the same M positions -- staged once (coalesced, every load in flight at the
        // same time) instead of 27 dependent rounds of L2 latency per warp
        extern __shared__ float4 s_simi_pts[];
        if ((int)(blockIdx.x * WARPS_PER_BLOCK) < M) {
            for (int j = threadIdx.x; j < M; j += blockDim.x) s_simi_pts[j] = pts[j];
            pts_scan = s_simi_pts;
        }
        __syncthreads();
    }
    if (i < M) simi_row<FILL>(st, pts, nrms, simi_off, simi_idx, tmp_j, tmp_d, simi_bits, M, W, i, lane, pts_scan);
    if (FILL) return;
    // the last block of the count pass turns the counts into CSR offsets (no separate scan launch)
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket_m, 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) { st->ticket_m = 0; simi_off[M] = 0; }
    __threadfence();
    __syncthreads();
    const int total = block_scan_inplace(simi_off, M + 1);
    if (threadIdx.x == 0) { st->S = total; if (total > Scap) st->status = OSEP_ERR_CAPACITY; }
}

// ------------------------------------------------------------------------------------------
// R13.  kNN(nb_knn) among the M points, brute force, canonical (d2, index) order, self first.

// One warp per point.  Candidates inside an adaptive radius are compacted into a small shared list
// (ballot + popc), which is then rank-sorted by (d2, index); the radius grows until the list holds at
// least k points (then the k nearest are all inside it: exact) and shrinks if it overflows the list.
constexpr int SURF_LIST = 128;
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_surf_knn(FrameState* st, const float4* __restrict__ pts, int* __restrict__ surf, int k_req) {
    __shared__ float s_d[WARPS_PER_BLOCK][SURF_LIST];
    __shared__ int s_j[WARPS_PER_BLOCK][SURF_LIST];
    frame_stamp(st, TS_SURF);
    if (st->status != 0) return;
    const int M = st->M;
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * WARPS_PER_BLOCK + wib;
    if (p >= M) return;
    const int k_eff = min(k_req, M);
    const f3 q = ld3(pts, p);
    float r = 2.5f * st->leaf_used;
    if (!(r > 0.0f) || !fin(r)) r = 1.0f;
    float lo = 0.0f, hi = __int_as_float(0x7f800000);      // radii known to be too small / too large
    int cnt = 0;
    for (int attempt = 0; attempt < 64; ++attempt) {
        const float r2 = (M <= SURF_LIST) ? __int_as_float(0x7f800000) : r * r;
        cnt = 0;
        for (int j0 = 0; j0 < M; j0 += 32) {
            const int j = j0 + lane;
            float d2 = 0.0f; bool in = false;
            if (j < M) { d2 = dist2(q, ld3(pts, j)); in = d2 < r2; }
            const unsigned m = __ballot_sync(FULL, in);
            const int pos = cnt + __popc(m & ((1u << lane) - 1u));
            if (in && pos < SURF_LIST) { s_d[wib][pos] = d2; s_j[wib][pos] = j; }
            cnt += __popc(m);
        }
        if (cnt >= k_eff && cnt <= SURF_LIST) break;
        if (cnt < k_eff) { lo = r; r = (hi < 3.0e38f) ? 0.5f * (lo + hi) : r * 1.6f; }
        else { hi = r; r = 0.5f * (lo + hi); }
    }
    __syncwarp();
    if (cnt < k_eff || cnt > SURF_LIST) { if (lane == 0) st->status = OSEP_ERR_CAPACITY; return; }    // > 128 coincident points
    for (int e = lane; e < cnt; e += 32) {
        const float d = s_d[wib][e]; const int j = s_j[wib][e];
        int rank = 0;
        for (int f = 0; f < cnt; ++f) rank += nb_less(s_d[wib][f], s_j[wib][f], d, j) ? 1 : 0;
        if (rank < k_eff) surf[p * k_req + rank] = j;
    }
    for (int u = k_eff + lane; u < k_req; u += 32) surf[p * k_req + u] = -1;
}

// R19 part 1 (single block): compact goodCenters, list poor points in ascending index.
__global__ void k_poor_fixup(FrameState* st, const int* __restrict__ good, int* good_rank, const float4* __restrict__ centers,
                             float4* __restrict__ good_centers, int* __restrict__ poor_idx, const float4* __restrict__ pts, float4* pset) {
    frame_stamp(st, TS_FIXUP);
    if (st->status != 0) return;
    const int M = st->M;
    for (int i = threadIdx.x; i < M; i += blockDim.x) good_rank[i] = good[i];
    if (threadIdx.x == 0) good_rank[M] = 0;
    __syncthreads();
    const int ng = block_scan_inplace(good_rank, M + 1);
    __syncthreads();
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        if (good[i]) good_centers[good_rank[i]] = centers[i];
        else poor_idx[i - good_rank[i]] = i;
    }
    if (threadIdx.x == 0) { st->n_good = ng; st->n_poor = M - ng; }
    __syncthreads();
    // R19 part 2 (same block, no second launch): one warp per poor point: nearest slot of tmp_pt_ (good slot = surface
    // point, poor slot = (0,0,0), rosa.cpp:314,359), then pset[poor] = goodCenters[gid] AS WRITTEN (rosa.cpp:374: gid is a
    // slot index applied to the compacted list).  Out of bounds = UB in the reference: pset is left unchanged and
    // OSEP_FLAG_POOR_FIXUP_OOB raised.  All fix-ups read pts / good / good_centers only, so they are independent.
    const int lane = threadIdx.x & 31, n_poor = M - ng;
    for (int w = threadIdx.x >> 5; w < n_poor; w += (int)(blockDim.x >> 5)) {
        const int poor = poor_idx[w];
        const f3 q = ld3(pts, poor);
        float bd = __int_as_float(0x7f800000); int bi = 0x7fffffff;
        for (int i = lane; i < M; i += 32) {
            const f3 c = good[i] ? ld3(pts, i) : f3{0.0f, 0.0f, 0.0f};
            const float d = dist2(q, c);
            if (nb_less(d, i, bd, bi)) { bd = d; bi = i; }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const float od = __shfl_xor_sync(FULL, bd, off); const int oi = __shfl_xor_sync(FULL, bi, off);
            if (nb_less(od, oi, bd, bi)) { bd = od; bi = oi; }
        }
        if (lane == 0 && bi != 0x7fffffff) {
            if (bi < ng) pset[poor] = good_centers[bi];
            else atomicOr(&st->flags, (unsigned)OSEP_FLAG_POOR_FIXUP_OOB);
        }
    }
}

// ------------------------------------------------------------------------------------------
// R20 dcrosa (rosa.cpp:382-422): two Jacobi passes per iteration, one thread per point, sums in
// similarity-list order.
__global__ void k_dcrosa_mean(const FrameState* __restrict__ st, const float4* __restrict__ pin, float4* __restrict__ pout,
                              const int* __restrict__ simi_off, const int* __restrict__ simi_idx) {
    if (st->status != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st->M) return;
    const int b = simi_off[i], e = simi_off[i + 1];
    if (e == b) { pout[i] = pin[i]; return; }
    f3 acc{0.0f, 0.0f, 0.0f};
    for (int t = b; t < e; ++t) acc = add3(acc, ld3(pin, simi_idx[t]));
    st3(pout, i, div3(acc, (float)(e - b)), 1.0f);
}
__global__ void k_dcrosa_line(const FrameState* __restrict__ st, const float4* __restrict__ pin, float4* __restrict__ pout,
                              const int* __restrict__ simi_off, const int* __restrict__ simi_idx) {
    if (st->status != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st->M) return;
    const int b = simi_off[i], e = simi_off[i + 1];
    const int m = e - b;
    const float4 self4 = pin[i];
    if (m < 3) { pout[i] = self4; return; }
    // local_line_fit (rosa.cpp:793-841)
    f3 mean{0.0f, 0.0f, 0.0f};
    for (int t = b; t < e; ++t) mean = add3(mean, ld3(pin, simi_idx[t]));
    mean = div3(mean, (float)m);
    float c00 = 0, c10 = 0, c11 = 0, c20 = 0, c21 = 0, c22 = 0;
    for (int t = b; t < e; ++t) {
        const f3 q = sub3(ld3(pin, simi_idx[t]), mean);
        c00 += q.x * q.x; c10 += q.y * q.x; c11 += q.y * q.y; c20 += q.z * q.x; c21 += q.z * q.y; c22 += q.z * q.z;
    }
    const float fm = (float)m;
    c00 /= fm; c10 /= fm; c11 /= fm; c20 /= fm; c21 /= fm; c22 /= fm;
    const Eig3 E = eig3_sym_dev(c00, c10, c11, c20, c21, c22);
    if (!E.ok) { pout[i] = self4; return; }
    const float denom = smax(1e-6f, sum3(E.ev[0], E.ev[1], E.ev[2]));
    f3 dir = eig_col(E, 2);
    const float nrm = norm3(dir);
    if (nrm > 1e-6f) dir = div3(dir, nrm); else dir = {1.0f, 0.0f, 0.0f};
    const float conf = E.ev[2] / denom;
    if (conf > 0.5f) {
        const f3 x{self4.x, self4.y, self4.z};
        const float t = dot3(dir, sub3(x, mean));
        st3(pout, i, add3(mean, mul3(dir, t)), 1.0f);
    } else pout[i] = self4;
}

// R20 fused: all niter_dcrosa iterations (two Jacobi passes each) in ONE launch on a thread-block CLUSTER of
// 8 CTAs.  Every CTA keeps a full copy of the pset ping-pong in its shared memory (16 B x M x 2); a thread owns
// one point, gathers its similarity neighbours from the local copy (sums in list order, as rosa.cpp:395-397 and
// :804-813), and broadcasts its result to the 8 copies through distributed shared memory; cluster.sync()
// separates the Jacobi passes.  One SM alone is shared-memory-bandwidth bound on the random gathers; 8 SMs
// with replicated operands are not, and no pass goes back to L2.
#ifndef OSEP_DC_THREADS
#define OSEP_DC_THREADS 256
#endif
#ifndef OSEP_DC_SPREAD
#define OSEP_DC_SPREAD 2
#endif
#ifndef OSEP_DC_UNROLL
#define OSEP_DC_UNROLL 8
#endif
#ifndef OSEP_DC_BALANCE
#define OSEP_DC_BALANCE 48
#endif
constexpr int DC_THREADS = OSEP_DC_THREADS;
constexpr int DC_UNROLL = OSEP_DC_UNROLL;
constexpr int DC_SPREAD = OSEP_DC_SPREAD;       // every DC_SPREAD-th thread owns a point: fewer points per warp = shorter per-warp maximum of list length / eigen iterations
constexpr int DC_BALANCE = OSEP_DC_BALANCE;     // cost of a point in list entries (eigen solve + fixed work) next to its list length, for the partition below
constexpr int DC_SMEM_BUDGET = 200 * 1024;

// A CTA is bound by the number of warp instructions it issues (three list sweeps + one eigen solve per point and iteration:
// more warps per CTA measured SLOWER, 68 -> 83 -> 124 us for 8 / 16 / 32 warps), so the lever is the number of SMs: CL = 16
// (non-portable cluster size) when a single frame owns the GPU, the portable 8 when several frames are in flight.  Points are
// dealt to the CTAs in contiguous ranges of equal COST (list length + DC_BALANCE per point), not equal count.
// vertex_sampling (R21, rosa.cpp:424-532) runs in the same launch (vs.on): when the last dcrosa pass is done every CTA of the
// cluster holds the final pset in shared memory, so the five stages that used to be five launches (55 us) become phases
// separated by cluster.sync(): finite compaction (every CTA, redundantly), cover rows (a slice of rows per CTA, written
// into CTA 0's shared memory through DSMEM), the sequential greedy cover (one warp of CTA 0; seeds pushed to every CTA),
// correspondences (a slice of points per CTA, pushed to every CTA) and recentring (vertices dealt to the warps of the cluster).
struct VsArgs {
    int on;
    const float4* pts; float alpha;
    unsigned* adj; int* seeds; int* corresp; float4* skel; float4* skel_sampled;
};

template <int CL>
__global__ void __launch_bounds__(DC_THREADS)
k_dcrosa_cluster(FrameState* st, float4* pset, const int* __restrict__ simi_off, const int* __restrict__ simi_idx, int niter,
                 unsigned smem_bytes, long long* dbg, const VsArgs vs) {
    frame_stamp(st, TS_DCROSA);
    int dk = 0;
    auto stamp = [&]() { if (dbg && threadIdx.x == 0 && blockIdx.x == 0 && dk < 60) { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[dk++] = t; } };
    stamp();
    extern __shared__ float4 dcs[];
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    if (st->status != 0) return;                       // uniform over the cluster
    const int M = st->M;
    float4* pa = dcs; float4* pb = dcs + M;
    const unsigned rank = cluster.block_rank();
    for (int i = threadIdx.x; i < M; i += blockDim.x) pa[i] = pset[i];
    // range [lo, hi) of this CTA: first point whose cost prefix simi_off[i] + DC_BALANCE * i reaches rank / CL of the total
    __shared__ int s_lohi[2];
    {
        const long long total = (long long)simi_off[M] + (long long)DC_BALANCE * M;
        const long long want0 = total * rank / CL, want1 = total * (rank + 1) / CL;
        for (int i = threadIdx.x; i <= M; i += blockDim.x) {      // the prefix is strictly increasing: exactly one i per boundary
            const long long pi = (long long)simi_off[i] + (long long)DC_BALANCE * i;
            const long long pp = i > 0 ? (long long)simi_off[i - 1] + (long long)DC_BALANCE * (i - 1) : -1;
            if (pi >= want0 && pp < want0) s_lohi[0] = i;
            if (pi >= want1 && pp < want1) s_lohi[1] = i;
        }
    }
    __syncthreads();
    const int lo = s_lohi[0], hi = s_lohi[1];
    // this CTA's slice of the similarity CSR, staged in shared memory when it fits
    int* sidx = reinterpret_cast<int*>(dcs + 2 * (size_t)M);
    const int seg_b = simi_off[lo], seg_e = simi_off[hi];
    const bool csr_in = ((size_t)32 * M + 4 * (size_t)(seg_e - seg_b)) <= smem_bytes;
    if (csr_in) for (int t = seg_b + threadIdx.x; t < seg_e; t += blockDim.x) sidx[t - seg_b] = simi_idx[t];
    float4* pa_r[CL]; float4* pb_r[CL];
#pragma unroll
    for (int r = 0; r < CL; ++r) { pa_r[r] = cluster.map_shared_rank(pa, r); pb_r[r] = cluster.map_shared_rank(pb, r); }
    // list bounds of this thread's points do not change over the iterations: the first point's are kept in registers
    const bool owner = (threadIdx.x % DC_SPREAD) == 0;
    const int i_first = lo + (int)(threadIdx.x / DC_SPREAD), i_step = (int)(blockDim.x / DC_SPREAD);
    int b_first = 0, e_first = 0;
    if (owner && i_first < hi) { b_first = simi_off[i_first]; e_first = simi_off[i_first + 1]; }
    cluster.sync();
    stamp();
    auto run = [&](auto nbr) {              // nbr[t] for t in [seg_b, seg_e): a shared-memory or a global pointer (typed, so the staged lists are read with LDS)
        for (int it = 0; it < niter; ++it) {
            for (int i = i_first; i < hi && owner; i += i_step) {          // neighbour mean (rosa.cpp:388-399)
                const int b = (i == i_first) ? b_first : simi_off[i], e = (i == i_first) ? e_first : simi_off[i + 1];
                float4 out = pa[i];
                if (e > b) {
                    f3 acc{0.0f, 0.0f, 0.0f};
#pragma unroll (DC_UNROLL)
                    for (int t = b; t < e; ++t) { const float4 v = pa[nbr[t]]; acc = add3(acc, f3{v.x, v.y, v.z}); }
                    const f3 mth = div3(acc, (float)(e - b));
                    out = make_float4(mth.x, mth.y, mth.z, 1.0f);
                }
#pragma unroll
                for (int r = 0; r < CL; ++r) pb_r[r][i] = out;
            }
            stamp();
            cluster.sync();
            stamp();
            for (int i = i_first; i < hi && owner; i += i_step) {          // local_line_fit + projection (rosa.cpp:402-418, 793-841)
                const int b = (i == i_first) ? b_first : simi_off[i], e = (i == i_first) ? e_first : simi_off[i + 1];
                const int m = e - b;
                const float4 self4 = pb[i];
                float4 out = self4;
                if (m >= 3) {
                    f3 mean{0.0f, 0.0f, 0.0f};
#pragma unroll (DC_UNROLL)
                    for (int t = b; t < e; ++t) { const float4 v = pb[nbr[t]]; mean = add3(mean, f3{v.x, v.y, v.z}); }
                    mean = div3(mean, (float)m);
                    float c00 = 0, c10 = 0, c11 = 0, c20 = 0, c21 = 0, c22 = 0;
#pragma unroll (DC_UNROLL)
                    for (int t = b; t < e; ++t) {
                        const float4 v = pb[nbr[t]];
                        const f3 q = sub3(f3{v.x, v.y, v.z}, mean);
                        c00 += q.x * q.x; c10 += q.y * q.x; c11 += q.y * q.y; c20 += q.z * q.x; c21 += q.z * q.y; c22 += q.z * q.z;
                    }
                    const float fm = (float)m;
                    c00 /= fm; c10 /= fm; c11 /= fm; c20 /= fm; c21 /= fm; c22 /= fm;
                    const Eig3 E = eig3_sym_dev(c00, c10, c11, c20, c21, c22);
                    if (E.ok) {
                        const float denom = smax(1e-6f, sum3(E.ev[0], E.ev[1], E.ev[2]));
                        f3 dir = eig_col(E, 2);
                        const float nrm = norm3(dir);
                        if (nrm > 1e-6f) dir = div3(dir, nrm); else dir = {1.0f, 0.0f, 0.0f};
                        const float conf = E.ev[2] / denom;
                        if (conf > 0.5f) {
                            const f3 x{self4.x, self4.y, self4.z};
                            const float t = dot3(dir, sub3(x, mean));
                            const f3 pr = add3(mean, mul3(dir, t));
                            out = make_float4(pr.x, pr.y, pr.z, 1.0f);
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < CL; ++r) pa_r[r][i] = out;
            }
            stamp();
            cluster.sync();
            stamp();
        }
    };
    if (csr_in) run(static_cast<const int*>(sidx - seg_b)); else run(simi_idx);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) pset[i] = pa[i];
    stamp();
    if (!vs.on) return;
    // ---- R21 vertex_sampling.  pa = final pset (complete in every CTA; the last cluster.sync() above also retired the staged lists)
    if (M == 0) { if (rank == 0 && threadIdx.x == 0) st->status = OSEP_STAGE_6; return; }      // uniform over the cluster
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarps = DC_THREADS / 32;
    unsigned char* R = reinterpret_cast<unsigned char*>(dcs + 2 * (size_t)M);
    float4* pcl = reinterpret_cast<float4*>(R);                                 // finite rows of pset, compacted (rosa.cpp:441-449)
    int* scan = reinterpret_cast<int*>(pcl + M);                                // M + 1 (16-byte padded)
    int* seeds_s = scan + ((M + 1 + 3) & ~3);
    int* corr_s = seeds_s + ((M + 3) & ~3);
    unsigned* rows_s = reinterpret_cast<unsigned*>(corr_s + ((M + 3) & ~3));     // CTA 0: Mc x Wn cover bit rows
    __shared__ int s_V;
    const size_t fixed_bytes = (size_t)32 * M + (size_t)16 * M + 4 * (size_t)((M + 4) & ~3) + 8 * (size_t)((M + 3) & ~3);
    const bool rows_in = fixed_bytes + (size_t)4 * M * ((M + 31) >> 5) <= smem_bytes;
    for (int i = threadIdx.x; i < M; i += blockDim.x) { const float4 v = pa[i]; scan[i] = (fin(v.x) && fin(v.y) && fin(v.z)) ? 1 : 0; }
    if (threadIdx.x == 0) scan[M] = 0;
    __syncthreads();
    const int Mc = block_scan_inplace(scan, M + 1);
    __syncthreads();
    for (int i = threadIdx.x; i < M; i += blockDim.x) { const float4 v = pa[i]; if (fin(v.x) && fin(v.y) && fin(v.z)) pcl[scan[i]] = v; }
    __syncthreads();
    stamp();
    const int Wn = (Mc + 31) >> 5;
    const float Rl = st->leaf_size; const float R2 = Rl * Rl;
    const int per = (Mc + CL - 1) / CL;
    const int i0 = min(Mc, (int)rank * per), i1 = min(Mc, i0 + per);
    {   // cover rows of this CTA's points: bit j of row i = (d2(i, j) < R^2) or i == j; one warp per row, lane w keeps word w
#ifdef OSEP_EXP2
        unsigned* rows0 = rows_s;
#else
        unsigned* rows0 = rows_in ? cluster.map_shared_rank(rows_s, 0) : vs.adj;
#endif
        for (int i = i0 + wib; i < i1; i += nwarps) {
            const f3 q = ld3(pcl, i);
            for (int w0 = 0; w0 < Wn; w0 += 32) {
                unsigned mine = 0;
                const int wn = min(32, Wn - w0);
#pragma unroll 4
                for (int u = 0; u < wn; ++u) {
                    const int j = (w0 + u) * 32 + lane;
                    const float dd = dist2(q, ld3(pcl, min(j, Mc - 1)));            // branch-free: the load is clamped, the verdict masked
                    const bool in = (j < Mc) && ((j == i) || (dd < R2));
                    const unsigned word = __ballot_sync(FULL, in);
                    mine = (lane == u) ? word : mine;
                }
                if (lane < wn) rows0[(size_t)i * Wn + w0 + lane] = mine;
            }
        }
    }
    stamp();
    if (!rows_in) __threadfence();
    cluster.sync();
    stamp();
    if (rank == 0) {
        // greedy cover with the reference's sequential semantics (lowest uncovered index becomes the next vertex, rosa.cpp:462-502):
        // each lane of warp 0 keeps its words of the covered set in registers, a step = ffs + warp min + one row OR
        if (wib == 0) {
            auto walk = [&](auto rows, auto WPLc) {
                constexpr int WPL = decltype(WPLc)::value;                   // words of the covered set per lane
                unsigned cov[WPL];
#pragma unroll
                for (int k = 0; k < WPL; ++k) {
                    const int w = lane + 32 * k;
                    unsigned c = 0xffffffffu;               // words beyond Mc count as covered
                    if (w < Wn) { const int rem = Mc - w * 32; c = (rem >= 32) ? 0u : ~((1u << rem) - 1u); }
                    cov[k] = c;
                }
                int V = 0;
                while (true) {
                    int cand = 0x7fffffff;
#pragma unroll
                    for (int k = WPL - 1; k >= 0; --k) { const unsigned u = ~cov[k]; if (u) cand = (lane + 32 * k) * 32 + __ffs(u) - 1; }
                    const int seed = __reduce_min_sync(FULL, cand);
                    if (seed == 0x7fffffff) break;
                    if (lane == 0) { seeds_s[V] = seed; vs.seeds[V] = seed; }
                    ++V;
                    auto row = rows + (size_t)seed * Wn;
#pragma unroll
                    for (int k = 0; k < WPL; ++k) { const int w = lane + 32 * k; if (w < Wn) cov[k] |= row[w]; }
                }
                if (lane == 0) { s_V = V; st->Mc = Mc; st->V = V; if (Mc != M) st->flags |= OSEP_FLAG_NONFINITE_PSET; }
            };
            if (rows_in) { if (Wn <= 32) walk(static_cast<const unsigned*>(rows_s), std::integral_constant<int, 1>{}); else walk(static_cast<const unsigned*>(rows_s), std::integral_constant<int, 4>{}); }
            else walk(static_cast<const unsigned*>(vs.adj), std::integral_constant<int, 4>{});          // up to 128 words = 4096 points
        }
        __syncthreads();
        const int V = s_V;
        stamp();
        for (int r = 1; r < CL; ++r) {                  // push the seeds and their count to the other CTAs
            int* sr = cluster.map_shared_rank(seeds_s, r);
            for (int t = threadIdx.x; t < V; t += blockDim.x) sr[t] = seeds_s[t];
            if (threadIdx.x == 0) *cluster.map_shared_rank(&s_V, r) = V;
        }
        stamp();
    }
    cluster.sync();
    stamp();
    const int V = s_V;
    // corresp[idx] = vertex with the smallest d2 among those covering idx, first wins ties (rosa.cpp:490-496); the seed
    // positions are staged contiguously (pb is free now) so that the scan has no dependent index load
    float4* seedpos = pb;
    for (int t = threadIdx.x; t < V; t += blockDim.x) { const int sd = seeds_s[t]; const float4 c = pcl[sd]; seedpos[t] = make_float4(c.x, c.y, c.z, __int_as_float(sd)); }
    __syncthreads();
    for (int idx = i0 + (int)threadIdx.x; idx < i1; idx += blockDim.x) {
        const f3 q = ld3(pcl, idx);
        float mind2 = __int_as_float(0x7f800000); int best = -1;
#pragma unroll 4
        for (int t = 0; t < V; ++t) {
            const float4 c = seedpos[t];
            const bool self = __float_as_int(c.w) == idx;
            const float dd = dist2(f3{c.x, c.y, c.z}, q);
            const float d2 = self ? 0.0f : dd;
            const bool take = (self || d2 < R2) && d2 < mind2;
            mind2 = take ? d2 : mind2; best = take ? t : best;
        }
        vs.corresp[idx] = best;
#ifdef OSEP_EXP1
        corr_s[idx] = best;
#else
#pragma unroll
        for (int r = 0; r < CL; ++r) cluster.map_shared_rank(corr_s, r)[idx] = best;
#endif
    }
    stamp();
    cluster.sync();
    stamp();
    // recentring (rosa.cpp:506-530): one warp per vertex; the sum over pts_ stays in ascending point index (lanes load 32
    // consecutive candidates, the matches are added in bit order)
    for (int v = (int)rank * nwarps + wib; v < V; v += CL * nwarps) {
        f3 sum{0.0f, 0.0f, 0.0f}; int cnt = 0;
        for (int j0 = 0; j0 < Mc; j0 += 32) {
            const int j = j0 + lane;
            bool hit = false; float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < Mc) { hit = (corr_s[j] == v); if (hit) p = vs.pts[j]; }
            unsigned m = __ballot_sync(FULL, hit);
            cnt += __popc(m);
            while (m) {
                const int b = __ffs(m) - 1; m &= m - 1;
                sum.x += __shfl_sync(FULL, p.x, b); sum.y += __shfl_sync(FULL, p.y, b); sum.z += __shfl_sync(FULL, p.z, b);
            }
        }
        if (lane == 0) {
            f3 cur = ld3(pa, seeds_s[v]);                    // rosa.cpp:475 (uncompacted pset, compacted index)
            if (cnt > 0) {
                const f3 ec = div3(sum, (float)cnt);
                const float b = (float)(1.0 - (double)vs.alpha);    // rosa.cpp:529
                cur = add3(mul3(cur, vs.alpha), mul3(ec, b));
            }
            st3(vs.skel, v, cur, 1.0f); st3(vs.skel_sampled, v, cur, 1.0f);
        }
    }
    stamp();
}

template <int CL>
static void launch_dcrosa_cluster(FrameState* st, float4* pset, const int* simi_off, const int* simi_idx, int niter, long long* dbg, const VsArgs& vs, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL); cfg.blockDim = dim3(DC_THREADS); cfg.dynamicSmemBytes = DC_SMEM_BUDGET; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_dcrosa_cluster<CL>, st, pset, simi_off, simi_idx, niter, (unsigned)DC_SMEM_BUDGET, dbg, vs);
}

// ------------------------------------------------------------------------------------------
// R21 vertex_sampling (rosa.cpp:424-532)
__global__ void k_vs_compact(FrameState* st, const float4* __restrict__ pset, float4* __restrict__ pcloud, int* scratch) {
    if (st->status != 0) return;
    const int M = st->M;
    if (M == 0) { if (threadIdx.x == 0) st->status = OSEP_STAGE_6; return; }
    for (int i = threadIdx.x; i < M; i += blockDim.x) { const float4 v = pset[i]; scratch[i] = (fin(v.x) && fin(v.y) && fin(v.z)) ? 1 : 0; }
    if (threadIdx.x == 0) scratch[M] = 0;
    __syncthreads();
    // keep flags: recompute after the scan from pset
    const int Mc = block_scan_inplace(scratch, M + 1);
    __syncthreads();
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        const float4 v = pset[i];
        if (fin(v.x) && fin(v.y) && fin(v.z)) pcloud[scratch[i]] = v;
    }
    if (threadIdx.x == 0) { st->Mc = Mc; st->V = 0; if (Mc != M) st->flags |= OSEP_FLAG_NONFINITE_PSET; }
}
// cover adjacency: bit j of row i = (d2(i,j) < R^2) or i == j; one warp per row
__global__ void k_vs_adj(const FrameState* __restrict__ st, const float4* __restrict__ pcloud, unsigned* __restrict__ adj, int W_cap) {
    if (st->status != 0) return;
    const int Mc = st->Mc;
    const int W = (Mc + 31) >> 5;          // dense row stride
    (void)W_cap;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= Mc) return;
    const float R = st->leaf_size; const float R2 = R * R;
    const f3 q = ld3(pcloud, i);
    const int Wn = (Mc + 31) >> 5;
    for (int w = 0; w < Wn; ++w) {
        const int j = w * 32 + lane;
        bool in = false;
        if (j < Mc) in = (j == i) || (dist2(q, ld3(pcloud, j)) < R2);
        const unsigned word = __ballot_sync(FULL, in);
        if (lane == 0) adj[(size_t)i * W + w] = word;
    }
}
// greedy cover with the reference's sequential semantics (lowest uncovered index becomes the next vertex,
// rosa.cpp:462-502).  The cover bit matrix is staged in shared memory by the whole block; warp 0 then walks:
// each lane keeps its words of the covered set in registers, a step = ffs + warp min + one row OR.
constexpr int VS_SMEM_BUDGET = 200 * 1024;
__global__ void __launch_bounds__(1024) k_vs_greedy(FrameState* st, const unsigned* __restrict__ adj, int W, int* __restrict__ seeds) {
    extern __shared__ unsigned vs_rows[];
    if (st->status != 0) return;
    const int Mc = st->Mc;
    const int Wn = (Mc + 31) >> 5;
    const bool in_smem = (size_t)4 * Mc * Wn <= (size_t)VS_SMEM_BUDGET;
    if (in_smem) stage_bit_rows(vs_rows, adj, Mc, Wn);
    (void)W;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const unsigned* rows = in_smem ? vs_rows : adj;
    const int stride = Wn;
    constexpr int WPL = 4;                      // up to 128 words = 4096 points
    unsigned cov[WPL];
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
        const int w = lane + 32 * k;
        unsigned c = 0xffffffffu;               // words beyond Mc count as covered
        if (w < Wn) { const int rem = Mc - w * 32; c = (rem >= 32) ? 0u : ~((1u << rem) - 1u); }
        cov[k] = c;
    }
    int V = 0;
    while (true) {
        int cand = 0x7fffffff;
#pragma unroll
        for (int k = WPL - 1; k >= 0; --k) { const unsigned u = ~cov[k]; if (u) cand = (lane + 32 * k) * 32 + __ffs(u) - 1; }
        const int seed = __reduce_min_sync(FULL, cand);
        if (seed == 0x7fffffff) break;
        if (lane == 0) seeds[V] = seed;
        ++V;
        const unsigned* row = rows + (size_t)seed * stride;
#pragma unroll
        for (int k = 0; k < WPL; ++k) { const int w = lane + 32 * k; if (w < Wn) cov[k] |= row[w]; }
    }
    if (lane == 0) st->V = V;
}
// corresp[idx] = vertex with the smallest d2 among those covering idx, first wins ties (rosa.cpp:490-496);
// seed coordinates are staged in shared memory tile by tile
__global__ void __launch_bounds__(128) k_vs_corresp(const FrameState* __restrict__ st, const float4* __restrict__ pcloud, const int* __restrict__ seeds,
                                                    int* __restrict__ corresp) {
    __shared__ float4 sseed[512];
    __shared__ int sidx[512];
    if (st->status != 0) return;
    const int Mc = st->Mc, V = st->V;
    if ((int)(blockIdx.x * blockDim.x) >= Mc) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const float R = st->leaf_size; const float R2 = R * R;
    const f3 q = idx < Mc ? ld3(pcloud, idx) : f3{0.f, 0.f, 0.f};
    float mind2 = __int_as_float(0x7f800000); int best = -1;
    for (int v0 = 0; v0 < V; v0 += 512) {
        const int nv = min(512, V - v0);
        __syncthreads();
        for (int t = threadIdx.x; t < nv; t += blockDim.x) { const int s = seeds[v0 + t]; sidx[t] = s; sseed[t] = pcloud[s]; }
        __syncthreads();
        if (idx < Mc) {
            for (int t = 0; t < nv; ++t) {
                const int s = sidx[t];
                const float4 c = sseed[t];
                const float d2 = (s == idx) ? 0.0f : dist2(f3{c.x, c.y, c.z}, q);
                if ((s == idx || d2 < R2) && d2 < mind2) { mind2 = d2; best = v0 + t; }
            }
        }
    }
    if (idx < Mc) corresp[idx] = best;
}
// recentering (rosa.cpp:506-530): one warp per vertex; the sum over pts_ stays in ascending point index
// (lanes load 32 consecutive candidates, the matches are added in bit order)
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_vs_recenter(FrameState* st, const float4* __restrict__ pts, const float4* __restrict__ pset, const int* __restrict__ seeds,
              const int* __restrict__ corresp, float alpha, float4* __restrict__ skel, float4* __restrict__ skel_sampled) {
    if (st->status != 0) return;
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (v >= st->V) return;
    const int Mc = st->Mc;
    f3 sum{0.0f, 0.0f, 0.0f}; int cnt = 0;
    for (int i0 = 0; i0 < Mc; i0 += 32) {
        const int i = i0 + lane;
        bool hit = false; float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < Mc) { hit = (corresp[i] == v); if (hit) p = pts[i]; }
        unsigned m = __ballot_sync(FULL, hit);
        cnt += __popc(m);
        while (m) {
            const int b = __ffs(m) - 1; m &= m - 1;
            sum.x += __shfl_sync(FULL, p.x, b); sum.y += __shfl_sync(FULL, p.y, b); sum.z += __shfl_sync(FULL, p.z, b);
        }
    }
    if (lane != 0) return;
    f3 cur = ld3(pset, seeds[v]);                        // rosa.cpp:475 (uncompacted pset, compacted index)
    if (cnt > 0) {
        const f3 ec = div3(sum, (float)cnt);
        const float b = (float)(1.0 - (double)alpha);    // rosa.cpp:529
        cur = add3(mul3(cur, alpha), mul3(ec, b));
    }
    st3(skel, v, cur, 1.0f); st3(skel_sampled, v, cur, 1.0f);
}

// R22 vertex_smooth (rosa.cpp:534-620): tail.cuh vertex_smooth_block, launched as part of k_frame_tail (graph.cu)

// ------------------------------------------------------------------------------------------
void drosa_set_attributes();
void launch_drosa(osep_rosa_impl* h, int sms);
void launch_drosa_prep(osep_rosa_impl* h, cudaStream_t sk);
bool dcrosa_cluster16_ok() {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16); cfg.blockDim = dim3(DC_THREADS); cfg.dynamicSmemBytes = DC_SMEM_BUDGET;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_dcrosa_cluster<16>, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
    return n >= 1;
}
void mscale_set_attributes(int Mcap) {
    drosa_set_attributes();
    cudaFuncSetAttribute(k_dcrosa_cluster<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, DC_SMEM_BUDGET);
    cudaFuncSetAttribute(k_dcrosa_cluster<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, DC_SMEM_BUDGET);
    cudaFuncSetAttribute(k_dcrosa_cluster<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(k_vs_greedy, cudaFuncAttributeMaxDynamicSharedMemorySize, VS_SMEM_BUDGET);
    cudaFuncSetAttribute(k_simi<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float4) * (size_t)Mcap));
}

void launch_mscale(osep_rosa_impl* h) {
    RosaDev& d = h->d;
    cudaStream_t s = h->stream;
    const osep_rosa_config& c = h->cfg;
    const int warp_blocks = div_up(d.Mcap, WARPS_PER_BLOCK);
    const int tb = WARPS_PER_BLOCK * 32;
    const int tblocks = div_up(d.Mcap, 128);
    // surf kNN + smoothing schedule need only the downsampled POSITIONS (phase 0 of the voxel reduce, ev_pts): they run on a
    // third stream while the kNN / normals branch is still busy, and have long finished when the similarity lists are done
    if (h->overlap) {
        cudaStream_t sk = h->stream3;
        cudaStreamWaitEvent(sk, h->ev_pts, 0);
        k_surf_knn<<<warp_blocks, tb, 0, sk>>>(d.st, d.pts, d.surf, d.nbk);   // R13
        launch_drosa_prep(h, sk);
        cudaEventRecord(h->ev_join3, sk);
    }
    // R12
    k_simi<false><<<warp_blocks, tb, sizeof(float4) * (size_t)d.Mcap, s>>>(d.st, d.pts, d.nrms, d.simi_off, d.simi_idx, nullptr, nullptr, d.simi_bits, d.Scap);   // count + bit rows; its last block scans
    // the similarity LISTS (CSR, in (d2, index) order) are only read by dcrosa: the fill pass runs on the side stream next to
    // the fused drosa kernel (which needs the bit rows of the count pass only) and is joined before dcrosa
    const bool simi_side = h->overlap && (h->simi_side >= 0 ? h->simi_side == 1 : h->fused_grid == 0);
    if (simi_side) {
        cudaEventRecord(h->ev_fork[1], s); cudaStreamWaitEvent(h->stream2, h->ev_fork[1], 0);
        k_simi<true><<<warp_blocks, tb, 0, h->stream2>>>(d.st, d.pts, d.nrms, d.simi_off, d.simi_idx, d.rv1, (float*)d.rk1, d.simi_bits, d.Scap);
        cudaEventRecord(h->ev_join[1], h->stream2);
    } else k_simi<true><<<warp_blocks, tb, 0, s>>>(d.st, d.pts, d.nrms, d.simi_off, d.simi_idx, d.rv1, (float*)d.rk1, d.simi_bits, d.Scap);
    stage_mark(h, SM_SIMI);
    if (h->overlap) cudaStreamWaitEvent(s, h->ev_join3, 0);
    else {                                                                     // profiling: sequential, attributable
        k_surf_knn<<<warp_blocks, tb, 0, s>>>(d.st, d.pts, d.surf, d.nbk);
        launch_drosa_prep(h, s);
    }
    launch_drosa(h, h->sms);
    k_poor_fixup<<<1, 1024, 0, s>>>(d.st, d.good, d.good_rank, d.centers, d.pset2, d.poor_idx, d.pts, d.pset);
    if (h->debug) cudaMemcpyAsync(d.centers, d.pset, sizeof(float4) * d.Mcap, cudaMemcpyDeviceToDevice, s);   // tap: pset after drosa
    stage_mark(h, SM_POSITION);
    if (simi_side) cudaStreamWaitEvent(s, h->ev_join[1], 0);                   // similarity lists
    // R20 (+ R21 in the same cluster launch when its per-point arrays fit next to the pset copies)
    bool vs_done = false;
    if ((size_t)32 * d.Mcap <= (size_t)DC_SMEM_BUDGET) {
        long long* dc_dbg = h->fused_detail ? d.fused_dbg + 64 : nullptr;
        const int cl = h->dc_cluster > 0 ? h->dc_cluster : ((h->fused_grid == 0 && h->dc16_ok) ? 16 : 8);
        VsArgs vs{};
        vs.on = ((size_t)60 * d.Mcap + 64 <= (size_t)DC_SMEM_BUDGET && !h->vs_separate) ? 1 : 0;
        vs.pts = d.pts; vs.alpha = c.alpha_recenter; vs.adj = d.adj_bits; vs.seeds = d.seeds; vs.corresp = d.corresp; vs.skel = d.skel; vs.skel_sampled = d.skel_sampled;
        vs_done = vs.on != 0;
        if (cl == 16) launch_dcrosa_cluster<16>(d.st, d.pset, d.simi_off, d.simi_idx, c.niter_dcrosa, dc_dbg, vs, s);
        else launch_dcrosa_cluster<8>(d.st, d.pset, d.simi_off, d.simi_idx, c.niter_dcrosa, dc_dbg, vs, s);
    } else {
        for (int it = 0; it < c.niter_dcrosa; ++it) {
            k_dcrosa_mean<<<tblocks, 128, 0, s>>>(d.st, d.pset, d.pset2, d.simi_off, d.simi_idx);
            k_dcrosa_line<<<tblocks, 128, 0, s>>>(d.st, d.pset2, d.pset, d.simi_off, d.simi_idx);
        }
    }
    stage_mark(h, SM_DCROSA);
    // R21 as its own launches (large arenas, OSEP_VS_SEPARATE=1)
    if (!vs_done) {
        k_vs_compact<<<1, 1024, 0, s>>>(d.st, d.pset, d.pcloud, d.good_rank);
        k_vs_adj<<<warp_blocks, tb, 0, s>>>(d.st, d.pcloud, d.adj_bits, d.W);
        k_vs_greedy<<<1, 1024, VS_SMEM_BUDGET, s>>>(d.st, d.adj_bits, d.W, d.seeds);
        k_vs_corresp<<<tblocks, 128, 0, s>>>(d.st, d.pcloud, d.seeds, d.corresp);
        k_vs_recenter<<<warp_blocks, tb, 0, s>>>(d.st, d.pts, d.pset, d.seeds, d.corresp, c.alpha_recenter, d.skel, d.skel_sampled);
    }
    stage_mark(h, SM_VSAMPLE);
    // R22 vertex_smooth runs in the frame's tail kernel (graph.cu frame_tail_enqueue)
    h->n_launches += 2 + 1 + 1 + 1 + (vs_done ? 0 : 5);   // + launch_drosa (dcrosa + vertex_sampling = one cluster kernel in the common case)
}

}  // namespace osep
