// csrc/tail.cuh -- the last stages of a frame as BLOCK-LEVEL device functions:
//   R22 vertex_smooth (rosa.cpp:534-620), G3 graph_adj + G4 mst (gskel.cpp:201-281) and the result export.
// They are latency-bound single-CTA work over V ~ 10^2 vertices; as four launches they cost 96 us of a 1.13 ms frame
// (ncu: 32 + 12 + 47 + 5), most of it cold global-memory round trips.  k_frame_tail (graph.cu) runs them back to back in
// ONE launch with the vertex positions, neighbour lists, edge keys and union-find in shared memory; the stand-alone
// kernels (GSkel ticks, seam graphs of tiled maps, frames without a graph) call the same functions.
// Every workspace pointer may point to shared memory (when the frame's vertex set fits the arena) or to the global scratch
// arrays -- same code, same results.  Nothing here is read through the non-coherent path: the inputs of one phase are the
// outputs of the previous one in the same launch.
#pragma once
#include "osep_common.cuh"

namespace osep {

constexpr unsigned TAIL_FULL = 0xffffffffu;
constexpr int GK = 10;                 // kNN size of graph_adj (gskel.cpp:206)
constexpr int GACC = GK - 1;           // accepted neighbours per vertex (rank 0 = self is skipped)
constexpr int SMOOTH_WIN = 32;         // vertex_smooth: window of a vertex's radius list in the single-pass path
constexpr int SMOOTH_STRIDE = 33;      // its stride in shared memory (odd: the Jacobi passes read one window per thread)
struct GraphDevState { int n_edges; int n_mst; int overflow; int pad; };

__device__ __forceinline__ f3 ld3(const float4* a, int i) { const float4 v = a[i]; return {v.x, v.y, v.z}; }
__device__ __forceinline__ void st3(float4* a, int i, const f3& v, float w = 0.0f) { a[i] = make_float4(v.x, v.y, v.z, w); }

// block-wide exclusive scan of one int per thread-strided element (single block helper, in place)
static __device__ int block_scan_inplace(int* data, int n) {
    __shared__ int wsum[32];
    __shared__ int s_carry, s_tile;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + tid;
        const int v = (i < n) ? data[i] : 0;
        int incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(TAIL_FULL, incl, off); if (lane >= off) incl += t; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int ws = (lane < nw) ? wsum[lane] : 0;
            int wi = ws;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(TAIL_FULL, wi, off); if (lane >= off) wi += t; }
            wsum[lane] = wi - ws;
            if (lane == 31) s_tile = wi;
        }
        __syncthreads();
        if (i < n) data[i] = s_carry + wsum[wid] + incl - v;
        __syncthreads();
        if (tid == 0) s_carry += s_tile;
        __syncthreads();
    }
    return s_carry;
}

__host__ __device__ __forceinline__ size_t tail_align16(size_t b) { return (b + 15) & ~(size_t)15; }
// bytes of the arena the vertex positions of vertex_smooth_block occupy (cur | nxt | compacted | offsets); what follows is free
// for the lists and, afterwards, for the graph phases
__host__ __device__ __forceinline__ size_t smooth_pos_bytes(int V) { return tail_align16((size_t)48 * V + 4 * ((size_t)V + 1)); }

// ------------------------------------------------------------------------------------------
// R22 vertex_smooth (rosa.cpp:534-620) by one block: finite compaction, radius neighbours (sorted (d2, index), self
// forced first), niter Jacobi PCA-line projections.  Caller guarantees status == 0 and V > 0.
// Returns the final positions (a pointer into the arena when they were kept there, else skel); nullptr after a capacity
// failure (status set).  skel holds the result in either case.
static __device__ const float4* vertex_smooth_block(FrameState* st, float4* skel, float4* skel2, float4* vtmp, int* vs_off, int* vs_idx,
                                                    int* tmp_j, float* tmp_d, int cap, float radius, float alpha, int niter,
                                                    unsigned char* arena, size_t arena_bytes, long long* dbg = nullptr) {
    int dk = 8;
    auto stamp = [&]() { if (dbg && threadIdx.x == 0 && dk < 16) { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[dk++] = t; } };
    const int V = st->V;
    const size_t pos_bytes = smooth_pos_bytes(V);
    const bool small = arena != nullptr && pos_bytes + 12 * 64 <= arena_bytes;
    float4* A = skel; float4* B = skel2; float4* C = vtmp; int* off = vs_off;
    int ecap_s = 0; int* s_lists = nullptr;
    if (small) {
        A = reinterpret_cast<float4*>(arena); B = A + V; C = B + V; off = reinterpret_cast<int*>(C + V);
        s_lists = reinterpret_cast<int*>(arena + pos_bytes);
        ecap_s = (int)((arena_bytes - pos_bytes) / 12);
    }
    // finite compaction (rosa.cpp:543-552)
    for (int i = threadIdx.x; i < V; i += blockDim.x) {
        const float4 v = skel[i];
        off[i] = (fin(v.x) && fin(v.y) && fin(v.z)) ? 1 : 0;
        if (small) A[i] = v;
        B[i] = v;
    }
    if (threadIdx.x == 0) off[V] = 0;
    __syncthreads();
    const int Vf = block_scan_inplace(off, V + 1);
    __syncthreads();
    for (int i = threadIdx.x; i < V; i += blockDim.x) { const float4 v = A[i]; if (fin(v.x) && fin(v.y) && fin(v.z)) C[off[i]] = v; }
    __syncthreads();
    stamp();
    const float r2 = radius * radius;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // Frames (V ~ 10^2, lists of ~10): ONE pass per vertex into a fixed window of SMOOTH_WIN entries in shared memory -- no
    // count pass, no offset scan (15 -> 7.6 us of a single-SM kernel that is bound by its instruction count).
    // A list that does not fit its window sends the whole block to the general two-pass CSR path below.
    bool fixed = false;
    int* idx = vs_idx;
    if (small && (size_t)Vf * SMOOTH_STRIDE * 12 <= arena_bytes - pos_bytes) {
        int* fidx = s_lists; int* ftj = fidx + (size_t)Vf * SMOOTH_STRIDE; float* ftd = reinterpret_cast<float*>(ftj + (size_t)Vf * SMOOTH_STRIDE);
        bool over = false;
        for (int i = warp; i < Vf; i += nwarps) {
            const f3 q = ld3(C, i);
            const int b = i * SMOOTH_STRIDE;
            int c = 0;
            for (int j0 = 0; j0 < Vf; j0 += 32) {
                const int j = j0 + lane;
                const float d = dist2(q, ld3(C, min(j, Vf - 1)));
                const bool in = j < Vf && d < r2;
                const unsigned m = __ballot_sync(TAIL_FULL, in);
                const int pos = c + __popc(m & ((1u << lane) - 1u));
                if (in && pos < SMOOTH_WIN) { ftj[b + pos] = j; ftd[b + pos] = d; }
                c += __popc(m);
            }
            __syncwarp();
            if (c >= SMOOTH_WIN) { over = true; continue; }          // (one more entry when the vertex itself has to be put in front)
            if (c == 0) { if (lane == 0) { fidx[b] = i; off[i] = 1; } continue; }
            // rank sort by (d2, index): lane e holds entry e; if the first sorted entry is not i, i is inserted in front (rosa.cpp:567-569)
            int rank = 0, je = 0x7fffffff;
            if (lane < c) {
                const float de = ftd[b + lane]; je = ftj[b + lane];
#pragma unroll 4
                for (int f = 0; f < c; ++f) rank += nb_less(ftd[b + f], ftj[b + f], de, je) ? 1 : 0;
            }
            const int first = __reduce_min_sync(TAIL_FULL, (lane < c && rank == 0) ? je : 0x7fffffff);
            const int shift = (first != i) ? 1 : 0;
            if (shift && lane == 0) fidx[b] = i;
            if (lane < c) fidx[b + shift + rank] = je;
            if (lane == 0) off[i] = c + shift;                       // fixed windows: off holds the list LENGTH
        }
        fixed = __syncthreads_or(over ? 1 : 0) == 0;
        if (fixed) idx = fidx;
    }
    if (!fixed) {
        // neighbour counts: one warp per vertex, ballots over 32 candidates at a time; `first` = smallest (d2, index)
        for (int i = warp; i < Vf; i += nwarps) {
            const f3 q = ld3(C, i);
            int c = 0; float bd = __int_as_float(0x7f800000); int bi = 0x7fffffff;
            for (int j0 = 0; j0 < Vf; j0 += 32) {
                const int j = j0 + lane;
                float d = 0.0f; bool in = false;
                if (j < Vf) { d = dist2(q, ld3(C, j)); in = d < r2; }
                c += __popc(__ballot_sync(TAIL_FULL, in));
                if (in && nb_less(d, j, bd, bi)) { bd = d; bi = j; }
            }
    #pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const float od = __shfl_xor_sync(TAIL_FULL, bd, o); const int oi = __shfl_xor_sync(TAIL_FULL, bi, o);
                if (nb_less(od, oi, bd, bi)) { bd = od; bi = oi; }
            }
            if (c == 0) c = 1; else if (bi != i) c += 1;      // self forced first (rosa.cpp:567-573)
            __syncwarp();
            if (lane == 0) off[i] = c;
        }
        __syncthreads();                                       // (the offsets of the compaction are dead: off is reused for the list offsets)
        if (threadIdx.x == 0) off[Vf] = 0;
        __syncthreads();
        const int total = block_scan_inplace(off, Vf + 1);
        __syncthreads();
        stamp();
        if (total > cap) { if (threadIdx.x == 0) st->status = OSEP_ERR_CAPACITY; return nullptr; }
        idx = vs_idx; int* tj = tmp_j; float* td = tmp_d;
        if (small && total <= ecap_s) { idx = s_lists; tj = idx + ecap_s; td = reinterpret_cast<float*>(tj + ecap_s); }
        for (int i = warp; i < Vf; i += nwarps) {
            const f3 q = ld3(C, i);
            const int b = off[i];
            int c = 0;
            for (int j0 = 0; j0 < Vf; j0 += 32) {
                const int j = j0 + lane;
                float d = 0.0f; bool in = false;
                if (j < Vf) { d = dist2(q, ld3(C, j)); in = d < r2; }
                const unsigned m = __ballot_sync(TAIL_FULL, in);
                if (in) { const int pos = b + c + __popc(m & ((1u << lane) - 1u)); tj[pos] = j; td[pos] = d; }
                c += __popc(m);
            }
            __syncwarp();
            if (c == 0) { if (lane == 0) idx[b] = i; continue; }
            // rank sort by (d2, index); if the first sorted entry is not i, i is inserted in front (rosa.cpp:567-569)
            int myfirst = 0x7fffffff;
            int rank_e[4];                                     // ranks of this lane's entries e = lane + 32 u, u < 4 (longer lists recompute)
    #pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = lane + 32 * u;
                rank_e[u] = 0;
                if (e < c) {
                    const float de = td[b + e]; const int je = tj[b + e];
                    int rank = 0;
                    for (int f = 0; f < c; ++f) rank += nb_less(td[b + f], tj[b + f], de, je) ? 1 : 0;
                    rank_e[u] = rank;
                    if (rank == 0) myfirst = je;
                }
            }
            for (int e = lane + 128; e < c; e += 32) {
                int rank = 0;
                for (int f = 0; f < c; ++f) rank += nb_less(td[b + f], tj[b + f], td[b + e], tj[b + e]) ? 1 : 0;
                if (rank == 0) myfirst = tj[b + e];
            }
            const int first = __reduce_min_sync(TAIL_FULL, myfirst);
            const int shift = (first != i) ? 1 : 0;
            if (shift && lane == 0) idx[b] = i;
    #pragma unroll
            for (int u = 0; u < 4; ++u) { const int e = lane + 32 * u; if (e < c) idx[b + shift + rank_e[u]] = tj[b + e]; }
            for (int e = lane + 128; e < c; e += 32) {
                int rank = 0;
                for (int f = 0; f < c; ++f) rank += nb_less(td[b + f], tj[b + f], td[b + e], tj[b + e]) ? 1 : 0;
                idx[b + shift + rank] = tj[b + e];
            }
        }
    }
    __syncthreads();
    stamp();
    float4* cur = A; float4* nxt = B;
    for (int it = 0; it < niter; ++it) {
        for (int i = threadIdx.x; i < Vf; i += blockDim.x) {
            const int b = fixed ? i * SMOOTH_STRIDE : off[i];
            const int n = fixed ? off[i] : off[i + 1] - off[i];
            const int e = b + n;
            if (n <= 1) { nxt[i] = cur[i]; continue; }
            f3 mean{0.0f, 0.0f, 0.0f};
            for (int t = b; t < e; ++t) mean = add3(mean, ld3(cur, idx[t]));
            mean = div3(mean, (float)n);
            float c00 = 0, c10 = 0, c11 = 0, c20 = 0, c21 = 0, c22 = 0;
            for (int t = b; t < e; ++t) {
                const f3 d = sub3(ld3(cur, idx[t]), mean);
                c00 += d.x * d.x; c10 += d.y * d.x; c11 += d.y * d.y; c20 += d.z * d.x; c21 += d.z * d.y; c22 += d.z * d.z;
            }
            const float fn = (float)n;
            c00 /= fn; c10 /= fn; c11 /= fn; c20 /= fn; c21 /= fn; c22 /= fn;
            const Eig3 E = eig3_sym_dev(c00, c10, c11, c20, c21, c22);
            const f3 dir = eig_col(E, 2);
            const f3 p = ld3(cur, i);
            const f3 delta = sub3(p, mean);
            const f3 proj = add3(mean, mul3(dir, dot3(dir, delta)));
            st3(nxt, i, add3(mul3(p, alpha), mul3(proj, 1.0f - alpha)), 1.0f);
        }
        __syncthreads();
        stamp();
        float4* t = cur; cur = nxt; nxt = t;
    }
    // RD.skelver = cur
    if (cur != skel) { for (int i = threadIdx.x; i < V; i += blockDim.x) skel[i] = cur[i]; }
    if (threadIdx.x == 0) st->Vf = Vf;
    return cur;
}

// ------------------------------------------------------------------------------------------
// G3 graph_adj (gskel.cpp:201-246): brute-force kNN(10) by (d2, index) + RNG test restricted to the kNN list.
// One THREAD per vertex (vertices first, first + stride, ...).  The thread scans all vertices in ascending index (every
// thread of a warp reads the same position: one broadcast load) and keeps its ten smallest (distance, index) pairs sorted
// in registers: a candidate below the tenth distance is inserted by ten independent compare / select groups.  Candidates
// arrive in ascending index, so a strict '<' on the distance alone is exactly the (d2, index) order.  Then the
// relative-neighbourhood test of gskel.cpp:221-235 over the 36 member pairs (the distance of the vertex to a member is
// computed once).
// The single-CTA frame tail is bound by the issue rate of ONE SM, so what counts is the number of warp instructions:
// warp per vertex with ten rounds of warp arg-min (shuffle butterflies) 55 us for 124 vertices, with REDUX 37 us (the unit
// serialises the 32 warps), four lanes per vertex + butterfly merge of 64-bit keys 13.8 us (16 warps x 5 k instructions).
static __constant__ unsigned char GADJ_PA[36] = {1,1,1,1,1,1,1,1, 2,2,2,2,2,2,2, 3,3,3,3,3,3, 4,4,4,4,4, 5,5,5,5, 6,6,6, 7,7, 8};
static __constant__ unsigned char GADJ_PB[36] = {2,3,4,5,6,7,8,9, 3,4,5,6,7,8,9, 4,5,6,7,8,9, 5,6,7,8,9, 6,7,8,9, 7,8,9, 8,9, 9};
static_assert(GK == 10, "pair tables of graph_adj_thread are written for 9 members");

static __device__ __forceinline__ void graph_adj_thread(const float4* pos, int G, int first, int stride, float max_dist_th, int* acc, int* acc_cnt) {
    for (int i = first; i < G; i += stride) {
        const float4 pi4 = pos[i];
        const f3 pi{pi4.x, pi4.y, pi4.z};
        float kd[GK]; int kj[GK];
        const float INF = __int_as_float(0x7f800000);
#pragma unroll
        for (int s = 0; s < GK; ++s) { kd[s] = INF; kj[s] = -1; }                   // index -1: "no such neighbour"
        int filled = 0;                                                             // an infinite distance is a valid pick while the list is short
        float4 q = pos[0];
        for (int j = 0; j < G; ++j) {
            const float dd = dist2(pi.x, pi.y, pi.z, q.x, q.y, q.z);
            q = pos[min(j + 1, G - 1)];                                              // next candidate in flight during the insertion
            if (dd < kd[GK - 1] || (dd == INF && filled < GK)) {                    // (a NaN distance fails both tests: never picked)
                bool c[GK];
#pragma unroll
                for (int s = 0; s < GK; ++s) c[s] = dd < kd[s] || (kj[s] < 0);      // empty slots sort after everything
#pragma unroll
                for (int s = GK - 1; s >= 1; --s) { kd[s] = c[s] ? (c[s - 1] ? kd[s - 1] : dd) : kd[s]; kj[s] = c[s] ? (c[s - 1] ? kj[s - 1] : j) : kj[s]; }
                kd[0] = c[0] ? dd : kd[0]; kj[0] = c[0] ? j : kj[0];
                filled = min(filled + 1, GK);
            }
        }
        // members 1..9 (rank 0 is skipped like gskel.cpp:217): position and distance to vertex i
        float4 mem[GK];
        unsigned okmask = 0;
#pragma unroll
        for (int m = 1; m < GK; ++m) {
            const bool ok = kj[m] >= 0;
            const float4 q4 = pos[ok ? kj[m] : i];
            mem[m] = make_float4(q4.x, q4.y, q4.z, norm3(sub3(pi, f3{q4.x, q4.y, q4.z})));
            okmask |= ok ? (1u << m) : 0u;
        }
        // member a is dropped when another member b is closer to both i and a than i and a are to each other
        unsigned killed = 0;
#pragma unroll 6
        for (int p = 0; p < 36; ++p) {
            const int a = GADJ_PA[p], b = GADJ_PB[p];
            const float4 A = mem[a], B = mem[b];
            const float d_ab = norm3(sub3(f3{A.x, A.y, A.z}, f3{B.x, B.y, B.z}));
            const bool both = ((okmask >> a) & (okmask >> b) & 1u) != 0u;
            if (both && d_ab < A.w && B.w < A.w) killed |= 1u << a;
            if (both && d_ab < B.w && A.w < B.w) killed |= 1u << b;
        }
        int cnt = 0;
#pragma unroll
        for (int m = 1; m < GK; ++m) {
            const bool good = ((okmask >> m) & 1u) && !(mem[m].w > max_dist_th) && !((killed >> m) & 1u);
            if (good) { acc[i * GACC + cnt] = kj[m]; ++cnt; }                       // ascending rank = the reference's push order
        }
        acc_cnt[i] = cnt;
    }
}

// ------------------------------------------------------------------------------------------
// G4 mst (gskel.cpp:248-281) by one block.  Edge keys = (weight bits << 32 | u << 16 | v) for every accepted pair, u < v
// (the reference's "nb <= i" skip, gskel.cpp:254); bitonic sort = the canonical (w,u,v) order (C6); duplicates (i accepts
// nb and nb accepts i) sit next to each other and the second is ignored exactly as the reference's second unite() fails.
// Under a strict total order of the edges the minimum spanning forest is UNIQUE, so the edges Kruskal's loop accepts
// (gskel.cpp:258-278) are that forest's edges in key order, whatever algorithm finds them: the block runs Boruvka rounds
// (every component takes its smallest outgoing edge -- the smallest SORTED POSITION, a 32-bit atomicMin --, hooks onto the
// component at the other end, the two ends of a mutual choice keep the smaller label as root, labels are refreshed by
// pointer chasing; <= log2 G rounds of a few hundred cycles), marks the forest's edges in bit 63 of their keys (weights
// are non-negative) and a stable compaction writes them in key order.  (The first version ran Kruskal on one warp: 72 us
// for 124 vertices.)
// skeys / swork: shared-memory workspaces (capacities in entries), used when the frame fits; gkeys / gwork: global ones
// (gwork: 3 G ints).
static __device__ void graph_mst_block(const float4* pos, int G, const int* acc, const int* acc_cnt,
                                       unsigned long long* skeys, int skeys_cap, unsigned long long* gkeys, int keycap,
                                       int* swork, int swork_cap, int* gwork, int* mst_uv, float* mst_w, GraphDevState* gs, long long* dbg) {
    __shared__ int s_n, s_tot, s_added, s_base;
    __shared__ int s_wcnt[32];
    auto stamp = [&](int k) { if (dbg && threadIdx.x == 0) { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); dbg[k] = t; } };
    if (threadIdx.x == 0) { s_n = 0; s_tot = 0; s_base = 0; }
    __syncthreads();
    {
        int local = 0;
        for (int i = threadIdx.x; i < G; i += blockDim.x) local += acc_cnt[i];
        local = __reduce_add_sync(TAIL_FULL, local);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_tot, local);
    }
    __syncthreads();
    int P2t = 1; while (P2t < s_tot) P2t <<= 1;
    unsigned long long* keys = (P2t <= skeys_cap) ? skeys : gkeys;
    int* work = (3 * G <= swork_cap) ? swork : gwork;
    int* comp = work; int* parent = work + G; int* best = work + 2 * G;
    // one thread per (vertex, accepted slot); a warp reserves its key slots with one shared-memory atomic
    for (int e0 = 0; e0 < G * GACC; e0 += blockDim.x) {
        const int e = e0 + threadIdx.x;
        bool have = false; int u = 0, v = 0;
        if (e < G * GACC) {
            const int i = e / GACC, t = e - i * GACC;
            if (t < acc_cnt[i]) { const int nb = acc[e]; u = min(i, nb); v = max(i, nb); have = u != v; }
        }
        const unsigned hm = __ballot_sync(TAIL_FULL, have);
        int base = 0;
        if ((threadIdx.x & 31) == 0 && hm) base = atomicAdd(&s_n, __popc(hm));
        base = __shfl_sync(TAIL_FULL, base, 0);
        if (have) {
            const float4 pu = pos[u], pv = pos[v];
            // weight = (ver_i - ver_nb).norm() with i = u (the smaller index owns the edge in gskel.cpp:252-263)
            const float w = norm3(sub3(f3{pu.x, pu.y, pu.z}, f3{pv.x, pv.y, pv.z}));
            const int slot = base + __popc(hm & ((1u << (threadIdx.x & 31)) - 1u));
            if (slot < keycap) keys[slot] = ((unsigned long long)__float_as_uint(w) << 32) | ((unsigned long long)u << 16) | (unsigned long long)v;
        }
    }
    __syncthreads();
    stamp(0);
    const int n = s_n;
    if (n > keycap) { if (threadIdx.x == 0) { gs->overflow = 1; gs->n_edges = 0; gs->n_mst = 0; gs->pad = G; } return; }
    int P2 = 1; while (P2 < n) P2 <<= 1;
    for (int i = n + threadIdx.x; i < P2; i += blockDim.x) keys[i] = ~0ull;
    for (int i = threadIdx.x; i < G; i += blockDim.x) { comp[i] = i; parent[i] = i; }
    __syncthreads();
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool up = ((i & k) == 0);
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    stamp(1);
    constexpr unsigned long long IN_MST = 1ull << 63, KEY = ~IN_MST;
    for (int round = 0; round < 32; ++round) {
        for (int c = threadIdx.x; c < G; c += blockDim.x) best[c] = 0x7fffffff;
        if (threadIdx.x == 0) s_added = 0;
        __syncthreads();
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const unsigned long long key = keys[e];
            if (key & IN_MST) continue;                                         // its ends are in one component already
            if (e > 0 && (keys[e - 1] & KEY) == key) continue;                  // the same pair accepted from both ends
            const int cu = comp[(int)((key >> 16) & 0xffffu)], cv = comp[(int)(key & 0xffffu)];
            if (cu != cv) { atomicMin(&best[cu], e); atomicMin(&best[cv], e); }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < G; c += blockDim.x) {
            const int e = best[c];
            if (e == 0x7fffffff) continue;                                      // (only current roots receive a best edge)
            const unsigned long long key = keys[e];
            const int cu = comp[(int)((key >> 16) & 0xffffu)], cv = comp[(int)(key & 0xffffu)];
            const int other = (cu == c) ? cv : cu;
            if (!(best[other] == e && c < other)) parent[c] = other;            // a mutual choice keeps the smaller label as the root
            keys[e] = key | IN_MST;                                             // (both ends of a mutual choice store the same value)
            s_added = 1;
        }
        __syncthreads();
        if (!s_added) break;
        for (int v = threadIdx.x; v < G; v += blockDim.x) { int r = comp[v]; while (parent[r] != r) r = parent[r]; comp[v] = r; }
        __syncthreads();
    }
    // the forest's edges in key order
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        const int e = base + threadIdx.x;
        const unsigned long long key = e < n ? keys[e] : 0ull;
        const bool in = (key & IN_MST) != 0ull;
        const unsigned bm = __ballot_sync(TAIL_FULL, in);
        if (lane == 0) s_wcnt[wid] = __popc(bm);
        __syncthreads();
        int before = s_base, total = 0;
        for (int w = 0; w < nw; ++w) { const int c = s_wcnt[w]; if (w < wid) before += c; total += c; }
        if (in) {
            const int pos_e = before + __popc(bm & ((1u << lane) - 1u));
            mst_uv[2 * pos_e] = (int)((key >> 16) & 0xffffu); mst_uv[2 * pos_e + 1] = (int)(key & 0xffffu);
            mst_w[pos_e] = __uint_as_float((unsigned)((key & KEY) >> 32));
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) { gs->n_edges = n; gs->n_mst = s_base; gs->overflow = 0; gs->pad = G; }
}

// ------------------------------------------------------------------------------------------
// One result block per frame: [FrameState | first Mq vertices | GraphDevState | acc_cnt[Mq] | acc[Mq*9] | mst_uv[2 Mq] | mst_w[Mq]].
// A single block gathers it at the end of the frame, so the frame needs ONE device-to-host copy node instead of seven.
__host__ __device__ inline size_t exp_off_skel() { return 512; }
__host__ __device__ inline size_t exp_off_graph(int Mq) { return exp_off_skel() + sizeof(float4) * (size_t)Mq; }
__host__ __device__ inline size_t exp_bytes(int Mq, bool with_graph) { return exp_off_graph(Mq) + (with_graph ? 16 + (size_t)Mq * (4 + 4 * GACC + 8 + 4) : 0); }
static __device__ void export_block(const FrameState* st, const float4* skel, int Mq, const GraphDevState* gs, const int* acc_cnt, const int* acc,
                                    const int* mst_uv, const float* mst_w, unsigned char* out) {
    static_assert(sizeof(FrameState) <= 512, "FrameState outgrew its slot of the export block");
    const int* s32 = reinterpret_cast<const int*>(st); int* o32 = reinterpret_cast<int*>(out);
    for (int i = threadIdx.x; i < (int)(sizeof(FrameState) / 4); i += blockDim.x) o32[i] = s32[i];
    const int V = (st->status == 0 || st->status == OSEP_STAGE_7) ? min(st->V, Mq) : 0;
    float4* os = reinterpret_cast<float4*>(out + exp_off_skel());
    for (int i = threadIdx.x; i < V; i += blockDim.x) os[i] = skel[i];
    if (!gs) return;
    int* og = reinterpret_cast<int*>(out + exp_off_graph(Mq));
    if (threadIdx.x < 4) og[threadIdx.x] = reinterpret_cast<const int*>(gs)[threadIdx.x];
    const int G = min(gs->pad, Mq), nm = min(gs->n_mst, Mq);
    int* o_cnt = og + 4; int* o_acc = o_cnt + Mq; int* o_uv = o_acc + (size_t)Mq * GACC; float* o_w = reinterpret_cast<float*>(o_uv + 2 * (size_t)Mq);
    for (int i = threadIdx.x; i < G; i += blockDim.x) o_cnt[i] = acc_cnt[i];
    for (int i = threadIdx.x; i < G * GACC; i += blockDim.x) o_acc[i] = acc[i];
    for (int i = threadIdx.x; i < 2 * nm; i += blockDim.x) o_uv[i] = mst_uv[i];
    for (int i = threadIdx.x; i < nm; i += blockDim.x) o_w[i] = mst_w[i];
}

}  // namespace osep
