// csrc/drosa.cu -- the drosa stage (rosa.cpp:258-380) on sm_100a: R14-R19.
//
// Data layout.  M ~ 1000 points; everything a seed touches is staged in shared memory once
// per CTA (one CTA per SM, 148 CTAs):
//   * the similarity graph as an M x ceil(M/32) BIT MATRIX (row i = simi_nbs[i]; 128 KB at
//     M = 1003) -- the active set of a seed (connected component of the seed inside its slab,
//     rosa.cpp:653-683) is then a bit-parallel BFS: one 32-bit word of slab / visited /
//     frontier per LANE in registers, one LDS + OR per visited node, no atomics, no queue;
//   * points, raw normals and the per-point unit normals (hoisted: the reference recomputes
//     v / |v| inside every sum, rosa.cpp:693-696 -- same value every time).
// One warp per seed; float sums run in the canonical lane-owner order (C4): lane l owns the
// members of words w = l (mod 32), ascending, then an xor butterfly.
// The in-place smoothing pass (Gauss-Seidel, rosa.cpp:295-308) is a dependent chain: one warp
// walks its dependency DAG with all operands in shared memory (k_drosa_fused: a pull-model
// walker over a modelled schedule, six chains as overlapping wavefronts; k_drosa_smooth2: the
// level-synchronous fallback).
#include "rosa_impl.cuh"
#include <type_traits>

namespace osep {

constexpr unsigned FULLM = 0xffffffffu;
#ifndef OSEP_SEED_WARPS
#define OSEP_SEED_WARPS 8
#endif
constexpr int SEED_WARPS = OSEP_SEED_WARPS;
constexpr int SEED_SMEM_BUDGET = 226 * 1024;       // unfused seed kernels and the upper bound of the fused kernel's request
constexpr int FUSED_SMEM_SINGLE = 208 * 1024;      // fused kernel, one frame at a time: leaves ~19 KB of every SM for the small kernels that run next to the persistent one (similarity list fill)
constexpr int GS_THREADS = 256;
constexpr int GS_SMEM_BUDGET = 200 * 1024;
constexpr int MAX_FUSED_ITERS = 8;

__device__ __forceinline__ f3 ld3f(const float4* a, int i) { const float4 v = a[i]; return {v.x, v.y, v.z}; }

template <int K> __device__ __forceinline__ void bfly(float (&v)[K]) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = v[k] + __shfl_xor_sync(FULLM, v[k], off);
    }
}

// per-point normal quantities used inside the active-set sums, computed once per frame:
//   nraw[i] = (n, il_dd)  il_dd = (s2 > 1e-20f) ? (float)(1.0 / (double)sqrtf(s2)) : 0      rosa.cpp:764-766
//   vnd[i]  = (n * ((double)s2 > 1e-20 ? 1/sqrtf(s2) : 0), il_f)                             rosa.cpp:694-696, 740-742
//             il_f = (s2 > 1e-20f) ? 1.0f / sqrtf(s2) : 0                                    rosa.cpp:340-341, 717-718
__global__ void k_normal_prep(const FrameState* __restrict__ st, const float4* __restrict__ nrms, float4* __restrict__ nraw, float4* __restrict__ vnd) {
    if (st->status != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st->M) return;
    float4 a, b2;
    normal_prep_vals(ld3f(nrms, i), a, b2);
    nraw[i] = a; vnd[i] = b2;
}

__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// shared-memory access by 32-bit shared-window address: LDS/STS instead of generic loads, 32-bit address arithmetic
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds_f4(unsigned a) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_u4(unsigned a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ unsigned lds_u16v(unsigned a) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u16v(unsigned a, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((unsigned short)v) : "memory"); }

// What a seed warp reads.  Points, raw normals and hoisted unit normals are always staged (48 bytes per point: 147 KB at the
// arena's 3072 points); the bit matrix is staged when it fits, with its rows padded to a multiple of 4 words so that a
// row is read with 16-byte loads, plus one member list per warp for the breadth-first search.
struct SeedCtx {
    unsigned pts_s, nraw_s, vnd_s;      // shared-window byte addresses of float4[M]
    unsigned bits_s; int Wp;            // staged bit matrix (0 = not staged), row stride in words
    unsigned list_s;                    // this warp's member list, ushort[M]
    const unsigned* bits; int bstride;  // the matrix in global memory (dense rows), used when it is not staged
    int M, Wn; float leaf;
};
__device__ __forceinline__ float4 seed_pt(const SeedCtx& c, int i) { return lds_f4(c.pts_s + 16u * (unsigned)i); }
__device__ __forceinline__ float4 seed_nraw(const SeedCtx& c, int i) { return lds_f4(c.nraw_s + 16u * (unsigned)i); }
__device__ __forceinline__ float4 seed_vnd(const SeedCtx& c, int i) { return lds_f4(c.vnd_s + 16u * (unsigned)i); }

__host__ __device__ __forceinline__ size_t seed_small_bytes(int M) { return (size_t)3 * 16 * M; }
__host__ __device__ __forceinline__ size_t seed_mat_bytes(int M) { const int Wp = (((M + 31) >> 5) + 3) & ~3; return (size_t)4 * M * Wp + (size_t)SEED_WARPS * (((size_t)2 * M + 15) & ~(size_t)15); }

// stage the per-frame read-only data of the seed warps; every thread of the CTA calls it (ends with __syncthreads)
__device__ __forceinline__ SeedCtx seed_stage(uint4* smem, const float4* __restrict__ pts, const float4* __restrict__ nraw, const float4* __restrict__ vnd,
                                              const unsigned* __restrict__ bits, int M, float leaf, int budget = SEED_SMEM_BUDGET) {
    const int Wn = (M + 31) >> 5, Wp = (Wn + 3) & ~3;
    float4* s_pts = reinterpret_cast<float4*>(smem);
    float4* s_nraw = s_pts + M; float4* s_vnd = s_nraw + M;
    unsigned* s_bits = reinterpret_cast<unsigned*>(s_vnd + M);
    const bool mat_in = seed_small_bytes(M) + seed_mat_bytes(M) <= (size_t)budget;
    for (int i = threadIdx.x; i < M; i += blockDim.x) { s_pts[i] = pts[i]; s_nraw[i] = nraw[i]; s_vnd[i] = vnd[i]; }
    if (mat_in) stage_bit_rows(s_bits, bits, M, Wp);          // rows are padded to Wp words in global memory too (k_simi)
    __syncthreads();
    SeedCtx c;
    c.pts_s = smem_addr(s_pts); c.nraw_s = smem_addr(s_nraw); c.vnd_s = smem_addr(s_vnd);
    c.bits_s = mat_in ? smem_addr(s_bits) : 0u; c.Wp = Wp;
    // one member list per warp: behind the staged matrix, or right behind the per-point arrays when the matrix stays in
    // global memory (0 = no room for lists either: the per-member loop over global rows)
    const size_t list_bytes = (size_t)SEED_WARPS * (((size_t)2 * M + 15) & ~(size_t)15);
    const bool list_in = mat_in || seed_small_bytes(M) + list_bytes <= (size_t)budget;
    c.list_s = list_in ? smem_addr(s_bits + (mat_in ? (size_t)M * Wp : 0)) + (unsigned)(threadIdx.x >> 5) * (unsigned)((2 * M + 15) & ~15) : 0u;
    c.bits = bits; c.bstride = Wp;
    c.M = M; c.Wn = Wn; c.leaf = leaf;
    return c;
}

// R14 compute_active_samples: returns |active|; vis[k] = this lane's word (lane + 32 k) of the member set
template <int WPL>
__device__ __forceinline__ int active_set(const SeedCtx& c, int seed, const f3& p, const f3& n, unsigned (&vis)[WPL], long long* tl = nullptr) {
    const int lane = threadIdx.x & 31;
    unsigned slab[WPL], front[WPL];
#pragma unroll
    for (int k = 0; k < WPL; ++k) slab[k] = 0u;
    // four words per round: the four loads are issued back to back, then the four slab tests, then the four votes (one
    // word at a time the loop was a chain of load -> 8 dependent FP operations -> vote per word)
    for (int t0 = 0; t0 < c.Wn; t0 += 4) {
        float4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = (t0 + u) * 32 + lane; q[u] = seed_pt(c, i < c.M ? i : 0); }
        bool in[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float d = n.x * (p.x - q[u].x) + n.y * (p.y - q[u].y) + n.z * (p.z - q[u].z);      // rosa.cpp:660
            in[u] = ((t0 + u) * 32 + lane < c.M) && (fabsf(d) < c.leaf);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned word = __ballot_sync(FULLM, in[u]);
            const int t = t0 + u;
#pragma unroll
            for (int k = 0; k < WPL; ++k) if (t == lane + 32 * k) slab[k] = word;
        }
    }
    if (tl && lane == 0) tl[7] = gtime();
    const int sw = seed >> 5; const unsigned sbit = 1u << (seed & 31);
    bool seed_in = false;
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
        const bool mine = (sw == lane + 32 * k);
        vis[k] = front[k] = (mine && (slab[k] & sbit)) ? sbit : 0u;
        seed_in |= (vis[k] != 0u);
    }
    if (!__any_sync(FULLM, seed_in)) return 0;                                              // rosa.cpp:664
    if (c.list_s) {
        // the frontier is listed in shared memory (warp scan of the per-lane popcounts), then LPR lanes read one member's
        // row with 16-byte loads -- 32 / LPR independent rows per round of loads instead of one dependent
        // ffs -> address -> load -> or chain per member.  Rows come from the staged matrix, or -- when it does not fit
        // (M > ~1100: frames whose leaf loop did not converge) -- from global memory through the same loop with twice the
        // loads in flight (an L2 round trip per row: serialised per member it made these tasks 44 us instead of 6)
        constexpr int LPR = 8 * WPL;                      // lanes per row: 4 * LPR words >= 32 * WPL >= Wn
        const int g = lane / LPR, ch = lane % LPR;
        const bool chv = 4 * ch < c.Wp;
        int f = 1;                                        // members in the list
        if (seed_in) sts_u16v(c.list_s, (unsigned)seed);
        __syncwarp();
        auto bfs = [&](auto load_row, auto unr) {
            constexpr int UNR = decltype(unr)::value;
            while (true) {
                uint4 a4 = make_uint4(0u, 0u, 0u, 0u);
                if (chv && UNR <= 4) {
                    // staged rows: short lists are the common case, one entry -> row pair per step
#pragma unroll UNR
                    for (int i = g; i < f; i += 32 / LPR) {
                        const unsigned mb = lds_u16v(c.list_s + 2u * (unsigned)i);
                        const uint4 r4 = load_row(mb);
                        a4.x |= r4.x; a4.y |= r4.y; a4.z |= r4.z; a4.w |= r4.w;
                    }
                } else if (chv) {
                    // rows in global memory: UNR members per round, their list entries read back to back, then their rows
                    // (the list loads are volatile asm statements and keep their order: entry -> row -> entry -> row would
                    // expose the latency of every entry in front of an L2 round trip).  Positions past the end repeat the
                    // last member: OR-ing a row twice changes nothing
                    for (int i = g; i < f; i += UNR * (32 / LPR)) {
                        unsigned mb[UNR];
#pragma unroll
                        for (int u = 0; u < UNR; ++u) { const int j = i + u * (32 / LPR); mb[u] = lds_u16v(c.list_s + 2u * (unsigned)(j < f ? j : f - 1)); }
#pragma unroll
                        for (int u = 0; u < UNR; ++u) { const uint4 r4 = load_row(mb[u]); a4.x |= r4.x; a4.y |= r4.y; a4.z |= r4.z; a4.w |= r4.w; }
                    }
                }
#pragma unroll
                for (int off = LPR; off < 32; off <<= 1) {
                    a4.x |= __shfl_xor_sync(FULLM, a4.x, off); a4.y |= __shfl_xor_sync(FULLM, a4.y, off);
                    a4.z |= __shfl_xor_sync(FULLM, a4.z, off); a4.w |= __shfl_xor_sync(FULLM, a4.w, off);
                }
                bool any = false; int cnt = 0;
#pragma unroll
                for (int k = 0; k < WPL; ++k) {
                    const int src = (lane + 32 * k) >> 2;     // lane `src` of group 0 holds chunk src = words 4 src .. 4 src + 3
                    const unsigned x = __shfl_sync(FULLM, a4.x, src), y = __shfl_sync(FULLM, a4.y, src);
                    const unsigned z = __shfl_sync(FULLM, a4.z, src), w = __shfl_sync(FULLM, a4.w, src);
                    const unsigned acc = (lane & 2) ? ((lane & 1) ? w : z) : ((lane & 1) ? y : x);
                    const unsigned nw = acc & slab[k] & ~vis[k];
                    vis[k] |= nw; front[k] = nw; any |= (nw != 0u); cnt += __popc(nw);
                }
                if (!__any_sync(FULLM, any)) break;
                // list the new frontier
                int incl = cnt;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(FULLM, incl, off); if (lane >= off) incl += t; }
                f = __shfl_sync(FULLM, incl, 31);
                unsigned pos = c.list_s + 2u * (unsigned)(incl - cnt);
#pragma unroll
                for (int k = 0; k < WPL; ++k) {
                    unsigned b = front[k];
                    const unsigned base = (unsigned)(lane + 32 * k) * 32u;
                    while (b) { sts_u16v(pos, base + (unsigned)(__ffs(b) - 1)); b &= b - 1; pos += 2u; }
                }
                __syncwarp();
            }
        };
        if (c.bits_s) {
            const unsigned rowb = c.bits_s + 16u * (unsigned)ch;
            const unsigned rstride = 4u * (unsigned)c.Wp;
            bfs([&](unsigned mb) { return lds_u4(rowb + mb * rstride); }, std::integral_constant<int, 4>{});
        } else {
            const uint4* grow = reinterpret_cast<const uint4*>(c.bits) + ch;      // rows are 16-byte aligned: bstride = Wp, a multiple of 4 words
            const unsigned gstride = (unsigned)c.bstride >> 2;
            bfs([&](unsigned mb) { return __ldg(grow + (size_t)mb * gstride); }, std::integral_constant<int, 8>{});
        }
    } else {
        while (true) {
            unsigned acc[WPL];
#pragma unroll
            for (int k = 0; k < WPL; ++k) acc[k] = 0u;
#pragma unroll
            for (int k = 0; k < WPL; ++k) {
                unsigned nz = __ballot_sync(FULLM, front[k] != 0u);
                while (nz) {
                    const int sl = __ffs(nz) - 1; nz &= nz - 1;
                    unsigned fw = __shfl_sync(FULLM, front[k], sl);
                    const int nbase = (sl + 32 * k) * 32;
                    while (fw) {
                        const int b = __ffs(fw) - 1; fw &= fw - 1;
                        const unsigned* row = c.bits + (size_t)(nbase + b) * c.bstride;
#pragma unroll
                        for (int kk = 0; kk < WPL; ++kk) { const int w = lane + 32 * kk; if (w < c.Wn) acc[kk] |= __ldg(row + w); }
                    }
                }
            }
            bool any = false;
#pragma unroll
            for (int k = 0; k < WPL; ++k) { const unsigned nw = acc[k] & slab[k] & ~vis[k]; vis[k] |= nw; front[k] = nw; any |= (nw != 0u); }
            if (!__any_sync(FULLM, any)) break;
        }
    }
    int m = 0;
#pragma unroll
    for (int k = 0; k < WPL; ++k) m += __popc(vis[k]);
    return __reduce_add_sync(FULLM, m);
}

template <int WPL, class F> __device__ __forceinline__ void for_members(const unsigned (&vis)[WPL], F body) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
        unsigned bits = vis[k];
        const int base = (lane + 32 * k) * 32;
        while (bits) { const int i = base + (__ffs(bits) - 1); bits &= bits - 1; body(i); }
    }
}

// sums for the normal covariance (rosa.cpp:689-699, 735-745)
template <int WPL> __device__ __forceinline__ void normal_cov_sums(const SeedCtx& c, const unsigned (&vis)[WPL], float (&s)[9]) {
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] = 0.0f;
    for_members<WPL>(vis, [&](int i) {
        const float4 u = seed_vnd(c, i);
        s[0] += u.x; s[1] += u.y; s[2] += u.z;
        s[3] += u.x * u.x; s[4] += u.y * u.x; s[5] += u.y * u.y;
        s[6] += u.z * u.x; s[7] += u.z * u.y; s[8] += u.z * u.z;
    });
    bfly<9>(s);
}
// covariance = sum_outer * inv_m - mu mu^T (rosa.cpp:701-703, 747-749), lower triangle -> eigen
__device__ __forceinline__ Eig3 normal_cov_eig(const float (&s)[9], float inv_m) {
    const float mx = s[0] * inv_m, my = s[1] * inv_m, mz = s[2] * inv_m;
    const float c00 = s[3] * inv_m - mx * mx;
    const float c10 = s[4] * inv_m - my * mx;
    const float c11 = s[5] * inv_m - my * my;
    const float c20 = s[6] * inv_m - mz * mx;
    const float c21 = s[7] * inv_m - mz * my;
    const float c22 = s[8] * inv_m - mz * mz;
    return eig3_sym_dev(c00, c10, c11, c20, c21, c22);
}

struct SeedOut {
    float4* vnew; float* vvar;                                    // orientation pass
    const float4* prev;                                           // fused path: previous iteration's vnew row (kept when the active set is empty)
    int pack_w;                                                   // fused path: weight travels in vnew[].w, no vvar array
    float4* pset; int* good; float4* centers; int* active_size;   // position pass
    float max_proj_range;
    long long* tl = nullptr;                                      // detail mode: stamps of the current task (slots 5, 6)
};

// MODE 0: R15-R17 Jacobi half of one drosa iteration (rosa.cpp:275-292)
// MODE 1: R18 position step (rosa.cpp:317-365)
template <int MODE, int WPL>
__device__ __forceinline__ void seed_body(const SeedCtx& c, int seed, const f3& p, const f3& n, const SeedOut& o) {
    const int lane = threadIdx.x & 31;
    unsigned vis[WPL];
    const int m = active_set<WPL>(c, seed, p, n, vis, MODE == 0 ? o.tl : nullptr);
    if (MODE == 0 && o.tl && lane == 0) { o.tl[5] = gtime(); o.tl[8] = clock64(); }
    if (MODE == 0) {
        if (m == 0) {                                     // vvar = 0 (rosa.cpp:285-287) -> 1/(0^4+eps); vnew row kept
            if (lane == 0) {
                if (o.pack_w) { float4 pv = o.prev ? __ldcg(o.prev + seed) : make_float4(0.f, 0.f, 0.f, 0.f); pv.w = 1.0f / (0.0f + 1e-5f); o.vnew[seed] = pv; }
                else o.vvar[seed] = 1.0f / (0.0f + 1e-5f);
            }
            return;
        }
        float s[9];
        normal_cov_sums<WPL>(c, vis, s);
        const float inv_m = 1.0f / (float)m;
        const Eig3 E = normal_cov_eig(s, inv_m);
        const f3 sym = eig_col(E, 0);                                                   // compute_symmetrynormal (rosa.cpp:685-706)
        if (o.tl && lane == 0) { o.tl[6] = gtime() + (sym.x == 123.0f ? 1 : 0); o.tl[9] = clock64(); }
        float a2[2] = {0.0f, 0.0f};                                                     // symmnormal_variance (rosa.cpp:708-732)
        for_members<WPL>(vis, [&](int i) {
            const float4 v = seed_nraw(c, i);
            const float il = seed_vnd(c, i).w;
            const float a = dot3(f3{v.x, v.y, v.z}, sym) * il;
            a2[0] += a; a2[1] += a * a;
        });
        bfly<2>(a2);
        const float mu = a2[0] * inv_m;
        float var = a2[1] * inv_m - mu * mu;
        if (m > 1) var *= (float)m / (float)(m - 1);
        if (lane == 0) {
            float a = var * var; a = a * a;
            const float wgt = 1.0f / (a + 1e-5f);                                       // rosa.cpp:291-292
            if (o.pack_w) o.vnew[seed] = make_float4(sym.x, sym.y, sym.z, wgt);
            else { o.vnew[seed] = make_float4(sym.x, sym.y, sym.z, 0.0f); o.vvar[seed] = wgt; }
        }
    } else {
        if (lane == 0) o.active_size[seed] = m;
        if (m == 0) { if (lane == 0) o.good[seed] = 0; return; }
        float s[9];
        normal_cov_sums<WPL>(c, vis, s);                                                // cov_eigs_from_normals (rosa.cpp:734-753)
        const float inv_m = 1.0f / (float)m;
        const Eig3 E = normal_cov_eig(s, inv_m);
        const float planar_score = (E.ev[1] - E.ev[0]) / smax(E.ev[2], 1e-6f);
        f3 center;
        if (planar_score < 0.1f) {                                                      // rosa.cpp:334-349
            float t[6] = {0, 0, 0, 0, 0, 0};
            for_members<WPL>(vis, [&](int i) {
                const float4 q = seed_pt(c, i); const float4 v = seed_nraw(c, i); const float il = seed_vnd(c, i).w;
                t[0] += q.x; t[1] += q.y; t[2] += q.z; t[3] += v.x * il; t[4] += v.y * il; t[5] += v.z * il;
            });
            bfly<6>(t);
            const f3 mean_p{t[0] * inv_m, t[1] * inv_m, t[2] * inv_m};
            const f3 mean_n{t[3] * inv_m, t[4] * inv_m, t[5] * inv_m};
            center = sub3(mean_p, mul3(mean_n, 0.5f * o.max_proj_range));
        } else {                                                                        // closest_projection_point (rosa.cpp:755-791)
            float t[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) t[k] = 0.0f;
            for_members<WPL>(vis, [&](int i) {
                const float4 q = seed_pt(c, i); const float4 v = seed_nraw(c, i);
                const f3 vn{v.x * v.w, v.y * v.w, v.z * v.w};
                const float P00 = 1.0f - vn.x * vn.x, P01 = 0.0f - vn.x * vn.y, P02 = 0.0f - vn.x * vn.z;
                const float P10 = 0.0f - vn.y * vn.x, P11 = 1.0f - vn.y * vn.y, P12 = 0.0f - vn.y * vn.z;
                const float P20 = 0.0f - vn.z * vn.x, P21 = 0.0f - vn.z * vn.y, P22 = 1.0f - vn.z * vn.z;
                t[0] += P00; t[1] += P01; t[2] += P02; t[3] += P10; t[4] += P11; t[5] += P12; t[6] += P20; t[7] += P21; t[8] += P22;
                t[9] += sum3(P00 * q.x, P01 * q.y, P02 * q.z);
                t[10] += sum3(P10 * q.x, P11 * q.y, P12 * q.z);
                t[11] += sum3(P20 * q.x, P21 * q.y, P22 * q.z);
            });
            bfly<12>(t);
            const float tr = sum3(t[0], t[4], t[8]);
            const float reg = 1e-6f * smax(tr, 1e-6f);
            float Mm[9] = {t[0] + reg, t[1], t[2], t[3], t[4] + reg, t[5], t[6], t[7], t[8] + reg};
            const f3 B{t[9], t[10], t[11]};
            if (!llt3_solve(Mm, B, center)) { if (!ldlt3_solve(Mm, B, center)) center = {1e8f, 1e8f, 1e8f}; }
        }
        const bool ok = fin3(center) && (norm3(sub3(center, p)) < o.max_proj_range);
        if (lane == 0) {
            o.good[seed] = ok ? 1 : 0;
            if (ok) { o.pset[seed] = make_float4(center.x, center.y, center.z, 1.0f); o.centers[seed] = make_float4(center.x, center.y, center.z, 1.0f); }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(SEED_WARPS * 32)
k_drosa_seed(FrameState* st, const float4* __restrict__ pts, const float4* __restrict__ nraw, const float4* __restrict__ vnd,
             float4* pset, const float4* __restrict__ vset, const unsigned* __restrict__ bits, int W, SeedOut o) {
    extern __shared__ uint4 seed_smem[];
    if (st->status != 0) return;
    const int M = st->M;
    if ((int)blockIdx.x >= M) return;
    const int Wn = (M + 31) >> 5;
    (void)W;
    const SeedCtx c = seed_stage(seed_smem, pts, nraw, vnd, bits, M, st->leaf_size);
    const int wib = threadIdx.x >> 5;
    for (int seed = blockIdx.x + gridDim.x * wib; seed < M; seed += gridDim.x * SEED_WARPS) {
        const f3 p = ld3f(pset, seed), n = ld3f(vset, seed);
        if (Wn <= 32) seed_body<MODE, 1>(c, seed, p, n, o);
        else if (Wn <= 64) seed_body<MODE, 2>(c, seed, p, n, o);
        else seed_body<MODE, 4>(c, seed, p, n, o);
    }
}

// ------------------------------------------------------------------------------------------
// Dependency levels of the in-place smoothing pass (rosa.cpp:295-308) for the unfused, level-synchronous kernel:
// p reads the updated orientation of every surf neighbour j < p, so level[p] = 1 + max level[j < p].  One block,
// chaotic relaxation in shared memory (values only grow towards the fixpoint), then a counting sort by level.
constexpr int SCHED_BINS = 4096;

// val[p] = 1 + max(base[p], max_{j<p in nbs(p)} val[j]) to the fixpoint; returns with all threads synchronised
template <int NBK>
__device__ __forceinline__ void sched_relax(int M, int nbk, int k_eff, const unsigned short* ss, int* val, const int* base) {
    while (true) {
        int changed = 0;
        for (int sweep = 0; sweep < 8; ++sweep) {
            for (int p = threadIdx.x; p < M; p += blockDim.x) {
                int v = base ? base[p] : 0;
                if (NBK > 0) {
#pragma unroll
                    for (int u = 0; u < NBK; ++u) { const int j = ss[p * NBK + u]; if (j < p) v = max(v, ((volatile int*)val)[j]); }
                } else {
                    for (int u = 0; u < k_eff; ++u) { const int j = ss[p * nbk + u]; if (j < p) v = max(v, ((volatile int*)val)[j]); }
                }
                v += 1;
                if (v != val[p]) { ((volatile int*)val)[p] = v; changed = 1; }
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
}

// counting sort of the points by key[p] >= 1 (ties in arbitrary order): order[] and optionally the bin offsets out
__device__ __forceinline__ void sched_sort(int M, const int* key, int* hist, int* s_tot, int* __restrict__ order, int* __restrict__ off_out, int* n_bins_out) {
    int mx = 0;
    for (int p = threadIdx.x; p < M; p += blockDim.x) mx = max(mx, key[p]);
    mx = __reduce_max_sync(FULLM, mx);
    if (threadIdx.x == 0) *s_tot = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicMax(s_tot, mx);
    __syncthreads();
    const int nb = min(*s_tot, SCHED_BINS - 1);           // bin = key - 1; the clamp is never reached in practice (T < 8 * (depth + lag + 1))
    for (int l = threadIdx.x; l <= nb; l += blockDim.x) hist[l] = 0;
    __syncthreads();
    for (int p = threadIdx.x; p < M; p += blockDim.x) atomicAdd(&hist[min(key[p], nb) - 1], 1);
    __syncthreads();
    if (threadIdx.x < 32) {               // exclusive scan of hist[0..nb] by one warp
        int carry = 0;
        for (int b = 0; b <= nb; b += 32) {
            const int l = b + threadIdx.x;
            const int v = (l <= nb) ? hist[l] : 0;
            int incl = v;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(FULLM, incl, off); if ((int)threadIdx.x >= off) incl += t; }
            if (l <= nb) { hist[l] = carry + incl - v; if (off_out) off_out[l] = carry + incl - v; }
            carry += __shfl_sync(FULLM, incl, 31);
        }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < M; p += blockDim.x) { const int pos = atomicAdd(&hist[min(key[p], nb) - 1], 1); order[pos] = p; }
    if (n_bins_out && threadIdx.x == 0) *n_bins_out = nb;
    __syncthreads();
}

template <int NBK>
__device__ __forceinline__ void chain_schedule_pass(int w, int M, int nbk, int k_eff, const unsigned short* ssurf,
                                                    unsigned short* sT, volatile int* prog, volatile int* tmax, int lag);

// Dependency levels of the smoothing pass (T_0 of the fused kernel's schedule model, see chain_schedule_pass) and the points
// sorted by level, on the side stream next to the similarity stage: one warp walks the index chunks, the block sorts.
__global__ void __launch_bounds__(256) k_gs_schedule(FrameState* st, const int* __restrict__ surf, int nbk, int* __restrict__ level,
                                                     int* __restrict__ level_off, int* __restrict__ level_pts) {
    extern __shared__ int lsm[];          // lvl[M] | hist[SCHED_BINS] | surf as ushort [M*nbk] | T_0 as ushort [M]
    __shared__ int s_tot;
    __shared__ int s_prog[2];
    frame_stamp(st, TS_SCHED);
    if (st->status != 0) return;
    const int M = st->M;
    int* ta = lsm; int* hist = ta + M;
    unsigned short* ss = reinterpret_cast<unsigned short*>(hist + SCHED_BINS);
    unsigned short* sT = ss + (size_t)M * nbk;
    const int k_eff = min(nbk, M);
    for (int i = threadIdx.x; i < M * nbk; i += blockDim.x) { const int j = surf[i]; ss[i] = (unsigned short)(j < 0 ? 0 : j); }
    __syncthreads();
    if (threadIdx.x < 32) {
        if (nbk == 10 && k_eff == 10) chain_schedule_pass<10>(0, M, nbk, k_eff, ss, sT, s_prog, nullptr, 0);
        else chain_schedule_pass<0>(0, M, nbk, k_eff, ss, sT, s_prog, nullptr, 0);
    }
    __syncthreads();
    for (int p = threadIdx.x; p < M; p += blockDim.x) { ta[p] = sT[p]; level[p] = sT[p]; }
    __syncthreads();
    sched_sort(M, ta, hist, &s_tot, level_pts, level_off, &st->n_levels);
    __syncthreads();
    frame_stamp(st, TS_SCHED_END);
}

// R17 smoothing half, level by level.  All operands in shared memory; one warp walks the chain.
// sold = orientation after the Jacobi half (vset = vnew, rosa.cpp:289) with the weight in .w;
// snew = orientations already rewritten by this pass.  p reads snew[j] for j < p, sold[j] otherwise.
template <int NBK>
__device__ __forceinline__ float4 gs_point(int p, const float4* sold, const float4* snew, const unsigned short* ssurf, int nbk, int k_eff, int* iters = nullptr) {
    float m00 = 0, m10 = 0, m11 = 0, m20 = 0, m21 = 0, m22 = 0;
    if (NBK > 0) {
        int j[NBK > 0 ? NBK : 1];
#pragma unroll
        for (int u = 0; u < NBK; ++u) j[u] = ssurf[p * NBK + u];
        float4 v[NBK > 0 ? NBK : 1];
#pragma unroll
        for (int u = 0; u < NBK; ++u) v[u] = (j[u] < p) ? snew[j[u]] : sold[j[u]];
#pragma unroll
        for (int u = 0; u < NBK; ++u) {
            const float wx = v[u].x * v[u].w, wy = v[u].y * v[u].w, wz = v[u].z * v[u].w;      // (w v) v^T, lower triangle (rosa.cpp:300-304)
            m00 += wx * v[u].x; m10 += wy * v[u].x; m11 += wy * v[u].y; m20 += wz * v[u].x; m21 += wz * v[u].y; m22 += wz * v[u].z;
        }
    } else {
        for (int u = 0; u < k_eff; ++u) {
            const int jj = ssurf[p * nbk + u];
            const float4 v = (jj < p) ? snew[jj] : sold[jj];
            const float wx = v.x * v.w, wy = v.y * v.w, wz = v.z * v.w;
            m00 += wx * v.x; m10 += wy * v.x; m11 += wy * v.y; m20 += wz * v.x; m21 += wz * v.y; m22 += wz * v.z;
        }
    }
    const Eig3 E = eig3_sym_dev(m00, m10, m11, m20, m21, m22);
    if (iters) *iters = E.iters;
    const f3 d = normalize3_dev(eig_col(E, 2));                                                    // rosa.cpp:305-307
    return make_float4(d.x, d.y, d.z, sold[p].w);
}

__global__ void __launch_bounds__(GS_THREADS)
k_drosa_smooth2(FrameState* st, const float4* __restrict__ vnew, const float* __restrict__ vvar, const int* __restrict__ surf, int nbk,
                const int* __restrict__ level_off, const int* __restrict__ level_pts, float4* __restrict__ vset) {
    extern __shared__ uint4 gs_smem[];
    if (st->status != 0) return;
    const int M = st->M;
    const int nl = st->n_levels;
    if ((size_t)32 * M + 4 * ((size_t)nl + 2) + 2 * (size_t)M * nbk + 2 * (size_t)M + 64 > (size_t)GS_SMEM_BUDGET) {
        if (threadIdx.x == 0) st->status = OSEP_ERR_CAPACITY;
        return;
    }
    // carve: sold[M] snew[M] (float4) | soff[nl+1] (int) | ssurf[M*nbk] slist[M] (ushort)
    float4* sold = reinterpret_cast<float4*>(gs_smem);
    float4* snew = sold + M;
    int* soff = reinterpret_cast<int*>(snew + M);
    unsigned short* ssurf = reinterpret_cast<unsigned short*>(soff + nl + 1);
    unsigned short* slist = ssurf + (size_t)M * nbk;
    for (int i = threadIdx.x; i < M; i += blockDim.x) { float4 v = vnew[i]; v.w = vvar[i]; sold[i] = v; snew[i] = v; }
    for (int i = threadIdx.x; i <= nl; i += blockDim.x) soff[i] = level_off[i];
    for (int i = threadIdx.x; i < M * nbk; i += blockDim.x) { const int j = surf[i]; ssurf[i] = (unsigned short)(j < 0 ? 0 : j); }
    for (int i = threadIdx.x; i < M; i += blockDim.x) slist[i] = (unsigned short)level_pts[i];
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const int k_eff = min(nbk, M);
    const bool fast10 = (nbk == 10 && k_eff == 10);
    for (int l = 0; l < nl; ++l) {
        const int b = soff[l], e = soff[l + 1];
        for (int t = b + lane; t < e; t += 32) {
            const int p = slist[t];
            snew[p] = fast10 ? gs_point<10>(p, sold, snew, ssurf, nbk, k_eff) : gs_point<0>(p, sold, snew, ssurf, nbk, k_eff);
        }
        __syncwarp();
    }
    for (int i = lane; i < M; i += 32) { float4 v = snew[i]; v.w = 0.0f; vset[i] = v; }
}


// ------------------------------------------------------------------------------------------
// Fused drosa: ONE persistent cooperative kernel for the niter_drosa orientation iterations and
// the position step.  The reference's loop (rosa.cpp:274-309) is
//     for it: { Jacobi: O(it,p) for all p ;  Gauss-Seidel: G(it,p) in index order }
// and O(it+1,p) reads only G(it,p) (its own orientation), while G(it,p) reads O(it,j) of its surf
// neighbours and G(it,j<p).  So the six Gauss-Seidel chains (each ~230 dependent eigen solves) do
// not have to run back to back: they run as overlapping WAVEFRONTS, iteration it+1 trailing
// iteration it by a few levels.  Results are bit-identical (same operands, same operations).
//   CTA it < nit   : chain walker of iteration it (warp 0) + prefetch warps (stage O(it,*) results
//                    from L2 into shared memory as they appear) + publish warp (G(it,p) -> L2 + flag)
//   other CTAs     : workers; warps claim the next slot of one ready queue (slots [0, M) = the iteration-0
//                    tasks, later slots are filled by the chain publishers as G(it,p) becomes final), run the
//                    active-set body and publish flagO[it][p].
// Deadlock freedom: only ready tasks are ever queued, and all CTAs are co-resident (cooperative launch).
struct FusedArgs {
    FrameState* st;
    const float4* pts; const float4* nraw; const float4* vnd; float4* pset;
    const unsigned* bits; int W;
    float4* vset_it;       // [nit+1][Mcap]  vset_it[0] = rosa_init orientation, vset_it[it+1] = after iteration it
    float4* vnew_it;       // [nit][Mcap]    xyz = symmetry normal, w = weight 1/(var^4+eps)
    int* flagO; int* taskq; int* qctr;      // qctr[0] = claim counter, qctr[1] = publish counter
    int Mcap; int nit;
    const int* surf; int nbk; int dbg_detail; int lag;
    int* good; float4* centers; int* active_size; float max_proj_range;
    long long* dbg;        // per chain: [t_begin, t_first_step, t_end, steps, spins, points] (globaltimer ns)
    const int* level; const int* level_pts;   // T_0 = dependency level of every point and the points sorted by it (k_gs_schedule, side stream)
    int hot_ctas;          // worker CTAs that serve the ready queue after the iteration-0 tasks
    unsigned spin_ns;      // idle workers sleep this long between two polls of their queue slot (0 = busy poll)
    int smem_budget;       // dynamic shared memory of this launch (bytes)
    long long* tl;         // detail mode: per (it, p) 8 globaltimer stamps [G solved, task published, worker start, O flag, O staged] (tools/fused_dbg.py)
};

// -DOSEP_STRESS (make stress -> libosep_skel_stress.so): a pseudo-random delay of 0..4 us in front of every flag / queue /
// state-byte access of the fused kernel, so that orderings a normal run never produces do occur; tools/soak.py then has
// to reproduce every frame bit for bit (profiles/r2_stress_soak.log).  compute-sanitizer is not usable on this pool.
#ifdef OSEP_STRESS
__device__ __forceinline__ void stress_delay() {
    unsigned x = (unsigned)clock64() * 2654435761u + threadIdx.x * 40503u + blockIdx.x * 9176u;
    x ^= x >> 13; x *= 2246822519u; x ^= x >> 15;
    if ((x & 3u) == 0u) __nanosleep((x >> 8) & 4095u);
}
#else
__device__ __forceinline__ void stress_delay() {}
#endif
__device__ __forceinline__ int ld_vol(const int* p) { stress_delay(); return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_vol(int* p, int v) { stress_delay(); *reinterpret_cast<volatile int*>(p) = v; }

__device__ __forceinline__ void fused_worker(const FusedArgs& a, uint4* smem, int M) {
    const int Wn = (M + 31) >> 5;
    const SeedCtx c = seed_stage(smem, a.pts, a.nraw, a.vnd, a.bits, M, a.st->leaf_size, a.smem_budget);
    const int lane = threadIdx.x & 31;
    const int total = (a.nit + 1) * M;
    SeedOut o; o.vvar = nullptr; o.pack_w = 1; o.pset = a.pset; o.good = a.good; o.centers = a.centers; o.active_size = a.active_size;
    o.max_proj_range = a.max_proj_range;
    const bool hot = ((int)blockIdx.x - a.nit) < a.hot_ctas;
    bool phase2 = false;
    while (true) {
        // phase 1: every worker CTA takes iteration-0 tasks (all ready at start; counter qctr[2]).  phase 2: only the first
        // `hot_ctas` worker CTAs keep claiming slots of the ready queue, which the chain publishers fill as orientations
        // get final (counter qctr[0]).  A lone task warp on an otherwise idle SM runs ~3x slower than the same warp on an
        // SM where several warps execute the same code (instruction supply: measured, tools/fused_timeline.py), and the
        // chains only release ~6 tasks per microsecond, so the tasks are concentrated on few SMs.
        int it = 0, p = 0;
        if (!phase2) {
            int t = 0;
            if (lane == 0) t = atomicAdd(a.qctr + 2, 1);
            t = __shfl_sync(FULLM, t, 0);
            if (t >= M) { if (!hot) break; phase2 = true; continue; }
            p = t;
        } else {
            int t = 0;
            if (lane == 0) t = atomicAdd(a.qctr, 1);
            t = __shfl_sync(FULLM, t, 0);
            if (t >= total - M) break;
            int e = 0;
            // every lane polls the slot (one broadcast transaction): a warp whose lane 0 spins while the other 31 wait at the
            // reconvergence point takes issue slots from the task warp that shares its scheduler (measured: tasks ran 3x slower)
            { const int* q = a.taskq + t; while ((e = ld_vol(q)) == 0) { if (a.spin_ns) __nanosleep(a.spin_ns); } }
            e = e - 1;
            __threadfence();
            it = e / a.Mcap; p = e - it * a.Mcap;
        }
        if (a.tl && lane == 0) a.tl[((size_t)it * a.Mcap + p) * 16 + 2] = gtime();
        const float4 n4 = __ldcg(a.vset_it + (size_t)it * a.Mcap + p);
        const float4 p4 = __ldcg(a.pset + p);
        const f3 pp{p4.x, p4.y, p4.z}, nn{n4.x, n4.y, n4.z};
        if (it < a.nit) {
            o.vnew = a.vnew_it + (size_t)it * a.Mcap;
            o.prev = it > 0 ? a.vnew_it + (size_t)(it - 1) * a.Mcap : nullptr;
            o.tl = a.tl ? a.tl + ((size_t)it * a.Mcap + p) * 16 : nullptr;
            if (Wn <= 32) seed_body<0, 1>(c, p, pp, nn, o);
            else if (Wn <= 64) seed_body<0, 2>(c, p, pp, nn, o);
            else seed_body<0, 4>(c, p, pp, nn, o);
            __syncwarp();
            if (lane == 0) { __threadfence(); st_vol(a.flagO + (size_t)it * a.Mcap + p, 1); if (a.tl) a.tl[((size_t)it * a.Mcap + p) * 16 + 3] = gtime(); }
        } else {
            if (Wn <= 32) seed_body<1, 1>(c, p, pp, nn, o);
            else if (Wn <= 64) seed_body<1, 2>(c, p, pp, nn, o);
            else seed_body<1, 4>(c, p, pp, nn, o);
        }
    }
}

constexpr int O_LAG_STEPS = 3;            // modelled latency (walker steps) from G(k-1,j) to the staged O(k,j)

__host__ __device__ __forceinline__ size_t fused_chain_bytes(int M, int nbk, int nit) {
    return (size_t)32 * M + 64 + (size_t)2 * M * nbk + (size_t)4 * M + (size_t)M + 4 + (size_t)2 * M * nit + 4 + (size_t)4 * SCHED_BINS + 64;
}

// One pass of the schedule model (see fused_chain), run by ONE warp over index chunks of 32 points:
//   w == 0: T_0(p) = 1 + max_{j<p in nbs(p)} T_0(j)                                   (the dependency level)
//   w >= 1: T_w(p) = 1 + max( max_{j in nbs(p)+p} T_{w-1}(j) + lag, max_{j<p in nbs(p)} T_w(j) )
// Dependencies below the chunk are final; the ones inside the chunk are resolved by a register fixpoint
// (a shuffle per neighbour and round, about as many rounds as the chunk is deep).  Pass w trails pass w-1
// through prog[]: it waits until every T_{w-1} it reads has been published.
template <int NBK>
__device__ __forceinline__ void chain_schedule_pass(int w, int M, int nbk, int k_eff, const unsigned short* ssurf,
                                                    unsigned short* sT, volatile int* prog, volatile int* tmax, int lag) {
    const int lane = threadIdx.x & 31;
    unsigned short* Tc = sT + (size_t)w * M;
    const unsigned short* Tp = sT + (size_t)(w > 0 ? w - 1 : 0) * M;
    int vmax = 0;
    for (int base = 0; base < M; base += 32) {
        const int p = min(base + lane, M - 1);
        const bool valid = base + lane < M;
        int j[NBK > 0 ? NBK : 1];
        int jmax = p;
        if (NBK > 0) {
#pragma unroll
            for (int u = 0; u < NBK; ++u) { j[u] = ssurf[p * NBK + u]; jmax = max(jmax, j[u]); }
        } else {
            for (int u = 0; u < k_eff; ++u) jmax = max(jmax, (int)ssurf[p * nbk + u]);
        }
        int vext = 0;
        if (w > 0) {
            jmax = __reduce_max_sync(FULLM, jmax);
            while (prog[w - 1] <= jmax) { }
            __syncwarp();
            int e = Tp[p];
            if (NBK > 0) {
#pragma unroll
                for (int u = 0; u < NBK; ++u) e = max(e, (int)((const volatile unsigned short*)Tp)[j[u]]);
            } else {
                for (int u = 0; u < k_eff; ++u) e = max(e, (int)((const volatile unsigned short*)Tp)[ssurf[p * nbk + u]]);
            }
            vext = e + lag;
        }
        if (NBK > 0) {
#pragma unroll
            for (int u = 0; u < NBK; ++u) if (j[u] < base) vext = max(vext, (int)Tc[j[u]]);
        } else {
            for (int u = 0; u < k_eff; ++u) { const int jj = ssurf[p * nbk + u]; if (jj < base) vext = max(vext, (int)Tc[jj]); }
        }
        int v = vext + 1;
        if (NBK == 10) {
            // in-chunk dependencies: a register fixpoint; per round 10 independent shuffles, a max TREE (not a chain) and a vote
            int src[10]; bool in[10];
#pragma unroll
            for (int u = 0; u < 10; ++u) { in[u] = (j[u] >= base && j[u] < p); src[u] = j[u] & 31; }
            bool mine = false;
#pragma unroll
            for (int u = 0; u < 10; ++u) mine |= in[u];
            if (__any_sync(FULLM, mine)) {
                while (true) {
                    int tv[10];
#pragma unroll
                    for (int u = 0; u < 10; ++u) { const int t = __shfl_sync(FULLM, v, src[u]); tv[u] = in[u] ? t : 0; }
                    const int m01 = max(tv[0], tv[1]), m23 = max(tv[2], tv[3]), m45 = max(tv[4], tv[5]), m67 = max(tv[6], tv[7]), m89 = max(tv[8], tv[9]);
                    const int nv = max(max(max(m01, m23), max(m45, m67)), max(m89, vext)) + 1;
                    const bool changed = nv != v;
                    v = nv;
                    if (!__any_sync(FULLM, changed)) break;
                }
            }
        } else
        while (true) {
            int nv = vext;
            for (int u = 0; u < k_eff; ++u) { const int jj = ssurf[p * nbk + u]; const int tv = __shfl_sync(FULLM, v, jj & 31); if (jj >= base && jj < p) nv = max(nv, tv); }
            nv += 1;
            const bool changed = nv != v;
            v = nv;
            if (!__any_sync(FULLM, changed)) break;
        }
        if (valid) { Tc[p] = (unsigned short)min(v, 65535); vmax = max(vmax, v); }
        __syncwarp();
        __threadfence_block();
        if (lane == 0) prog[w] = min(base + 32, M);
    }
    vmax = __reduce_max_sync(FULLM, vmax);
    if (lane == 0 && tmax) *tmax = vmax;
}

__device__ __forceinline__ void fused_chain(const FusedArgs& a, uint4* smem, int M, int it) {
    // Walker of the Gauss-Seidel pass of iteration `it` (rosa.cpp:295-308).  Point q may be rewritten once
    // (a) the Jacobi results O(it,j) of all its surf neighbours (and its own) are staged and (b) every
    // neighbour j < q has been rewritten.  The walker PULLS: each lane holds one candidate of a 32-point
    // window over a topologically sorted point list and tests the candidate's dependencies against a
    // per-point state byte in shared memory (bit 0 = O staged, bit 1 = G done); ready candidates are
    // rewritten in the same step and their lanes refill from the list.  No counters, no atomics and no
    // release phase on the critical path.  Any order that respects the dependencies produces the same
    // bits (every operand is fixed by them).
    //
    // Window order.  Iteration 0 has every O available at once, so its order is the dependency level T_0.
    // For it >= 1 the O results arrive as chain it-1 completes them, so the CTA first models its own
    // schedule (chain_schedule_pass; warps 0..it run passes 0..it as a pipeline while the workers compute
    // the first Jacobi results) and sorts the points by T_it.  T_it(p) > T_it(j) for every dependency
    // j < p, so the order is topological (progress is guaranteed for any window size) and follows the
    // arrival of the operands.
    const int nbk = a.nbk;
    float4* sold = reinterpret_cast<float4*>(smem);
    float4* snew = sold + M;
    int* s_ctl = reinterpret_cast<int*>(snew + M);        // [1] points rewritten (publisher progress) [2] max T  [4..12) pass progress
    unsigned short* ssurf = reinterpret_cast<unsigned short*>(s_ctl + 16);
    unsigned short* sorder = ssurf + (size_t)M * nbk;     // window order
    unsigned short* solist = sorder + M;                  // completion order (what the publisher walks)
    unsigned char* sstate = reinterpret_cast<unsigned char*>(solist + M);
    unsigned short* sT = reinterpret_cast<unsigned short*>((reinterpret_cast<uintptr_t>(sstate + M) + 3) & ~(uintptr_t)3);   // (it+1) x M
    int* shist = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(sT + (size_t)(it + 1) * M) + 3) & ~(uintptr_t)3);
    for (int i = threadIdx.x; i < M; i += blockDim.x) sstate[i] = 0;
    for (int i = threadIdx.x; i < M * nbk; i += blockDim.x) { const int j = a.surf[i]; ssurf[i] = (unsigned short)(j < 0 ? 0 : j); }
    if (threadIdx.x < 16) s_ctl[threadIdx.x] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int k_eff = min(nbk, M);
    const bool fast10 = (nbk == 10 && k_eff == 10);
    // pass 0 (the dependency levels) was computed by k_gs_schedule on the side stream while the similarity stage ran
    for (int i = threadIdx.x; i < M; i += blockDim.x) sT[i] = (unsigned short)min(a.level[i], 65535);
    if (threadIdx.x == 0) s_ctl[4] = M;
    __syncthreads();
    if (warp >= 1 && warp <= it) {
        int* tmax = (warp == it) ? s_ctl + 2 : nullptr;      // only the last pass reports its maximum (the sort's bin count)
        if (fast10) chain_schedule_pass<10>(warp, M, nbk, k_eff, ssurf, sT, s_ctl + 4, tmax, a.lag);
        else chain_schedule_pass<0>(warp, M, nbk, k_eff, ssurf, sT, s_ctl + 4, tmax, a.lag);
    }
    __syncthreads();
    {
        // counting sort by T_it (ties in any order); a schedule deeper than the histogram falls back to the index order
        const unsigned short* T = sT + (size_t)it * M;
        const int nb = s_ctl[2];                          // written by warp `it` only
        if (it == 0) {
            for (int i = threadIdx.x; i < M; i += blockDim.x) sorder[i] = (unsigned short)a.level_pts[i];
        } else if (nb >= SCHED_BINS) {
            for (int i = threadIdx.x; i < M; i += blockDim.x) sorder[i] = (unsigned short)i;
        } else {
            for (int l = threadIdx.x; l <= nb; l += blockDim.x) shist[l] = 0;
            __syncthreads();
            for (int p = threadIdx.x; p < M; p += blockDim.x) atomicAdd(&shist[T[p] - 1], 1);
            __syncthreads();
            if (threadIdx.x < 32) {
                int carry = 0;
                for (int b = 0; b <= nb; b += 32) {
                    const int l = b + threadIdx.x;
                    const int v = (l <= nb) ? shist[l] : 0;
                    int incl = v;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(FULLM, incl, off); if ((int)threadIdx.x >= off) incl += t; }
                    if (l <= nb) shist[l] = carry + incl - v;
                    carry += __shfl_sync(FULLM, incl, 31);
                }
            }
            __syncthreads();
            for (int p = threadIdx.x; p < M; p += blockDim.x) sorder[atomicAdd(&shist[T[p] - 1], 1)] = (unsigned short)p;
        }
    }
    __syncthreads();
    const int* fO = a.flagO + (size_t)it * a.Mcap;
    const float4* vn = a.vnew_it + (size_t)it * a.Mcap;
    volatile unsigned char* vstate = sstate;
    if (warp == 0) {
        // ---- chain walker
        const unsigned lt = (1u << lane) - 1u;
        const bool detail = a.dbg_detail != 0;
        int next = 32;
        int cand = lane < M ? (int)sorder[lane] : -1;
        unsigned nb[5] = {0u, 0u, 0u, 0u, 0u};             // fast path: the candidate's 10 neighbour indices, 2 per register
        if (fast10 && cand >= 0) {
            const unsigned* wsrc = reinterpret_cast<const unsigned*>(ssurf + cand * 10);
#pragma unroll
            for (int k = 0; k < 5; ++k) nb[k] = wsrc[k];
        }
        int ndone = 0;
        long long t_begin = gtime(), t_first = 0; int steps = 0, spins = 0;
        long long c_comp = 0, c_rel = 0, c_pre = 0; int it_sum = 0, it_max = 0;
        // fast path (nb_knn = 10).  One poll = the 11 state bytes of the candidate AND, right behind them, its 10 operand
        // rows (speculative: shared memory serves a warp's loads in order, so a row read after a state byte that says
        // "final" is final) -- the dependency test, the ballot and the operand latency overlap instead of following one
        // another, and the test is a tree instead of a chain.
        while (fast10 && ndone < M) {
            const long long c0 = detail ? clock64() : 0;
            asm volatile("" ::: "memory");
            unsigned ok = 0u; float4 v[10]; float wself = 0.0f;
            if (cand >= 0) {
                int j[10]; unsigned sj[10];
#pragma unroll
                for (int k = 0; k < 5; ++k) { j[2 * k] = (int)(nb[k] & 0xffffu); j[2 * k + 1] = (int)(nb[k] >> 16); }
                stress_delay();
                const unsigned sc = vstate[cand];
#pragma unroll
                for (int u = 0; u < 10; ++u) sj[u] = vstate[j[u]];
                asm volatile("" ::: "memory");
#pragma unroll
                for (int u = 0; u < 10; ++u) v[u] = (j[u] < cand ? snew : sold)[j[u]];
                wself = sold[cand].w;
#pragma unroll
                for (int u = 0; u < 10; ++u) sj[u] = (j[u] >= cand) ? sj[u] : (sj[u] >> 1);
                ok = sc & ((sj[0] & sj[1]) & (sj[2] & sj[3])) & ((sj[4] & sj[5]) & (sj[6] & sj[7])) & (sj[8] & sj[9]) & 1u;
            }
            const bool ready = ok != 0u;
            const unsigned rm = __ballot_sync(FULLM, ready);
            if (rm == 0u) { ++spins; continue; }
            if (steps == 0) t_first = gtime();
            ++steps;
            const long long c1 = detail ? clock64() : 0;
            const int r = __popc(rm & lt);
            int my_it = 0;
            if (ready) {
                // refill loads first: they do not depend on the solve below and hide behind it
                const int pos = next + r;
                const int ncand = pos < M ? (int)sorder[pos] : -1;
                unsigned nnb[5] = {0u, 0u, 0u, 0u, 0u};
                if (ncand >= 0) {
                    const unsigned* wsrc = reinterpret_cast<const unsigned*>(ssurf + ncand * 10);
#pragma unroll
                    for (int k = 0; k < 5; ++k) nnb[k] = wsrc[k];
                }
                float m00 = 0, m10 = 0, m11 = 0, m20 = 0, m21 = 0, m22 = 0;
#pragma unroll
                for (int u = 0; u < 10; ++u) {
                    const float wx = v[u].x * v[u].w, wy = v[u].y * v[u].w, wz = v[u].z * v[u].w;      // (w v) v^T, lower triangle (rosa.cpp:300-304)
                    m00 += wx * v[u].x; m10 += wy * v[u].x; m11 += wy * v[u].y; m20 += wz * v[u].x; m21 += wz * v[u].y; m22 += wz * v[u].z;
                }
                const Eig3 E = eig3_sym_dev(m00, m10, m11, m20, m21, m22);
                my_it = E.iters;
                const f3 d = normalize3_dev(eig_col(E, 2));                                         // rosa.cpp:305-307
                snew[cand] = make_float4(d.x, d.y, d.z, wself);
                stress_delay();
                vstate[cand] = 3;
                if (a.tl) a.tl[((size_t)it * a.Mcap + cand) * 16 + 0] = gtime();
                solist[ndone + r] = (unsigned short)cand;
                cand = ncand;
#pragma unroll
                for (int k = 0; k < 5; ++k) nb[k] = nnb[k];
            }
            const long long c2 = detail ? clock64() : 0;
            if (detail) { it_sum += __reduce_add_sync(FULLM, ready ? my_it : 0); it_max += __reduce_max_sync(FULLM, ready ? my_it : 0); }
            const int c = __popc(rm);
            ndone += c; next += c;
            __syncwarp();
            __threadfence_block();
            if (lane == 0) st_vol(s_ctl + 1, ndone);
            if (detail) { c_comp += c2 - c1; c_rel += (c1 - c0) + (clock64() - c2); c_pre += c1 - c0; }
        }
        while (ndone < M) {
            const long long c0 = detail ? clock64() : 0;
            bool ready = false;
            if (cand >= 0) {
                unsigned ok = vstate[cand] & 1u;
                if (fast10) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const int j0 = (int)(nb[k] & 0xffffu), j1 = (int)(nb[k] >> 16);
                        const unsigned s0 = vstate[j0], s1 = vstate[j1];
                        ok &= s0 & ((j0 >= cand) ? 1u : (s0 >> 1));
                        ok &= s1 & ((j1 >= cand) ? 1u : (s1 >> 1));
                    }
                } else {
                    for (int u = 0; u < k_eff; ++u) {
                        const int j = ssurf[cand * nbk + u];
                        const unsigned sj = vstate[j];
                        ok &= sj & ((j >= cand) ? 1u : (sj >> 1));
                    }
                }
                ready = ok != 0u;
            }
            const unsigned rm = __ballot_sync(FULLM, ready);
            if (rm == 0u) { ++spins; continue; }
            if (steps == 0) t_first = gtime();
            ++steps;
            __threadfence_block();
            const long long c1 = detail ? clock64() : 0;
            // refill loads first: they do not depend on the solve below and hide behind it
            const int r = __popc(rm & lt);
            int ncand = cand; int my_it = 0;
            unsigned nnb[5] = {nb[0], nb[1], nb[2], nb[3], nb[4]};
            if (ready) {
                const int pos = next + r;
                ncand = pos < M ? (int)sorder[pos] : -1;
                if (fast10 && ncand >= 0) {
                    const unsigned* wsrc = reinterpret_cast<const unsigned*>(ssurf + ncand * 10);
#pragma unroll
                    for (int k = 0; k < 5; ++k) nnb[k] = wsrc[k];
                }
                int qi = 0;
                snew[cand] = fast10 ? gs_point<10>(cand, sold, snew, ssurf, nbk, k_eff, &qi) : gs_point<0>(cand, sold, snew, ssurf, nbk, k_eff, &qi);
                my_it = qi;
            }
            __syncwarp();
            const long long c2 = detail ? clock64() : 0;
            if (detail) { it_sum += __reduce_add_sync(FULLM, ready ? my_it : 0); it_max += __reduce_max_sync(FULLM, ready ? my_it : 0); }
            if (ready) {
                vstate[cand] = 3;
                solist[ndone + r] = (unsigned short)cand;
                cand = ncand;
#pragma unroll
                for (int k = 0; k < 5; ++k) nb[k] = nnb[k];
            }
            const int c = __popc(rm);
            ndone += c; next += c;
            __syncwarp();
            __threadfence_block();
            if (lane == 0) st_vol(s_ctl + 1, ndone);
            if (detail) { c_comp += c2 - c1; c_rel += (c1 - c0) + (clock64() - c2); c_pre += c1 - c0; }
        }
        if (lane == 0 && a.dbg) { long long* g = a.dbg + it * 8; g[0] = t_begin; g[1] = t_first; g[2] = gtime(); g[3] = steps; g[4] = detail ? -c_pre : spins; g[5] = ((long long)it_sum << 32) | (unsigned)it_max; g[6] = c_comp; g[7] = c_rel; }
    } else if (warp == nw - 1) {
        // ---- publisher: G(it,p) -> vset_it[it+1][p] in L2, then task (it+1, p) goes on the workers' ready queue
        float4* dst = a.vset_it + (size_t)(it + 1) * a.Mcap;
        const volatile unsigned short* vol = solist;
        int idx = lane;
        while (__any_sync(FULLM, idx < M)) {
            const int upto = ld_vol(s_ctl + 1);
            if (idx < upto) {
                __threadfence_block();
                const int p = vol[idx];
                float4 v = snew[p]; v.w = 0.0f;
                dst[p] = v;
                __threadfence();
                const int slot = atomicAdd(a.qctr + 1, 1);
                st_vol(a.taskq + slot, (it + 1) * a.Mcap + p + 1);
                if (a.tl) a.tl[((size_t)(it + 1) * a.Mcap + p) * 16 + 1] = gtime();
                idx += 32;
            }
        }
    } else if ((warp & 3) != 0) {
        // ---- prefetchers: stage O(it,j) = (symmetry normal, weight) as soon as its flag is up (window order).
        // Warps 4, 8, ... share the walker's scheduler (warp id mod 4) and leave the kernel instead: a spinning warp -- above
        // all a DIVERGED one, some lanes looping while the others wait at the reconvergence point -- takes issue slots from
        // the warps of its scheduler (measured on the task warps: 3x).  The poll loop below is uniform: every lane runs the
        // same instruction stream over its (at most 32) entries, all flag loads of a round in flight together.
        const int npf = (nw - 2) - ((nw - 2) >> 2);
        const int pfi = (warp - 1) - (warp >> 2);
        const int stride = npf * 32;
        const int first = pfi * 32 + lane;
        const int kmax = (M + stride - 1) / stride;
        unsigned pend = 0u;
        for (int k = 0; k < 32; ++k) if (first + k * stride < M) pend |= 1u << k;
        while (__any_sync(FULLM, pend != 0u)) {
            for (int kb = 0; kb < kmax; kb += 8) {
                int j[8], f[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const bool want = (pend >> (kb + u)) & 1u;
                    j[u] = want ? (int)sorder[first + (kb + u) * stride] : 0;
                    f[u] = want ? ld_vol(fO + j[u]) : 0;
                }
                __threadfence();
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) if (f[u] != 0) v[u] = __ldcg(vn + j[u]);
#pragma unroll
                for (int u = 0; u < 8; ++u) if (f[u] != 0) sold[j[u]] = v[u];
                __threadfence_block();
#pragma unroll
                for (int u = 0; u < 8; ++u) if (f[u] != 0) {
                    stress_delay();
                    vstate[j[u]] = 1;
                    if (a.tl) a.tl[((size_t)it * a.Mcap + j[u]) * 16 + 4] = gtime();
                    pend &= ~(1u << (kb + u));
                }
            }
        }
    }
}

__global__ void __launch_bounds__(SEED_WARPS * 32) k_drosa_fused(const __grid_constant__ FusedArgs a) {
    extern __shared__ uint4 fused_smem[];
    frame_stamp(a.st, TS_DROSA);
    if (a.st->status != 0) return;
    const int M = a.st->M;
    // shared-memory footprint of a chain CTA (see fused_chain); evaluated identically by every CTA
    const size_t chain_bytes = fused_chain_bytes(M, a.nbk, a.nit);
    if (chain_bytes > (size_t)a.smem_budget) {
        if (blockIdx.x == 0 && threadIdx.x == 0) a.st->status = OSEP_ERR_CAPACITY;
        return;
    }
    // detail mode (OSEP_FUSED_DETAIL=1): kernel entry / last exit stamps next to the per-chain records (tools/fused_dbg.py)
    if (a.dbg_detail && threadIdx.x == 0) atomicMin(reinterpret_cast<unsigned long long*>(a.dbg + 56), (unsigned long long)gtime());
    if ((int)blockIdx.x < a.nit) fused_chain(a, fused_smem, M, (int)blockIdx.x);
    else fused_worker(a, fused_smem, M);
    if (a.dbg_detail && threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long*>(a.dbg + 57), (unsigned long long)gtime());
}

// ------------------------------------------------------------------------------------------
static size_t gs_smem_bytes(int M, int nbk) { return (size_t)32 * M + 4 * ((size_t)M + 2) + 2 * (size_t)M * nbk + 2 * (size_t)M + 64; }

// can `grid` CTAs of the fused kernel be co-resident on this device? (cooperative launch requirement)
bool drosa_fused_supported(int sms) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_drosa_fused, SEED_WARPS * 32, (size_t)SEED_SMEM_BUDGET) != cudaSuccess) { cudaGetLastError(); return false; }
    return per_sm >= 1 && sms >= MAX_FUSED_ITERS + 1;
}

void drosa_set_attributes() {
    cudaFuncSetAttribute(k_drosa_seed<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SEED_SMEM_BUDGET + 1024);
    cudaFuncSetAttribute(k_drosa_seed<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SEED_SMEM_BUDGET + 1024);
    cudaFuncSetAttribute(k_drosa_smooth2, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM_BUDGET);
    cudaFuncSetAttribute(k_gs_schedule, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM_BUDGET);
    cudaFuncSetAttribute(k_drosa_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, SEED_SMEM_BUDGET);
}

// hoisted normals + schedule of the smoothing pass: needs only the downsampled cloud and surf_nbs, so it runs on the
// side stream next to the similarity stage
void launch_drosa_prep(osep_rosa_impl* h, cudaStream_t sk) {
    RosaDev& d = h->d;
    const osep_rosa_config& c = h->cfg;
    const bool fused = h->fused && c.niter_drosa <= MAX_FUSED_ITERS;
    // (the hoisted normals nraw / vnd and the iteration-0 orientations vset_it[0] are written by the voxel reduce, nscale.cu)
    if (fused) cudaMemsetAsync(d.fused_flags, 0, sizeof(int) * ((size_t)2 * MAX_FUSED_ITERS * d.Mcap + 16), sk);     // off the critical path
    k_gs_schedule<<<1, 256, sizeof(int) * (d.Mcap + SCHED_BINS) + sizeof(unsigned short) * d.Mcap * (d.nbk + 1), sk>>>(d.st, d.surf, d.nbk, d.level, d.level_off, d.level_pts);
    h->n_launches += 1;
}

// R14-R18; `sms` = number of SMs (one seed CTA per SM)
void launch_drosa(osep_rosa_impl* h, int sms) {
    RosaDev& d = h->d;
    cudaStream_t s = h->stream;
    const osep_rosa_config& c = h->cfg;
    stage_mark(h, SM_SURF);
    stage_mark(h, SM_DROSA_ORIENT);
    SeedOut o; o.vnew = d.vnew; o.vvar = d.vvar; o.prev = nullptr; o.pack_w = 0; o.pset = d.pset; o.good = d.good; o.centers = d.centers; o.active_size = d.active_size;
    o.max_proj_range = c.max_proj_range;
    if (h->fused && c.niter_drosa <= MAX_FUSED_ITERS) {
        // capacity of the chain CTA's shared memory: 62 bytes per point (+ level offsets)
        FusedArgs a;
        a.st = d.st; a.pts = d.pts; a.nraw = d.nraw; a.vnd = d.vnd; a.pset = d.pset; a.bits = d.simi_bits; a.W = d.W;
        a.vset_it = d.vset_it; a.vnew_it = d.vnew_it;
        a.flagO = d.fused_flags; a.taskq = d.fused_flags + (size_t)MAX_FUSED_ITERS * d.Mcap; a.qctr = d.fused_flags + (size_t)2 * MAX_FUSED_ITERS * d.Mcap;
        a.Mcap = d.Mcap; a.nit = c.niter_drosa; a.surf = d.surf; a.nbk = d.nbk; a.dbg_detail = h->fused_detail ? 1 : 0; a.lag = h->sched_lag; a.spin_ns = (unsigned)h->spin_ns; a.hot_ctas = h->hot_ctas; a.level = d.level; a.level_pts = d.level_pts;
        a.good = d.good; a.centers = d.centers; a.active_size = d.active_size; a.max_proj_range = c.max_proj_range;
        a.dbg = d.fused_dbg; a.tl = h->fused_detail ? d.fused_tl : nullptr;
        if (h->fused_detail) { cudaMemsetAsync(d.fused_dbg + 56, 0xff, 8, s); cudaMemsetAsync(d.fused_dbg + 57, 0, 8, s); }
        int grid = h->fused_grid > 0 ? h->fused_grid : sms;
        if (grid < c.niter_drosa + 1) grid = c.niter_drosa + 1;
        void* args[] = {(void*)&a};
        if (grid > sms) grid = sms;
        a.smem_budget = h->fused_smem > 0 ? std::min(h->fused_smem, SEED_SMEM_BUDGET) : (h->fused_grid > 0 ? SEED_SMEM_BUDGET : FUSED_SMEM_SINGLE);
        cudaLaunchCooperativeKernel((const void*)k_drosa_fused, dim3(grid), dim3(SEED_WARPS * 32), args, (size_t)a.smem_budget, s);
        if (h->debug) cudaMemcpyAsync(d.vset, d.vset_it + (size_t)c.niter_drosa * d.Mcap, sizeof(float4) * d.Mcap, cudaMemcpyDeviceToDevice, s);   // tap "vset"
        if (h->debug && d.vset_iters)
            for (int it = 0; it < c.niter_drosa; ++it)
                cudaMemcpyAsync(d.vset_iters + (size_t)it * d.Mcap, d.vset_it + (size_t)(it + 1) * d.Mcap, sizeof(float4) * d.Mcap, cudaMemcpyDeviceToDevice, s);
        stage_mark(h, SM_DROSA_SMOOTH);
        h->n_launches += 1;
        return;
    }
    cudaMemsetAsync(d.vnew, 0, sizeof(float4) * d.Mcap, s);        // vnew = Zero (rosa.cpp:272)
    const size_t gs_bytes = gs_smem_bytes(d.Mcap, d.nbk) <= (size_t)GS_SMEM_BUDGET ? gs_smem_bytes(d.Mcap, d.nbk) : (size_t)GS_SMEM_BUDGET;
    for (int it = 0; it < c.niter_drosa; ++it) {
        k_drosa_seed<0><<<sms, SEED_WARPS * 32, SEED_SMEM_BUDGET, s>>>(d.st, d.pts, d.nraw, d.vnd, d.pset, d.vset, d.simi_bits, d.W, o);
        k_drosa_smooth2<<<1, GS_THREADS, gs_bytes, s>>>(d.st, d.vnew, d.vvar, d.surf, d.nbk, d.level_off, d.level_pts, d.vset);
        if (h->debug && d.vset_iters) cudaMemcpyAsync(d.vset_iters + (size_t)it * d.Mcap, d.vset, sizeof(float4) * d.Mcap, cudaMemcpyDeviceToDevice, s);
    }
    stage_mark(h, SM_DROSA_SMOOTH);
    k_drosa_seed<1><<<sms, SEED_WARPS * 32, SEED_SMEM_BUDGET, s>>>(d.st, d.pts, d.nraw, d.vnd, d.pset, d.vset, d.simi_bits, d.W, o);
    h->n_launches += 2 * c.niter_drosa + 1;
}

}  // namespace osep
