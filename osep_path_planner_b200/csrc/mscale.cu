// csrc/mscale.cu -- M-scale ROSA stages (M ~ 1000 downsampled points) on sm_100a:
//   R12 similarity_neighbor_extraction   (rosa.cpp:201-256, 625-651)
//   R13 surf_nbs kNN(nb_knn)             (rosa.cpp:259-267)
//   R14-R17 drosa orientation loop       (rosa.cpp:270-309, 653-732)
//   R18 position step, R19 poor fix-up   (rosa.cpp:311-376, 734-791)
//   R20 dcrosa                           (rosa.cpp:382-422, 793-841)
//   R21 vertex_sampling                  (rosa.cpp:424-532)
//   R22 vertex_smooth                    (rosa.cpp:534-620)
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

namespace osep {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ f3 ld3(const float4* a, int i) { const float4 v = a[i]; return {v.x, v.y, v.z}; }
__device__ __forceinline__ void st3(float4* a, int i, const f3& v, float w = 0.0f) { a[i] = make_float4(v.x, v.y, v.z, w); }

template <int K> __device__ __forceinline__ void butterfly(float (&v)[K]) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = v[k] + __shfl_xor_sync(FULL, v[k], off);
    }
}

// block-wide exclusive scan of one int per thread-strided element (single block helper, in place)
__device__ int block_scan_inplace(int* data, int n) {
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
        for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(FULL, incl, off); if (lane >= off) incl += t; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int ws = (lane < nw) ? wsum[lane] : 0;
            int wi = ws;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(FULL, wi, off); if (lane >= off) wi += t; }
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

// stage gate: rosa_init always succeeds; similarity_neighbor_extraction fails on an empty cloud (rosa.cpp:205)
__global__ void k_mscale_begin(FrameState* st) {
    if (st->status == 0 && st->M == 0) st->status = OSEP_STAGE_3;
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
__global__ void k_simi(FrameState* st, const float4* __restrict__ pts, const float4* __restrict__ nrms,
                       int* simi_off, int* __restrict__ simi_idx, int* __restrict__ tmp_j, float* __restrict__ tmp_d) {
    if (st->status != 0) return;
    const int M = st->M;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= M) return;
    const float leaf = st->leaf_size;
    const float radius_r = 10.0f * leaf;
    const float th_sim = 0.1f * leaf;
    const float r2 = radius_r * radius_r;
    const f3 P1 = ld3(pts, i), N1 = ld3(nrms, i);
    const bool n1ok = fin3(N1);
    const f3 v1 = normalize3(N1);
    const int base = FILL ? simi_off[i] : 0;
    int cnt = 0;
    if (n1ok) {
        for (int j0 = 0; j0 < M; j0 += 32) {
            const int j = j0 + lane;
            bool keep = false; float d2 = 0.0f;
            if (j < M && j != i) {
                const f3 P2 = ld3(pts, j);
                d2 = dist2(P1, P2);
                if (d2 < r2) {
                    const f3 N2 = ld3(nrms, j);
                    if (fin3(N2)) {
                        const f3 v2 = normalize3(N2);
                        keep = similarity_metric(P1, v1, P2, v2, radius_r) > th_sim;
                    }
                }
            }
            const unsigned m = __ballot_sync(FULL, keep);
            if (FILL && keep) { const int pos = base + cnt + __popc(m & ((1u << lane) - 1u)); tmp_j[pos] = j; tmp_d[pos] = d2; }
            cnt += __popc(m);
        }
    }
    if (!FILL) { if (lane == 0) simi_off[i] = cnt; return; }
    __syncwarp();
    for (int e = lane; e < cnt; e += 32) {
        const float d = tmp_d[base + e]; const int j = tmp_j[base + e];
        int rank = 0;
        for (int f = 0; f < cnt; ++f) rank += nb_less(tmp_d[base + f], tmp_j[base + f], d, j) ? 1 : 0;
        simi_idx[base + rank] = j;
    }
}

__global__ void k_simi_scan(FrameState* st, int* simi_off, int Scap) {
    if (st->status != 0) return;
    const int M = st->M;
    if (threadIdx.x == 0) simi_off[M] = 0;
    __syncthreads();
    const int total = block_scan_inplace(simi_off, M + 1);
    if (threadIdx.x == 0) { st->S = total; if (total > Scap) st->status = OSEP_ERR_CAPACITY; }
}

// ------------------------------------------------------------------------------------------
// R13.  kNN(nb_knn) among the M points, brute force, canonical (d2, index) order, self first.
template <int K> struct TopKM {
    float d[K]; int id[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int t = 0; t < K; ++t) { d[t] = __int_as_float(0x7f800000); id[t] = 0x7fffffff; }
    }
    __device__ __forceinline__ void push(float dd, int ii) {
        if (!nb_less(dd, ii, d[K - 1], id[K - 1])) return;
#pragma unroll
        for (int t = K - 1; t > 0; --t) {
            const bool lt_prev = nb_less(dd, ii, d[t - 1], id[t - 1]);
            const bool lt_cur = nb_less(dd, ii, d[t], id[t]);
            d[t] = lt_prev ? d[t - 1] : (lt_cur ? dd : d[t]);
            id[t] = lt_prev ? id[t - 1] : (lt_cur ? ii : id[t]);
        }
        if (nb_less(dd, ii, d[0], id[0])) { d[0] = dd; id[0] = ii; }
    }
};

template <int K>
__global__ void k_surf_knn(const FrameState* __restrict__ st, const float4* __restrict__ pts, int* __restrict__ surf, int k_req) {
    if (st->status != 0) return;
    const int M = st->M;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= M) return;
    const f3 q = ld3(pts, p);
    TopKM<K> tk; tk.init();
    for (int j = 0; j < M; ++j) { const f3 c = ld3(pts, j); tk.push(dist2(q, c), j); }
    const int k_eff = min(k_req, M);
#pragma unroll
    for (int u = 0; u < K; ++u) if (u < k_req) surf[p * k_req + u] = (u < k_eff) ? tk.id[u] : -1;
}

// dependency levels of the in-place smoothing pass (rosa.cpp:295-308): point p reads the
// already-updated orientation of every surf neighbour j < p.  level[p] = 1 + max level[j<p].
// One warp; then points are bucketed by level (order inside a level is irrelevant).
__global__ void k_gs_levels(FrameState* st, const int* __restrict__ surf, int nbk, int* level, int* level_off, int* level_pts) {
    extern __shared__ int lv[];          // M ints
    if (st->status != 0) return;
    const int M = st->M;
    const int lane = threadIdx.x;
    int maxlv = 0;
    int jn = (lane < nbk && M > 0) ? surf[lane] : -1;
    for (int p = 0; p < M; ++p) {
        const int j = jn;
        if (p + 1 < M) jn = (lane < nbk) ? surf[(p + 1) * nbk + lane] : -1;
        int v = 0;
        if (j >= 0 && j < p) v = lv[j];
        v = __reduce_max_sync(FULL, v) + 1;
        if (lane == 0) lv[p] = v;
        maxlv = max(maxlv, v);
        __syncwarp();
    }
    // bucket by level: level_off has maxlv + 1 entries (levels are 1..maxlv)
    for (int l = lane; l <= maxlv; l += 32) level_off[l] = 0;
    __syncwarp();
    for (int p = lane; p < M; p += 32) { level[p] = lv[p]; atomicAdd(&level_off[lv[p] - 1], 1); }
    __syncwarp();
    // exclusive scan over maxlv+1 entries (single warp, tiles of 32)
    int carry = 0;
    for (int b = 0; b <= maxlv; b += 32) {
        const int l = b + lane;
        const int v = (l <= maxlv) ? level_off[l] : 0;
        int incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(FULL, incl, off); if (lane >= off) incl += t; }
        if (l <= maxlv) level_off[l] = carry + incl - v;
        carry += __shfl_sync(FULL, incl, 31);
    }
    __syncwarp();
    // fill (reuse lv[] as per-level cursors is not possible: use atomics on a copy kept in level[] high bits)
    for (int p = lane; p < M; p += 32) lv[p] = 0;     // lv now: cursor per level index
    __syncwarp();
    for (int p = lane; p < M; p += 32) {
        const int l = level[p] - 1;
        const int pos = level_off[l] + atomicAdd(&lv[l], 1);
        level_pts[pos] = p;
    }
    if (lane == 0) st->n_levels = maxlv;
}

// ------------------------------------------------------------------------------------------
// R14: active samples of one seed = component of the seed in the similarity graph restricted
// to the slab |n.(p-q)| < leaf (rosa.cpp:653-683), as a bitmask in shared memory.
// Returns the number of members (0 when the seed itself is outside the slab).
__device__ __forceinline__ int active_bfs(int seed, const f3& p, const f3& n, float leaf, int M, int Wn,
                                          const float4* __restrict__ pts, const int* __restrict__ simi_off,
                                          const int* __restrict__ simi_idx, unsigned* slab, unsigned* vis, unsigned* front, unsigned* next) {
    const int lane = threadIdx.x & 31;
    for (int w = 0; w < Wn; ++w) {
        const int i = w * 32 + lane;
        bool in = false;
        if (i < M) {
            const float4 q = pts[i];
            const float d = n.x * (p.x - q.x) + n.y * (p.y - q.y) + n.z * (p.z - q.z);
            in = fabsf(d) < leaf;
        }
        const unsigned word = __ballot_sync(FULL, in);
        if (lane == 0) { slab[w] = word; vis[w] = 0u; front[w] = 0u; next[w] = 0u; }
    }
    __syncwarp();
    const unsigned sbit = 1u << (seed & 31);
    if (!(slab[seed >> 5] & sbit)) return 0;
    __syncwarp();
    if (lane == 0) { vis[seed >> 5] = sbit; front[seed >> 5] = sbit; }
    __syncwarp();
    while (true) {
        for (int w = 0; w < Wn; ++w) {
            unsigned fw = front[w];
            while (fw) {
                const int b = __ffs(fw) - 1; fw &= fw - 1;
                const int cur = w * 32 + b;
                const int rs = simi_off[cur], re = simi_off[cur + 1];
                for (int e = rs + lane; e < re; e += 32) {
                    const int nb = simi_idx[e];
                    const unsigned bit = 1u << (nb & 31); const int ww = nb >> 5;
                    if ((slab[ww] & bit) && !(vis[ww] & bit)) { atomicOr(&vis[ww], bit); atomicOr(&next[ww], bit); }
                }
            }
        }
        __syncwarp();
        unsigned any = 0u;
        for (int w = lane; w < Wn; w += 32) { const unsigned f = next[w]; front[w] = f; next[w] = 0u; any |= f; }
        __syncwarp();
        if (!__any_sync(FULL, any != 0u)) break;
    }
    int m = 0;
    for (int w = lane; w < Wn; w += 32) m += __popc(vis[w]);
    m = __reduce_add_sync(FULL, m);
    return m;
}

// unit normals as written in the reference: double literal 1e-20 at rosa.cpp:695,741; float
// literal at :341,:718; double division at :765
__device__ __forceinline__ f3 unit_normal_d(const f3& v) { const float s2 = sqn3(v); const float il = ((double)s2 > 1e-20) ? 1.0f / sqrtf(s2) : 0.0f; return mul3(v, il); }
__device__ __forceinline__ f3 unit_normal_f(const f3& v) { const float s2 = sqn3(v); const float il = (s2 > 1e-20f) ? 1.0f / sqrtf(s2) : 0.0f; return mul3(v, il); }

// canonical reduction C4: lane l owns the members whose 32-index word w satisfies w % 32 == l,
// adds them in ascending index, then xor butterfly.
template <class F> __device__ __forceinline__ void for_owned(const unsigned* vis, int Wn, int lane, F body) {
    for (int w = lane; w < Wn; w += 32) {
        unsigned bits = vis[w];
        while (bits) { const int i = w * 32 + (__ffs(bits) - 1); bits &= bits - 1; body(i); }
    }
}

__device__ __forceinline__ void normal_cov_sums(const unsigned* vis, int Wn, int lane, const float4* __restrict__ nrms, float (&s)[9]) {
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] = 0.0f;
    for_owned(vis, Wn, lane, [&](int i) {
        const f3 vn = unit_normal_d(ld3(nrms, i));
        s[0] += vn.x; s[1] += vn.y; s[2] += vn.z;
        s[3] += vn.x * vn.x; s[4] += vn.y * vn.x; s[5] += vn.y * vn.y;
        s[6] += vn.z * vn.x; s[7] += vn.z * vn.y; s[8] += vn.z * vn.z;
    });
    butterfly<9>(s);
}
// covariance = sum_outer * inv_m - mu mu^T (rosa.cpp:701-703, 747-749), lower triangle
__device__ __forceinline__ Eig3 normal_cov_eig(const float (&s)[9], float inv_m) {
    const float mx = s[0] * inv_m, my = s[1] * inv_m, mz = s[2] * inv_m;
    const float c00 = s[3] * inv_m - mx * mx;
    const float c10 = s[4] * inv_m - my * mx;
    const float c11 = s[5] * inv_m - my * my;
    const float c20 = s[6] * inv_m - mz * mx;
    const float c21 = s[7] * inv_m - mz * my;
    const float c22 = s[8] * inv_m - mz * mz;
    return eig3_sym(c00, c10, c11, c20, c21, c22);
}

// R15-R17 Jacobi half of one drosa iteration: one warp per seed.
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_drosa_orient(FrameState* st, const float4* __restrict__ pts, const float4* __restrict__ nrms, const float4* __restrict__ pset,
               const float4* __restrict__ vset, const int* __restrict__ simi_off, const int* __restrict__ simi_idx,
               float4* __restrict__ vnew, float* __restrict__ vvar, int W) {
    extern __shared__ unsigned smem_u[];
    if (st->status != 0) return;
    const int M = st->M;
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seed = blockIdx.x * WARPS_PER_BLOCK + wib;
    if (seed >= M) return;
    unsigned* slab = smem_u + (size_t)wib * 4 * W; unsigned* vis = slab + W; unsigned* front = vis + W; unsigned* next = front + W;
    const int Wn = (M + 31) >> 5;
    const f3 p = ld3(pset, seed), n = ld3(vset, seed);
    const int m = active_bfs(seed, p, n, st->leaf_size, M, Wn, pts, simi_off, simi_idx, slab, vis, front, next);
    if (m == 0) { if (lane == 0) vvar[seed] = 1.0f / (0.0f + 1e-5f); return; }   // vvar = 0 -> 1/(0^4 + eps), vnew keeps its row (rosa.cpp:285-287)
    // compute_symmetrynormal (rosa.cpp:685-706)
    float s[9];
    normal_cov_sums(vis, Wn, lane, nrms, s);
    const float inv_m = 1.0f / (float)m;
    const Eig3 E = normal_cov_eig(s, inv_m);
    const f3 sym = eig_col(E, 0);
    // symmnormal_variance (rosa.cpp:708-732)
    float a2[2] = {0.0f, 0.0f};
    for_owned(vis, Wn, lane, [&](int i) {
        const f3 v = ld3(nrms, i);
        const float s2 = sqn3(v);
        const float il = (s2 > 1e-20f) ? 1.0f / sqrtf(s2) : 0.0f;
        const float a = dot3(v, sym) * il;
        a2[0] += a; a2[1] += a * a;
    });
    butterfly<2>(a2);
    const float mu = a2[0] * inv_m;
    float var = a2[1] * inv_m - mu * mu;
    if (m > 1) var *= (float)m / (float)(m - 1);
    if (lane == 0) {
        st3(vnew, seed, sym);
        // vvar = 1 / (vvar^4 + 1e-5)   (rosa.cpp:291-292)
        float a = var * var; a = a * a;
        vvar[seed] = 1.0f / (a + 1e-5f);
    }
}

// R17 smoothing half: the in-place pass of rosa.cpp:295-308 with the reference's sequential
// (Gauss-Seidel) semantics: point p reads the UPDATED orientation of neighbours j < p and the
// OLD one of neighbours j >= p (itself included).  Old and new orientations are kept in two
// shared-memory arrays, so only the true dependencies (j < p) order the work: points are
// processed level by level of that DAG.  One warp.
__global__ void k_drosa_smooth(FrameState* st, const float4* __restrict__ vnew, const float* __restrict__ vvar,
                               const int* __restrict__ surf, int nbk, const int* __restrict__ level_off,
                               const int* __restrict__ level_pts, float4* __restrict__ vset, int Mcap) {
    extern __shared__ float4 sv[];       // [0, Mcap): old (xyz, w = weight); [Mcap, 2 Mcap): new
    if (st->status != 0) return;
    const int M = st->M;
    float4* sold = sv; float4* snew = sv + Mcap;
    const int lane = threadIdx.x;
    for (int i = lane; i < M; i += 32) { float4 v = vnew[i]; v.w = vvar[i]; sold[i] = v; snew[i] = v; }     // vset = vnew (rosa.cpp:289)
    __syncwarp();
    const int nl = st->n_levels;
    const int k_eff = min(nbk, M);
    for (int l = 0; l < nl; ++l) {
        const int b = level_off[l], e = level_off[l + 1];
        for (int t = b + lane; t < e; t += 32) {
            const int p = level_pts[t];
            float m00 = 0, m10 = 0, m11 = 0, m20 = 0, m21 = 0, m22 = 0;
            for (int u = 0; u < k_eff; ++u) {
                const int j = surf[p * nbk + u];
                const float4 v = (j < p) ? snew[j] : sold[j];
                const float wx = v.x * v.w, wy = v.y * v.w, wz = v.z * v.w;     // (w * v) v^T, lower triangle
                m00 += wx * v.x; m10 += wy * v.x; m11 += wy * v.y; m20 += wz * v.x; m21 += wz * v.y; m22 += wz * v.z;
            }
            const Eig3 E = eig3_sym(m00, m10, m11, m20, m21, m22);
            const f3 d = normalize3(eig_col(E, 2));
            snew[p] = make_float4(d.x, d.y, d.z, sold[p].w);
        }
        __syncwarp();
    }
    for (int i = lane; i < M; i += 32) { float4 v = snew[i]; v.w = 0.0f; vset[i] = v; }
}

// R18 position step: one warp per seed.
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
k_drosa_position(FrameState* st, const float4* __restrict__ pts, const float4* __restrict__ nrms, float4* pset,
                 const float4* __restrict__ vset, const int* __restrict__ simi_off, const int* __restrict__ simi_idx,
                 int* __restrict__ good, float4* __restrict__ centers, int* __restrict__ active_size, float max_proj_range, int W) {
    extern __shared__ unsigned smem_u[];
    if (st->status != 0) return;
    const int M = st->M;
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seed = blockIdx.x * WARPS_PER_BLOCK + wib;
    if (seed >= M) return;
    unsigned* slab = smem_u + (size_t)wib * 4 * W; unsigned* vis = slab + W; unsigned* front = vis + W; unsigned* next = front + W;
    const int Wn = (M + 31) >> 5;
    const f3 p = ld3(pset, seed), n = ld3(vset, seed);
    const int m = active_bfs(seed, p, n, st->leaf_size, M, Wn, pts, simi_off, simi_idx, slab, vis, front, next);
    if (lane == 0) active_size[seed] = m;
    if (m == 0) { if (lane == 0) good[seed] = 0; return; }
    // cov_eigs_from_normals (rosa.cpp:734-753)
    float s[9];
    normal_cov_sums(vis, Wn, lane, nrms, s);
    const float inv_m = 1.0f / (float)m;
    const Eig3 E = normal_cov_eig(s, inv_m);
    const float planar_score = (E.ev[1] - E.ev[0]) / smax(E.ev[2], 1e-6f);
    f3 center;
    if (planar_score < 0.1f) {                                     // rosa.cpp:334-349
        float t[6] = {0, 0, 0, 0, 0, 0};
        for_owned(vis, Wn, lane, [&](int i) {
            const f3 q = ld3(pts, i);
            const f3 vn = unit_normal_f(ld3(nrms, i));
            t[0] += q.x; t[1] += q.y; t[2] += q.z; t[3] += vn.x; t[4] += vn.y; t[5] += vn.z;
        });
        butterfly<6>(t);
        const f3 mean_p{t[0] * inv_m, t[1] * inv_m, t[2] * inv_m};
        const f3 mean_n{t[3] * inv_m, t[4] * inv_m, t[5] * inv_m};
        center = sub3(mean_p, mul3(mean_n, 0.5f * max_proj_range));
    } else {                                                       // closest_projection_point (rosa.cpp:755-791)
        float t[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) t[k] = 0.0f;
        for_owned(vis, Wn, lane, [&](int i) {
            const f3 q = ld3(pts, i);
            const f3 v = ld3(nrms, i);
            const float s2 = sqn3(v);
            const float il = (s2 > 1e-20f) ? (float)(1.0 / (double)sqrtf(s2)) : 0.0f;
            const f3 vn = mul3(v, il);
            const float P00 = 1.0f - vn.x * vn.x, P01 = 0.0f - vn.x * vn.y, P02 = 0.0f - vn.x * vn.z;
            const float P10 = 0.0f - vn.y * vn.x, P11 = 1.0f - vn.y * vn.y, P12 = 0.0f - vn.y * vn.z;
            const float P20 = 0.0f - vn.z * vn.x, P21 = 0.0f - vn.z * vn.y, P22 = 1.0f - vn.z * vn.z;
            t[0] += P00; t[1] += P01; t[2] += P02; t[3] += P10; t[4] += P11; t[5] += P12; t[6] += P20; t[7] += P21; t[8] += P22;
            t[9] += sum3(P00 * q.x, P01 * q.y, P02 * q.z);
            t[10] += sum3(P10 * q.x, P11 * q.y, P12 * q.z);
            t[11] += sum3(P20 * q.x, P21 * q.y, P22 * q.z);
        });
        butterfly<12>(t);
        const float tr = sum3(t[0], t[4], t[8]);
        const float reg = 1e-6f * smax(tr, 1e-6f);
        float Mm[9] = {t[0] + reg, t[1], t[2], t[3], t[4] + reg, t[5], t[6], t[7], t[8] + reg};
        const f3 B{t[9], t[10], t[11]};
        if (!llt3_solve(Mm, B, center)) { if (!ldlt3_solve(Mm, B, center)) center = {1e8f, 1e8f, 1e8f}; }
    }
    const bool ok = fin3(center) && (norm3(sub3(center, p)) < max_proj_range);
    if (lane == 0) {
        good[seed] = ok ? 1 : 0;
        if (ok) { st3(pset, seed, center, 1.0f); st3(centers, seed, center, 1.0f); }
    }
}

// R19 part 1 (single block): compact goodCenters, list poor points in ascending index.
__global__ void k_poor_prepare(FrameState* st, const int* __restrict__ good, int* good_rank, const float4* __restrict__ centers,
                               float4* __restrict__ good_centers, int* __restrict__ poor_idx, int* scratch) {
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
}
// R19 part 2: one warp per poor point: nearest slot of tmp_pt_ (good slot = surface point, poor
// slot = (0,0,0), rosa.cpp:314,359), then pset[poor] = goodCenters[gid] AS WRITTEN (rosa.cpp:374:
// gid is a slot index applied to the compacted list).  Out of bounds = UB in the reference:
// pset is left unchanged and OSEP_FLAG_POOR_FIXUP_OOB raised.
__global__ void k_poor_fix(FrameState* st, const float4* __restrict__ pts, const int* __restrict__ good,
                           const float4* __restrict__ good_centers, const int* __restrict__ poor_idx, float4* pset) {
    if (st->status != 0) return;
    const int M = st->M;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= st->n_poor) return;
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
        if (bi < st->n_good) pset[poor] = good_centers[bi];
        else atomicOr(&st->flags, (unsigned)OSEP_FLAG_POOR_FIXUP_OOB);
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
    const Eig3 E = eig3_sym(c00, c10, c11, c20, c21, c22);
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
__global__ void k_vs_adj(const FrameState* __restrict__ st, const float4* __restrict__ pcloud, unsigned* __restrict__ adj, int W) {
    if (st->status != 0) return;
    const int Mc = st->Mc;
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
// greedy cover with the reference's sequential semantics (lowest uncovered index becomes the
// next vertex, rosa.cpp:462-502); one warp, covered set in shared memory.
__global__ void k_vs_greedy(FrameState* st, const unsigned* __restrict__ adj, int W, int* __restrict__ seeds) {
    extern __shared__ unsigned cov[];
    if (st->status != 0) return;
    const int Mc = st->Mc;
    const int lane = threadIdx.x;
    const int Wn = (Mc + 31) >> 5;
    for (int w = lane; w < Wn; w += 32) {
        unsigned c = 0u;
        const int rem = Mc - w * 32;
        if (rem < 32) c = ~((1u << rem) - 1u);       // bits beyond Mc count as covered
        cov[w] = c;
    }
    __syncwarp();
    int V = 0;
    while (true) {
        int cand = 0x7fffffff;
        for (int w = lane; w < Wn; w += 32) { const unsigned u = ~cov[w]; if (u) { cand = min(cand, w * 32 + __ffs(u) - 1); break; } }
        const int seed = __reduce_min_sync(FULL, cand);
        if (seed == 0x7fffffff) break;
        if (lane == 0) seeds[V] = seed;
        ++V;
        for (int w = lane; w < Wn; w += 32) cov[w] |= adj[(size_t)seed * W + w];
        __syncwarp();
    }
    if (lane == 0) st->V = V;
}
// corresp[idx] = vertex with the smallest d2 among those covering idx, first wins ties (rosa.cpp:490-496)
__global__ void k_vs_corresp(const FrameState* __restrict__ st, const float4* __restrict__ pcloud, const int* __restrict__ seeds,
                             int* __restrict__ corresp) {
    if (st->status != 0) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= st->Mc) return;
    const float R = st->leaf_size; const float R2 = R * R;
    const f3 q = ld3(pcloud, idx);
    const int V = st->V;
    float mind2 = __int_as_float(0x7f800000); int best = -1;
    for (int v = 0; v < V; ++v) {
        const int s = seeds[v];
        const float d2 = (s == idx) ? 0.0f : dist2(ld3(pcloud, s), q);
        if ((s == idx || d2 < R2) && d2 < mind2) { mind2 = d2; best = v; }
    }
    corresp[idx] = best;
}
// recentering (rosa.cpp:506-530): sums over pts_ in ascending point index
__global__ void k_vs_recenter(FrameState* st, const float4* __restrict__ pts, const float4* __restrict__ pset, const int* __restrict__ seeds,
                              const int* __restrict__ corresp, float alpha, float4* __restrict__ skel, float4* __restrict__ skel_sampled) {
    if (st->status != 0) return;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= st->V) return;
    const int Mc = st->Mc;
    f3 sum{0.0f, 0.0f, 0.0f}; int cnt = 0;
    for (int i = 0; i < Mc; ++i) if (corresp[i] == v) { sum = add3(sum, ld3(pts, i)); ++cnt; }
    f3 cur = ld3(pset, seeds[v]);                        // rosa.cpp:475 (uncompacted pset, compacted index)
    if (cnt > 0) {
        const f3 ec = div3(sum, (float)cnt);
        const float b = (float)(1.0 - (double)alpha);    // rosa.cpp:529
        cur = add3(mul3(cur, alpha), mul3(ec, b));
    }
    st3(skel, v, cur, 1.0f); st3(skel_sampled, v, cur, 1.0f);
}

// ------------------------------------------------------------------------------------------
// R22 vertex_smooth (rosa.cpp:534-620), single block: finite compaction, radius neighbours
// (sorted (d2, index), self forced first), niter Jacobi PCA-line projections.
__global__ void k_vertex_smooth(FrameState* st, float4* skel, float4* skel2, float4* __restrict__ vtmp, int* vs_off, int* __restrict__ vs_idx,
                                int* __restrict__ tmp_j, float* __restrict__ tmp_d, int cap, float radius, float alpha, int niter) {
    if (st->status != 0) return;
    const int V = st->V;
    if (V == 0) { if (threadIdx.x == 0) st->status = OSEP_STAGE_7; return; }
    // finite compaction (rosa.cpp:543-552)
    for (int i = threadIdx.x; i < V; i += blockDim.x) { const float4 v = skel[i]; vs_off[i] = (fin(v.x) && fin(v.y) && fin(v.z)) ? 1 : 0; skel2[i] = v; }
    if (threadIdx.x == 0) vs_off[V] = 0;
    __syncthreads();
    const int Vf = block_scan_inplace(vs_off, V + 1);
    __syncthreads();
    for (int i = threadIdx.x; i < V; i += blockDim.x) { const float4 v = skel[i]; if (fin(v.x) && fin(v.y) && fin(v.z)) vtmp[vs_off[i]] = v; }
    __syncthreads();
    const float r2 = radius * radius;
    // neighbour counts
    for (int i = threadIdx.x; i < Vf; i += blockDim.x) {
        const f3 q = ld3(vtmp, i);
        int c = 0; float bd = __int_as_float(0x7f800000); int bi = 0x7fffffff;
        for (int j = 0; j < Vf; ++j) { const float d = dist2(q, ld3(vtmp, j)); if (d < r2) { ++c; if (nb_less(d, j, bd, bi)) { bd = d; bi = j; } } }
        if (c == 0) c = 1; else if (bi != i) c += 1;
        vs_off[i] = c;
    }
    if (threadIdx.x == 0) vs_off[Vf] = 0;
    __syncthreads();
    const int total = block_scan_inplace(vs_off, Vf + 1);
    __syncthreads();
    if (total > cap) { if (threadIdx.x == 0) st->status = OSEP_ERR_CAPACITY; return; }
    for (int i = threadIdx.x; i < Vf; i += blockDim.x) {
        const f3 q = ld3(vtmp, i);
        const int b = vs_off[i];
        int c = 0;
        for (int j = 0; j < Vf; ++j) { const float d = dist2(q, ld3(vtmp, j)); if (d < r2) { tmp_j[b + c] = j; tmp_d[b + c] = d; ++c; } }
        if (c == 0) { vs_idx[b] = i; continue; }
        // rank sort; if the first sorted entry is not i, i is inserted in front (rosa.cpp:567-569)
        int first = -1;
        for (int e = 0; e < c; ++e) {
            int rank = 0;
            for (int f = 0; f < c; ++f) rank += nb_less(tmp_d[b + f], tmp_j[b + f], tmp_d[b + e], tmp_j[b + e]) ? 1 : 0;
            if (rank == 0) first = tmp_j[b + e];
        }
        const int shift = (first != i) ? 1 : 0;
        if (shift) vs_idx[b] = i;
        for (int e = 0; e < c; ++e) {
            int rank = 0;
            for (int f = 0; f < c; ++f) rank += nb_less(tmp_d[b + f], tmp_j[b + f], tmp_d[b + e], tmp_j[b + e]) ? 1 : 0;
            vs_idx[b + shift + rank] = tmp_j[b + e];
        }
    }
    __syncthreads();
    float4* cur = skel; float4* nxt = skel2;
    for (int it = 0; it < niter; ++it) {
        for (int i = threadIdx.x; i < Vf; i += blockDim.x) {
            const int b = vs_off[i], e = vs_off[i + 1];
            const int n = e - b;
            if (n <= 1) { nxt[i] = cur[i]; continue; }
            f3 mean{0.0f, 0.0f, 0.0f};
            for (int t = b; t < e; ++t) mean = add3(mean, ld3(cur, vs_idx[t]));
            mean = div3(mean, (float)n);
            float c00 = 0, c10 = 0, c11 = 0, c20 = 0, c21 = 0, c22 = 0;
            for (int t = b; t < e; ++t) {
                const f3 d = sub3(ld3(cur, vs_idx[t]), mean);
                c00 += d.x * d.x; c10 += d.y * d.x; c11 += d.y * d.y; c20 += d.z * d.x; c21 += d.z * d.y; c22 += d.z * d.z;
            }
            const float fn = (float)n;
            c00 /= fn; c10 /= fn; c11 /= fn; c20 /= fn; c21 /= fn; c22 /= fn;
            const Eig3 E = eig3_sym(c00, c10, c11, c20, c21, c22);
            const f3 dir = eig_col(E, 2);
            const f3 p = ld3(cur, i);
            const f3 delta = sub3(p, mean);
            const f3 proj = add3(mean, mul3(dir, dot3(dir, delta)));
            st3(nxt, i, add3(mul3(p, alpha), mul3(proj, 1.0f - alpha)), 1.0f);
        }
        __syncthreads();
        float4* t = cur; cur = nxt; nxt = t;
    }
    // RD.skelver = cur
    if (cur != skel) { for (int i = threadIdx.x; i < V; i += blockDim.x) skel[i] = cur[i]; }
    if (threadIdx.x == 0) st->Vf = Vf;
}

// ------------------------------------------------------------------------------------------
template <int K> static void launch_surf(osep_rosa_impl* h) {
    RosaDev& d = h->d;
    k_surf_knn<K><<<div_up(d.Mcap, 128), 128, 0, h->stream>>>(d.st, d.pts, d.surf, d.nbk);
}

void mscale_set_attributes(int Mcap) {
    cudaFuncSetAttribute(k_drosa_smooth, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sizeof(float4) * Mcap));
}

void launch_mscale(osep_rosa_impl* h) {
    RosaDev& d = h->d;
    cudaStream_t s = h->stream;
    const osep_rosa_config& c = h->cfg;
    const int warp_blocks = div_up(d.Mcap, WARPS_PER_BLOCK);
    const int tb = WARPS_PER_BLOCK * 32;
    const int tblocks = div_up(d.Mcap, 128);
    k_mscale_begin<<<1, 1, 0, s>>>(d.st);
    // R12
    k_simi<false><<<warp_blocks, tb, 0, s>>>(d.st, d.pts, d.nrms, d.simi_off, d.simi_idx, nullptr, nullptr);
    k_simi_scan<<<1, 1024, 0, s>>>(d.st, d.simi_off, d.Scap);
    k_simi<true><<<warp_blocks, tb, 0, s>>>(d.st, d.pts, d.nrms, d.simi_off, d.simi_idx, d.rv1, (float*)d.rk1);
    stage_mark(h, SM_SIMI);
    // R13 + level schedule
    if (d.nbk <= 10) launch_surf<10>(h); else if (d.nbk <= 16) launch_surf<16>(h); else if (d.nbk <= 32) launch_surf<32>(h); else launch_surf<64>(h);
    k_gs_levels<<<1, 32, sizeof(int) * d.Mcap, s>>>(d.st, d.surf, d.nbk, d.level, d.level_off, d.level_pts);
    stage_mark(h, SM_SURF);
    stage_mark(h, SM_DROSA_ORIENT);
    // R14-R17
    const size_t bfs_smem = (size_t)WARPS_PER_BLOCK * 4 * d.W * sizeof(unsigned);
    cudaMemsetAsync(d.vnew, 0, sizeof(float4) * d.Mcap, s);        // vnew = Zero (rosa.cpp:272)
    for (int it = 0; it < c.niter_drosa; ++it) {
        k_drosa_orient<<<warp_blocks, tb, bfs_smem, s>>>(d.st, d.pts, d.nrms, d.pset, d.vset, d.simi_off, d.simi_idx, d.vnew, d.vvar, d.W);
        k_drosa_smooth<<<1, 32, 2 * sizeof(float4) * d.Mcap, s>>>(d.st, d.vnew, d.vvar, d.surf, d.nbk, d.level_off, d.level_pts, d.vset, d.Mcap);
        if (h->debug && d.vset_iters) cudaMemcpyAsync(d.vset_iters + (size_t)it * d.Mcap, d.vset, sizeof(float4) * d.Mcap, cudaMemcpyDeviceToDevice, s);
    }
    stage_mark(h, SM_DROSA_SMOOTH);
    // R18-R19
    k_drosa_position<<<warp_blocks, tb, bfs_smem, s>>>(d.st, d.pts, d.nrms, d.pset, d.vset, d.simi_off, d.simi_idx, d.good, d.centers,
                                                      d.active_size, c.max_proj_range, d.W);
    k_poor_prepare<<<1, 1024, 0, s>>>(d.st, d.good, d.good_rank, d.centers, d.pset2, d.poor_idx, nullptr);
    k_poor_fix<<<warp_blocks, tb, 0, s>>>(d.st, d.pts, d.good, d.pset2, d.poor_idx, d.pset);
    if (h->debug) cudaMemcpyAsync(d.centers, d.pset, sizeof(float4) * d.Mcap, cudaMemcpyDeviceToDevice, s);   // tap: pset after drosa
    stage_mark(h, SM_POSITION);
    // R20
    for (int it = 0; it < c.niter_dcrosa; ++it) {
        k_dcrosa_mean<<<tblocks, 128, 0, s>>>(d.st, d.pset, d.pset2, d.simi_off, d.simi_idx);
        k_dcrosa_line<<<tblocks, 128, 0, s>>>(d.st, d.pset2, d.pset, d.simi_off, d.simi_idx);
    }
    stage_mark(h, SM_DCROSA);
    // R21
    k_vs_compact<<<1, 1024, 0, s>>>(d.st, d.pset, d.pcloud, d.good_rank);
    k_vs_adj<<<warp_blocks, tb, 0, s>>>(d.st, d.pcloud, d.adj_bits, d.W);
    k_vs_greedy<<<1, 32, sizeof(unsigned) * d.W, s>>>(d.st, d.adj_bits, d.W, d.seeds);
    k_vs_corresp<<<tblocks, 128, 0, s>>>(d.st, d.pcloud, d.seeds, d.corresp);
    k_vs_recenter<<<tblocks, 128, 0, s>>>(d.st, d.pts, d.pset, d.seeds, d.corresp, c.alpha_recenter, d.skel, d.skel_sampled);
    stage_mark(h, SM_VSAMPLE);
    // R22
    k_vertex_smooth<<<1, 1024, 0, s>>>(d.st, d.skel, d.skel2, d.vtmp, d.vs_off, d.vs_idx, d.rv1, (float*)d.rk1, d.Scap,
                                       c.radius_smooth, c.alpha_recenter, c.niter_smooth);
    stage_mark(h, SM_VSMOOTH);
    h->n_launches += 1 + 3 + 1 + 1 + 2 * c.niter_drosa + 1 + 2 + 2 * c.niter_dcrosa + 5 + 1;
}

}  // namespace osep
