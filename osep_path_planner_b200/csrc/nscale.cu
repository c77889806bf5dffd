// csrc/nscale.cu -- N-scale stages of Rosa::preprocess (rosa.cpp:57-181) on sm_100a:
//   R3  distance/finite filter, stable compaction            (rosa.cpp:67-88)
//   R4  min_points gate                                      (rosa.cpp:91-95)
//   R7  estimate_leaf_from_bbox                              (rosa.hpp:99-110)
//   R8  leaf-adaptation loop, device resident                (rosa.cpp:120-148)
//   R5  kNN(ne_knn) normals on a sparse hash grid            (rosa.cpp:97-101, PCL NormalEstimation)
//   R9  pcl::VoxelGrid<PointNormal>::filter                  (rosa.cpp:126-128, PCL VoxelGrid)
//   R10 unpack + renormalise, R11 rosa_init                  (rosa.cpp:150-199)
//
// Order of execution differs from the reference's source order on purpose: the leaf loop
// depends on positions only, so it runs BEFORE the normals (it also supplies the density
// estimate that sizes the kNN grid), and only the final pass carries normals.  Results are
// identical: VoxelGrid keys/counts never read the normals (PCL applyFilter steps 2-7).
//
// HBM layout: filtered cloud P as float4 AoS {x,y,z,1} (pcl::PointXYZ's own 16-byte layout:
// one coalesced 16 B load per point); the kNN copy Pc is the same points regrouped by grid
// cell with the filtered index in .w; normals Nrm float4 {nx,ny,nz,0}.  All tables (2 hash
// tables, voxel groups) are L2 resident at 100k points (< 10 MB of the 126 MB L2).
#include "rosa_impl.cuh"

namespace osep {

constexpr int NT = 256;                 // threads per block for N-scale kernels

// ------------------------------------------------------------------------------------------
// single-block exclusive scan (in place) over n ints; returns the total to every thread.
// 1024 threads, 4 items per thread per tile, carry across tiles.
__device__ int block_excl_scan(int* data, int n) {
    __shared__ int wsum[32];
    __shared__ int s_carry, s_tile;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    const int tile = blockDim.x * 4;
    for (int base = 0; base < n; base += tile) {
        const int i0 = base + tid * 4;
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? data[i0 + k] : 0;
        const int tsum = v[0] + v[1] + v[2] + v[3];
        int incl = tsum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int ws = (lane < nw) ? wsum[lane] : 0;
            int wi = ws;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, off); if (lane >= off) wi += t; }
            wsum[lane] = wi - ws;
            if (lane == 31) s_tile = wi;
        }
        __syncthreads();
        int excl = s_carry + wsum[wid] + (incl - tsum);
#pragma unroll
        for (int k = 0; k < 4; ++k) { if (i0 + k < n) data[i0 + k] = excl; excl += v[k]; }
        __syncthreads();
        if (tid == 0) s_carry += s_tile;
        __syncthreads();
    }
    return s_carry;
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void frame_init(FrameState* st, int n_in) {
    const unsigned epoch = st->epoch;
    FrameState z;
    memset(&z, 0, sizeof(z));
    z.epoch = epoch;                 // table epochs persist across frames
    z.g_h_last = st->g_h_last;       // so does the cell size of the kNN grid (a warm start: never changes a result)
    z.n_in = n_in;
    for (int d = 0; d < 3; ++d) { z.bb_lo[d] = 0x7fffffff; z.bb_hi[d] = (int)0x80000000; }
    *st = z;
}

// ---- R3: filter (rosa.cpp:76-82): keep iff finite and x*x + y*y + z*z <= lim^2
__device__ __forceinline__ bool keep_point(float x, float y, float z, float r2) {
    const float d2 = x * x + y * y + z * z;
    return fin(x) && fin(y) && fin(z) && d2 <= r2;
}

__global__ void k_filter_count(const FrameParams* __restrict__ prm, float r2, int* __restrict__ blk_cnt) {
    const float* __restrict__ in = prm->in; const int n = prm->n, stride = prm->stride;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    if (i < n) {
        const float* p = in + (size_t)i * stride;
        keep = keep_point(p[0], p[1], p[2], r2);
    }
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) blk_cnt[blockIdx.x] = c;
}

// single block: scan the per-block counts; R4 gate (rosa.cpp:91-95)
__global__ void k_filter_scan(int* blk_cnt, int nblk, FrameState* st, int min_points, const FrameParams* __restrict__ prm) {
    if (threadIdx.x == 0) frame_init(st, prm->n);          // first writer of the frame state (k_filter_count does not touch it)
    frame_stamp(st, TS_FILTER);
    const int total = block_excl_scan(blk_cnt, nblk);
    if (threadIdx.x == 0) {
        st->n_filt = total;
        if (total < min_points) st->status = OSEP_STAGE_1;
    }
}

__device__ void leaf_step(FrameState* st, int p, int max_points);
__device__ void grid_params(FrameState* st, int target_ppc);
__global__ void k_filter_scatter(const FrameParams* __restrict__ prm, float r2, const int* __restrict__ blk_off,
                                 float4* __restrict__ P, int* __restrict__ filt_idx, FrameState* st, int max_points, int target_ppc) {
    const float* __restrict__ in = prm->in; const int n = prm->n, stride = prm->stride;
    __shared__ int warp_cnt[NT / 32];
    __shared__ int s_last;
    __shared__ int s_lo[3], s_hi[3];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x < 3) { s_lo[threadIdx.x] = 0x7fffffff; s_hi[threadIdx.x] = (int)0x80000000; }
    float x = 0, y = 0, z = 0; bool keep = false;
    if (i < n) {
        const float* p = in + (size_t)i * stride;
        x = p[0]; y = p[1]; z = p[2];
        keep = keep_point(x, y, z, r2);
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[wid] = __popc(m);
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; ++w) woff += warp_cnt[w];
    if (keep) {
        const int pos = blk_off[blockIdx.x] + woff + __popc(m & ((1u << lane) - 1u));
        P[pos] = make_float4(x, y, z, 1.0f);
        filt_idx[pos] = i;
        atomicMin(&s_lo[0], f2ord(x)); atomicMax(&s_hi[0], f2ord(x));
        atomicMin(&s_lo[1], f2ord(y)); atomicMax(&s_hi[1], f2ord(y));
        atomicMin(&s_lo[2], f2ord(z)); atomicMax(&s_hi[2], f2ord(z));
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        if (s_lo[threadIdx.x] != 0x7fffffff) {
            atomicMin(&st->bb_lo[threadIdx.x], s_lo[threadIdx.x]);
            atomicMax(&st->bb_hi[threadIdx.x], s_hi[threadIdx.x]);
        }
    }
    // the last block sees the complete bounding box: R7 + set-up of VoxelGrid pass 0 here (no extra launch)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last && threadIdx.x == 0) { st->ticket = 0; __threadfence(); leaf_step(st, 0, max_points); grid_params(st, target_ppc); }
}

// ------------------------------------------------------------------------------------------
// R7 + R8: the leaf-size control law, one thread, no host round trip.
// step p = 0..8: consume the voxel count of pass p-1 exactly as the body of the loop at
// rosa.cpp:125-146 does, then (p < 8, not finished) set up pass p.
__device__ void leaf_step(FrameState* st, int p, int max_points) {
    if (st->status != 0) return;
    const float tol = 0.05f, leaf_min = 1e-4f, leaf_max = 1e4f;
    if (p == 0) {
        for (int d = 0; d < 3; ++d) { st->bmin[d] = ord2f(st->bb_lo[d]); st->bmax[d] = ord2f(st->bb_hi[d]); }
        // estimate_leaf_from_bbox (rosa.hpp:99-110)
        const float dx = smax(1e-6f, st->bmax[0] - st->bmin[0]);
        const float dy = smax(1e-6f, st->bmax[1] - st->bmin[1]);
        const float dz = smax(1e-6f, st->bmax[2] - st->bmin[2]);
        const float vol = dx * dy * dz;
        st->leaf = (max_points == 0) ? 0.0f : canon_cbrtf(vol / smax(1.0f, (float)(size_t)max_points));
        st->leaf_done = 0;
    }
    float leaf = st->leaf;
    int done = st->leaf_done;
    if (!done && p > 0) {
        const int N = st->pass_overflow ? st->n_filt : st->nvox;
        st->nvox_hist[p - 1] = N;
        if (N == 0) {
            leaf *= 0.5f;
            if (leaf < leaf_min) { leaf = leaf_min; done = 1; }
        } else {
            const float ratio_out = (float)N / (float)max_points;
            if (fabsf(ratio_out - 1.0f) <= tol) { done = 2; }      // converged: break before the update
            else {
                leaf *= canon_cbrtf(ratio_out);
                if (!fin(leaf)) leaf = 0.05f;
                if (leaf < leaf_min) leaf = leaf_min;
                if (leaf > leaf_max) leaf = leaf_max;
            }
        }
    }
    int active = 0;
    if (!done && p < LEAF_MAX_ITERS) {
        active = 1;
        st->leaf_used = leaf;
        st->leaf_hist[p] = leaf;
        st->n_pass = p + 1;
        // PCL VoxelGrid::applyFilter set-up
        const float inv = 1.0f / leaf;
        st->inv_leaf = inv;
        const long long ddx = (long long)((st->bmax[0] - st->bmin[0]) * inv) + 1;
        const long long ddy = (long long)((st->bmax[1] - st->bmin[1]) * inv) + 1;
        const long long ddz = (long long)((st->bmax[2] - st->bmin[2]) * inv) + 1;
        st->pass_overflow = (ddx * ddy * ddz > 2147483647LL) ? 1 : 0;
        for (int d = 0; d < 3; ++d) {
            const int mn = (int)floorf(st->bmin[d] * inv);
            const int mx = (int)floorf(st->bmax[d] * inv);
            st->minb[d] = mn; st->divb[d] = mx - mn + 1;
        }
        st->nvox = 0;
        st->epoch += 1;
    }
    st->pass_active = active;           // "pass active" flag read by k_vox_count
    st->leaf = leaf;
    st->leaf_done = done;
    if (p == LEAF_MAX_ITERS || done) {
        st->leaf_size = leaf;                                   // rosa.cpp:148
        if (done != 2) st->flags |= OSEP_FLAG_LEAF_NOT_CONVERGED;
        else st->flags &= ~OSEP_FLAG_LEAF_NOT_CONVERGED;
        if (st->pass_overflow) st->flags |= OSEP_FLAG_VOXEL_OVERFLOW;
    }
}

// voxel key of PCL VoxelGrid::applyFilter step 5
__device__ __forceinline__ unsigned voxel_key(const float4& p, float inv, int mb0, int mb1, int mb2, int db0, int db1) {
    const int i0 = (int)(floorf(p.x * inv) - (float)mb0);
    const int i1 = (int)(floorf(p.y * inv) - (float)mb1);
    const int i2 = (int)(floorf(p.z * inv) - (float)mb2);
    return (unsigned)(i0 + i1 * db0 + i2 * (db0 * db1));
}

// insert (epoch, key) into the open-addressing set; returns true when this thread created it
__device__ __forceinline__ bool vset_insert(unsigned long long* tab, unsigned mask, unsigned epoch, unsigned key) {
    const unsigned long long want = ((unsigned long long)epoch << 32) | key;
    unsigned s = hash_u32(key) & mask;
    while (true) {
        unsigned long long cur = tab[s];
        if (cur == want) return false;
        if ((unsigned)(cur >> 32) != epoch) {
            const unsigned long long old = atomicCAS(&tab[s], cur, want);
            if (old == cur) return true;
            if (old == want) return false;
            if ((unsigned)(old >> 32) != epoch) continue;       // stale value changed under us: retry this slot
        }
        s = (s + 1) & mask;
    }
}

// count-only VoxelGrid pass: number of distinct voxel keys at the current leaf
__global__ void k_vox_count(const float4* __restrict__ P, FrameState* st, unsigned long long* vtab, unsigned mask, int next_step, int max_points, int target_ppc) {
    if (next_step == 1) frame_stamp(st, TS_LEAF0);
    if (st->status != 0) return;
    if (st->pass_active && !st->pass_overflow) {
        const int n = st->n_filt;
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        bool isnew = false;
        if (i < n) {
            const unsigned key = voxel_key(P[i], st->inv_leaf, st->minb[0], st->minb[1], st->minb[2], st->divb[0], st->divb[1]);
            isnew = vset_insert(vtab, mask, st->epoch, key);
        }
        const unsigned m = __ballot_sync(0xffffffffu, isnew);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&st->nvox, __popc(m));
    }
    // the last block to finish runs the control law for the next pass (rosa.cpp:130-146): no extra launch
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        st->ticket = 0; __threadfence(); leaf_step(st, next_step, max_points);
    }
}

// ------------------------------------------------------------------------------------------
// kNN grid: sparse uniform grid in a hash table keyed by the linear cell id.
__device__ void grid_set_h(FrameState* st, float h) {
    if (!(h > 1e-6f) || !fin(h)) h = 1.0f;
    float ext = 0.0f;
    for (int d = 0; d < 3; ++d) ext = fmaxf(ext, st->bmax[d] - st->bmin[d]);
    if (ext / h > 2.0e6f) h = ext / 2.0e6f;         // 21 bits per axis in the 64-bit cell key
    st->g_h = h; st->g_inv_h = 1.0f / h;
    for (int d = 0; d < 3; ++d) {
        int dim = (int)floorf((st->bmax[d] - st->bmin[d]) * st->g_inv_h) + 1;
        if (dim < 1) dim = 1; if (dim > 2097151) dim = 2097151;
        st->g_dim[d] = dim;
    }
}

__device__ __forceinline__ unsigned long long cell_key1(int x, int y, int z) {
    return (((unsigned long long)z << 42) | ((unsigned long long)y << 21) | (unsigned long long)x) + 1ull;
}
__device__ __forceinline__ unsigned hash_u64(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}
// Home slot of a grid cell: the 4x4x4 block of cells is hashed, the position inside the block (6 bits) selects the
// slot inside a 64-slot group.  Cells of one block are distinct slots of one group, so (a) the 27 cells a query block
// touches sit in a few groups of the table and (b) the cell data, allocated in slot order, is spatially coherent:
// neighbouring cells share cache lines and neighbouring work units share L2 lines (maps larger than L2: tiled path).
__device__ __forceinline__ unsigned cell_slot(int x, int y, int z, unsigned mask) {
    const unsigned fine = (unsigned)((x & 3) | ((y & 3) << 2) | ((z & 3) << 4));
    const unsigned long long coarse = (((unsigned long long)(z >> 2) << 42) | ((unsigned long long)(y >> 2) << 21) | (unsigned long long)(x >> 2));
    return ((hash_u64(coarse + 0x9e3779b97f4a7c15ull) << 6) | fine) & mask;
}

// Cell size of the kNN grid, chosen by the last block of the filter (the bounding box is complete there), so that the
// kNN / normals branch starts right after the filter instead of after VoxelGrid pass 0: a stream's frames resemble each
// other, so the size that hit the target occupancy on the previous frame is reused; a cold handle guesses from the faces
// of the bounding box (surface-like cloud).  grid_refine checks the real occupancy after the first insertion pass and
// asks for one rebuild when the guess was off (cold start, scene change): 14 us, the exception.
__device__ void grid_params(FrameState* st, int target_ppc) {
    if (st->status != 0) return;
    const int n = st->n_filt;
    float h = st->g_h_last;
    if (!(h > 0.0f)) {
        const float dx = st->bmax[0] - st->bmin[0], dy = st->bmax[1] - st->bmin[1], dz = st->bmax[2] - st->bmin[2];
        h = sqrtf((float)target_ppc * fmaxf(1e-12f, (dx * dy + dy * dz + dx * dz) * (1.0f / 3.0f)) / (float)(n > 0 ? n : 1));
    }
    grid_set_h(st, h);
    st->g_cells_used = 0; st->n_units = 0; st->n_slow = 0; st->n_over = 0; st->unit_ctr = 0; st->n_retry = 0; st->n_grow = 0; st->n_cells = 0;
    st->g_ncells = 0; st->g_redo = 0;
}

// After the first insertion pass the real occupancy is known: if the estimate was off (the surface-like scaling
// breaks when leaf0 is larger than the structures, e.g. multi-structure maps), pick the cell size that would
// give the target and ask for one rebuild.  Only the speed of the search depends on h, never its result.
__device__ __forceinline__ void grid_refine(FrameState* st, int target_ppc) {
    const int n = st->n_filt, nc = *(volatile int*)&st->g_ncells;
    if (nc <= 0) return;
    const float occ = (float)n / (float)nc;
    const float h_fit = st->g_h * sqrtf((float)target_ppc / occ);
    st->g_h_last = h_fit;
    if (occ > 1.4f * (float)target_ppc || occ < 0.6f * (float)target_ppc) {
        grid_set_h(st, h_fit);
        st->g_redo = 1; st->g_ncells = 0;
    }
}
__global__ void k_grid_clear_if_redo(const FrameState* __restrict__ st, unsigned long long* gtab, int* g_cnt, int T) {
    if (st->status != 0 || !st->g_redo) return;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < T) { gtab[s] = 0ull; g_cnt[s] = 0; }
}

__device__ __forceinline__ void cell_of(const FrameState* st, float x, float y, float z, int& cx, int& cy, int& cz) {
    cx = (int)floorf((x - st->bmin[0]) * st->g_inv_h);
    cy = (int)floorf((y - st->bmin[1]) * st->g_inv_h);
    cz = (int)floorf((z - st->bmin[2]) * st->g_inv_h);
    cx = max(0, min(cx, st->g_dim[0] - 1)); cy = max(0, min(cy, st->g_dim[1] - 1)); cz = max(0, min(cz, st->g_dim[2] - 1));
}

// gtab stores key+1 (0 = empty) so one memset clears keys, counts and fills together.
// pass 0 = first build; pass 1 = rebuild, runs only when k_grid_refine asked for it.
__global__ void k_grid_insert(const float4* __restrict__ P, FrameState* st, unsigned long long* gtab, int* g_cnt,
                              int* __restrict__ p_slot, unsigned mask, int pass, int target_ppc) {
    frame_stamp(st, pass == 0 ? TS_GRID : TS_COUNT);
    __shared__ int s_last;
    if (st->status != 0) return;
    if (pass == 1 && !st->g_redo) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool created = false;
    if (i < st->n_filt) {
        const float4 p = P[i];
        int cx, cy, cz; cell_of(st, p.x, p.y, p.z, cx, cy, cz);
        const unsigned long long key1 = cell_key1(cx, cy, cz);
        unsigned s = cell_slot(cx, cy, cz, mask);
        while (true) {
            const unsigned long long cur = gtab[s];
            if (cur == key1) break;
            if (cur == 0ull) { const unsigned long long old = atomicCAS(&gtab[s], 0ull, key1); if (old == 0ull) { created = true; break; } if (old == key1) break; }
            s = (s + 1) & mask;
        }
        atomicAdd(&g_cnt[s], 1);
        p_slot[i] = (int)s;
    }
    const unsigned m = __ballot_sync(0xffffffffu, created);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&st->g_ncells, __popc(m));
    if (pass != 0) return;
    // the last block of the first pass knows the real occupancy: decide about the rebuild here (no extra launch)
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); s_last = (atomicAdd(&st->ticket_grid, 1) == (int)gridDim.x - 1) ? 1 : 0; }
    __syncthreads();
    if (s_last && threadIdx.x == 0) { st->ticket_grid = 0; __threadfence(); grid_refine(st, target_ppc); }
}
#ifndef OSEP_QCH
#define OSEP_QCH 12
#endif
constexpr int QCH = OSEP_QCH;            // queries per kNN work unit
__global__ void k_grid_alloc(FrameState* st, const int* __restrict__ g_cnt, int* __restrict__ g_start, int T, int2* __restrict__ units,
                             int* __restrict__ g_cell, int* __restrict__ cell_slots, int cell_cap) {
    __shared__ int s_pts, s_units, s_cells, b_pts, b_units, b_cells;
    if (st->status != 0) return;
    if (threadIdx.x == 0) { s_pts = 0; s_units = 0; s_cells = 0; }
    __syncthreads();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = s < T ? g_cnt[s] : 0;
    // dense cells have long candidate lists: fewer queries per unit there, so more warps share the work
    const int chunk = (c >= 64) ? 2 : QCH;
    const int nu = (c + chunk - 1) / chunk;
    int op = 0, ou = 0, oc = 0;
    if (c > 0) { op = atomicAdd(&s_pts, c); ou = atomicAdd(&s_units, nu); oc = atomicAdd(&s_cells, 1); }     // block-local cursors: three global atomics per block
    __syncthreads();
    if (threadIdx.x == 0 && s_pts > 0) { b_pts = atomicAdd(&st->g_cells_used, s_pts); b_units = atomicAdd(&st->n_units, s_units); b_cells = atomicAdd(&st->n_cells, s_cells); }
    __syncthreads();
    if (c > 0) {
        g_start[s] = b_pts + op;
        const int ub = b_units + ou;
        for (int u = 0; u < nu; ++u) units[ub + u] = make_int2(s, (u * chunk) | (chunk << 24));
        const int ci = b_cells + oc;                      // compact cell id (thread-per-query kNN: row of the neighbour table)
        g_cell[s] = ci;
        if (ci < cell_cap) cell_slots[ci] = s;
    }
}
// neighbour table: row ci = (start, count) of the occupied cells among the 27 around cell ci, one warp per cell (lanes probe
// the hash table in parallel), so that the query threads of k_knn_scan read one packed row instead of running 27 dependent
// probes each.  The warp also runs a PILOT query for the cell's first point: the radius^2 that holds `target` of the
// block's points (bisection over the staged distances) -- the scan radius of every query of the cell (nbr_tau).
constexpr int PILOT_CAP = 256;
__global__ void __launch_bounds__(256) k_grid_nbrs(const FrameState* __restrict__ st, const float4* __restrict__ Pc, const unsigned long long* __restrict__ gtab,
                            const int* __restrict__ g_cnt, const int* __restrict__ g_start, unsigned mask, const int* __restrict__ cell_slots, int cell_cap,
                            int2* __restrict__ nbr, float* __restrict__ nbr_tau, float target) {
    __shared__ float sd[8][PILOT_CAP];
    if (st->status != 0) return;
    const int nc = min(st->n_cells, cell_cap);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int dimx = st->g_dim[0], dimy = st->g_dim[1], dimz = st->g_dim[2];
    for (int ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; ci < nc; ci += (gridDim.x * blockDim.x) >> 5) {
        const unsigned long long key = gtab[cell_slots[ci]] - 1ull;
        const int cx = (int)(key & 0x1fffffull), cy = (int)((key >> 21) & 0x1fffffull), cz = (int)(key >> 42);
        int2 e = make_int2(0, 0);
        if (lane < 27) {
            const int x = cx + (lane % 3) - 1, y = cy + ((lane / 3) % 3) - 1, z = cz + (lane / 9) - 1;
            if (x >= 0 && x < dimx && y >= 0 && y < dimy && z >= 0 && z < dimz) {
                const unsigned long long key1 = cell_key1(x, y, z);
                unsigned s = cell_slot(x, y, z, mask);
                unsigned long long cur;
                while ((cur = gtab[s]) != key1 && cur != 0ull) s = (s + 1) & mask;
                if (cur != 0ull) e = make_int2(g_start[s], g_cnt[s]);
            }
        }
        // occupied cells first (rows are scanned until the first empty entry)
        const unsigned occ = __ballot_sync(0xffffffffu, e.y > 0);
        if (e.y > 0) nbr[(size_t)ci * 27 + __popc(occ & ((1u << lane) - 1u))] = e;
        if (lane >= __popc(occ) && lane < 27) nbr[(size_t)ci * 27 + lane] = make_int2(0, 0);
        // pilot: distances from the cell's first point to the block's points (the first PILOT_CAP of them)
        const int own = __shfl_sync(0xffffffffu, e.x, 13);
        const float4 q = Pc[own];
        int total = 0;
        __syncwarp();
        for (unsigned mm = occ; mm; mm &= mm - 1) {
            const int src = __ffs(mm) - 1;
            const int b = __shfl_sync(0xffffffffu, e.x, src), m = __shfl_sync(0xffffffffu, e.y, src);
            for (int j = lane; j < m && total + j < PILOT_CAP; j += 32) { const float4 p = __ldg(Pc + b + j); sd[wib][total + j] = dist2(q.x, q.y, q.z, p.x, p.y, p.z); }
            total = min(total + m, PILOT_CAP);
        }
        __syncwarp();
        float v[PILOT_CAP / 32];
        float vmax = 0.0f;
#pragma unroll
        for (int r = 0; r < PILOT_CAP / 32; ++r) { const int t = r * 32 + lane; v[r] = t < total ? sd[wib][t] : 3.0e38f; if (t < total) vmax = fmaxf(vmax, v[r]); }
        vmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(vmax)));
        float tau = 3.0e38f;                              // fewer than `target` points in the block: the scan radius is the block's bound
        if ((float)total > target) {
            float lo = 0.0f, hi = vmax * 1.0001f + 1e-30f;
            for (int it = 0; it < 14; ++it) {
                const float mid = 0.5f * (lo + hi);
                int c = 0;
#pragma unroll
                for (int r = 0; r < PILOT_CAP / 32; ++r) c += (v[r] < mid) ? 1 : 0;
                c = __reduce_add_sync(0xffffffffu, c);
                if ((float)c >= target) hi = mid; else lo = mid;
            }
            tau = hi;
        }
        if (lane == 0) nbr_tau[ci] = tau;
    }
}

__global__ void k_grid_scatter(const float4* __restrict__ P, const FrameState* __restrict__ st, const int* __restrict__ p_slot,
                               const int* __restrict__ g_start, int* g_fill, float4* __restrict__ Pc, const int* __restrict__ g_cell, int* __restrict__ pc_cell) {
    if (st->status != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st->n_filt) return;
    const int s = p_slot[i];
    const int pos = g_start[s] + atomicAdd(&g_fill[s], 1);
    float4 p = P[i]; p.w = __int_as_float(i);
    Pc[pos] = p;
    pc_cell[pos] = g_cell[s];
}

// ---- register-resident top-K by (d2, index); sentinels (inf, INT_MAX) mark empty slots
template <int K> struct TopK {
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

// PCL computeMeanAndCovarianceMatrix (float, shifted by the nearest neighbour, sums in neighbour order)
// -> smallest eigenvector -> flipNormalTowardsViewpoint(point, 0,0,0) (PCL NormalEstimation::computeFeature)
template <class IDX>
__device__ __forceinline__ float4 normal_from_neighbours(const float4* __restrict__ P, const float4& q, IDX idx, int k_eff) {
    const float qnan = __int_as_float(0x7fc00000);
    if (k_eff < 3) return make_float4(qnan, qnan, qnan, 0.0f);
    const float4 Kp = P[idx(0)];
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0;
    for (int u = 0; u < k_eff; ++u) {
        const float4 c = P[idx(u)];
        const float x = c.x - Kp.x, y = c.y - Kp.y, z = c.z - Kp.z;
        a0 += x * x; a1 += x * y; a2 += x * z; a3 += y * y; a4 += y * z; a5 += z * z; a6 += x; a7 += y; a8 += z;
    }
    const float cnt = (float)k_eff;
    a0 /= cnt; a1 /= cnt; a2 /= cnt; a3 /= cnt; a4 /= cnt; a5 /= cnt; a6 /= cnt; a7 /= cnt; a8 /= cnt;
    const float c00 = a0 - a6 * a6, c01 = a1 - a6 * a7, c02 = a2 - a6 * a8;
    const float c11 = a3 - a7 * a7, c12 = a4 - a7 * a8, c22 = a5 - a8 * a8;
    const Eig3 E = eig3_sym_dev(c00, c01, c11, c02, c12, c22);
    f3 nv = eig_col(E, 0);
    const float vx = 0.0f - q.x, vy = 0.0f - q.y, vz = 0.0f - q.z;
    const float cos_theta = (vx * nv.x + vy * nv.y + vz * nv.z);
    if (cos_theta < 0.0f) { nv.x *= -1.0f; nv.y *= -1.0f; nv.z *= -1.0f; }
    return make_float4(nv.x, nv.y, nv.z, 0.0f);
}

// distance from q to the faces of the scanned cell block [x0..x1]x[y0..y1]x[z0..z1] that still have unscanned
// cells behind them, minus a float-safety margin: every point closer than this is inside the block.
__device__ __forceinline__ float block_bound(const FrameState* st, float rx, float ry, float rz, int x0, int x1, int y0, int y1, int z0, int z1) {
    const float h = st->g_h;
    float bound = 3.0e38f;
    if (x0 > 0) bound = fminf(bound, rx - (float)x0 * h);
    if (x1 < st->g_dim[0] - 1) bound = fminf(bound, (float)(x1 + 1) * h - rx);
    if (y0 > 0) bound = fminf(bound, ry - (float)y0 * h);
    if (y1 < st->g_dim[1] - 1) bound = fminf(bound, (float)(y1 + 1) * h - ry);
    if (z0 > 0) bound = fminf(bound, rz - (float)z0 * h);
    if (z1 < st->g_dim[2] - 1) bound = fminf(bound, (float)(z1 + 1) * h - rz);
    return bound - 1e-3f * h;
}

// R5 slow path: one WARP per query, exact kNN by ring expansion (R = 1, 2, 3, 4, 6, 8, 12, ...), then everything.
// Handles whatever the cell-cooperative kernels could not certify: points whose k nearest neighbours lie many cells away
// (the sparse far end of a LiDAR frame; beyond a block of n / 8 cells the lanes stride over the whole cloud instead).  The lanes probe the cells of the NEW SHELL of a ring in parallel (the block
// scanned so far is not rescanned) and push the points they find into a register-resident sorted top-K of their own; the 32
// lists are then merged into the frame's (d2, index) order by k rounds of a warp arg-min over the list heads, which gives
// the k-th distance for the certification test and re-seeds the lists.  One thread per query (the first version) walked
// 15 k cells with a dependent hash probe each for such a point: 1.9 ms for 323 queries of a 42 k-point frame.
template <int K>
__global__ void __launch_bounds__(128) k_knn_exact(const float4* __restrict__ Pc, const FrameState* __restrict__ st,
                                                   const unsigned long long* __restrict__ gtab, const int* __restrict__ g_cnt,
                                                   const int* __restrict__ g_start, unsigned mask,
                                                   int* __restrict__ knn_idx, int k_req, const int* __restrict__ slow_list) {
    if (st->status != 0) return;
    constexpr int RS = (K + 31) / 32;                // merged entries per lane: entry r lives in lane r & 31, slot r >> 5
    const int n = st->n_filt;
    const int nwork = st->n_slow;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int k_eff = min(k_req, n);
    const float finf = __int_as_float(0x7f800000);
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nwork; w += warps) {
        const int t = slow_list[w];
        const float4 q = Pc[t];
        const int qi = __float_as_int(q.w);
        const int dimx = st->g_dim[0], dimy = st->g_dim[1], dimz = st->g_dim[2];
        int cx, cy, cz; cell_of(st, q.x, q.y, q.z, cx, cy, cz);
        const float rx = q.x - st->bmin[0], ry = q.y - st->bmin[1], rz = q.z - st->bmin[2];
        TopK<K> tk; tk.init();
        float resd[RS]; int resi[RS];
        // merge the 32 per-lane lists: the k_eff smallest (d2, index) pairs, entry r to lane r & 31; returns the k-th pair
        auto merge = [&](float& wd, int& wi) {
#pragma unroll
            for (int u = 0; u < RS; ++u) { resd[u] = finf; resi[u] = 0x7fffffff; }
            for (int r = 0; r < k_eff; ++r) {
                float bd = tk.d[0]; int bi = tk.id[0]; int bl = lane;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const float od = __shfl_xor_sync(0xffffffffu, bd, off); const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    const int ol = __shfl_xor_sync(0xffffffffu, bl, off);
                    if (nb_less(od, oi, bd, bi)) { bd = od; bi = oi; bl = ol; }
                }
                if (bl == lane && bi == tk.id[0]) {       // this lane's head won: pop it
#pragma unroll
                    for (int u = 0; u < K - 1; ++u) { tk.d[u] = tk.d[u + 1]; tk.id[u] = tk.id[u + 1]; }
                    tk.d[K - 1] = finf; tk.id[K - 1] = 0x7fffffff;
                }
#pragma unroll
                for (int u = 0; u < RS; ++u) if ((r >> 5) == u && (r & 31) == lane) { resd[u] = bd; resi[u] = bi; }
            }
            const int kl = (k_eff - 1) & 31, ks = (k_eff - 1) >> 5;
            float md = finf; int mi = 0x7fffffff;
#pragma unroll
            for (int u = 0; u < RS; ++u) if (u == ks) { md = resd[u]; mi = resi[u]; }
            wd = __shfl_sync(0xffffffffu, md, kl); wi = __shfl_sync(0xffffffffu, mi, kl);
            tk.init();                                     // the merged list is the state carried into the next ring
#pragma unroll
            for (int u = 0; u < RS; ++u) if (resi[u] != 0x7fffffff) tk.push(resd[u], resi[u]);
        };
        int ox0 = 1, ox1 = 0, oy0 = 1, oy1 = 0, oz0 = 1, oz1 = 0;      // block scanned so far (empty)
        // the k-th pair found so far bounds the answer: candidates beyond it are dropped before they reach a list (once a
        // ring has found k points almost every later candidate is; without it every lane's list, re-seeded with <= 2
        // entries, took every candidate through the insertion network)
        float wd = finf; int wi = 0x7fffffff;
        for (int R = 1;;) {
            const long long side = 2ll * R + 1;
            if (R > 3 && side * side * side > (long long)n / 8) {
                // the block would cost more than every point: lanes stride over the cloud
                tk.init();
                for (int e0 = lane; e0 < n; e0 += 32 * 8) {              // eight independent loads in flight per lane
                    float4 c8[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { const int e = e0 + 32 * u; c8[u] = Pc[e < n ? e : e0]; }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const float dd = dist2(q.x, q.y, q.z, c8[u].x, c8[u].y, c8[u].z); const int ii = __float_as_int(c8[u].w);
                        if (e0 + 32 * u < n && !nb_less(wd, wi, dd, ii)) tk.push(dd, ii);
                    }
                }
                merge(wd, wi);
                break;
            }
            const int x0 = max(cx - R, 0), x1 = min(cx + R, dimx - 1);
            const int y0 = max(cy - R, 0), y1 = min(cy + R, dimy - 1);
            const int z0 = max(cz - R, 0), z1 = min(cz + R, dimz - 1);
            const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1;
            const int total = nx * ny * nz;
            for (int c0 = lane; c0 < total; c0 += 32 * 4) {            // four cells per lane and round: their first probes are independent loads
                unsigned s4[4]; unsigned long long key4[4], cur4[4]; bool v4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = c0 + 32 * u;
                    const int x = x0 + c % nx, y = y0 + (c / nx) % ny, z = z0 + c / (nx * ny);
                    v4[u] = c < total && !(x >= ox0 && x <= ox1 && y >= oy0 && y <= oy1 && z >= oz0 && z <= oz1);      // inside: scanned by an earlier ring
                    key4[u] = cell_key1(x, y, z); s4[u] = cell_slot(x, y, z, mask);
                    cur4[u] = v4[u] ? gtab[s4[u]] : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!v4[u]) continue;
                    unsigned s = s4[u]; unsigned long long cur = cur4[u];
                    while (cur != key4[u] && cur != 0ull) { s = (s + 1) & mask; cur = gtab[s]; }
                    if (cur == 0ull) continue;
                    const int b = g_start[s], cn = g_cnt[s];
#pragma unroll 4
                    for (int e = b; e < b + cn; ++e) {
                        const float4 cc = Pc[e];
                        const float dd = dist2(q.x, q.y, q.z, cc.x, cc.y, cc.z); const int ii = __float_as_int(cc.w);
                        if (!nb_less(wd, wi, dd, ii)) tk.push(dd, ii);
                    }
                }
            }
            merge(wd, wi);
            if (x0 == 0 && y0 == 0 && z0 == 0 && x1 == dimx - 1 && y1 == dimy - 1 && z1 == dimz - 1) break;
            const float bound = block_bound(st, rx, ry, rz, x0, x1, y0, y1, z0, z1);
            if (bound > 0.0f && wi != 0x7fffffff && wd < bound * bound) break;
            ox0 = x0; ox1 = x1; oy0 = y0; oy1 = y1; oz0 = z0; oz1 = z1;
            // next ring: 2, 3, 4, 6, 9, ... while fewer than k points are known; once k are, their k-th distance bounds the
            // answer and the block that certifies it is known -- go there directly (the test above stays the only exit)
            int Rn = (R < 4) ? R + 1 : R + (R >> 1);
            if (wi != 0x7fffffff) { const float need = sqrtf(wd) * st->g_inv_h + 1.0f; if (need < 1.0e6f) Rn = max(R + 1, (int)need); }
            R = Rn;
        }
#pragma unroll
        for (int u = 0; u < RS; ++u) { const int r = u * 32 + lane; if (r < k_eff) knn_idx[(size_t)qi * k_eff + r] = resi[u]; }
    }
}

// R5 fast path: one WARP per work unit = up to QCH queries of one grid cell.  The queries of a cell share
// their candidate cells: the (start, count) table of the non-empty cells of the (2R+1)^3 block is built
// once per unit (lanes probe the hash table in parallel) and the candidate points are staged in shared
// memory with coalesced float4 loads (units with more than KNN_STAGE candidates hand their queries to
// k_knn_big).  Per query the warp keeps the distances of the staged candidates in registers and
//   1. tries the radius just above the previous query's k-th distance, else counts the candidates inside
//      8 nested radii (bound^2 / 2^m) with per-lane compares,
//   2. compacts the candidates of the smallest radius that still holds >= k points (exact: every point
//      closer than `bound` is inside the scanned block, and the k nearest by (d2, index) are all inside
//      any radius that holds >= k points),
//   3. sorts them as packed (d2, index) keys with a register bitonic network.
// Queries that cannot be certified within R <= 2 (isolated points) go to k_knn_exact.
constexpr int KNN_LIST = 64;
constexpr int KNN_TAB = 125;
constexpr int KNN_WARPS = 4;
constexpr int KNN_STAGE = 512;
#ifndef OSEP_KNN_GUESS
#define OSEP_KNN_GUESS 1.35f
#endif

// flat candidate index t -> (cell entry, offset): largest ci with pre[ci] <= t (pre = exclusive prefix of the cell
// counts, nt <= 125 entries => 7 halvings).  Lets the staging loops issue independent loads for 32 candidates per
// round instead of walking the cells one after the other.
__device__ __forceinline__ int flat_cell(const int* pre, int nt, int t) {
    int lo = 0, hi = nt - 1;
#pragma unroll
    for (int step = 0; step < 7; ++step) { const int mid = (lo + hi + 1) >> 1; if (pre[mid] <= t) lo = mid; else hi = mid - 1; }
    return lo;
}

// ascending bitonic sort of 64 packed keys (d2 bits << 32 | index; d2 >= 0 so the bit pattern orders like the
// float and the packed compare is exactly the canonical (d2, index) order), two keys per lane: element r*32+lane.
__device__ __forceinline__ void bitonic64(unsigned long long& a0, unsigned long long& a1) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j == 32) {                       // partner = other register of the same lane (k == 64: ascending)
                const unsigned long long lo = a0 < a1 ? a0 : a1, hi = a0 < a1 ? a1 : a0;
                a0 = lo; a1 = hi;
            } else {
                const unsigned long long b0 = __shfl_xor_sync(0xffffffffu, a0, j), b1 = __shfl_xor_sync(0xffffffffu, a1, j);
                const bool lower = (lane & j) == 0;
                const bool up0 = ((lane & k) == 0) || (k == 64);
                const bool up1 = (k == 64) ? true : (((lane + 32) & k) == 0);
                const bool keepmin0 = (lower == up0), keepmin1 = (lower == up1);
                a0 = keepmin0 ? (a0 < b0 ? a0 : b0) : (a0 < b0 ? b0 : a0);
                a1 = keepmin1 ? (a1 < b1 ? a1 : b1) : (a1 < b1 ? b1 : a1);
            }
        }
    }
}

// 32-key variant (one key per lane) for lists that fit one register
__device__ __forceinline__ void bitonic32(unsigned long long& a0) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const unsigned long long b0 = __shfl_xor_sync(0xffffffffu, a0, j);
            const bool lower = (lane & j) == 0;
            const bool up = ((lane & k) == 0) || (k == 32);
            a0 = (lower == up) ? (a0 < b0 ? a0 : b0) : (a0 < b0 ? b0 : a0);
        }
    }
}

template <int K>
__global__ void __launch_bounds__(KNN_WARPS * 32) k_knn_cells(const float4* __restrict__ Pc, FrameState* st,
                                                             const unsigned long long* __restrict__ gtab, const int* __restrict__ g_cnt,
                                                             const int* __restrict__ g_start, unsigned mask, const int2* __restrict__ units,
                                                             int* __restrict__ knn_idx, int k_req, int* __restrict__ slow_list,
                                                             int* __restrict__ over_list) {
    __shared__ int2 tab[KNN_WARPS][KNN_TAB];
    __shared__ int pre[KNN_WARPS][KNN_TAB + 1];
    __shared__ unsigned long long keys[KNN_WARPS][KNN_LIST];
    __shared__ float4 stage[KNN_WARPS][KNN_STAGE];
    if (st->status != 0) return;
    const int n = st->n_filt;
    const int k_eff = min(k_req, n);
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int dimx = st->g_dim[0], dimy = st->g_dim[1], dimz = st->g_dim[2];
    const float h = st->g_h;
    const float diag2 = ((float)dimx * h) * ((float)dimx * h) + ((float)dimy * h) * ((float)dimy * h) + ((float)dimz * h) * ((float)dimz * h);
    const int n_work = st->n_units;
    const unsigned long long KINF = ~0ull;
    while (true) {
        int unit = 0;
        if (lane == 0) unit = atomicAdd(&st->unit_ctr, 1);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= n_work) break;
        const int2 un = units[unit];
        const int cstart = g_start[un.x], ccnt = g_cnt[un.x];
        const int q0 = un.y & 0xffffff;
        const int nq = min(un.y >> 24, ccnt - q0);
        int cx, cy, cz;
        { const float4 f = Pc[cstart]; cell_of(st, f.x, f.y, f.z, cx, cy, cz); }
        unsigned todo = (1u << nq) - 1u;
        bool overflow = false;
        for (int R = 1; R <= 2 && todo && !overflow; ++R) {
            const int x0 = max(cx - R, 0), x1 = min(cx + R, dimx - 1);
            const int y0 = max(cy - R, 0), y1 = min(cy + R, dimy - 1);
            const int z0 = max(cz - R, 0), z1 = min(cz + R, dimz - 1);
            const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1;
            const int ncell = nx * ny * nz;
            const bool whole = (x0 == 0 && y0 == 0 && z0 == 0 && x1 == dimx - 1 && y1 == dimy - 1 && z1 == dimz - 1);
            int nt = 0, total = 0;
            __syncwarp();
            for (int c0 = 0; c0 < ncell; c0 += 32) {
                const int c = c0 + lane;
                int2 e = make_int2(0, 0);
                if (c < ncell) {
                    const int x = x0 + c % nx, y = y0 + (c / nx) % ny, z = z0 + c / (nx * ny);
                    const unsigned long long key1 = cell_key1(x, y, z);
                    unsigned s = cell_slot(x, y, z, mask);
                    unsigned long long cur;
                    while ((cur = gtab[s]) != key1 && cur != 0ull) s = (s + 1) & mask;
                    if (cur != 0ull) e = make_int2(g_start[s], g_cnt[s]);
                }
                const unsigned m = __ballot_sync(0xffffffffu, e.y > 0);
                int incl = e.y;                                   // inclusive scan of the cell counts over the lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                if (e.y > 0) { const int slot = nt + __popc(m & lt); tab[wib][slot] = e; pre[wib][slot] = total + incl - e.y; }
                nt += __popc(m);
                total += __shfl_sync(0xffffffffu, incl, 31);
            }
            __syncwarp();
            if (total > KNN_STAGE) { overflow = true; break; }
            // stage: flat copy of the candidate cells, 32 candidates per round, all rounds' loads independent
#pragma unroll 4
            for (int t0 = 0; t0 < total; t0 += 32) {
                const int t = t0 + lane;
                if (t < total) { const int ci = flat_cell(pre[wib], nt, t); stage[wib][t] = Pc[tab[wib][ci].x + (t - pre[wib][ci])]; }
            }
            __syncwarp();
            constexpr int RPL = KNN_STAGE / 32;
            const int nr = (total + 31) >> 5;                   // rounds of 32 staged candidates actually in use (uniform)
            float guess = 0.0f;                               // radius^2 just above the previous query's k-th distance
            for (int qi = 0; qi < nq; ++qi) {
                if (!((todo >> qi) & 1u)) continue;
                const float4 q = Pc[cstart + q0 + qi];
                float b2;
                if (whole) b2 = 4.0f * diag2 + 1.0f;
                else {
                    const float bound = block_bound(st, q.x - st->bmin[0], q.y - st->bmin[1], q.z - st->bmin[2], x0, x1, y0, y1, z0, z1);
                    b2 = bound > 0.0f ? bound * bound : 0.0f;
                }
                unsigned long long k0 = KINF, k1 = KINF;
                // the distances of this lane's candidates (slots r*32+lane of the staged list) live in registers for
                // all passes of the query
                float dq[RPL]; int iq[RPL];
#pragma unroll
                for (int r = 0; r < RPL; ++r) {
                    const int t = r * 32 + lane;
                    dq[r] = 3.0e38f; iq[r] = 0;
                    if (r < nr && t < total) { const float4 c = stage[wib][t]; dq[r] = dist2(q.x, q.y, q.z, c.x, c.y, c.z); iq[r] = __float_as_int(c.w); }
                }
                auto collect = [&](float thr) -> int {
                    int ln = 0;
#pragma unroll
                    for (int r = 0; r < RPL; ++r) {
                        if (r >= nr) break;
                        const bool in = dq[r] < thr;
                        const unsigned m = __ballot_sync(0xffffffffu, in);
                        const int pos = ln + __popc(m & lt);
                        if (in && pos < KNN_LIST) keys[wib][pos] = ((unsigned long long)__float_as_uint(dq[r]) << 32) | (unsigned)iq[r];
                        ln += __popc(m);
                    }
                    __syncwarp();
                    k0 = (lane < ln && lane < KNN_LIST) ? keys[wib][lane] : KINF;
                    k1 = (lane + 32 < ln && lane + 32 < KNN_LIST) ? keys[wib][lane + 32] : KINF;
                    __syncwarp();
                    return ln;
                };
                auto population = [&](float thr) -> int {
                    int cm = 0;
#pragma unroll
                    for (int r = 0; r < RPL; ++r) if (r < nr) cm += (dq[r] < thr) ? 1 : 0;
                    return __reduce_add_sync(0xffffffffu, cm);
                };
                int ln = -1; float hi = 0.0f;
                if (guess > 0.0f && guess < b2) { hi = guess; ln = collect(hi); if (ln < k_eff || ln > KNN_LIST) ln = -1; }
                if (ln < 0) {
                    // population of 8 nested radii b2 / 2^m: the smallest one that still holds >= k points usually holds
                    // fewer than 2k (area halves per level), so the list fits without any bisection pass
                    constexpr int NR = 8;
                    int cr[NR];
                    float tr[NR];
#pragma unroll
                    for (int m = 0; m < NR; ++m) { cr[m] = 0; tr[m] = b2 * (1.0f / (float)(1 << m)); }
#pragma unroll
                    for (int r = 0; r < RPL; ++r) {
                        if (r >= nr) break;
#pragma unroll
                        for (int m = 0; m < NR; ++m) cr[m] += (dq[r] < tr[m]) ? 1 : 0;
                    }
#pragma unroll
                    for (int m = 0; m < NR; ++m) cr[m] = __reduce_add_sync(0xffffffffu, cr[m]);
                    if (cr[0] < k_eff) continue;              // not certifiable at this R
                    float lo = 0.0f; int cnt = cr[0];
                    hi = tr[0];
#pragma unroll
                    for (int m = 1; m < NR; ++m) if (cr[m] >= k_eff) { hi = tr[m]; cnt = cr[m]; }
#pragma unroll
                    for (int m = 1; m < NR; ++m) if (cr[m] < k_eff && tr[m] > lo) lo = tr[m];   // largest radius known to hold < k
                    int guardc = 0;
                    while (cnt > KNN_LIST && guardc < 48) {   // bisection: population(hi) >= k > population(lo)
                        const float mid = 0.5f * (lo + hi);
                        if (!(mid > lo && mid < hi)) break;
                        const int cm = population(mid);
                        if (cm >= k_eff) { hi = mid; cnt = cm; } else lo = mid;
                        ++guardc;
                    }
                    if (cnt > KNN_LIST) continue;             // > 64 candidates tie at the k-th distance: slow path
                    ln = collect(hi);
                }
                if (ln <= 32) bitonic32(k0); else bitonic64(k0, k1);
                {   // next query of the cell: a radius^2 a little above this query's k-th distance usually holds k..32 points,
                    // so it is accepted without counting and sorted with the one-register network
                    const int kl = (k_eff - 1) & 31;
                    const unsigned long long kth = __shfl_sync(0xffffffffu, (k_eff - 1) < 32 ? k0 : k1, kl);
                    guess = __uint_as_float((unsigned)(kth >> 32)) * OSEP_KNN_GUESS;
                }
                const int qidx = __float_as_int(q.w);
                if (lane < k_eff) knn_idx[(size_t)qidx * k_eff + lane] = (int)(unsigned)(k0 & 0xffffffffull);
                if (lane + 32 < k_eff) knn_idx[(size_t)qidx * k_eff + lane + 32] = (int)(unsigned)(k1 & 0xffffffffull);
                todo &= ~(1u << qi);
            }
        }
        __syncwarp();
        // overflow work items are single queries (unit * 32 + query): k_knn_big gives each its own warp, so a few
        // oversized blocks do not serialise 8-12 long queries on one warp at the end of the stage
        if (overflow) { if (lane < nq && ((todo >> lane) & 1u)) over_list[atomicAdd(&st->n_over, 1)] = unit * 32 + lane; }
        else if (lane < nq && ((todo >> lane) & 1u)) slow_list[atomicAdd(&st->n_slow, 1)] = cstart + q0 + lane;
        __syncwarp();
    }
}

// R5 overflow path: queries whose candidate block does not fit the per-warp stage of k_knn_cells (dense spots: more
// than KNN_STAGE points in 27 cells).  One WARP per QUERY: the (d2, index) pairs of all candidates
// are computed once into a large shared-memory list (KNN_BIG entries), every selection pass then runs from shared
// memory with full lane utilisation.  Same certification and the same canonical result as the fast path; what does
// not fit or cannot be certified within R <= 2 goes to k_knn_exact.
#ifndef OSEP_KNN_BIG
#define OSEP_KNN_BIG 1280
#endif
#ifndef OSEP_BIG_WARPS
#define OSEP_BIG_WARPS 4
#endif
constexpr int KNN_BIG = OSEP_KNN_BIG;        // candidates per query (one warp); more go to k_knn_exact
constexpr int BIG_WARPS = OSEP_BIG_WARPS;    // queries (warps) per CTA
// mode 0: work items are (unit, query) codes of k_knn_cells' overflow list.  mode 1: positions in Pc (k_knn_scan's fall-back
// chain).  mode 2: positions in Pc, and the R = 1 candidate cells come from the neighbour table row of the query's cell
// instead of 27 hash probes; a query that needs more (5x5x5 block, or more than CAPQ candidates) goes on `next_list`.
template <int K, int CAPQ, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_knn_big(const float4* __restrict__ Pc, FrameState* st,
                                                const unsigned long long* __restrict__ gtab, const int* __restrict__ g_cnt,
                                                const int* __restrict__ g_start, unsigned mask, const int2* __restrict__ units,
                                                int* __restrict__ knn_idx, int k_req, int* __restrict__ slow_list,
                                                const int* __restrict__ over_list, const int* __restrict__ n_work_ptr, int mode,
                                                const int* __restrict__ pc_cell, const int2* __restrict__ nbr,
                                                int* __restrict__ next_list, int* __restrict__ n_next, int r_first) {
    __shared__ int2 tab_[WARPS][KNN_TAB];
    __shared__ int pre_[WARPS][KNN_TAB + 1];
    __shared__ float sd_[WARPS][CAPQ];
    __shared__ int si_[WARPS][CAPQ];
    __shared__ unsigned long long keys_[WARPS][KNN_LIST];
    const int wib = threadIdx.x >> 5;
    int2* tab = tab_[wib]; int* pre = pre_[wib]; float* sd = sd_[wib]; int* si = si_[wib]; unsigned long long* keys = keys_[wib];
    if (st->status != 0) return;
    const int n = st->n_filt;
    const int k_eff = min(k_req, n);
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int dimx = st->g_dim[0], dimy = st->g_dim[1], dimz = st->g_dim[2];
    const float h = st->g_h;
    const float diag2 = ((float)dimx * h) * ((float)dimx * h) + ((float)dimy * h) * ((float)dimy * h) + ((float)dimz * h) * ((float)dimz * h);
    const int n_work = *n_work_ptr;
    const unsigned long long KINF = ~0ull;
    for (int w = blockIdx.x * WARPS + wib; w < n_work; w += gridDim.x * WARPS) {
        const int code = over_list[w];
        int qpos = code & 0x7fffffff;            // bit 31 (modes 1, 2): the 27-cell block overflowed the previous kernel's stage -- it has not been shown insufficient
        if (mode == 0) { const int2 un = units[code >> 5]; qpos = g_start[un.x] + (un.y & 0xffffff) + (code & 31); }
        const int r0 = (mode != 0 && r_first == 2 && code >= 0) ? 2 : 1;
        bool stage_overflow = false;
        const float4 q = Pc[qpos];
        int cx, cy, cz; cell_of(st, q.x, q.y, q.z, cx, cy, cz);
        bool done = false;
        for (int R = r0; R <= (mode == 2 ? 1 : 2) && !done; ++R) {      // r0 = 2: the 27-cell block is already known not to certify the query
            const int x0 = max(cx - R, 0), x1 = min(cx + R, dimx - 1);
            const int y0 = max(cy - R, 0), y1 = min(cy + R, dimy - 1);
            const int z0 = max(cz - R, 0), z1 = min(cz + R, dimz - 1);
            const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1;
            const int ncell = nx * ny * nz;
            const bool whole = (x0 == 0 && y0 == 0 && z0 == 0 && x1 == dimx - 1 && y1 == dimy - 1 && z1 == dimz - 1);
            int nt = 0, total = 0;
            __syncwarp();
            for (int c0 = 0; c0 < ncell; c0 += 32) {
                const int c = c0 + lane;
                int2 e = make_int2(0, 0);
                if (mode == 2) { if (c < 27) e = __ldg(nbr + (size_t)pc_cell[qpos] * 27 + c); }      // packed row: occupied cells first, in any order
                else if (c < ncell) {
                    const int x = x0 + c % nx, y = y0 + (c / nx) % ny, z = z0 + c / (nx * ny);
                    const unsigned long long key1 = cell_key1(x, y, z);
                    unsigned s = cell_slot(x, y, z, mask);
                    unsigned long long cur;
                    while ((cur = gtab[s]) != key1 && cur != 0ull) s = (s + 1) & mask;
                    if (cur != 0ull) e = make_int2(g_start[s], g_cnt[s]);
                }
                const unsigned m = __ballot_sync(0xffffffffu, e.y > 0);
                int incl = e.y;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                if (e.y > 0) { const int slot = nt + __popc(m & lt); tab[slot] = e; pre[slot] = total + incl - e.y; }
                nt += __popc(m);
                total += __shfl_sync(0xffffffffu, incl, 31);
            }
            __syncwarp();
            if (total > CAPQ) { stage_overflow = true; break; }      // next kernel of the chain
            // every candidate's (d2, index) once, 32 candidates per round, all rounds' loads independent
#pragma unroll 4
            for (int t0 = 0; t0 < total; t0 += 32) {
                const int t = t0 + lane;
                if (t < total) {
                    const int ci = flat_cell(pre, nt, t);
                    const float4 c = Pc[tab[ci].x + (t - pre[ci])];
                    sd[t] = dist2(q.x, q.y, q.z, c.x, c.y, c.z); si[t] = __float_as_int(c.w);
                }
            }
            __syncwarp();
            float b2;
            if (whole) b2 = 4.0f * diag2 + 1.0f;
            else {
                const float bound = block_bound(st, q.x - st->bmin[0], q.y - st->bmin[1], q.z - st->bmin[2], x0, x1, y0, y1, z0, z1);
                b2 = bound > 0.0f ? bound * bound : 0.0f;
            }
            auto population = [&](float thr) -> int {
                int cm = 0;
                for (int t = lane; t < total; t += 32) cm += (sd[t] < thr) ? 1 : 0;
                return __reduce_add_sync(0xffffffffu, cm);
            };
            constexpr int NR = 8;
            int cr[NR]; float tr[NR];
#pragma unroll
            for (int m = 0; m < NR; ++m) { cr[m] = 0; tr[m] = b2 * (1.0f / (float)(1 << m)); }
            for (int t = lane; t < total; t += 32) {
                const float d2 = sd[t];
#pragma unroll
                for (int m = 0; m < NR; ++m) cr[m] += (d2 < tr[m]) ? 1 : 0;
            }
#pragma unroll
            for (int m = 0; m < NR; ++m) cr[m] = __reduce_add_sync(0xffffffffu, cr[m]);
            if (cr[0] < k_eff) continue;                      // not certifiable at this R
            float lo = 0.0f, hi = tr[0]; int cnt = cr[0];
#pragma unroll
            for (int m = 1; m < NR; ++m) if (cr[m] >= k_eff) { hi = tr[m]; cnt = cr[m]; }
#pragma unroll
            for (int m = 1; m < NR; ++m) if (cr[m] < k_eff && tr[m] > lo) lo = tr[m];
            int guardc = 0;
            while (cnt > KNN_LIST && guardc < 48) {           // bisection: population(hi) >= k > population(lo)
                const float mid = 0.5f * (lo + hi);
                if (!(mid > lo && mid < hi)) break;
                const int cm = population(mid);
                if (cm >= k_eff) { hi = mid; cnt = cm; } else lo = mid;
                ++guardc;
            }
            if (cnt > KNN_LIST) break;                        // > 64 candidates tie at the k-th distance: slow path
            int ln = 0;
            for (int t0 = 0; t0 < total; t0 += 32) {
                const int t = t0 + lane;
                const float d2 = t < total ? sd[t] : 3.0e38f;
                const bool in = d2 < hi;
                const unsigned m = __ballot_sync(0xffffffffu, in);
                const int pos = ln + __popc(m & lt);
                if (in && pos < KNN_LIST) keys[pos] = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)si[t];
                ln += __popc(m);
            }
            __syncwarp();
            unsigned long long k0 = (lane < ln) ? keys[lane] : KINF;
            unsigned long long k1 = (lane + 32 < ln) ? keys[lane + 32] : KINF;
            __syncwarp();
            if (ln <= 32) bitonic32(k0); else bitonic64(k0, k1);
            const int qidx = __float_as_int(q.w);
            if (lane < k_eff) knn_idx[(size_t)qidx * k_eff + lane] = (int)(unsigned)(k0 & 0xffffffffull);
            if (lane + 32 < k_eff) knn_idx[(size_t)qidx * k_eff + lane + 32] = (int)(unsigned)(k1 & 0xffffffffull);
            done = true;
        }
        if (!done && lane == 0) { if (mode == 2) next_list[atomicAdd(n_next, 1)] = qpos | (stage_overflow ? (int)0x80000000 : 0); else slow_list[atomicAdd(&st->n_slow, 1)] = qpos; }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// R5, K <= 24: one THREAD per query.  k_knn_cells above spends most of its ~720 warp instructions per query on warp-wide
// selection (ballot compaction, 64-bit shuffle sorts, one query at a time); here 32 queries run their selection in
// lockstep, each in its own registers:
//   1. the (start, count) of the (2R+1)^3 cells around the query's cell -> registers (R = 1) / local memory (R = 2);
//   2. ONE scan of the candidates with a radius^2 `tau` chosen from the block's point density (surface model:
//      N(r) ~ n_block * pi r^2 / (block side)^2), capped by the certification bound of the block; candidates inside tau
//      are appended to the thread's column of a shared-memory list (SCAN_CAP entries);
//   3. SCAN_K <= count <= SCAN_CAP (and tau <= bound^2) certifies the list: it holds the k nearest by (d2, index).  It is
//      sorted as 32-bit keys (d2 bits with the low 6 bits replaced by the list slot: one IMNMX pair per compare-exchange,
//      a 32-key bitonic network in registers + a bubble pass for slots 32..47); if two of the first k+1 keys agree in
//      the 26 kept bits the thread redoes its selection exactly by (d2, index).
// A miss (count outside the window) does not loop in place -- 31 lanes would idle: the query goes on a RETRY list with a
// rescaled tau and a second launch of the same kernel runs all retries together (up to 3 attempts each); queries whose
// block cannot certify k points go on the GROW list (R = 2 launch), what is left goes to k_knn_exact.
#ifndef OSEP_SCAN_THREADS
#define OSEP_SCAN_THREADS 128
#endif
constexpr int SCAN_THREADS = OSEP_SCAN_THREADS;
constexpr int SCAN_CAP = 40;

__device__ __forceinline__ void cex(unsigned& a, unsigned& b) { const unsigned lo = min(a, b), hi = max(a, b); a = lo; b = hi; }
__device__ __forceinline__ void sts_u32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

template <int K>
__global__ void __launch_bounds__(SCAN_THREADS) k_knn_scan(const float4* __restrict__ Pc, FrameState* st, const int* __restrict__ pc_cell,
                                                           const int2* __restrict__ nbr, const float* __restrict__ nbr_tau, int cell_cap,
                                                           int* __restrict__ knn_idx, int k_req, int* __restrict__ grow_list) {
    constexpr int R = 1, NC = 27;
    frame_stamp(st, TS_KNN);
    __shared__ unsigned lst[2 * (SCAN_CAP + 1)][SCAN_THREADS];    // rows [0, CAP]: d2 bits (row CAP = overflow sink); rows [CAP + 1, 2 CAP + 1]: point index
    unsigned (*ld)[SCAN_THREADS] = lst;
    unsigned (*li)[SCAN_THREADS] = lst + (SCAN_CAP + 1);
    if (st->status != 0) return;
    const int n = st->n_filt;
    const int k_eff = min(k_req, n);
    const int tid = threadIdx.x;
    const int dimx = st->g_dim[0], dimy = st->g_dim[1], dimz = st->g_dim[2];
    const float h = st->g_h;
    const float diag2 = ((float)dimx * h) * ((float)dimx * h) + ((float)dimy * h) * ((float)dimy * h) + ((float)dimz * h) * ((float)dimz * h);
    const bool table_ok = st->n_cells <= cell_cap;       // more occupied cells than the neighbour table holds: every query takes the warp-per-query path
    const unsigned ld0 = (unsigned)__cvta_generic_to_shared(&lst[0][tid]);
    constexpr unsigned ROW = SCAN_THREADS * 4u, LI_OFF = (SCAN_CAP + 1) * ROW;
    const unsigned ld_end = ld0 + SCAN_CAP * ROW;
    for (int wbase = (blockIdx.x * SCAN_THREADS + (tid & ~31)); wbase < n; wbase += gridDim.x * SCAN_THREADS) {
        const int t = wbase + (tid & 31);
        const bool valid = t < n;
        if (!table_ok) { if (valid) grow_list[atomicAdd(&st->n_grow, 1)] = t; continue; }
        const float4 q = valid ? Pc[t] : make_float4(0.f, 0.f, 0.f, 0.f);
        int cx, cy, cz; cell_of(st, q.x, q.y, q.z, cx, cy, cz);
        const int x0 = max(cx - R, 0), x1 = min(cx + R, dimx - 1);
        const int y0 = max(cy - R, 0), y1 = min(cy + R, dimy - 1);
        const int z0 = max(cz - R, 0), z1 = min(cz + R, dimz - 1);
        const bool whole = (x0 == 0 && y0 == 0 && z0 == 0 && x1 == dimx - 1 && y1 == dimy - 1 && z1 == dimz - 1);
        float b2;
        if (whole) b2 = 4.0f * diag2 + 1.0f;
        else {
            const float bound = block_bound(st, q.x - st->bmin[0], q.y - st->bmin[1], q.z - st->bmin[2], x0, x1, y0, y1, z0, z1);
            b2 = bound > 0.0f ? bound * bound : 0.0f;
        }
        // 1. the query's row of the neighbour table and the cell's pilot radius
        const int ci = valid ? pc_cell[t] : 0;
        const int2* row = nbr + (size_t)ci * NC;
        const float tau = valid ? fminf(nbr_tau[ci], b2) : 0.0f;
        // 2. one scan: candidates inside tau go to the thread's list column
        unsigned pl = ld0;
        auto visit = [&](const float4& p, bool on) {
            const float d2 = dist2(q.x, q.y, q.z, p.x, p.y, p.z);
            if (on && d2 < tau) { sts_u32(pl, __float_as_uint(d2)); sts_u32(pl + LI_OFF, __float_as_uint(p.w)); pl = min(pl + ROW, ld_end); }
        };
        int2 cur = valid ? __ldg(row) : make_int2(0, 0);
#pragma unroll 1
        for (int c = 0; c < NC && cur.y > 0; ++c) {
            const float4* base = Pc + cur.x; const int m = cur.y;
            cur = (c + 1 < NC) ? __ldg(row + c + 1) : make_int2(0, 0);      // next cell's entry is in flight during this cell's scan
            int e = 0;
#pragma unroll 1
            for (; e + 4 <= m; e += 4) {                  // 4 candidates per round: their loads are in flight together
                const float4 p0 = __ldg(base + e), p1 = __ldg(base + e + 1), p2 = __ldg(base + e + 2), p3 = __ldg(base + e + 3);
                visit(p0, true); visit(p1, true); visit(p2, true); visit(p3, true);
            }
            if (e < m) {
                const float4 p0 = __ldg(base + e), p1 = __ldg(base + min(e + 1, m - 1)), p2 = __ldg(base + min(e + 2, m - 1));
                visit(p0, true); visit(p1, e + 1 < m); visit(p2, e + 2 < m);
            }
        }
        const int cnt = (int)((pl - ld0) / ROW);          // SCAN_CAP = "SCAN_CAP or more": one slot of the window is given up for the test
        const bool sortme = valid && cnt >= k_eff && cnt < SCAN_CAP;
        // the count missed the window or the block cannot certify k points: warp-per-query kernel
        if (valid && !sortme) grow_list[atomicAdd(&st->n_grow, 1)] = t;
        // 3. sort: 32-bit keys = d2 bits (>= 0: ordered like the floats) with the low 6 bits replaced by the list slot
        unsigned key[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) key[i] = (sortme && i < cnt) ? ((ld[i][tid] & ~63u) | (unsigned)i) : 0xffffffffu;
#pragma unroll
        for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (int j = kk >> 1; j > 0; j >>= 1) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int l = i ^ j;
                    if (l > i) { if ((i & kk) == 0) cex(key[i], key[l]); else cex(key[l], key[i]); }
                }
            }
        }
        {   // slots 32.. bubble through the first K + 1 positions (all that the output and the tie test read)
            const int extra = sortme ? cnt - 32 : 0;
            const int wmax = __reduce_max_sync(0xffffffffu, extra);
            for (int i = 0; i < wmax; ++i) {
                unsigned x = (i < extra) ? ((ld[32 + i][tid] & ~63u) | (unsigned)(32 + i)) : 0xffffffffu;
#pragma unroll
                for (int pz = 0; pz <= K; ++pz) cex(key[pz], x);
            }
        }
        bool tie = false;
#pragma unroll
        for (int pz = 0; pz < K; ++pz) if (pz < k_eff && key[pz + 1] != 0xffffffffu && ((key[pz] ^ key[pz + 1]) < 64u)) tie = true;
        if (sortme) {
            int* out = knn_idx + (size_t)__float_as_int(q.w) * k_eff;
            if (!tie) {
#pragma unroll
                for (int pz = 0; pz < K; ++pz) if (pz < k_eff) out[pz] = (int)li[key[pz] & 63u][tid];
            } else {
                // exact selection by (d2, index) over the thread's list (rare: two of the first k+1 candidates within 2^-17 relative)
                for (int r = 0; r < k_eff; ++r) {
                    int bi = -1; float bd = 0.0f; int bid = 0;
                    for (int i = 0; i < cnt; ++i) {
                        const unsigned u = ld[i][tid];
                        if (u == 0xffffffffu) continue;
                        const float dd = __uint_as_float(u); const int id = (int)li[i][tid];
                        if (bi < 0 || nb_less(dd, id, bd, bid)) { bi = i; bd = dd; bid = id; }
                    }
                    out[r] = bid; ld[bi][tid] = 0xffffffffu;
                }
            }
        }
    }
}

// normals: one thread per point, neighbours from knn_idx (k_eff per point, canonical order)
__global__ void __launch_bounds__(128) k_normals(const float4* __restrict__ P, const FrameState* __restrict__ st,
                                                 const int* __restrict__ knn_idx, int k_req, float4* __restrict__ Nrm) {
    frame_stamp(st, TS_NORMALS);
    if (st->status != 0) return;
    const int n = st->n_filt;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k_eff = min(k_req, n);
    const int* row = knn_idx + (size_t)i * k_eff;
    Nrm[i] = normal_from_neighbours(P, P[i], [&](int u) { return row[u]; }, k_eff);
}

// ------------------------------------------------------------------------------------------
// R9 final pass: distinct voxel keys -> sorted -> dense voxel ids -> points grouped by voxel (atomic
// cursors) -> per-voxel index sort + sums in ascending point index (k_vox_reduce).
__global__ void k_vox_collect(FrameState* st, const unsigned long long* __restrict__ vtab, int T, unsigned* __restrict__ vkeys, int Mcap) {
    frame_stamp(st, TS_VOXFINAL);
    if (st->status != 0) return;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= T) return;
    const unsigned long long e = vtab[s];
    if ((unsigned)(e >> 32) == st->epoch) {
        const int pos = atomicAdd(&st->n_keys, 1);
        if (pos < Mcap) vkeys[pos] = (unsigned)e;
    }
}

// single block: finalise M, capacity check, bitonic sort of the keys in shared memory
__global__ void k_vox_sortkeys(FrameState* st, unsigned* vkeys, int Mcap, int* vox_cnt) {
    extern __shared__ unsigned sk[];
    if (st->status != 0) return;
    __shared__ int sM;
    if (threadIdx.x == 0) {
        int M = st->n_keys;
        if (st->pass_overflow) {
            // PCL's guard ("Leaf size is too small for the input dataset. Integer indices would overflow."): VoxelGrid returns the
            // input cloud unchanged.  Every filtered point becomes its own voxel as long as the downsampled arena holds them
            // (small clouds whose leaf the control law shrank; a 100k-point frame would make the reference run ROSA on 100k points).
            if (st->n_filt <= Mcap) { M = st->n_filt; st->vox_identity = 1; }
            else { st->status = OSEP_ERR_CAPACITY; M = 0; }
        }
        else if (M > Mcap) { st->status = OSEP_ERR_CAPACITY; M = 0; }
        st->M = M;
        if (M == 0 && st->status == 0) st->status = OSEP_STAGE_3;   // stage gate: rosa_init always succeeds, similarity_neighbor_extraction fails on an empty cloud (rosa.cpp:205)
        sM = M;
    }
    __syncthreads();
    const int M = sM;
    int P2 = 1; while (P2 < M) P2 <<= 1;
    for (int i = threadIdx.x; i < P2; i += blockDim.x) sk[i] = (i < M) ? vkeys[i] : 0xffffffffu;
    for (int i = threadIdx.x; i <= M; i += blockDim.x) vox_cnt[i] = 0;
    __syncthreads();
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned a = sk[i], b = sk[ixj];
                    const bool up = ((i & k) == 0);
                    if ((a > b) == up) { sk[i] = b; sk[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < M; i += blockDim.x) vkeys[i] = sk[i];
}

__global__ void k_vox_assign(const float4* __restrict__ P, const FrameState* __restrict__ st, const unsigned* __restrict__ vkeys,
                             unsigned* __restrict__ rk, int* __restrict__ rv, int* vox_cnt) {
    if (st->status != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st->n_filt) return;
    if (st->vox_identity) { rk[i] = (unsigned)i; rv[i] = i; atomicAdd(&vox_cnt[i], 1); return; }      // output = input
    const unsigned key = voxel_key(P[i], st->inv_leaf, st->minb[0], st->minb[1], st->minb[2], st->divb[0], st->divb[1]);
    int lo = 0, hi = st->M - 1;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (vkeys[mid] < key) lo = mid + 1; else hi = mid; }
    rk[i] = (unsigned)lo; rv[i] = i;
    atomicAdd(&vox_cnt[lo], 1);
}

__global__ void k_scan_counts(const FrameState* __restrict__ st, int* data, int which /*0: M+1 entries*/) {
    if (st->status != 0) return;
    block_excl_scan(data, st->M + 1);
}

// group the point indices by voxel: slot = vox_off[v] + atomic cursor (order inside a voxel is fixed later)
__global__ void k_vox_scatter(const FrameState* __restrict__ st, const unsigned* __restrict__ vid, const int* __restrict__ vox_off,
                              int* fill, int* __restrict__ grouped) {
    if (st->status != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st->n_filt) return;
    const int v = (int)vid[i];
    grouped[vox_off[v] + atomicAdd(&fill[v], 1)] = i;
}

// ---- stable partition of the points by voxel (frames whose block-count matrix fits: blk_hist != nullptr) -------------------------
// The atomic cursors of k_vox_scatter leave a voxel's points in arbitrary order, so k_vox_reduce had to sort every voxel's
// index list first (6.5 M warp instructions per 100k-point frame, a third of the final pass).  Here the order is right from
// the start: a block of 256 consecutive points ranks each point among the block's earlier points of the same voxel (warp by
// warp: __match_any groups + a shared count per voxel), writes its per-voxel counts as one row of a [block][voxel] matrix,
// a thread per voxel turns the matrix's columns into exclusive prefixes over the blocks (coalesced across voxels), and the
// scatter adds voxel offset + block prefix + rank: ascending point index inside every voxel, no atomics on global memory.
__global__ void __launch_bounds__(NT) k_vox_assign_ranked(const float4* __restrict__ P, const FrameState* __restrict__ st, const unsigned* __restrict__ vkeys,
                                                          unsigned* __restrict__ vid_out, int* __restrict__ rank_out, int* __restrict__ blk_hist, int Mcap) {
    extern __shared__ int vhist[];
    if (st->status != 0) return;
    const int n = st->n_filt, M = st->M;
    if ((int)(blockIdx.x * blockDim.x) >= n) return;
    for (int v = threadIdx.x; v < M; v += blockDim.x) vhist[v] = 0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    int vid = 0;
    if (valid) {
        if (st->vox_identity) vid = i;                                     // output = input: every filtered point is its own voxel
        else {
            const unsigned key = voxel_key(P[i], st->inv_leaf, st->minb[0], st->minb[1], st->minb[2], st->divb[0], st->divb[1]);
            int lo = 0, hi = M - 1;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (vkeys[mid] < key) lo = mid + 1; else hi = mid; }
            vid = lo;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int rank = 0;
    for (int w = 0; w < NT / 32; ++w) {                                    // warps in index order
        if (wib == w) {
            const unsigned grp = __match_any_sync(0xffffffffu, valid ? vid : (M + lane));
            const int base = valid ? vhist[vid] : 0;
            __syncwarp();
            if (valid && lane == __ffs(grp) - 1) vhist[vid] = base + __popc(grp);
            rank = base + __popc(grp & ((1u << lane) - 1u));
        }
        __syncthreads();
    }
    if (valid) { vid_out[i] = (unsigned)vid; rank_out[i] = rank; }
    int* row = blk_hist + (size_t)blockIdx.x * Mcap;
    for (int v = threadIdx.x; v < M; v += blockDim.x) row[v] = vhist[v];
}
// Column-wise exclusive prefix of the [block][voxel] matrix over the blocks (in place), column totals -> vox_cnt.  A CTA takes 32
// voxels (lane = voxel: every access is coalesced across voxels), warp c the c-th chunk of 32 blocks: its 32 loads are
// independent (one latency), the prefix inside the chunk stays in registers, chunk totals meet in shared memory.
// (A thread per voxel walking its column -- load, add, store, next -- took 100 us: 391 dependent L2 round trips.)
__global__ void __launch_bounds__(1024) k_vox_blockscan(const FrameState* __restrict__ st, int* __restrict__ blk_hist, int Mcap, int* __restrict__ vox_cnt) {
    __shared__ int tot[32][33];
    __shared__ int carry[32];
    if (st->status != 0) return;
    const int M = st->M;
    if ((int)blockIdx.x * 32 > M) return;
    const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
    const int v = blockIdx.x * 32 + lane;
    const int nblk = (st->n_filt + NT - 1) / NT;
    if (c == 0) carry[lane] = 0;
    __syncthreads();
    int* col = blk_hist + v;
    for (int g0 = 0; g0 < nblk; g0 += 32 * 32) {                          // 32 chunks of 32 blocks per round
        const int b0 = g0 + c * 32;
        int val[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) { const int b = b0 + k; val[k] = (v < M && b < nblk) ? col[(size_t)b * Mcap] : 0; }
        int run = 0;
#pragma unroll
        for (int k = 0; k < 32; ++k) { const int t = val[k]; val[k] = run; run += t; }
        tot[c][lane] = run;
        __syncthreads();
        int off = carry[lane];
        for (int cc = 0; cc < c; ++cc) off += tot[cc][lane];
#pragma unroll
        for (int k = 0; k < 32; ++k) { const int b = b0 + k; if (v < M && b < nblk) col[(size_t)b * Mcap] = off + val[k]; }
        __syncthreads();
        if (c == 31) carry[lane] = off + run;                              // (warp 31's offset covers chunks 0..30)
        __syncthreads();
    }
    if (c == 0) { if (v < M) vox_cnt[v] = carry[lane]; else if (v == M) vox_cnt[M] = 0; }
}
__global__ void k_vox_scatter_ranked(const FrameState* __restrict__ st, const unsigned* __restrict__ vid, const int* __restrict__ rank, const int* __restrict__ vox_off,
                                     const int* __restrict__ blk_hist, int Mcap, int* __restrict__ grouped) {
    if (st->status != 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st->n_filt) return;
    const int v = (int)vid[i];
    grouped[vox_off[v] + blk_hist[(size_t)blockIdx.x * Mcap + v] + rank[i]] = i;
}

// per-voxel centroid (pcl::CentroidPoint<PointNormal>), then R10 + R11 fused.
// One CTA per voxel: (1) the voxel's point indices are sorted ascending in shared memory (bitonic; lists
// longer than VOX_SORT are rank-sorted through global memory) -- this restores the canonical order C3 that a
// stable sort of (key, index) would give; (2) all threads stage points + normals in that order (many gathers
// in flight); (3) ONE thread adds them, so the float sums stay strictly sequential in ascending point index.
constexpr int VOX_SORT = 2048;
constexpr int VOX_CHUNK = 256;
// PHASE 0 (needs only the filtered cloud: runs while the kNN / normals branch is still busy): sort the voxel's point indices
// (left in `grouped` for phase 1), sum the POSITIONS -> pts, pset.  PHASE 1 (after the normals): sum the NORMALS in the stored
// order -> nrms, vset (R10 + R11), plus what drosa hoists per point (nraw, vnd) and its iteration-0 orientations (vset_it).
template <int PHASE, bool PRESORTED = false>
__global__ void __launch_bounds__(256)
k_vox_reduce(FrameState* st, int* grouped, int* __restrict__ scratch, const int* __restrict__ vox_off,
             const float4* __restrict__ P, const float4* __restrict__ Nrm,
             float4* __restrict__ pts, float4* __restrict__ nrms, float4* __restrict__ pset, float4* __restrict__ vset,
             float4* __restrict__ nraw, float4* __restrict__ vnd, float4* __restrict__ vset_it) {
    __shared__ int sidx[VOX_SORT];
    __shared__ float4 sp[VOX_CHUNK];
    frame_stamp(st, PHASE == 0 ? TS_REDUCE0 : TS_REDUCE1);
    if (st->status != 0) return;
    const int M = st->M;
    for (int v = blockIdx.x; v < M; v += gridDim.x) {
        const int b = vox_off[v], e = vox_off[v + 1];
        const int cnt = e - b;
        const int* order = grouped + b;
        __syncthreads();
        if (PHASE == 0 && !PRESORTED) {
            if (cnt <= VOX_SORT) {
                int P2 = 1; while (P2 < cnt) P2 <<= 1;
                for (int t = threadIdx.x; t < P2; t += blockDim.x) sidx[t] = (t < cnt) ? grouped[b + t] : 0x7fffffff;
                __syncthreads();
                for (int k = 2; k <= P2; k <<= 1) {
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        for (int i = threadIdx.x; i < P2; i += blockDim.x) {
                            const int ixj = i ^ j;
                            if (ixj > i) {
                                const int x = sidx[i], y = sidx[ixj];
                                const bool up = ((i & k) == 0);
                                if ((x > y) == up) { sidx[i] = y; sidx[ixj] = x; }
                            }
                        }
                        __syncthreads();
                    }
                }
                for (int t = threadIdx.x; t < cnt; t += blockDim.x) grouped[b + t] = sidx[t];      // phase 1 reads the order from here
                order = sidx;
            } else {
                // voxels with more than VOX_SORT points (large maps): bitonic network with all comparators ascending
                // ("flip" first stage), which needs no power-of-two padding -- missing partners act as +inf -- in place
                // in global memory (the segment is private to this CTA)
                int* seg = grouped + b;
                for (int k = 2; (k >> 1) < cnt; k <<= 1) {
                    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
                        const int ixj = i ^ (k - 1);
                        if (ixj > i && ixj < cnt) { const int x = seg[i], y = seg[ixj]; if (x > y) { seg[i] = y; seg[ixj] = x; } }
                    }
                    __syncthreads();
                    for (int j = k >> 2; j > 0; j >>= 1) {
                        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
                            const int ixj = i ^ j;
                            if (ixj > i && ixj < cnt) { const int x = seg[i], y = seg[ixj]; if (x > y) { seg[i] = y; seg[ixj] = x; } }
                        }
                        __syncthreads();
                    }
                }
                order = seg;
            }
        }
        const float4* __restrict__ src = PHASE == 0 ? P : Nrm;
        float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int c0 = 0; c0 < cnt; c0 += VOX_CHUNK) {
            const int cn = min(VOX_CHUNK, cnt - c0);
            __syncthreads();
            for (int t = threadIdx.x; t < cn; t += blockDim.x) sp[t] = src[order[c0 + t]];
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int t = 0; t < cn; ++t) { const float4 q = sp[t]; s0 += q.x; s1 += q.y; s2 += q.z; s3 += 0.0f; }
            }
        }
        if (threadIdx.x != 0) continue;
        if (PHASE == 0) {
            const float cntf = (float)cnt;
            const float px = s0 / cntf, py = s1 / cntf, pz = s2 / cntf;
            pts[v] = make_float4(px, py, pz, 1.0f);
            pset[v] = make_float4(px, py, pz, 1.0f);         // R11: pset = pts (rosa.cpp:183-199)
            continue;
        }
        const float n0 = s0, n1 = s1, n2 = s2, n3 = s3;
        // AccumulatorNormal::get: Vector4f::normalized(); squaredNorm is packet-reduced (a0+a2)+(a1+a3)
        const float zz = (n0 * n0 + n2 * n2) + (n1 * n1 + n3 * n3);
        f3 nv{n0, n1, n2};
        if (zz > 0.0f && !st->vox_identity) { const float r = sqrtf(zz); nv = {n0 / r, n1 / r, n2 / r}; }      // output = input keeps the normal as estimated
        // R10 (rosa.cpp:156-165)
        const float L = norm3(nv);
        if (L > 1e-6f) nv = div3(nv, L); else nv = {0.0f, 0.0f, 0.0f};
        nrms[v] = make_float4(nv.x, nv.y, nv.z, 0.0f);
        { float4 a4, b4; normal_prep_vals(nv, a4, b4); nraw[v] = a4; vnd[v] = b4; }
        // R11 (rosa.cpp:183-199, rosa.hpp:112-128): vset = column 0 of the orthonormal frame
        f3 vs;
        const float nn_ = norm3(nv);
        if (nn_ == 0.0f) vs = {1.0f, 0.0f, 0.0f};
        else {
            const f3 z = div3(nv, nn_);
            const bool use_x = (double)fabsf(z.x) < 0.9;
            const f3 a = use_x ? f3{1.0f, 0.0f, 0.0f} : f3{0.0f, 1.0f, 0.0f};
            const float az = dot3(a, z);
            vs = normalize3(sub3(a, mul3(z, az)));
        }
        vset[v] = make_float4(vs.x, vs.y, vs.z, 0.0f);
        vset_it[v] = make_float4(vs.x, vs.y, vs.z, 0.0f);   // iteration-0 input of the fused drosa kernel
    }
}

// ------------------------------------------------------------------------------------------
template <int K>
static void launch_knn(osep_rosa_impl* h, int n, int k_req, cudaStream_t s) {
    RosaDev& d = h->d;
    const unsigned mask = (unsigned)(d.T - 1);
    if constexpr (K <= 24) if (!h->knn_warp_path) {
        // thread-per-query scan: all queries (one attempt), the retries, then the warp-per-query kernel for what the 27-cell
        // block cannot certify, then the exact ring search
        k_grid_nbrs<<<h->sms * 8, 256, 0, s>>>(d.st, d.Pc, d.gtab, d.g_cnt, d.g_start, mask, d.cell_slots, d.cell_cap, d.nbr, d.nbr_tau, h->knn_target);
        k_knn_scan<K><<<div_up(d.Ncap, SCAN_THREADS), SCAN_THREADS, 0, s>>>(d.Pc, d.st, d.pc_cell, d.nbr, d.nbr_tau, d.cell_cap, d.knn_full, k_req, d.knn_grow);
        k_knn_big<K, 384, 8><<<h->sms * 5, 256, 0, s>>>(d.Pc, d.st, d.gtab, d.g_cnt, d.g_start, mask, d.units, d.knn_full, k_req, d.slow_list, d.knn_grow, &d.st->n_grow, 2,
                                                        d.pc_cell, d.nbr, d.over_list, &d.st->n_over, 1);
        k_knn_big<K, KNN_BIG, BIG_WARPS><<<h->sms * 4, BIG_WARPS * 32, 0, s>>>(d.Pc, d.st, d.gtab, d.g_cnt, d.g_start, mask, d.units, d.knn_full, k_req, d.slow_list, d.over_list, &d.st->n_over, 1,
                                                                              nullptr, nullptr, nullptr, nullptr, 2);
        k_knn_exact<K><<<h->sms * 2, 128, 0, s>>>(d.Pc, d.st, d.gtab, d.g_cnt, d.g_start, mask, d.knn_full, k_req, d.slow_list);
        k_normals<<<div_up(d.Ncap, 128), 128, 0, s>>>(d.P, d.st, d.knn_full, k_req, d.Nrm);
        return;
    }
    // 3 CTAs per SM, not the 4 that fit: the register file would be full and the short kernels of the main stream (leaf
    // passes, voxel grouping) could not co-run; the stage alone is 13 % slower, the frame 1 % faster
    // (large maps: the stage is long, nothing competes for long, so the 4th CTA pays)
    const int ctas_per_sm = (h->knn_ctas_per_sm > 0) ? h->knn_ctas_per_sm : (n > 400000 ? 4 : 3);
    k_knn_cells<K><<<h->sms * ctas_per_sm, KNN_WARPS * 32, 0, s>>>(d.Pc, d.st, d.gtab, d.g_cnt, d.g_start, mask, d.units,
                                                              d.knn_full, k_req, d.slow_list, d.over_list);
    k_knn_big<K, KNN_BIG, BIG_WARPS><<<h->sms * 4, BIG_WARPS * 32, 0, s>>>(d.Pc, d.st, d.gtab, d.g_cnt, d.g_start, mask, d.units, d.knn_full, k_req, d.slow_list, d.over_list, &d.st->n_over, 0,
                                                                          nullptr, nullptr, nullptr, nullptr, 1);
    k_knn_exact<K><<<h->sms * 2, 128, 0, s>>>(d.Pc, d.st, d.gtab, d.g_cnt, d.g_start, mask, d.knn_full, k_req, d.slow_list);
    k_normals<<<div_up(d.Ncap, 128), 128, 0, s>>>(d.P, d.st, d.knn_full, k_req, d.Nrm);
}

// ------------------------------------------------------------------------------------------
// tiled maps: stable box / range compaction (osep_points_select_box, osep_points_select_range) and the leaf-only pass (osep_rosa_leaf_device)
struct Box3 { float lo[3], hi[3]; };
__device__ __forceinline__ bool in_box(const float* p, const Box3& b) {
    return p[0] >= b.lo[0] && p[0] < b.hi[0] && p[1] >= b.lo[1] && p[1] < b.hi[1] && p[2] >= b.lo[2] && p[2] < b.hi[2];
}
__global__ void k_range_count(const float* __restrict__ in, int n, int stride, Box3 box, int* __restrict__ blk_cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool keep = i < n && in_box(in + (size_t)i * stride, box);
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) blk_cnt[blockIdx.x] = c;
}
__global__ void k_range_scan(int* blk_cnt, int nblk, int* total) {
    const int t = block_excl_scan(blk_cnt, nblk);
    if (threadIdx.x == 0) *total = t;
}
__global__ void k_range_scatter(const float* __restrict__ in, int n, int stride, Box3 box, const int* __restrict__ blk_off,
                                float* __restrict__ out, int out_stride, int out_cap) {
    __shared__ int warp_cnt[NT / 32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const float* p = in + (size_t)min(i, n - 1) * stride;
    const bool keep = i < n && in_box(p, box);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[wid] = __popc(m);
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; ++w) woff += warp_cnt[w];
    if (keep) {
        const int pos = blk_off[blockIdx.x] + woff + __popc(m & ((1u << lane) - 1u));
        if (pos < out_cap) {
            float* o = out + (size_t)pos * out_stride;
            o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
            if (out_stride > 3) o[3] = 1.0f;
        }
    }
}
int select_box_enqueue(const float* in, int n, int stride, const float* lo3, const float* hi3, float* out, int out_stride, int out_cap,
                       int* blk_cnt, int* total_dev, cudaStream_t s) {
    Box3 box;
    for (int d = 0; d < 3; ++d) { box.lo[d] = lo3[d]; box.hi[d] = hi3[d]; }
    const int nb = div_up(n, NT);
    k_range_count<<<nb, NT, 0, s>>>(in, n, stride, box, blk_cnt);
    k_range_scan<<<1, 1024, 0, s>>>(blk_cnt, nb, total_dev);
    k_range_scatter<<<nb, NT, 0, s>>>(in, n, stride, box, blk_cnt, out, out_stride, out_cap);
    return cudaGetLastError() == cudaSuccess ? OSEP_OK : OSEP_ERR_CUDA;
}
// R3/R4 + R7/R8 only: the frame stops after the leaf-size control law
void launch_leaf_only(osep_rosa_impl* h) {
    RosaDev& d = h->d;
    cudaStream_t s = h->stream;
    const osep_rosa_config& c = h->cfg;
    const int nb = div_up(d.Ncap, NT);
    const float r2 = c.pts_dist_lim * c.pts_dist_lim;
    k_filter_count<<<nb, NT, 0, s>>>(d.prm, r2, d.blk_cnt);
    k_filter_scan<<<1, 1024, 0, s>>>(d.blk_cnt, nb, d.st, c.min_points, d.prm);
    k_filter_scatter<<<nb, NT, 0, s>>>(d.prm, r2, d.blk_cnt, d.P, d.filt_idx, d.st, c.max_points, 9);
    for (int p = 0; p < LEAF_MAX_ITERS; ++p)
        k_vox_count<<<nb, NT, 0, s>>>(d.P, d.st, d.vtab, (unsigned)(d.T - 1), p + 1, c.max_points, 9);
    h->n_launches += 3 + LEAF_MAX_ITERS;
}

void launch_preprocess(osep_rosa_impl* h, int n) {
    RosaDev& d = h->d;
    cudaStream_t s = h->stream;
    const osep_rosa_config& c = h->cfg;
    // capacity-sized grids: every kernel reads the frame's sizes from FrameParams / FrameState and surplus blocks exit at
    // once, so the launch sequence (and its CUDA graph) is the same for every frame this handle can take
    const int nb = div_up(d.Ncap, NT);
    const float r2 = c.pts_dist_lim * c.pts_dist_lim;
    // target points per kNN grid cell: large (dense, heavy-tailed) maps overflow the 512-point stage too often at 9
    const int ppc = h->knn_ppc > 0 ? h->knn_ppc : (n > 400000 ? 6 : 9);
    cudaStream_t sk = s;
    if (h->overlap) {                                      // the kNN grid tables are cleared while the filter runs
        sk = h->stream2;
        cudaEventRecord(h->ev_fork[0], s); cudaStreamWaitEvent(sk, h->ev_fork[0], 0);
        cudaMemsetAsync(d.gtab, 0, sizeof(unsigned) * 4 * (size_t)d.T, sk);   // gtab (8 B) | g_cnt | g_fill are contiguous
    }
    // R3/R4
    k_filter_count<<<nb, NT, 0, s>>>(d.prm, r2, d.blk_cnt);
    k_filter_scan<<<1, 1024, 0, s>>>(d.blk_cnt, nb, d.st, c.min_points, d.prm);
    k_filter_scatter<<<nb, NT, 0, s>>>(d.prm, r2, d.blk_cnt, d.P, d.filt_idx, d.st, c.max_points, ppc);
    if (h->overlap) { cudaEventRecord(h->ev_fork[0], s); cudaStreamWaitEvent(sk, h->ev_fork[0], 0); }      // the kNN branch needs the filtered cloud and its bounding box only
    stage_mark(h, SM_FILTER);
    // R7/R8: device-resident leaf loop (count-only passes).  The kNN grid is sized from the density seen by
    // pass 0, so the whole normals branch (grid build + kNN + covariance/eigen) only depends on pass 0 and runs
    // on the side stream while the remaining passes iterate the control law.
    k_vox_count<<<nb, NT, 0, s>>>(d.P, d.st, d.vtab, (unsigned)(d.T - 1), 1, c.max_points, ppc);
    // (a persistent cooperative kernel for passes 1..7 -- points in registers, one 64-bit atomic per block and pass, the
    // next pass's parameters returned by the barrier poll -- was measured: 7.1 us per pass instead of 9.3, but its spinning
    // blocks slow the kNN branch on the other stream down by more than that; not kept)
    for (int p = 1; p < LEAF_MAX_ITERS; ++p)
        k_vox_count<<<nb, NT, 0, s>>>(d.P, d.st, d.vtab, (unsigned)(d.T - 1), p + 1, c.max_points, ppc);
    stage_mark(h, SM_LEAF);
    // R5: grid + kNN normals
    if (!h->overlap) cudaMemsetAsync(d.gtab, 0, sizeof(unsigned) * 4 * (size_t)d.T, sk);
    k_grid_insert<<<nb, NT, 0, sk>>>(d.P, d.st, d.gtab, d.g_cnt, d.p_slot, (unsigned)(d.T - 1), 0, ppc);
    k_grid_clear_if_redo<<<div_up(d.T, NT), NT, 0, sk>>>(d.st, d.gtab, d.g_cnt, d.T);
    k_grid_insert<<<nb, NT, 0, sk>>>(d.P, d.st, d.gtab, d.g_cnt, d.p_slot, (unsigned)(d.T - 1), 1, ppc);
    k_grid_alloc<<<div_up(d.T, NT), NT, 0, sk>>>(d.st, d.g_cnt, d.g_start, d.T, d.units, d.g_cell, d.cell_slots, d.cell_cap);
    k_grid_scatter<<<nb, NT, 0, sk>>>(d.P, d.st, d.p_slot, d.g_start, d.g_fill, d.Pc, d.g_cell, d.pc_cell);
    stage_mark(h, SM_GRID);
    const int k = c.ne_knn;
    if (k == 20) launch_knn<20>(h, n, k, sk);
    else if (k <= 10) launch_knn<10>(h, n, k, sk);
    else if (k <= 16) launch_knn<16>(h, n, k, sk);
    else if (k <= 32) launch_knn<32>(h, n, k, sk);
    else launch_knn<64>(h, n, k, sk);
    if (h->overlap) cudaEventRecord(h->ev_join[0], sk);
    stage_mark(h, SM_KNN);
    // R9 final pass
    k_vox_collect<<<div_up(d.T, NT), NT, 0, s>>>(d.st, d.vtab, d.T, d.vkeys, d.Mcap);
    int P2 = 1; while (P2 < d.Mcap) P2 <<= 1;
    k_vox_sortkeys<<<1, 1024, sizeof(unsigned) * P2, s>>>(d.st, d.vkeys, d.Mcap, d.vox_cnt);
    if (d.blk_hist) {
        // stable partition by voxel: the reduce kernel finds every voxel's points in ascending index, nothing to sort
        k_vox_assign_ranked<<<nb, NT, sizeof(int) * d.Mcap, s>>>(d.P, d.st, d.vkeys, d.rk0, (int*)d.rk1, d.blk_hist, d.Mcap);
        k_vox_blockscan<<<div_up(d.Mcap + 1, 32), 1024, 0, s>>>(d.st, d.blk_hist, d.Mcap, d.vox_cnt);
        k_scan_counts<<<1, 1024, 0, s>>>(d.st, d.vox_cnt, 0);
        k_vox_scatter_ranked<<<nb, NT, 0, s>>>(d.st, d.rk0, (const int*)d.rk1, d.vox_cnt, d.blk_hist, d.Mcap, d.rv0);
        k_vox_reduce<0, true><<<1024, 256, 0, s>>>(d.st, d.rv0, d.rv1, d.vox_cnt, d.P, d.Nrm, d.pts, d.nrms, d.pset, d.vset, d.nraw, d.vnd, d.vset_it);
    } else {
        k_vox_assign<<<nb, NT, 0, s>>>(d.P, d.st, d.vkeys, d.rk0, d.rv0, d.vox_cnt);
        k_scan_counts<<<1, 1024, 0, s>>>(d.st, d.vox_cnt, 0);
        // large arenas (maps): group by voxel with atomics, then fix the order inside every voxel in the reduce kernel (no global sort)
        cudaMemsetAsync(d.vox_fill, 0, sizeof(int) * (d.Mcap + 2), s);
        k_vox_scatter<<<nb, NT, 0, s>>>(d.st, d.rk0, d.vox_cnt, d.vox_fill, d.rv0);
        k_vox_reduce<0><<<1024, 256, 0, s>>>(d.st, d.rv0, d.rv1, d.vox_cnt, d.P, d.Nrm, d.pts, d.nrms, d.pset, d.vset, d.nraw, d.vnd, d.vset_it);
    }
    if (h->overlap) cudaEventRecord(h->ev_pts, s);                         // the downsampled positions exist: surf kNN + schedule may start (stream3, launch_mscale)
    if (h->overlap) cudaStreamWaitEvent(s, h->ev_join[0], 0);              // phase 1 of the reduce is the first consumer of the normals
    k_vox_reduce<1><<<1024, 256, 0, s>>>(d.st, d.rv0, d.rv1, d.vox_cnt, d.P, d.Nrm, d.pts, d.nrms, d.pset, d.vset, d.nraw, d.vnd, d.vset_it);
    stage_mark(h, SM_VOXEL);
    h->n_launches += 3 + LEAF_MAX_ITERS + 5 + 4 + 4 + 2 + ((k <= 24 && !h->knn_warp_path) ? 2 : 0) + 1;
}

}  // namespace osep
