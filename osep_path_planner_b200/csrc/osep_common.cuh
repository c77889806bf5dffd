// csrc/osep_common.cuh -- device-resident frame state + launch helpers shared by the kernels.
//
// Design rule: one frame runs with ZERO host synchronisations.  Everything the reference
// decides on the host between stages (filtered count, leaf-size control law, M, S, V, stage
// failures) lives in one FrameState struct in HBM; kernels are launched with capacity-sized
// grids, read their sizes from it and exit early.  That makes the whole frame one stream of
// launches (CUDA-graph capturable) and lets many frames overlap on many streams.
#pragma once
#include "dev_math.cuh"
#include "eig_fast.cuh"
#include "../../include/osep_skel.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace osep {

constexpr int LEAF_MAX_ITERS = 8;       // rosa.cpp:120
constexpr int KNN_MAX = 64;             // largest ne_knn / nb_knn supported
constexpr int WARPS_PER_BLOCK = 8;

struct FrameState {
    // ---- sizes
    int n_in;
    int n_filt;            // N'
    int M;                 // downsampled points (pcd_size_)
    int Mc;                // finite pset rows in vertex_sampling (rosa.cpp:448)
    int S;                 // total similarity neighbours
    int V;                 // vertices
    int Vf;                // finite vertices in vertex_smooth (rosa.cpp:552)
    int status;            // 0, failing stage 1..7, or OSEP_ERR_*
    unsigned flags;
    // ---- bbox of the filtered cloud (ordered-int encoding for atomics, then decoded)
    int bb_lo[3], bb_hi[3];
    float bmin[3], bmax[3];
    // ---- kNN grid
    float g_inv_h, g_h;
    float g_h_last;        // cell size that hits the target occupancy on the last frame of this handle (persists across frames: a warm start -- only the speed of the search depends on it)
    int g_dim[3];
    int g_cells_used;      // allocation cursor (points)
    int g_ncells, g_redo;  // occupied cells of the current build; rebuild requested
    int n_units;           // kNN work units (cell, chunk of queries)
    int n_slow;            // queries handed to the exact slow path
    int n_over;            // units whose candidate set does not fit the shared-memory stage
    int n_cells;           // occupied grid cells with a compact id (k_grid_alloc)
    int n_retry, n_grow;   // thread-per-query kNN: queries that repeat the scan with a rescaled radius / with the 5x5x5 block
    int unit_ctr;          // dynamic work counter of k_knn_cells
    // ---- leaf loop (rosa.cpp:118-148)
    float leaf, leaf_used, leaf_size;
    int leaf_done, n_pass, nvox;
    int pass_overflow;
    int vox_identity;      // final VoxelGrid pass hit PCL's int32 index guard: output = input (every filtered point is its own voxel)
    unsigned epoch;
    float leaf_hist[LEAF_MAX_ITERS];
    int nvox_hist[LEAF_MAX_ITERS];
    // voxel-grid parameters of the current pass (PCL VoxelGrid::applyFilter)
    float inv_leaf;
    int minb[3], divb[3];
    // ---- drosa bookkeeping
    int n_levels;
    int n_good, n_poor;
    // ---- voxel passes
    int pass_active;       // the VoxelGrid pass set up by leaf_step is live (loop not finished)
    int n_keys;            // distinct voxel keys collected for the final pass
    int ticket;            // last-block detection (voxel count passes)
    int ticket_m;          // last-block detection (similarity count pass)
    int ticket_grid;       // last-block detection (kNN grid build; runs concurrently with the voxel passes)
    // ---- where the frame's time goes: globaltimer (ns) at the entry of selected kernels, written by block 0 thread 0 (tap "timeline", tools/frame_timeline.py)
    unsigned long long tstamp[18];
};
enum FrameStamp { TS_FILTER = 0, TS_LEAF0, TS_GRID, TS_KNN, TS_NORMALS, TS_VOXFINAL, TS_REDUCE0, TS_SURF, TS_SCHED, TS_SCHED_END, TS_REDUCE1, TS_SIMI, TS_DROSA, TS_FIXUP, TS_DCROSA, TS_TAIL, TS_END, TS_COUNT };

#ifdef __CUDACC__
__device__ __forceinline__ void frame_stamp(const FrameState* st, int k) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); const_cast<FrameState*>(st)->tstamp[k] = t; }
}
#endif
static_assert(sizeof(FrameState) <= 512, "the frame export block reserves 512 bytes for FrameState");

// Per-frame launch parameters.  They live in device memory (refreshed by one H2D copy node from a pinned host mirror at the
// head of every frame) instead of being kernel arguments, so the captured CUDA graph of a frame does not depend on the
// frame's point count, stride or buffer address: a LiDAR stream whose n changes every frame replays ONE graph.
struct FrameParams {
    const float* in;       // input points on the device (x, y, z at [i*stride .. i*stride+2])
    int n;                 // points in this frame
    int stride;            // floats per input point
    int pad;
};

// per-point normal quantities used inside the active-set sums of drosa, computed once per frame (by the voxel reduce):
//   nraw = (n, il_dd)  il_dd = (s2 > 1e-20f) ? (float)(1.0 / (double)sqrtf(s2)) : 0      rosa.cpp:764-766
//   vnd  = (n * ((double)s2 > 1e-20 ? 1/sqrtf(s2) : 0), il_f)                             rosa.cpp:694-696, 740-742
//          il_f = (s2 > 1e-20f) ? 1.0f / sqrtf(s2) : 0                                    rosa.cpp:340-341, 717-718
__device__ __forceinline__ void normal_prep_vals(const f3& v, float4& nraw, float4& vnd) {
    const float s2 = sqn3(v);
    const float rt = sqrtf(s2);
    const float il_dd = (s2 > 1e-20f) ? (float)(1.0 / (double)rt) : 0.0f;
    const float il_d = ((double)s2 > 1e-20) ? 1.0f / rt : 0.0f;
    const float il_f = (s2 > 1e-20f) ? 1.0f / rt : 0.0f;
    nraw = make_float4(v.x, v.y, v.z, il_dd);
    const f3 u = mul3(v, il_d);
    vnd = make_float4(u.x, u.y, u.z, il_f);
}

// ordered-int encoding of a float: monotone map so atomicMin/atomicMax(int) order floats
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned hash_u32(unsigned k) { k *= 2654435761u; k ^= k >> 15; k *= 2246822519u; k ^= k >> 13; return k; }

// stage a dense bit matrix (rows x Wn words, contiguous) into shared memory with 16-byte loads
__device__ __forceinline__ void stage_bit_rows(unsigned* __restrict__ dst, const unsigned* __restrict__ src, int rows, int Wn) {
    const int total = rows * Wn;
    const int n4 = total >> 2;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll 4
    for (int i = threadIdx.x; i < n4; i += blockDim.x) d4[i] = s4[i];
    for (int i = (n4 << 2) + threadIdx.x; i < total; i += blockDim.x) dst[i] = src[i];
}

inline int div_up(int a, int b) { return (a + b - 1) / b; }

}  // namespace osep
