// csrc/map.cu -- SURVEY.md 8f-3 / 8f-4: the stages either side of the hot path that keep the data on the device.
//   * k_tf_points: tf2::doTransform(PointCloud2) (gskel_node.cpp:137-138; tf2_sensor_msgs: float32
//     Translation3f * Quaternionf applied to every point) as a kernel -- Rosa's vertices go to GSkel, and streaming
//     frames go into the accumulated map, without a host round trip of the points.
//   * osep_map: the accumulated map as a voxel hash on the device.  Every frame is transformed into the map frame and
//     inserted; a voxel keeps the FIRST point that ever fell into it (lowest frame, then lowest index in the frame), and
//     the kept points form an append-only dense array in exactly that order -- deterministic whatever the thread
//     interleaving, so the CPU restatement (tests/pyref/map_ref.py) reproduces it bit for bit.  The dense array is what
//     the tiled path (BASELINE.json configs[4]) splits across the GPUs.
// The reference has no map accumulation (its RosaNode consumes one PointCloud2 per tick, rosa_node.cpp:120-140): this is
// the "next" row of SURVEY.md 8f, specified here and restated on the CPU for parity.
#include "osep_common.cuh"
#include <cstdio>
#include <vector>

namespace osep {

struct Tf { float r[9]; float t[3]; };

// Eigen 3.4 Quaternion::toRotationMatrix in float32 (the arithmetic of osep_gskel_run_tf / the CPU restatement)
static Tf make_tf(const double* translation, const double* q) {
    Tf T;
    if (!translation || !q) { for (int i = 0; i < 9; ++i) T.r[i] = (i % 4 == 0) ? 1.0f : 0.0f; T.t[0] = T.t[1] = T.t[2] = 0.0f; return T; }
    const float qx = (float)q[0], qy = (float)q[1], qz = (float)q[2], qw = (float)q[3];
    const float tx = 2.0f * qx, ty = 2.0f * qy, tz = 2.0f * qz;
    const float twx = tx * qw, twy = ty * qw, twz = tz * qw;
    const float txx = tx * qx, txy = ty * qx, txz = tz * qx;
    const float tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    T.r[0] = 1.0f - (tyy + tzz); T.r[1] = txy - twz; T.r[2] = txz + twy;
    T.r[3] = txy + twz; T.r[4] = 1.0f - (txx + tzz); T.r[5] = tyz - twx;
    T.r[6] = txz - twy; T.r[7] = tyz + twx; T.r[8] = 1.0f - (txx + tyy);
    T.t[0] = (float)translation[0]; T.t[1] = (float)translation[1]; T.t[2] = (float)translation[2];
    return T;
}
__device__ __forceinline__ f3 tf_apply(const Tf& T, float x, float y, float z) {
    return {sum3(T.r[0] * x, T.r[1] * y, T.r[2] * z) + T.t[0], sum3(T.r[3] * x, T.r[4] * y, T.r[5] * z) + T.t[1], sum3(T.r[6] * x, T.r[7] * y, T.r[8] * z) + T.t[2]};
}

__global__ void k_tf_points(const float* __restrict__ in, int n, int stride, Tf T, int identity, float* __restrict__ out, int out_stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = in + (size_t)i * stride;
    f3 g{p[0], p[1], p[2]};
    if (!identity) g = tf_apply(T, p[0], p[1], p[2]);
    float* o = out + (size_t)i * out_stride;
    o[0] = g.x; o[1] = g.y; o[2] = g.z;
    if (out_stride > 3) o[3] = 1.0f;
}

int tf_points_enqueue(const float* in_dev, int n, int stride, const double* translation, const double* rotation_xyzw, float* out_dev, int out_stride, cudaStream_t s) {
    if (n <= 0) return OSEP_OK;
    const Tf T = make_tf(translation, rotation_xyzw);
    k_tf_points<<<div_up(n, 256), 256, 0, s>>>(in_dev, n, stride, T, (translation && rotation_xyzw) ? 0 : 1, out_dev, out_stride);
    return cudaGetLastError() == cudaSuccess ? OSEP_OK : OSEP_ERR_CUDA;
}

// ---- accumulated map -------------------------------------------------------------------------------------------------
struct MapDev {
    int device = 0;
    cudaStream_t stream = nullptr;
    unsigned long long* keys = nullptr;       // open addressing, 0 = empty, else packed voxel coordinates + 1
    unsigned long long* ticket = nullptr;     // (frame << 32 | index in the frame) of the first point of the voxel; ~0 = none yet
    int* slot_of = nullptr;                   // per input point of the current frame: its voxel's slot (-1: dropped, non-finite)
    int* blk = nullptr;                       // per-block winner counts / offsets of the current frame
    int* ctr = nullptr;                       // [0] new voxels of the current frame  [1] table overflow flag
    float4* pts = nullptr;                    // dense, append only: the kept points in (frame, index) order
    unsigned T = 0; int cap_pts = 0; int cap_frame = 0;
    float voxel = 0.1f;
    int n_pts = 0; unsigned frame = 0;
    int* ctr_host = nullptr;
};

__device__ __forceinline__ unsigned long long map_key(int ix, int iy, int iz) {       // 21 bits per axis around 2^20
    return ((((unsigned long long)(unsigned)(iz + (1 << 20)) & 0x1fffffull) << 42) | (((unsigned long long)(unsigned)(iy + (1 << 20)) & 0x1fffffull) << 21) |
            ((unsigned long long)(unsigned)(ix + (1 << 20)) & 0x1fffffull)) + 1ull;
}
__device__ __forceinline__ unsigned map_hash(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k;
}

// per point: transform, voxel key, find-or-insert the voxel, race for the voxel's ticket with (frame, index)
__global__ void k_map_insert(const float* __restrict__ in, int n, int stride, Tf T, int identity, float inv_voxel, unsigned frame,
                             unsigned long long* keys, unsigned long long* ticket, unsigned mask, int* __restrict__ slot_of, int* ctr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = in + (size_t)i * stride;
    f3 g{p[0], p[1], p[2]};
    if (!identity) g = tf_apply(T, p[0], p[1], p[2]);
    int slot = -1;
    const float fx = floorf(g.x * inv_voxel), fy = floorf(g.y * inv_voxel), fz = floorf(g.z * inv_voxel);
    if (fin3(g) && fabsf(fx) < 1048575.0f && fabsf(fy) < 1048575.0f && fabsf(fz) < 1048575.0f) {
        const unsigned long long key = map_key((int)fx, (int)fy, (int)fz);
        unsigned s = map_hash(key) & mask;
        for (unsigned probes = 0; probes <= mask; ++probes) {
            const unsigned long long cur = keys[s];
            if (cur == key) { slot = (int)s; break; }
            if (cur == 0ull) {
                const unsigned long long old = atomicCAS(&keys[s], 0ull, key);
                if (old == 0ull || old == key) { slot = (int)s; break; }
            }
            s = (s + 1) & mask;
        }
        if (slot < 0) atomicExch(&ctr[1], 1);                     // table full
        else atomicMin(&ticket[slot], ((unsigned long long)frame << 32) | (unsigned)i);
    }
    slot_of[i] = slot;
}
__device__ __forceinline__ bool map_winner(const unsigned long long* __restrict__ ticket, const int* __restrict__ slot_of, unsigned frame, int i, int n) {
    if (i >= n) return false;
    const int s = slot_of[i];
    return s >= 0 && ticket[s] == (((unsigned long long)frame << 32) | (unsigned)i);
}
__global__ void k_map_count(const unsigned long long* __restrict__ ticket, const int* __restrict__ slot_of, unsigned frame, int n, int* __restrict__ blk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = __syncthreads_count(map_winner(ticket, slot_of, frame, i, n));
    if (threadIdx.x == 0) blk[blockIdx.x] = c;
}
__global__ void k_map_scan(int* blk, int nblk, int* ctr) {
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < nblk ? blk[i] : 0;
        int incl = v;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const int ws = wsum[lane]; int wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            wsum[lane] = wi - ws;
        }
        __syncthreads();
        if (i < nblk) blk[i] = carry + wsum[wid] + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += wsum[wid] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) ctr[0] = carry;
}
// the winners of the frame, in input order, are appended to the dense point array
__global__ void k_map_append(const float* __restrict__ in, int n, int stride, Tf T, int identity, const unsigned long long* __restrict__ ticket,
                             const int* __restrict__ slot_of, unsigned frame, const int* __restrict__ blk_off, float4* __restrict__ pts, int base, int cap) {
    __shared__ int warp_cnt[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool win = map_winner(ticket, slot_of, frame, i, n);
    const unsigned m = __ballot_sync(0xffffffffu, win);
    if (lane == 0) warp_cnt[wid] = __popc(m);
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < wid; ++w) woff += warp_cnt[w];
    if (win) {
        const int pos = base + blk_off[blockIdx.x] + woff + __popc(m & ((1u << lane) - 1u));
        if (pos < cap) {
            const float* p = in + (size_t)i * stride;
            f3 g{p[0], p[1], p[2]};
            if (!identity) g = tf_apply(T, p[0], p[1], p[2]);
            pts[pos] = make_float4(g.x, g.y, g.z, 1.0f);
        }
    }
}

}  // namespace osep

using namespace osep;
struct osep_map { MapDev d; };

extern "C" {

int osep_tf_points_device(int device, const float* in_dev, int n, int stride_floats, const double* translation, const double* rotation_xyzw,
                          float* out_dev, int out_stride, void* stream) {
    if (n < 0 || stride_floats < 3 || (out_stride != 3 && out_stride != 4) || (n > 0 && (!in_dev || !out_dev))) return OSEP_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return OSEP_ERR_CUDA;
    return tf_points_enqueue(in_dev, n, stride_floats, translation, rotation_xyzw, out_dev, out_stride, (cudaStream_t)stream);
}

int osep_map_create(int device, int capacity_points, int max_frame_points, float voxel_size, osep_map** out) {
    if (!out || capacity_points < 1 || max_frame_points < 1 || !(voxel_size > 0.0f)) return OSEP_ERR_ARG;
    *out = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return OSEP_ERR_CUDA;
    osep_map* m = new osep_map();
    MapDev& d = m->d;
    d.device = device; d.voxel = voxel_size; d.cap_pts = capacity_points; d.cap_frame = max_frame_points;
    d.T = 1; while (d.T < 2u * (unsigned)capacity_points) d.T <<= 1;
    bool ok = cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&d.keys, sizeof(unsigned long long) * d.T) == cudaSuccess && cudaMalloc((void**)&d.ticket, sizeof(unsigned long long) * d.T) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&d.slot_of, sizeof(int) * (size_t)max_frame_points) == cudaSuccess && cudaMalloc((void**)&d.blk, sizeof(int) * ((size_t)max_frame_points / 256 + 2)) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&d.ctr, sizeof(int) * 4) == cudaSuccess && cudaMalloc((void**)&d.pts, sizeof(float4) * (size_t)capacity_points) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&d.ctr_host, sizeof(int) * 4) == cudaSuccess;
    ok = ok && cudaMemset(d.keys, 0, sizeof(unsigned long long) * d.T) == cudaSuccess && cudaMemset(d.ticket, 0xff, sizeof(unsigned long long) * d.T) == cudaSuccess &&
         cudaMemset(d.ctr, 0, sizeof(int) * 4) == cudaSuccess;
    if (!ok) { cudaGetLastError(); osep_map_destroy(m); return OSEP_ERR_CUDA; }
    *out = m;
    return OSEP_OK;
}

int osep_map_destroy(osep_map* m) {
    if (!m) return OSEP_ERR_ARG;
    MapDev& d = m->d;
    cudaSetDevice(d.device);
    void* ptrs[] = {d.keys, d.ticket, d.slot_of, d.blk, d.ctr, d.pts};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (d.ctr_host) cudaFreeHost(d.ctr_host);
    if (d.stream) cudaStreamDestroy(d.stream);
    delete m;
    return OSEP_OK;
}

int osep_map_insert_device(osep_map* m, const float* xyz_dev, int n, int stride_floats, const double* translation, const double* rotation_xyzw, int* n_new) {
    if (!m || n < 0 || stride_floats < 3 || (n > 0 && !xyz_dev)) return OSEP_ERR_ARG;
    MapDev& d = m->d;
    if (n > d.cap_frame) return OSEP_ERR_CAPACITY;
    if (cudaSetDevice(d.device) != cudaSuccess) return OSEP_ERR_CUDA;
    if (n_new) *n_new = 0;
    const unsigned frame = d.frame++;
    if (n == 0) return OSEP_OK;
    const Tf T = make_tf(translation, rotation_xyzw);
    const int identity = (translation && rotation_xyzw) ? 0 : 1;
    const int nb = div_up(n, 256);
    const float inv = 1.0f / d.voxel;
    k_map_insert<<<nb, 256, 0, d.stream>>>(xyz_dev, n, stride_floats, T, identity, inv, frame, d.keys, d.ticket, d.T - 1, d.slot_of, d.ctr);
    k_map_count<<<nb, 256, 0, d.stream>>>(d.ticket, d.slot_of, frame, n, d.blk);
    k_map_scan<<<1, 1024, 0, d.stream>>>(d.blk, nb, d.ctr);
    k_map_append<<<nb, 256, 0, d.stream>>>(xyz_dev, n, stride_floats, T, identity, d.ticket, d.slot_of, frame, d.blk, d.pts, d.n_pts, d.cap_pts);
    if (cudaMemcpyAsync(d.ctr_host, d.ctr, sizeof(int) * 2, cudaMemcpyDeviceToHost, d.stream) != cudaSuccess || cudaStreamSynchronize(d.stream) != cudaSuccess) { cudaGetLastError(); return OSEP_ERR_CUDA; }
    const int fresh = d.ctr_host[0];
    if (d.ctr_host[1] || d.n_pts + fresh > d.cap_pts) return OSEP_ERR_CAPACITY;      // the map is full: nothing past the capacity was written
    d.n_pts += fresh;
    if (n_new) *n_new = fresh;
    return OSEP_OK;
}

int osep_map_insert(osep_map* m, const float* xyz_host, int n, int stride_floats, const double* translation, const double* rotation_xyzw, int* n_new) {
    if (!m || n < 0 || stride_floats < 3 || (n > 0 && !xyz_host)) return OSEP_ERR_ARG;
    MapDev& d = m->d;
    if (n > d.cap_frame) return OSEP_ERR_CAPACITY;
    if (cudaSetDevice(d.device) != cudaSuccess) return OSEP_ERR_CUDA;
    float* stage = nullptr;
    if (n == 0) return osep_map_insert_device(m, nullptr, 0, stride_floats, translation, rotation_xyzw, n_new);
    if (cudaMallocAsync((void**)&stage, sizeof(float) * (size_t)n * stride_floats, d.stream) != cudaSuccess) { cudaGetLastError(); return OSEP_ERR_CUDA; }
    cudaMemcpyAsync(stage, xyz_host, sizeof(float) * (size_t)n * stride_floats, cudaMemcpyHostToDevice, d.stream);
    const int rc = osep_map_insert_device(m, stage, n, stride_floats, translation, rotation_xyzw, n_new);
    cudaFreeAsync(stage, d.stream);
    cudaStreamSynchronize(d.stream);
    return rc;
}

int osep_map_points_device(osep_map* m, const float** xyzw_dev, int* n_points) {
    if (!m || !xyzw_dev || !n_points) return OSEP_ERR_ARG;
    *xyzw_dev = reinterpret_cast<const float*>(m->d.pts);
    *n_points = m->d.n_pts;
    return OSEP_OK;
}

int osep_map_clear(osep_map* m) {
    if (!m) return OSEP_ERR_ARG;
    MapDev& d = m->d;
    if (cudaSetDevice(d.device) != cudaSuccess) return OSEP_ERR_CUDA;
    cudaMemsetAsync(d.keys, 0, sizeof(unsigned long long) * d.T, d.stream);
    cudaMemsetAsync(d.ticket, 0xff, sizeof(unsigned long long) * d.T, d.stream);
    cudaMemsetAsync(d.ctr, 0, sizeof(int) * 4, d.stream);
    d.n_pts = 0; d.frame = 0;
    return cudaStreamSynchronize(d.stream) == cudaSuccess ? OSEP_OK : OSEP_ERR_CUDA;
}

}  // extern "C"
