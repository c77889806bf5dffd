// Exhaustive check behind eig_fast.cuh's "early seed" Givens: for EVERY float w in [1, 2] and every seed within
// +-K ulps of MUFU.RSQ(w):   sqrt-from-seed(w) == sqrtf(w)   and   rcp-from-seed(sqrtf(w)) == 1.0f / sqrtf(w).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float mufu_rsq(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__global__ void k(int K, unsigned long long* bad_sqrt, unsigned long long* bad_rcp, int* worst_s, int* worst_r) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > 0x800000u) return;
    const float w = __uint_as_float(0x3f800000u + i);
    const float u = sqrtf(w), iu = 1.0f / u;
    const float r0 = mufu_rsq(w);
    for (int kk = -K; kk <= K; ++kk) {
        const float rs = __uint_as_float(__float_as_uint(r0) + kk);
        const float y = w * rs, h = rs * 0.5f;
        const float e = fmaf(-y, y, w);
        const float s = fmaf(e, h, y);
        if (__float_as_uint(s) != __float_as_uint(u)) { atomicAdd(bad_sqrt, 1ull); atomicMin(worst_s, abs(kk)); }
        const float e3 = fmaf(-u, rs, 1.0f);
        const float inv = fmaf(rs, e3, rs);
        if (__float_as_uint(inv) != __float_as_uint(iu)) { atomicAdd(bad_rcp, 1ull); atomicMin(worst_r, abs(kk)); }
    }
}
int main() {
    unsigned long long* c; int* w; cudaMalloc(&c, 16); cudaMalloc(&w, 8);
    for (int K : {8, 16, 24, 32, 48, 64}) {
        cudaMemset(c, 0, 16); int init[2] = {1 << 30, 1 << 30}; cudaMemcpy(w, init, 8, cudaMemcpyHostToDevice);
        k<<<(0x800001 + 255) / 256, 256>>>(K, c, c + 1, w, w + 1);
        unsigned long long h[2]; int hw[2];
        cudaMemcpy(h, c, 16, cudaMemcpyDeviceToHost); cudaMemcpy(hw, w, 8, cudaMemcpyDeviceToHost);
        printf("K=%2d: sqrt mismatches %llu (smallest |k| %d), rcp mismatches %llu (smallest |k| %d)\n", K, h[0], hw[0], h[1], hw[1]);
    }
    return 0;
}
