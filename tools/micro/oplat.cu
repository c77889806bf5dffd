// dependent-chain latencies of the ops on the eigen solve's critical path (single warp, clock64)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float mufu_rcp(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_rsq(float x) { float r; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
template <int OP> __global__ void k(float x0, float y0, long long* out, float* sink) {
    float x = x0, y = y0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (OP == 0) x = fmaf(x, y, y);
            if (OP == 1) x = mufu_rcp(x);
            if (OP == 2) x = mufu_rsq(x);
            if (OP == 3) x = (x > y) ? y : x * 1.0001f;          // FSETP+FSEL / FMUL mix
            if (OP == 4) x = x / y;                               // full IEEE div
            if (OP == 5) x = sqrtf(x);                            // full IEEE sqrt
            if (OP == 6) x = __shfl_xor_sync(0xffffffffu, x, 1);
            if (OP == 7) x = fmaxf(x, y) + 1.0f;
            if (OP == 8) { x = mufu_rcp(x); x = fmaf(x, y, y); }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = x;
}
int main() {
    long long* o; float* s; cudaMalloc(&o, 8); cudaMalloc(&s, 128);
    const char* nm[] = {"FFMA", "MUFU.RCP", "MUFU.RSQ", "FSETP+FSEL|FMUL", "IEEE div", "IEEE sqrt", "SHFL", "FMNMX+FADD", "MUFU.RCP+FFMA"};
    long long c;
#define RUN(OP) k<OP><<<1, 32>>>(1.0001f, 0.99f, o, s); cudaDeviceSynchronize(); k<OP><<<1, 32>>>(1.0001f, 0.99f, o, s); cudaDeviceSynchronize(); cudaMemcpy(&c, o, 8, cudaMemcpyDeviceToHost); printf("%-18s %.1f cycles per dependent op\n", nm[OP], c / 4096.0);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    return 0;
}
