// tools/micro/lone_warp2.cu -- does an SM that has been idle (all warps asleep / polling) run its next burst of work slower?
// One CTA per SM, 8 warps.  Warp 0: repeat { idle for IDLE_US (nanosleep), then a dependent FP chain of ~2500 instructions }.
// Other warps: mode 0 exit, mode 1 poll a global flag (volatile load + nanosleep), mode 2 = one HEATER warp runs a
// register-only FMA loop until warp 0 is done (reads a shared flag every 256 FMAs), mode 3 = all 7 are heaters.
#include <cstdio>
#include <cuda_runtime.h>
template <int N> struct Chain { static __device__ __forceinline__ float run(float x, float a) { x = x * a + 0.25f; return Chain<N - 1>::run(x, a); } };
template <> struct Chain<0> { static __device__ __forceinline__ float run(float x, float) { return x; } };
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(256) k(int mode, int idle_us, int bursts, volatile int* flag, float* out, long long* cyc) {
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) done = 0;
    __syncthreads();
    float x = threadIdx.x * 1e-3f;
    if (warp == 0) {
        long long tot = 0, first = 0;
        for (int b = 0; b < bursts; ++b) {
            const long long w0 = gtime();
            while (gtime() - w0 < (long long)idle_us * 1000) __nanosleep(200);
            const long long t0 = clock64();
            x = Chain<1250>::run(x, 0.999f);
            const long long t1 = clock64();
            tot += t1 - t0; if (b == 0) first = t1 - t0;
        }
        if (threadIdx.x == 0) { cyc[blockIdx.x * 2] = tot; cyc[blockIdx.x * 2 + 1] = first; }
        out[blockIdx.x * 256 + threadIdx.x] = x;
        __syncwarp();
        if (threadIdx.x == 0) { done = 1; __threadfence(); atomicAdd((int*)flag, 1); }
    } else if (mode == 1) {
        if ((threadIdx.x & 31) == 0) while (*flag < (int)gridDim.x) __nanosleep(20);
    } else if ((mode == 2 && warp == 1) || mode == 3) {
        while (!done) { for (int u = 0; u < 256; ++u) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x)); }
        out[blockIdx.x * 256 + threadIdx.x] = x;
    }
}
int main() {
    int* flag; float* out; long long* cyc; cudaMalloc(&flag, 4); cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 16);
    const char* names[4] = {"others exit", "others poll", "one heater warp", "seven heater warps"};
    for (int idle = 0; idle <= 100; idle = idle == 0 ? 5 : idle * 2 + (idle == 20 ? 10 : 0)) {
        for (int mode = 0; mode < 4; ++mode) {
            double a = 0, f = 0;
            for (int trial = 0; trial < 2; ++trial) {
                cudaMemset(flag, 0, 4);
                k<<<148, 256>>>(mode, idle, 20, flag, out, cyc);
                cudaDeviceSynchronize();
                long long h[296]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
                a = 0; f = 0; for (int i = 0; i < 148; ++i) { a += h[2 * i]; f += h[2 * i + 1]; }
                a = a / 148 / (2500.0 * 20); f = f / 148 / 2500.0;
            }
            printf("idle %3d us, %-18s: %.2f cycles per dependent instruction (first burst %.2f)\n", idle, names[mode], a, f);
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
