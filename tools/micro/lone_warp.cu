// tools/micro/lone_warp.cu -- why does a lone task warp on an SM run ~3x slower than the same warp among busy warps?
// One CTA per SM, 8 warps.  Warp 0 runs a dependent FP chain (straight-line, BODY instructions, repeated REP times);
// the other 7 warps (mode 0) exit, (mode 1) poll a global flag with volatile loads, (mode 2) run the same chain,
// (mode 3) wait at a named barrier.  Reports cycles per instruction of warp 0.
#include <cstdio>
#include <cuda_runtime.h>
template <int N> struct Chain { static __device__ __forceinline__ float run(float x, float a) { x = x * a + 0.25f; return Chain<N - 1>::run(x, a); } };
template <> struct Chain<0> { static __device__ __forceinline__ float run(float x, float) { return x; } };
template <int BODY>
__global__ void __launch_bounds__(256) k(int mode, int rep, volatile int* flag, float* out, long long* cyc) {
    const int warp = threadIdx.x >> 5;
    float x = threadIdx.x * 1e-3f;
    if (warp == 0 || mode == 2) {
        const long long t0 = clock64();
        for (int r = 0; r < rep; ++r) x = Chain<BODY>::run(x, 0.999f);
        const long long t1 = clock64();
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
        out[blockIdx.x * 256 + threadIdx.x] = x;
        if (warp == 0 && threadIdx.x == 0 && mode == 1) { __threadfence(); atomicAdd((int*)flag, 1); }
        if (mode == 3 && warp == 0) asm volatile("bar.sync 1, 256;");
    } else if (mode == 1) {
        if ((threadIdx.x & 31) == 0) while (*flag < (int)gridDim.x) __nanosleep(20);
    } else if (mode == 3) {
        asm volatile("bar.sync 1, 256;");
    }
}
template <int BODY> void run(const char* name, int rep) {
    int* flag; float* out; long long* cyc; cudaMalloc(&flag, 4); cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int mode = 0; mode < 4; ++mode) {
        double best = 0;
        for (int trial = 0; trial < 3; ++trial) {
            cudaMemset(flag, 0, 4);
            k<BODY><<<148, 256>>>(mode, rep, flag, out, cyc);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
            best = s / 148 / ((double)BODY * 2 * rep);
        }
        printf("%s mode %d (%s): %.2f cycles per dependent instruction\n", name, mode, mode == 0 ? "others exit" : mode == 1 ? "others poll" : mode == 2 ? "others run the chain" : "others at a barrier", best);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<64>("loop body 2 KB x 64   ", 64);
    run<512>("loop body 16 KB x 8   ", 8);
    run<1024>("loop body 32 KB x 4   ", 4);
    run<2048>("straight 64 KB x 2    ", 2);
    return 0;
}
