// cost of the CTA-scope fences / release stores a single warp pays between shared-memory stores
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(long long* out, int* sink) {
    __shared__ int buf[64];
    volatile int* vb = buf;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) {
        vb[threadIdx.x] = i;
        if (OP == 1) __threadfence_block();
        if (OP == 2) asm volatile("fence.acq_rel.cta;" ::: "memory");
        if (OP == 3) { unsigned a = (unsigned)__cvta_generic_to_shared(&buf[32]); asm volatile("st.release.cta.shared.b32 [%0], %1;" :: "r"(a), "r"(i) : "memory"); }
        if (OP == 4) __syncwarp();
        if (OP == 5) { __syncwarp(); __threadfence_block(); }
        if (OP != 3) vb[32 + (threadIdx.x & 31)] = i;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    sink[threadIdx.x] = buf[threadIdx.x];
}
int main() {
    long long* o; int* s; cudaMalloc(&o, 8); cudaMalloc(&s, 256);
    const char* nm[] = {"2 STS (no fence)", "__threadfence_block", "fence.acq_rel.cta", "st.release.cta.shared", "__syncwarp", "__syncwarp+__threadfence_block"};
    long long c;
#define RUN(OP) k<OP><<<1, 32>>>(o, s); cudaDeviceSynchronize(); k<OP><<<1, 32>>>(o, s); cudaDeviceSynchronize(); cudaMemcpy(&c, o, 8, cudaMemcpyDeviceToHost); printf("%-32s %.1f cycles per iteration\n", nm[OP], c / 1024.0);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    return 0;
}
