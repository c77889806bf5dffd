// Micro-benchmark: latency of the device 3x3 eigen solve (dev_math.cuh::eig3_sym) on smoothing-like matrices,
// single lane vs several lanes of one warp with different matrices (divergence cost).
#include "../../osep_path_planner_b200/csrc/dev_math.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
using namespace osep;
__global__ void k(const float* mats, int nmat, int lanes, int reps, long long* out, float* sink) {
    const int lane = threadIdx.x;
    float acc = 0.f;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (lane < lanes) {
            const float* m = mats + 6 * ((r * 32 + lane) % nmat);
            // make the next solve depend on the previous result (dependent chain like the Gauss-Seidel walker)
            Eig3 E = eig3_sym(m[0] + acc * 0.f, m[1], m[2], m[3], m[4], m[5]);
            f3 d = normalize3_dynrow(eig_col(E, 2));
            acc = d.x;
        }
        __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    sink[lane] = acc;
}
int main() {
    const int nmat = 4096; std::vector<float> h(6 * nmat);
    srand(1);
    auto rnd = []() { return (rand() / (float)RAND_MAX) * 2.f - 1.f; };
    for (int t = 0; t < nmat; ++t) {
        float b[3] = {rnd(), rnd(), rnd()}; float nb = sqrtf(b[0]*b[0]+b[1]*b[1]+b[2]*b[2]); for (auto& x : b) x /= nb;
        float M[9] = {0};
        for (int j = 0; j < 10; ++j) {
            float v[3] = {b[0] + 0.15f * rnd(), b[1] + 0.15f * rnd(), b[2] + 0.15f * rnd()};
            float nv = sqrtf(v[0]*v[0]+v[1]*v[1]+v[2]*v[2]); for (auto& x : v) x /= nv;
            float w = 1.0f / (powf(fabsf(rnd()) * 0.3f, 4.f) + 1e-5f);
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r*3+c] += (w * v[r]) * v[c];
        }
        h[6*t+0]=M[0]; h[6*t+1]=M[3]; h[6*t+2]=M[4]; h[6*t+3]=M[6]; h[6*t+4]=M[7]; h[6*t+5]=M[8];
    }
    float* d; long long* o; float* sink;
    cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&o, 8); cudaMalloc(&sink, 128);
    for (int lanes : {1, 2, 4, 8, 32}) {
        k<<<1, 32>>>(d, nmat, lanes, 2000, o, sink); cudaDeviceSynchronize();
        k<<<1, 32>>>(d, nmat, lanes, 2000, o, sink); cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, o, 8, cudaMemcpyDeviceToHost);
        printf("lanes %2d: %.0f cycles per dependent eigen solve + normalize\n", lanes, c / 2000.0);
    }
    return 0;
}
