// Micro test + benchmark of csrc/eig_fast.cuh against the plain IEEE code of csrc/dev_math.cuh.
//   1. fdiv_nog / fsqrt_nog / frcp_nog vs `/`, sqrtf, 1.0f/x (compiled -prec-div=true -prec-sqrt=true) on every
//      guarded operand: sqrt and rcp exhaustively over all 2^32 bit patterns, div on 2^33 hashed operand pairs
//      plus every pair of a 4096-value edge set.
//   2. eig3_sym_dev vs eig3_sym, all outputs bit-for-bit, on several matrix families; reports the fallback rate.
//   3. cycles per DEPENDENT solve + normalize (the Gauss-Seidel walker's step), 1 and 4 active lanes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false
#include "../../osep_path_planner_b200/csrc/eig_fast.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
using namespace osep;

__device__ __forceinline__ unsigned hsh(unsigned k) { k ^= k >> 16; k *= 0x7feb352du; k ^= k >> 15; k *= 0x846ca68bu; k ^= k >> 16; return k; }

__global__ void k_unary(unsigned long long* bad_sqrt, unsigned long long* bad_rcp, unsigned long long* n_checked) {
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long ns = 0, nr = 0, nc = 0;
    for (unsigned long long b = tid; b < (1ull << 32); b += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)b);
        if (oor_den(x)) continue;
        ++nc;
        if (x > 0.0f) { const float a = fsqrt_nog(x), r = sqrtf(x); if (__float_as_uint(a) != __float_as_uint(r)) ++ns; }
        { const float a = frcp_nog(x), r = 1.0f / x; if (__float_as_uint(a) != __float_as_uint(r)) ++nr; }
    }
    atomicAdd(bad_sqrt, ns); atomicAdd(bad_rcp, nr); atomicAdd(n_checked, nc);
}

__global__ void k_div(unsigned long long* bad_div, unsigned long long* n_checked, unsigned seed, int per_thread) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long nb = 0, nc = 0;
    for (int i = 0; i < per_thread; ++i) {
        const unsigned h0 = hsh(tid * 2654435761u + i * 40503u + seed), h1 = hsh(h0 + 0x9e3779b9u);
        // exponents spread over the guarded ranges, random mantissas and signs
        unsigned ea = 42 + (h0 >> 8) % 146, eb = 67 + (h1 >> 8) % 101;
        if ((h0 & 255u) == 0u) ea = 0;      // zero numerators too
        float a = __uint_as_float((h0 & 0x80000000u) | (ea << 23) | (hsh(h0 ^ 0x1234567u) & 0x7fffffu));
        const float b = __uint_as_float((h1 & 0x80000000u) | (eb << 23) | (hsh(h1 ^ 0x89abcdeu) & 0x7fffffu));
        if (ea == 0) a = (h0 & 0x80000000u) ? -0.0f : 0.0f;
        if (oor_den(b) || oor_num(a)) continue;
        ++nc;
        const float q = fdiv_nog(a, b), r = a / b;
        if (__float_as_uint(q) != __float_as_uint(r)) ++nb;
    }
    atomicAdd(bad_div, nb); atomicAdd(n_checked, nc);
}
__global__ void k_div_edges(const float* vals, int n, unsigned long long* bad_div, unsigned long long* n_checked) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n) return;
    const float a = vals[i / n], b = vals[i % n];
    if (oor_den(b) || oor_num(a)) return;
    atomicAdd(n_checked, 1ull);
    const float q = fdiv_nog(a, b), r = a / b;
    if (__float_as_uint(q) != __float_as_uint(r)) atomicAdd(bad_div, 1ull);
}

__global__ void k_solver(const float* mats, int nmat, unsigned long long* mism, unsigned long long* fallback, unsigned long long* iters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nmat) return;
    const float* m = mats + 6 * (size_t)i;
    const Eig3 A = eig3_sym(m[0], m[1], m[2], m[3], m[4], m[5]);
    bool bad;
    Eig3 F = eig3_sym_fast(m[0], m[1], m[2], m[3], m[4], m[5], bad);
    if (bad) { atomicAdd(fallback, 1ull); F = eig3_sym(m[0], m[1], m[2], m[3], m[4], m[5]); }
    bool same = A.ok == F.ok && A.iters == F.iters;
    for (int k = 0; k < 3; ++k) same = same && __float_as_uint(A.ev[k]) == __float_as_uint(F.ev[k]);
    for (int k = 0; k < 9; ++k) same = same && __float_as_uint(A.q[k]) == __float_as_uint(F.q[k]);
    const f3 na = normalize3_dynrow(eig_col(A, 2)), nf = normalize3_dev(eig_col(F, 2));
    same = same && __float_as_uint(na.x) == __float_as_uint(nf.x) && __float_as_uint(na.y) == __float_as_uint(nf.y) && __float_as_uint(na.z) == __float_as_uint(nf.z);
    if (!same) atomicAdd(mism, 1ull);
    atomicAdd(iters, (unsigned long long)A.iters);
}

template <int FAST>
__global__ void k_lat(const float* mats, int nmat, int lanes, int reps, long long* out, float* sink) {
    const int lane = threadIdx.x;
    float acc = 0.f;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (lane < lanes) {
            const float* m = mats + 6 * ((r * 32 + lane) % nmat);
            f3 d;
            if (FAST) { const Eig3 E = eig3_sym_dev(m[0] + acc * 0.f, m[1], m[2], m[3], m[4], m[5]); d = normalize3_dev(eig_col(E, 2)); }
            else { const Eig3 E = eig3_sym(m[0] + acc * 0.f, m[1], m[2], m[3], m[4], m[5]); d = normalize3_dynrow(eig_col(E, 2)); }
            acc = d.x;
        }
        __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
    sink[lane] = acc;
}

static float rnd() { return (rand() / (float)RAND_MAX) * 2.f - 1.f; }

int main() {
    unsigned long long* cnt; cudaMalloc(&cnt, 64); cudaMemset(cnt, 0, 64);
    unsigned long long h[8];
    k_unary<<<148 * 16, 256>>>(cnt, cnt + 1, cnt + 2);
    cudaDeviceSynchronize(); cudaMemcpy(h, cnt, 64, cudaMemcpyDeviceToHost);
    printf("unary  : %llu guarded operands, sqrt mismatches %llu, rcp mismatches %llu\n", h[2], h[0], h[1]);
    cudaMemset(cnt, 0, 64);
    for (int s = 0; s < 8; ++s) k_div<<<1 << 14, 256>>>(cnt, cnt + 1, 0x51ed270bu * (s + 1), 256);
    cudaDeviceSynchronize(); cudaMemcpy(h, cnt, 64, cudaMemcpyDeviceToHost);
    printf("div    : %llu hashed pairs, mismatches %llu\n", h[1], h[0]);
    {
        std::vector<float> ev;
        for (int e = -90; e <= 62; ++e) for (float m : {1.0f, 1.0000001f, 1.5f, 1.9999999f, 1.3333334f, 1.7320508f}) for (float sg : {1.f, -1.f}) ev.push_back(sg * ldexpf(m, e));
        ev.push_back(0.f); ev.push_back(-0.f);
        while (ev.size() < 4096) { ev.push_back(ldexpf(1.0f + (rand() & 0x7fffff) / 8388608.0f, (rand() % 140) - 84) * ((rand() & 1) ? 1.f : -1.f)); }
        float* d; cudaMalloc(&d, ev.size() * 4); cudaMemcpy(d, ev.data(), ev.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(cnt, 0, 64);
        const int n = (int)ev.size();
        k_div_edges<<<(n * n + 255) / 256, 256>>>(d, n, cnt, cnt + 1);
        cudaDeviceSynchronize(); cudaMemcpy(h, cnt, 64, cudaMemcpyDeviceToHost);
        printf("div    : %llu edge pairs, mismatches %llu\n", h[1], h[0]);
    }
    // ---- solver families
    const int nmat = 1 << 20;
    std::vector<float> hm(6 * (size_t)nmat);
    float* dm; cudaMalloc(&dm, hm.size() * 4);
    const char* names[] = {"smoothing-like (sum of 10 weighted outer products)", "kNN covariance of a noisy planar patch", "random symmetric, entries 10^U(-6,6)",
                           "degenerate: diagonal / rank-1 / zero / repeated", "near-rank-1 with tiny perturbations"};
    srand(7);
    std::vector<float> smooth_m;
    for (int fam = 0; fam < 5; ++fam) {
        for (int t = 0; t < nmat; ++t) {
            float M[9] = {0};
            if (fam == 0 || fam == 4) {
                float b[3] = {rnd(), rnd(), rnd()}; float nb = sqrtf(b[0]*b[0]+b[1]*b[1]+b[2]*b[2]) + 1e-9f; for (auto& x : b) x /= nb;
                const float spread = (fam == 0) ? 0.15f : 1e-4f * fabsf(rnd());
                for (int j = 0; j < 10; ++j) {
                    float v[3] = {b[0] + spread * rnd(), b[1] + spread * rnd(), b[2] + spread * rnd()};
                    float nv = sqrtf(v[0]*v[0]+v[1]*v[1]+v[2]*v[2]); for (auto& x : v) x /= nv;
                    float w = 1.0f / (powf(fabsf(rnd()) * 0.3f, 4.f) + 1e-5f);
                    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r*3+c] += (w * v[r]) * v[c];
                }
            } else if (fam == 1) {
                float n[3] = {rnd(), rnd(), rnd()}; float nn = sqrtf(n[0]*n[0]+n[1]*n[1]+n[2]*n[2]) + 1e-9f; for (auto& x : n) x /= nn;
                float mean[3] = {0, 0, 0}; float pts[20][3];
                for (int j = 0; j < 20; ++j) { float p[3] = {0.3f * rnd(), 0.3f * rnd(), 0.3f * rnd()}; float dp = p[0]*n[0]+p[1]*n[1]+p[2]*n[2]; for (int r = 0; r < 3; ++r) { pts[j][r] = p[r] - dp * n[r] + 0.02f * rnd() * n[r]; mean[r] += pts[j][r] / 20.f; } }
                for (int j = 0; j < 20; ++j) for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r*3+c] += (pts[j][r] - mean[r]) * (pts[j][c] - mean[c]) / 20.f;
            } else if (fam == 2) {
                for (int r = 0; r < 3; ++r) for (int c = 0; c <= r; ++c) { float v = rnd() * powf(10.f, 6.f * rnd()); M[r*3+c] = v; M[c*3+r] = v; }
            } else {
                const int kind = t % 6;
                if (kind == 0) { M[0] = rnd(); M[4] = rnd(); M[8] = rnd(); }
                else if (kind == 1) { float v[3] = {rnd(), rnd(), rnd()}; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r*3+c] = v[r] * v[c]; }
                else if (kind == 2) { }
                else if (kind == 3) { M[0] = M[4] = M[8] = rnd(); }
                else if (kind == 4) { M[0] = M[4] = 1.f; M[8] = 2.f; M[3] = M[1] = 1e-20f * rnd(); }
                else { M[0] = 1.f; M[4] = 1.f + 1e-7f * rnd(); M[8] = 1.f; M[6] = M[2] = 1e-12f * rnd(); M[7] = M[5] = rnd() * 1e-3f; }
            }
            hm[6*(size_t)t+0]=M[0]; hm[6*(size_t)t+1]=M[3]; hm[6*(size_t)t+2]=M[4]; hm[6*(size_t)t+3]=M[6]; hm[6*(size_t)t+4]=M[7]; hm[6*(size_t)t+5]=M[8];
        }
        if (fam == 0) smooth_m.assign(hm.begin(), hm.begin() + 6 * 4096);
        cudaMemcpy(dm, hm.data(), hm.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(cnt, 0, 64);
        k_solver<<<(nmat + 127) / 128, 128>>>(dm, nmat, cnt, cnt + 1, cnt + 2);
        cudaDeviceSynchronize(); cudaMemcpy(h, cnt, 64, cudaMemcpyDeviceToHost);
        printf("solver : %-55s %d matrices, mismatches %llu, fallback %.4f %%, mean QR iterations %.2f\n", names[fam], nmat, h[0], 100.0 * h[1] / nmat, (double)h[2] / nmat);
    }
    // ---- latency
    float* d; long long* o; float* sink;
    cudaMalloc(&d, smooth_m.size() * 4); cudaMemcpy(d, smooth_m.data(), smooth_m.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&o, 8); cudaMalloc(&sink, 128);
    for (int lanes : {1, 4, 32}) {
        long long c0, c1;
        k_lat<0><<<1, 32>>>(d, 4096, lanes, 2000, o, sink); cudaDeviceSynchronize();
        k_lat<0><<<1, 32>>>(d, 4096, lanes, 2000, o, sink); cudaDeviceSynchronize(); cudaMemcpy(&c0, o, 8, cudaMemcpyDeviceToHost);
        k_lat<1><<<1, 32>>>(d, 4096, lanes, 2000, o, sink); cudaDeviceSynchronize();
        k_lat<1><<<1, 32>>>(d, 4096, lanes, 2000, o, sink); cudaDeviceSynchronize(); cudaMemcpy(&c1, o, 8, cudaMemcpyDeviceToHost);
        printf("latency: lanes %2d: IEEE %.0f cycles, fast %.0f cycles per dependent solve + normalize\n", lanes, c0 / 2000.0, c1 / 2000.0);
    }
    cudaError_t e = cudaGetLastError();
    printf("cuda: %s\n", cudaGetErrorString(e));
    return 0;
}
