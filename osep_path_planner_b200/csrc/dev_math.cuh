// csrc/dev_math.cuh -- float3 / 3x3 arithmetic of the skeleton path for sm_100a.
//
// Everything the hot path decides discretely (slab tests, similarity threshold, radius
// covers, kNN order) consumes these floats, and the pipeline is chaotic at ulp scale
// (SURVEY.md 7.2 H1), so every routine is a fixed sequence of separately rounded IEEE
// binary32 operations: compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false.
// The sequences follow the libraries the reference calls (Eigen 3.4 fixed-size kernels,
// PCL 1.12 centroid code, FLANN L2_Simple); each routine cites its reference call site in
// /root/reference/osep_skeleton_decomp/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#define OSEP_HD __host__ __device__ __forceinline__

namespace osep {

struct f3 { float x, y, z; };

// std::max / std::min semantics (NaN in the first operand wins), not fmaxf/fminf
OSEP_HD float smax(float a, float b) { return (a < b) ? b : a; }
OSEP_HD float smin(float a, float b) { return (b < a) ? b : a; }
OSEP_HD bool fin(float a) { return fabsf(a) <= FLT_MAX; }       // false for NaN and +-inf
OSEP_HD bool fin3(const f3& a) { return fin(a.x) && fin(a.y) && fin(a.z); }

// Eigen 3-element reductions are a binary tree: x0 + (x1 + x2)   (Redux.h redux_novec_unroller)
OSEP_HD float sum3(float a, float b, float c) { return a + (b + c); }
OSEP_HD float dot3(const f3& a, const f3& b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
OSEP_HD float sqn3(const f3& a) { return sum3(a.x * a.x, a.y * a.y, a.z * a.z); }
OSEP_HD float norm3(const f3& a) { return sqrtf(sqn3(a)); }
OSEP_HD f3 sub3(const f3& a, const f3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
OSEP_HD f3 add3(const f3& a, const f3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
OSEP_HD f3 mul3(const f3& a, float s) { return {a.x * s, a.y * s, a.z * s}; }
OSEP_HD f3 div3(const f3& a, float s) { return {a.x / s, a.y / s, a.z / s}; }
OSEP_HD f3 normalize3(const f3& a) {           // Eigen normalize(): only when squaredNorm() > 0
    const float z = sqn3(a);
    if (z > 0.0f) return div3(a, sqrtf(z));
    return a;
}
OSEP_HD f3 normalize3_dynrow(const f3& a) {    // normalize() of a dynamic-size row (`vset.row(p)`, rosa.cpp:307): squaredNorm summed left to right
    const float z = (a.x * a.x + a.y * a.y) + a.z * a.z;
    if (z > 0.0f) return div3(a, sqrtf(z));
    return a;
}
OSEP_HD float comp(const f3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// FLANN L2_Simple (kd-tree searches at rosa.cpp:99,231,265,372,480,566; gskel.cpp:77,216)
OSEP_HD float dist2(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    return (dx * dx + dy * dy) + dz * dz;
}
OSEP_HD float dist2(const f3& a, const f3& b) { return dist2(a.x, a.y, a.z, b.x, b.y, b.z); }
// canonical neighbour order: ascending (d2, index)
OSEP_HD bool nb_less(float d, int i, float d2, int i2) { return d < d2 || (d == d2 && i < i2); }

// Eigen numext::hypot (positive_real_hypot)
OSEP_HD float eig_hypot(float x, float y) {
    x = fabsf(x); y = fabsf(y);
    const float p = smax(x, y);
    if (p == 0.0f) return 0.0f;
    const float qp = smin(y, x) / p;
    return p * sqrtf(1.0f + qp * qp);
}

// Eigen JacobiRotation<float>::makeGivens (real).  Written branch-uniform (selects instead of a 4-way
// branch) so that the lanes of a warp solving different matrices stay converged; every lane still performs
// exactly the operations of its own case:
//   |p| > |q| : t = q/p, u = +-sqrt(1+t^2) (sign of p), c = 1/u,  s = -t*c        = -(t * (1/u))
//   otherwise : t = p/q, u = +-sqrt(1+t^2) (sign of q), s = -1/u, c = -t*s        =  (t * (1/u))
//   q == 0    : c = sign(p), s = 0          p == 0 : c = 0, s = -sign(q)
OSEP_HD void make_givens(float p, float q, float& c, float& s) {
    const bool A = fabsf(p) > fabsf(q);
    const float num = A ? q : p, den = A ? p : q;
    const float t = num / den;
    float u = sqrtf(1.0f + t * t);
    if (den < 0.0f) u = -u;
    const float inv = 1.0f / u;
    const float tinv = t * inv;
    float cc = A ? inv : tinv;
    float ss = A ? -tinv : -inv;
    if (p == 0.0f) { cc = 0.0f; ss = q < 0.0f ? 1.0f : -1.0f; }
    if (q == 0.0f) { cc = p < 0.0f ? -1.0f : 1.0f; ss = 0.0f; }
    c = cc; s = ss;
}

// Symmetric 3x3 eigen-decomposition = Eigen::SelfAdjointEigenSolver<Matrix3f>::compute
// (iterative path), the solver behind rosa.cpp:305,603,704,751,816; also used for the
// normals (rosa.cpp:97-101) in place of PCL's trigonometric eigen33.
// Input: LOWER triangle (a00,a10,a11,a20,a21,a22).  Output: eigenvalues ascending in ev[],
// eigenvectors in the columns of q (row-major q[r*3+c]).  Register resident: fixed-size
// arrays with compile-time indices only (the k-loops are fully unrolled over start/end).
struct Eig3 { float ev[3]; float q[9]; bool ok; int iters; };

// one Givens step of tridiagonal_qr_step at compile-time position k; `first` = (k == start), `last` = (k == end-1)
template <int k>
OSEP_HD void eig3_qr_rot(float* diag, float* sub, float* q, bool first, bool last, float& x, float& z) {
    float c, s;
    make_givens(x, z, c, s);
    const float sdk = s * diag[k] + c * sub[k];
    const float dkp1 = s * sub[k] + c * diag[k + 1];
    diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
    diag[k + 1] = s * sdk + c * dkp1;
    sub[k] = c * sdk - s * dkp1;
    if (k > 0) { if (!first) sub[k - 1] = c * sub[k - 1] - s * z; }
    x = sub[k];
    if (k < 1) { if (!last) { z = -s * sub[k + 1]; sub[k + 1] = c * sub[k + 1]; } }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float xi = q[r * 3 + k], yi = q[r * 3 + k + 1];
        q[r * 3 + k] = c * xi - s * yi;
        q[r * 3 + k + 1] = s * xi + c * yi;
    }
}

OSEP_HD Eig3 eig3_sym(float a00, float a10, float a11, float a20, float a21, float a22) {
    Eig3 R;
    float scale = fabsf(a00);
    scale = smax(scale, fabsf(a10)); scale = smax(scale, fabsf(a20));
    scale = smax(scale, fabsf(a11)); scale = smax(scale, fabsf(a21));
    scale = smax(scale, fabsf(a22));
    if (scale == 0.0f) scale = 1.0f;
    a00 /= scale; a10 /= scale; a11 /= scale; a20 /= scale; a21 /= scale; a22 /= scale;

    float d0, d1, d2, s0, s1;
    float* q = R.q;
    d0 = a00;
    const float v1norm2 = a20 * a20;
    if (v1norm2 <= FLT_MIN) {
        d1 = a11; d2 = a22; s0 = a10; s1 = a21;
        q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0; q[4] = 1; q[5] = 0; q[6] = 0; q[7] = 0; q[8] = 1;
    } else {
        const float beta = sqrtf(a10 * a10 + v1norm2);
        const float invBeta = 1.0f / beta;
        const float m01 = a10 * invBeta;
        const float m02 = a20 * invBeta;
        const float qq = 2.0f * m01 * a21 + m02 * (a22 - a11);
        d1 = a11 + m02 * qq;
        d2 = a22 - m02 * qq;
        s0 = beta;
        s1 = a21 - m01 * qq;
        q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0; q[4] = m01; q[5] = m02; q[6] = 0; q[7] = m02; q[8] = -m01;
    }
    float diag[3] = {d0, d1, d2};
    float sub[2] = {s0, s1};

    int end = 2, start = 0, iter = 0;
    const float precision_inv = 1.0f / FLT_EPSILON;
    while (end > 0) {
        // deflation test over i in [start, end)
        if (start <= 0 && 0 < end) {
            if (fabsf(sub[0]) < FLT_MIN) sub[0] = 0.0f;
            else { const float sc = precision_inv * sub[0]; if (sc * sc <= (fabsf(diag[0]) + fabsf(diag[1]))) sub[0] = 0.0f; }
        }
        if (start <= 1 && 1 < end) {
            if (fabsf(sub[1]) < FLT_MIN) sub[1] = 0.0f;
            else { const float sc = precision_inv * sub[1]; if (sc * sc <= (fabsf(diag[1]) + fabsf(diag[2]))) sub[1] = 0.0f; }
        }
        if (end == 2 && sub[1] == 0.0f) end = 1;
        if (end == 1 && sub[0] == 0.0f) end = 0;
        if (end <= 0) break;
        iter++;
        if (iter > 90) break;
        start = end - 1;
        if (start == 1 && sub[0] != 0.0f) start = 0;
        // tridiagonal_qr_step(diag, sub, start, end)
        const float dem1 = (end == 2) ? diag[1] : diag[0];
        const float de = (end == 2) ? diag[2] : diag[1];
        const float e = (end == 2) ? sub[1] : sub[0];
        const float td = (dem1 - de) * 0.5f;
        float mu = de;
        if (td == 0.0f) mu -= fabsf(e);
        else if (e != 0.0f) {
            const float e2 = e * e;
            const float h = eig_hypot(td, e);
            if (e2 == 0.0f) mu -= e / ((td + (td > 0.0f ? h : -h)) / e);
            else mu -= e2 / (td + (td > 0.0f ? h : -h));
        }
        float x = ((start == 0) ? diag[0] : diag[1]) - mu;
        float z = (start == 0) ? sub[0] : sub[1];
        // for (k = start; k < end && z != 0; ++k): one code copy per position, predicated, so lanes with
        // different (start, end) stay converged
        if (start == 0 && z != 0.0f) eig3_qr_rot<0>(diag, sub, q, true, end == 1, x, z);
        if (end == 2 && z != 0.0f) eig3_qr_rot<1>(diag, sub, q, start == 1, true, x, z);
    }
    R.ok = iter <= 90;
    R.iters = iter;
    if (R.ok) {
        // selection sort ascending, swapping eigenvector columns (first minimum wins)
        {
            int k = 0; float best = diag[0];
            if (diag[1] < best) { best = diag[1]; k = 1; }
            if (diag[2] < best) { best = diag[2]; k = 2; }
            if (k == 1) { float t = diag[0]; diag[0] = diag[1]; diag[1] = t;
#pragma unroll
                for (int r = 0; r < 3; ++r) { float u = q[r * 3]; q[r * 3] = q[r * 3 + 1]; q[r * 3 + 1] = u; } }
            else if (k == 2) { float t = diag[0]; diag[0] = diag[2]; diag[2] = t;
#pragma unroll
                for (int r = 0; r < 3; ++r) { float u = q[r * 3]; q[r * 3] = q[r * 3 + 2]; q[r * 3 + 2] = u; } }
        }
        if (diag[2] < diag[1]) { float t = diag[1]; diag[1] = diag[2]; diag[2] = t;
#pragma unroll
            for (int r = 0; r < 3; ++r) { float u = q[r * 3 + 1]; q[r * 3 + 1] = q[r * 3 + 2]; q[r * 3 + 2] = u; } }
    }
    R.ev[0] = diag[0] * scale; R.ev[1] = diag[1] * scale; R.ev[2] = diag[2] * scale;
    return R;
}
OSEP_HD f3 eig_col(const Eig3& e, int c) { return {e.q[c], e.q[3 + c], e.q[6 + c]}; }

// Eigen::LLT<Matrix3f> (lower, unblocked) factor + solve  (rosa.cpp:777-779).  m row-major 3x3.
OSEP_HD bool llt3_solve(const float* A, const f3& b, f3& x) {
    float l00 = A[0], l10 = A[3], l11 = A[4], l20 = A[6], l21 = A[7], l22 = A[8];
    if (l00 <= 0.0f) return false;
    l00 = sqrtf(l00);
    l10 /= l00; l20 /= l00;
    float d = l11 - l10 * l10;
    if (d <= 0.0f) return false;
    l11 = sqrtf(d);
    l21 -= l20 * l10;
    l21 /= l11;
    d = l22 - (l20 * l20 + l21 * l21);
    if (d <= 0.0f) return false;
    l22 = sqrtf(d);
    const float y0 = b.x / l00;
    const float y1 = (b.y - l10 * y0) / l11;
    const float y2 = (b.z - (l20 * y0 + l21 * y1)) / l22;
    const float x2 = y2 / l22;
    const float x1 = (y1 - l21 * x2) / l11;
    const float x0 = (y0 - (l10 * x1 + l20 * x2)) / l00;
    x = {x0, x1, x2};
    return true;
}

// Eigen::LDLT<Matrix3f> (lower, pivoted) factor + solve -- the fallback of rosa.cpp:782-785.
// Practically unreachable; kept in local memory (dynamic indices) on purpose.
OSEP_HD bool ldlt3_solve(const float* A, const f3& b, f3& x) {
    float m[3][3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m[r][c] = A[r * 3 + c];
    int tr[3] = {0, 1, 2};
    bool found_zero_pivot = false, ret = true;
    for (int k = 0; k < 3; ++k) {
        int big = k; float bv = fabsf(m[k][k]);
        for (int i = k + 1; i < 3; ++i) if (fabsf(m[i][i]) > bv) { bv = fabsf(m[i][i]); big = i; }
        tr[k] = big;
        if (k != big) {
            for (int c = 0; c < k; ++c) { float t = m[k][c]; m[k][c] = m[big][c]; m[big][c] = t; }
            for (int r = big + 1; r < 3; ++r) { float t = m[r][k]; m[r][k] = m[r][big]; m[r][big] = t; }
            { float t = m[k][k]; m[k][k] = m[big][big]; m[big][big] = t; }
            for (int i = k + 1; i < big; ++i) { float t = m[i][k]; m[i][k] = m[big][i]; m[big][i] = t; }
        }
        const int rs = 3 - k - 1;
        if (k > 0) {
            float temp[2] = {0.f, 0.f};
            for (int c = 0; c < k; ++c) temp[c] = m[c][c] * m[k][c];
            float acc = m[k][0] * temp[0];
            if (k == 2) acc = acc + m[k][1] * temp[1];
            m[k][k] -= acc;
            for (int r = k + 1; r < 3; ++r) {
                float a2 = m[r][0] * temp[0];
                if (k == 2) a2 = a2 + m[r][1] * temp[1];
                m[r][k] -= a2;
            }
        }
        const float akk = m[k][k];
        const bool pivot_is_valid = fabsf(akk) > 0.0f;
        if (k == 0 && !pivot_is_valid) {
            for (int j = 0; j < 3; ++j) { tr[j] = j; for (int r = j + 1; r < 3; ++r) ret = ret && (m[r][j] == 0.0f); }
            break;
        }
        if (rs > 0 && pivot_is_valid) { for (int r = k + 1; r < 3; ++r) m[r][k] /= akk; }
        else if (rs > 0) { for (int r = k + 1; r < 3; ++r) ret = ret && (m[r][k] == 0.0f); }
        if (found_zero_pivot && pivot_is_valid) ret = false;
        else if (!pivot_is_valid) found_zero_pivot = true;
    }
    if (!ret) return false;
    float v[3] = {b.x, b.y, b.z};
    for (int k = 0; k < 3; ++k) { float t = v[k]; v[k] = v[tr[k]]; v[tr[k]] = t; }
    v[1] = v[1] - m[1][0] * v[0];
    v[2] = v[2] - (m[2][0] * v[0] + m[2][1] * v[1]);
    for (int i = 0; i < 3; ++i) { if (fabsf(m[i][i]) > FLT_MIN) v[i] /= m[i][i]; else v[i] = 0.0f; }
    v[1] = v[1] - m[2][1] * v[2];
    v[0] = v[0] - (m[1][0] * v[1] + m[2][0] * v[2]);
    for (int k = 2; k >= 0; --k) { float t = v[k]; v[k] = v[tr[k]]; v[tr[k]] = t; }
    x = {v[0], v[1], v[2]};
    return true;
}

// Cube root for the leaf-size control law (rosa.hpp:108, rosa.cpp:142).  The reference calls
// libm cbrtf, whose last bit is glibc-version dependent and which does not exist on the
// device; this fixed sequence of IEEE binary64 operations returns the correctly rounded
// binary32 cube root (up to astronomically rare double roundings) so the leaf loop can run
// on the GPU without a host round trip.
OSEP_HD float canon_cbrtf(float xf) {
    if (!(xf == xf) || xf == 0.0f || !fin(xf)) return xf;
    const double x = fabs((double)xf);
    union { double d; unsigned long long u; } cv;
    cv.d = x;
    cv.u = cv.u / 3ull + 0x2A9F7893782DA1CEull;
    double y = cv.d;
#pragma unroll 1
    for (int i = 0; i < 6; ++i) {
        const double y2 = y * y;
        y = y - (y2 * y - x) / (3.0 * y2);
    }
    const float r = (float)y;
    return xf < 0.0f ? -r : r;
}

// Fixed-size 3x3 inverse = Eigen Matrix3f::inverse() (cofactors / determinant), used by the
// Kalman gain at lkf_vertex_fuse.hpp:24.  Row-major.
OSEP_HD float cof3(const float* m, int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}
OSEP_HD void mat3_inv(const float* m, float* out) {
    const float c0 = cof3(m, 0, 0), c1 = cof3(m, 1, 0), c2 = cof3(m, 2, 0);
    const float det = sum3(c0 * m[0], c1 * m[3], c2 * m[6]);
    const float invdet = 1.0f / det;
    for (int r = 1; r < 3; ++r) for (int c = 0; c < 3; ++c) out[r * 3 + c] = cof3(m, c, r) * invdet;
    out[0] = c0 * invdet; out[1] = c1 * invdet; out[2] = c2 * invdet;
}
OSEP_HD void mat3_mul(const float* A, const float* B, float* C) {
    float t[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
        t[r * 3 + c] = sum3(A[r * 3] * B[c], A[r * 3 + 1] * B[3 + c], A[r * 3 + 2] * B[6 + c]);
    for (int i = 0; i < 9; ++i) C[i] = t[i];
}

}  // namespace osep
