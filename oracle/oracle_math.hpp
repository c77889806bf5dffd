// oracle/oracle_math.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the third-party arithmetic the reference's skeleton path
// calls (Eigen 3.4.0 / PCL 1.12.1, neither vendored under /root/reference).
// PARITY UNPINNED: the reference holds no tests or golden vectors for these
// boundaries; the semantics below are restated from the
// published algorithms and anchored on the reference's call sites, which are
// cited per function.  Compile with -ffp-contract=off (the reference builds
// with -O3 and no -march / fast-math: osep_skeleton_decomp/CMakeLists.txt:6-9),
// so every + - * / sqrt below is a separately rounded IEEE binary32 operation.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>

namespace orc {

struct V3 {
    float x, y, z;
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};

// Eigen fixed-size-3 reductions (dot / squaredNorm / sum) are unrolled as a
// binary tree: redux(0,3) = f(x0, f(x1, x2))  [Eigen/src/Core/Redux.h,
// redux_novec_unroller: HalfLength = Length/2].
static inline float sum3(float a, float b, float c) { return a + (b + c); }
static inline float dot3(const V3& a, const V3& b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float sqn3(const V3& a) { return sum3(a.x * a.x, a.y * a.y, a.z * a.z); }
static inline float norm3(const V3& a) { return std::sqrt(sqn3(a)); }
static inline V3 sub3(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 add3(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 mul3(const V3& a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 div3(const V3& a, float s) { return {a.x / s, a.y / s, a.z / s}; }
// Eigen normalize(): z = squaredNorm(); if (z > 0) *this /= sqrt(z)   (true division)
static inline V3 normalize3(const V3& a) {
    float z = sqn3(a);
    if (z > 0.0f) return div3(a, std::sqrt(z));
    return a;
}
// The same on a DYNAMIC-size strided block -- `vset.row(p).normalize()` on a MatrixXf (rosa.cpp:307): Eigen's redux has no
// compile-time size to unroll by halving there and runs left to right (Redux.h: DefaultTraversal / NoUnrolling, and the
// too-small-to-vectorise branch of the linear traversal): squaredNorm = (x*x + y*y) + z*z, not x*x + (y*y + z*z).
static inline V3 normalize3_dynrow(const V3& a) {
    float z = (a.x * a.x + a.y * a.y) + a.z * a.z;
    if (z > 0.0f) return div3(a, std::sqrt(z));
    return a;
}
static inline bool finite3(const V3& a) { return std::isfinite(a.x) && std::isfinite(a.y) && std::isfinite(a.z); }

// FLANN L2_Simple squared distance: result=0; result += d*d per dimension in order.
static inline float dist2(const V3& a, const V3& b) {
    float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return (dx * dx + dy * dy) + dz * dz;
}

// Row-major 3x3.
struct M3 {
    float m[9];
    float& operator()(int r, int c) { return m[r * 3 + c]; }
    float operator()(int r, int c) const { return m[r * 3 + c]; }
};

// Eigen numext::hypot (positive_real_hypot) -- MathFunctionsImpl.h
static inline float eig_hypot(float x, float y) {
    x = std::fabs(x); y = std::fabs(y);
    float p = std::max(x, y);
    if (p == 0.0f) return 0.0f;
    float qp = std::min(y, x) / p;
    return p * std::sqrt(1.0f + qp * qp);
}

// Eigen JacobiRotation<float>::makeGivens(p, q) (real case) -- Jacobi.h
static inline void make_givens(float p, float q, float& c, float& s) {
    if (q == 0.0f) { c = p < 0.0f ? -1.0f : 1.0f; s = 0.0f; }
    else if (p == 0.0f) { c = 0.0f; s = q < 0.0f ? 1.0f : -1.0f; }
    else if (std::fabs(p) > std::fabs(q)) {
        float t = q / p; float u = std::sqrt(1.0f + t * t);
        if (p < 0.0f) u = -u;
        c = 1.0f / u; s = -t * c;
    } else {
        float t = p / q; float u = std::sqrt(1.0f + t * t);
        if (q < 0.0f) u = -u;
        s = -1.0f / u; c = -t * s;
    }
}

// Eigen::SelfAdjointEigenSolver<Matrix3f>::compute (iterative path) as used at
// rosa.cpp:305,603,704,751,816.  Reads the LOWER triangle of A.  Eigenvalues
// ascending in ev[0..2]; eigenvectors are the COLUMNS of Q (Q(r,c), c = index
// of eigenvalue).  Returns true when info()==Success.
// Steps: scale by max|coeff| -> closed-form 3x3 Householder tridiagonalisation
// (Tridiagonalization.h, tridiagonalization_inplace_selector<...,3,false>) ->
// implicit symmetric QR with Wilkinson shift (SelfAdjointEigenSolver.h,
// computeFromTridiagonal_impl / tridiagonal_qr_step, maxIterations = 30) ->
// selection-sort ascending with column swaps -> rescale.
static inline bool eig3_sym(const M3& A, float ev[3], M3& Q) {
    float a00 = A(0,0), a10 = A(1,0), a11 = A(1,1), a20 = A(2,0), a21 = A(2,1), a22 = A(2,2);
    float scale = std::fabs(a00);
    // mat.cwiseAbs().maxCoeff() over the lower-triangular copy (column-major visit order;
    // max is order independent except for NaN handling: NaN never wins a '<' visitor).
    scale = std::max(scale, std::fabs(a10)); scale = std::max(scale, std::fabs(a20));
    scale = std::max(scale, std::fabs(a11)); scale = std::max(scale, std::fabs(a21));
    scale = std::max(scale, std::fabs(a22));
    if (scale == 0.0f) scale = 1.0f;
    a00 /= scale; a10 /= scale; a11 /= scale; a20 /= scale; a21 /= scale; a22 /= scale;

    float diag[3], sub[2];
    const float tol = std::numeric_limits<float>::min();
    diag[0] = a00;
    float v1norm2 = a20 * a20;
    if (v1norm2 <= tol) {
        diag[1] = a11; diag[2] = a22; sub[0] = a10; sub[1] = a21;
        Q = {{1,0,0, 0,1,0, 0,0,1}};
    } else {
        float beta = std::sqrt(a10 * a10 + v1norm2);
        float invBeta = 1.0f / beta;
        float m01 = a10 * invBeta;
        float m02 = a20 * invBeta;
        float q = 2.0f * m01 * a21 + m02 * (a22 - a11);
        diag[1] = a11 + m02 * q;
        diag[2] = a22 - m02 * q;
        sub[0] = beta;
        sub[1] = a21 - m01 * q;
        Q = {{1,0,0, 0,m01,m02, 0,m02,-m01}};
    }

    const int n = 3;
    int end = n - 1, start = 0, iter = 0;
    const float considerAsZero = std::numeric_limits<float>::min();
    const float precision_inv = 1.0f / std::numeric_limits<float>::epsilon();
    const int maxIterations = 30;
    while (end > 0) {
        for (int i = start; i < end; ++i) {
            if (std::fabs(sub[i]) < considerAsZero) sub[i] = 0.0f;
            else {
                const float scaled = precision_inv * sub[i];
                if (scaled * scaled <= (std::fabs(diag[i]) + std::fabs(diag[i + 1]))) sub[i] = 0.0f;
            }
        }
        while (end > 0 && sub[end - 1] == 0.0f) end--;
        if (end <= 0) break;
        iter++;
        if (iter > maxIterations * n) break;
        start = end - 1;
        while (start > 0 && sub[start - 1] != 0.0f) start--;
        // tridiagonal_qr_step
        float td = (diag[end - 1] - diag[end]) * 0.5f;
        float e = sub[end - 1];
        float mu = diag[end];
        if (td == 0.0f) mu -= std::fabs(e);
        else if (e != 0.0f) {
            const float e2 = e * e;
            const float h = eig_hypot(td, e);
            if (e2 == 0.0f) mu -= e / ((td + (td > 0.0f ? h : -h)) / e);
            else mu -= e2 / (td + (td > 0.0f ? h : -h));
        }
        float x = diag[start] - mu;
        float z = sub[start];
        for (int k = start; k < end && z != 0.0f; ++k) {
            float c, s;
            make_givens(x, z, c, s);
            float sdk = s * diag[k] + c * sub[k];
            float dkp1 = s * sub[k] + c * diag[k + 1];
            diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
            diag[k + 1] = s * sdk + c * dkp1;
            sub[k] = c * sdk - s * dkp1;
            if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
            x = sub[k];
            if (k < end - 1) { z = -s * sub[k + 1]; sub[k + 1] = c * sub[k + 1]; }
            // Q = Q * G : applyOnTheRight(k, k+1, rot): col_k' = c*col_k - s*col_k1 ; col_k1' = s*col_k + c*col_k1
            for (int r = 0; r < 3; ++r) {
                float xi = Q(r, k), yi = Q(r, k + 1);
                Q(r, k) = c * xi - s * yi;
                Q(r, k + 1) = s * xi + c * yi;
            }
        }
    }
    bool ok = iter <= maxIterations * n;
    if (ok) {
        for (int i = 0; i < n - 1; ++i) {
            int k = 0; float best = diag[i];
            for (int j = 1; j < n - i; ++j) if (diag[i + j] < best) { best = diag[i + j]; k = j; }
            if (k > 0) {
                std::swap(diag[i], diag[k + i]);
                for (int r = 0; r < 3; ++r) std::swap(Q(r, i), Q(r, k + i));
            }
        }
    }
    ev[0] = diag[0] * scale; ev[1] = diag[1] * scale; ev[2] = diag[2] * scale;
    return ok;
}

static inline V3 col3(const M3& Q, int c) { return {Q(0, c), Q(1, c), Q(2, c)}; }

// ---------------------------------------------------------------------------
// pcl::eigen33 smallest-eigenvalue path (PCL 1.12.1 common/impl/eigen.hpp), used
// by NormalEstimation (rosa.cpp:97-101).  FAITHFUL mode only: it calls libm
// atan2/cos/sin, which no GPU can reproduce bit-for-bit; the canonical path
// replaces it by eig3_sym (DESIGN.md "canonical choices").
// ---------------------------------------------------------------------------
static inline void pcl_compute_roots2(float b, float c, float roots[3]) {
    roots[0] = 0.0f;
    float d = b * b - 4.0f * c;
    if (d < 0.0f) d = 0.0f;
    float sd = std::sqrt(d);
    roots[2] = 0.5f * (b + sd);
    roots[1] = 0.5f * (b - sd);
}
static inline void pcl_compute_roots(const M3& m, float roots[3]) {
    float c0 = m(0,0) * m(1,1) * m(2,2) + 2.0f * m(0,1) * m(0,2) * m(1,2)
             - m(0,0) * m(1,2) * m(1,2) - m(1,1) * m(0,2) * m(0,2) - m(2,2) * m(0,1) * m(0,1);
    float c1 = m(0,0) * m(1,1) - m(0,1) * m(0,1) + m(0,0) * m(2,2) - m(0,2) * m(0,2)
             + m(1,1) * m(2,2) - m(1,2) * m(1,2);
    float c2 = m(0,0) + m(1,1) + m(2,2);
    if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) {
        pcl_compute_roots2(c2, c1, roots);
    } else {
        const float s_inv3 = 1.0f / 3.0f;
        const float s_sqrt3 = std::sqrt(3.0f);
        float c2_over_3 = c2 * s_inv3;
        float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
        if (a_over_3 > 0.0f) a_over_3 = 0.0f;
        float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
        float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
        if (q > 0.0f) q = 0.0f;
        float rho = std::sqrt(-a_over_3);
        float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
        float cos_theta = std::cos(theta);
        float sin_theta = std::sin(theta);
        roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
        roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
        roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
        if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
        if (roots[1] >= roots[2]) {
            std::swap(roots[1], roots[2]);
            if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
        }
        if (roots[0] <= 0.0f) pcl_compute_roots2(c2, c1, roots);
    }
}
static inline V3 cross3(const V3& a, const V3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// eigen33(mat, eigenvalue, eigenvector): smallest eigenvalue + its eigenvector.
static inline V3 pcl_eigen33_smallest(const M3& mat, float& eigenvalue) {
    float scale = 0.0f;
    for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(mat.m[i]));
    if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
    M3 s;
    for (int i = 0; i < 9; ++i) s.m[i] = mat.m[i] / scale;
    float roots[3];
    pcl_compute_roots(s, roots);
    eigenvalue = roots[0] * scale;
    s(0,0) -= roots[0]; s(1,1) -= roots[0]; s(2,2) -= roots[0];
    V3 r0{s(0,0), s(0,1), s(0,2)}, r1{s(1,0), s(1,1), s(1,2)}, r2{s(2,0), s(2,1), s(2,2)};
    V3 v0 = cross3(r0, r1), v1 = cross3(r0, r2), v2 = cross3(r1, r2);
    float l0 = sqn3(v0), l1 = sqn3(v1), l2 = sqn3(v2);
    if (l0 >= l1 && l0 >= l2) return div3(v0, std::sqrt(l0));
    if (l1 >= l0 && l1 >= l2) return div3(v1, std::sqrt(l1));
    return div3(v2, std::sqrt(l2));
}

// ---------------------------------------------------------------------------
// Eigen::LLT<Matrix3f> (Lower, unblocked) + solve, as used at rosa.cpp:777-779.
// Returns false when a pivot is <= 0 (info()==NumericalIssue).
// ---------------------------------------------------------------------------
static inline bool llt3_solve(const M3& A, const V3& b, V3& x) {
    float L[3][3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) L[r][c] = A(r, c);
    for (int k = 0; k < 3; ++k) {
        int rs = 3 - k - 1;
        float d = L[k][k];
        if (k == 1) d -= L[1][0] * L[1][0];
        if (k == 2) d -= (L[2][0] * L[2][0] + L[2][1] * L[2][1]);
        if (d <= 0.0f) return false;
        d = std::sqrt(d);
        L[k][k] = d;
        if (k > 0 && rs > 0) {           // A21 -= A20 * A10^T
            for (int r = k + 1; r < 3; ++r) {
                float acc = L[r][0] * L[k][0];
                if (k == 2) acc = acc + L[r][1] * L[k][1];
                L[r][k] -= acc;
            }
        }
        if (rs > 0) for (int r = k + 1; r < 3; ++r) L[r][k] /= d;
    }
    // forward  L y = b
    float y0 = b.x / L[0][0];
    float y1 = (b.y - L[1][0] * y0) / L[1][1];
    float y2 = (b.z - (L[2][0] * y0 + L[2][1] * y1)) / L[2][2];
    // backward L^T x = y
    float x2 = y2 / L[2][2];
    float x1 = (y1 - L[2][1] * x2) / L[1][1];
    float x0 = (y0 - (L[1][0] * x1 + L[2][0] * x2)) / L[0][0];
    x = {x0, x1, x2};
    return true;
}

// Eigen::LDLT<Matrix3f> (Lower, pivoted, unblocked) + solve -- the fallback at
// rosa.cpp:782-785.  Practically unreachable (the LLT pivot test only fails on
// a non-positive pivot of a matrix that is >= 1e-6*trace*I by construction).
static inline bool ldlt3_solve(const M3& A, const V3& b, V3& x) {
    float m[3][3];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m[r][c] = A(r, c);
    int tr[3] = {0, 1, 2};
    bool found_zero_pivot = false, ret = true;
    for (int k = 0; k < 3; ++k) {
        int big = k; float bv = std::fabs(m[k][k]);
        for (int i = k + 1; i < 3; ++i) if (std::fabs(m[i][i]) > bv) { bv = std::fabs(m[i][i]); big = i; }
        tr[k] = big;
        if (k != big) {
            for (int c = 0; c < k; ++c) std::swap(m[k][c], m[big][c]);
            for (int r = big + 1; r < 3; ++r) std::swap(m[r][k], m[r][big]);
            std::swap(m[k][k], m[big][big]);
            for (int i = k + 1; i < big; ++i) { float t = m[i][k]; m[i][k] = m[big][i]; m[big][i] = t; }
        }
        int rs = 3 - k - 1;
        if (k > 0) {
            float temp[2];
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
        float akk = m[k][k];
        bool pivot_is_valid = std::fabs(akk) > 0.0f;
        if (k == 0 && !pivot_is_valid) {
            for (int j = 0; j < 3; ++j) {
                tr[j] = j;
                for (int r = j + 1; r < 3; ++r) ret = ret && (m[r][j] == 0.0f);
            }
            break;
        }
        if (rs > 0 && pivot_is_valid) for (int r = k + 1; r < 3; ++r) m[r][k] /= akk;
        else if (rs > 0) for (int r = k + 1; r < 3; ++r) ret = ret && (m[r][k] == 0.0f);
        if (found_zero_pivot && pivot_is_valid) ret = false;
        else if (!pivot_is_valid) found_zero_pivot = true;
    }
    if (!ret) return false;
    float v[3] = {b.x, b.y, b.z};
    for (int k = 0; k < 3; ++k) std::swap(v[k], v[tr[k]]);                 // P b
    v[1] = v[1] - m[1][0] * v[0];                                         // L^-1 (unit lower)
    v[2] = v[2] - (m[2][0] * v[0] + m[2][1] * v[1]);
    const float tolerance = std::numeric_limits<float>::min();
    for (int i = 0; i < 3; ++i) { if (std::fabs(m[i][i]) > tolerance) v[i] /= m[i][i]; else v[i] = 0.0f; }
    v[1] = v[1] - m[2][1] * v[2];                                         // L^-T
    v[0] = v[0] - (m[1][0] * v[1] + m[2][0] * v[2]);
    for (int k = 2; k >= 0; --k) std::swap(v[k], v[tr[k]]);               // P^T
    x = {v[0], v[1], v[2]};
    return true;
}

// Canonical cube root (DESIGN.md "canonical choices" C5): the reference calls
// libm cbrtf (rosa.hpp:108, rosa.cpp:142), whose last-bit behaviour is glibc
// version dependent.  This routine uses only IEEE binary64 * / + so CPU and GPU
// agree bit-for-bit; it returns the correctly rounded binary32 cube root except
// in astronomically rare double-rounding cases (checked against libm in tests).
static inline float canon_cbrtf(float xf) {
    if (!(xf == xf) || xf == 0.0f || std::isinf(xf)) return xf;
    double x = std::fabs((double)xf);
    // seed from the float exponent: y0 = 2^(e/3) with a linear mantissa guess
    uint64_t bits; std::memcpy(&bits, &x, 8);
    bits = bits / 3 + 0x2A9F7893782DA1CEull;    // classic bit-hack seed, ~5% error
    double y; std::memcpy(&y, &bits, 8);
    for (int i = 0; i < 6; ++i) {               // Newton: y <- y - (y^3 - x)/(3 y^2)
        double y2 = y * y;
        y = y - (y2 * y - x) / (3.0 * y2);
    }
    float r = (float)y;
    return xf < 0.0f ? -r : r;
}

}  // namespace orc
