// oracle/oracle_rosa.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See oracle_rosa.hpp.
// Every function cites the reference lines it restates (paths relative to
// /root/reference/osep_skeleton_decomp/).
#include "oracle_rosa.hpp"
#include <chrono>

namespace orc {

static double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// src/rosa.cpp:36-55 (rosa_run + RUN_STEP, include/rosa.hpp:12-18): stages in
// order, stop at the first one returning false.
bool RosaOracle::run(const float* xyz, int n, int stride) {
    orig.resize(n);
    for (int i = 0; i < n; ++i) orig[i] = {xyz[(size_t)i * stride], xyz[(size_t)i * stride + 1], xyz[(size_t)i * stride + 2]};
    flags = 0; failed_stage = 0;
    skelver.clear(); skelver_sampled.clear(); corresp.clear(); seeds.clear();
    vset_iters.clear(); pset_drosa.clear(); poor_idx.clear(); active_size_last.clear();
    typedef bool (RosaOracle::*Fn)();
    Fn steps[7] = {&RosaOracle::preprocess, &RosaOracle::rosa_init, &RosaOracle::similarity_neighbor_extraction,
                   &RosaOracle::drosa, &RosaOracle::dcrosa, &RosaOracle::vertex_sampling, &RosaOracle::vertex_smooth};
    for (int s = 0; s < 7 && s < opt.stop_after_stage; ++s) {
        double t0 = now_ms();
        bool ok = (this->*steps[s])();
        stage_ms[s + 1] = now_ms() - t0;
        if (!ok) { failed_stage = s + 1; return false; }
    }
    return true;
}

// pcl::VoxelGrid<PointNormal>::applyFilter (PCL 1.12.1 filters/impl/voxel_grid.hpp)
// as configured at src/rosa.cpp:126-128 (defaults: downsample_all_data, no filter
// field, min_points_per_voxel 0).  SURVEY.md Appendix A.1 steps 1-9.
// Canonical choice C3: the (key, index) sort is STABLE, so the per-voxel float
// sums run in ascending point index (PCL's own sort is unstable => unspecified).
int RosaOracle::voxel_grid(float leaf, bool emit, std::vector<V3>& out_p, std::vector<V3>& out_n) {
    const int n = (int)orig.size();
    const float inv = 1.0f / leaf;
    V3 mn = orig[0], mx = orig[0];
    for (int i = 1; i < n; ++i) for (int d = 0; d < 3; ++d) {
        mn[d] = std::min(mn[d], orig[i][d]); mx[d] = std::max(mx[d], orig[i][d]);
    }
    int64_t dx = (int64_t)((mx.x - mn.x) * inv) + 1;
    int64_t dy = (int64_t)((mx.y - mn.y) * inv) + 1;
    int64_t dz = (int64_t)((mx.z - mn.z) * inv) + 1;
    if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) {
        flags |= FLAG_VOXEL_OVERFLOW;          // "Leaf size is too small": output = input
        if (emit) { out_p = orig; out_n = normals_full; vox_keys.clear(); vox_count.clear(); }
        return n;
    }
    int minb[3], maxb[3], divb[3];
    for (int d = 0; d < 3; ++d) {
        minb[d] = (int)std::floor(mn[d] * inv);
        maxb[d] = (int)std::floor(mx[d] * inv);
        divb[d] = maxb[d] - minb[d] + 1;
    }
    const int mul1 = divb[0], mul2 = divb[0] * divb[1];
    std::vector<std::pair<unsigned, int>> kv(n);
    for (int i = 0; i < n; ++i) {
        int i0 = (int)(std::floor(orig[i].x * inv) - (float)minb[0]);
        int i1 = (int)(std::floor(orig[i].y * inv) - (float)minb[1]);
        int i2 = (int)(std::floor(orig[i].z * inv) - (float)minb[2]);
        kv[i] = {(unsigned)(i0 + i1 * mul1 + i2 * mul2), i};
    }
    std::stable_sort(kv.begin(), kv.end(), [](const std::pair<unsigned, int>& a, const std::pair<unsigned, int>& b) { return a.first < b.first; });
    int nvox = 0;
    if (emit) { out_p.clear(); out_n.clear(); vox_keys.clear(); vox_count.clear(); }
    for (int i = 0; i < n;) {
        int j = i + 1;
        while (j < n && kv[j].first == kv[i].first) ++j;
        if (emit) {
            // pcl::CentroidPoint<PointNormal>: AccumulatorXYZ (Vector3f sum, / n) and
            // AccumulatorNormal (Vector4f sum, .normalized()); curvature stays 0.
            V3 sp{0, 0, 0}; float s4[4] = {0, 0, 0, 0};
            for (int t = i; t < j; ++t) {
                const V3& p = orig[kv[t].second]; const V3& nn = normals_full[kv[t].second];
                sp.x += p.x; sp.y += p.y; sp.z += p.z;
                s4[0] += nn.x; s4[1] += nn.y; s4[2] += nn.z; s4[3] += 0.0f;
            }
            const float cnt = (float)(size_t)(j - i);
            out_p.push_back({sp.x / cnt, sp.y / cnt, sp.z / cnt});
            // Vector4f::squaredNorm() is packet-reduced: (a0+a2)+(a1+a3)  [Eigen SSE predux]
            float z = (s4[0] * s4[0] + s4[2] * s4[2]) + (s4[1] * s4[1] + s4[3] * s4[3]);
            V3 nv{s4[0], s4[1], s4[2]};
            if (z > 0.0f) { float r = std::sqrt(z); nv = {s4[0] / r, s4[1] / r, s4[2] / r}; }
            out_n.push_back(nv);
            vox_keys.push_back(kv[i].first); vox_count.push_back(j - i);
        }
        ++nvox; i = j;
    }
    return nvox;
}

// src/rosa.cpp:57-181
bool RosaOracle::preprocess() {
    pts.clear(); nrms.clear(); surf_nbs.clear(); simi_nbs.clear();
    filt_idx.clear(); normals_full.clear(); knn_full.clear(); leaf_hist.clear(); nvox_hist.clear();
    if (orig.empty()) return false;                                     // :63-65

    // distance + finite filter, order preserving (:67-88)
    const float r2 = cfg.pts_dist_lim * cfg.pts_dist_lim;
    std::vector<V3> kept; kept.reserve(orig.size());
    for (size_t i = 0; i < orig.size(); ++i) {
        const V3& p = orig[i];
        const float d2 = p.x * p.x + p.y * p.y + p.z * p.z;
        if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z) && d2 <= r2) { kept.push_back(p); filt_idx.push_back((int)i); }
    }
    orig.swap(kept);
    pcd_size = orig.size();
    if ((int)pcd_size < cfg.min_points) return false;                   // :91-95

    // pcl::NormalEstimation::compute (:97-101; PCL 1.12.1 features/impl/normal_3d.hpp,
    // common/impl/centroid.hpp computeMeanAndCovarianceMatrix)
    const int n = (int)orig.size();
    kd_.build(orig.data(), n);
    normals_full.resize(n);
    knn_k = std::min(cfg.ne_knn, n);
    knn_full.assign((size_t)n * knn_k, -1);
    std::vector<int> ki; std::vector<float> kd;
    for (int i = 0; i < n; ++i) {
        int found = kd_.knn(orig[i], cfg.ne_knn, ki, kd);
        for (int t = 0; t < found; ++t) knn_full[(size_t)i * knn_k + t] = ki[t];
        const float qnan = std::numeric_limits<float>::quiet_NaN();
        if (found < 3) { normals_full[i] = {qnan, qnan, qnan}; continue; }
        float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        const V3 K = orig[ki[0]];                 // shift by the first neighbour
        for (int t = 0; t < found; ++t) {
            const V3& q = orig[ki[t]];
            float x = q.x - K.x, y = q.y - K.y, z = q.z - K.z;
            accu[0] += x * x; accu[1] += x * y; accu[2] += x * z;
            accu[3] += y * y; accu[4] += y * z; accu[5] += z * z;
            accu[6] += x; accu[7] += y; accu[8] += z;
        }
        const float cnt = (float)found;
        for (int t = 0; t < 9; ++t) accu[t] /= cnt;
        M3 C;
        C(0,0) = accu[0] - accu[6] * accu[6];
        C(0,1) = accu[1] - accu[6] * accu[7];
        C(0,2) = accu[2] - accu[6] * accu[8];
        C(1,1) = accu[3] - accu[7] * accu[7];
        C(1,2) = accu[4] - accu[7] * accu[8];
        C(2,2) = accu[5] - accu[8] * accu[8];
        C(1,0) = C(0,1); C(2,0) = C(0,2); C(2,1) = C(1,2);
        V3 nv;
        if (opt.faithful_normals) { float ev; nv = pcl_eigen33_smallest(C, ev); }
        else { float ev[3]; M3 Q; eig3_sym(C, ev, Q); nv = col3(Q, 0); }   // canonical choice C2
        // flipNormalTowardsViewpoint(point, 0,0,0, nx,ny,nz)
        float vx = 0.0f - orig[i].x, vy = 0.0f - orig[i].y, vz = 0.0f - orig[i].z;
        float cos_theta = (vx * nv.x + vy * nv.y + vz * nv.z);
        if (cos_theta < 0.0f) { nv.x *= -1; nv.y *= -1; nv.z *= -1; }
        normals_full[i] = nv;
    }

    // estimate_leaf_from_bbox (include/rosa.hpp:99-110)
    auto cbrt_ = [&](float v) { return opt.libm_cbrt ? std::cbrt(v) : canon_cbrtf(v); };
    V3 mn = orig[0], mx = orig[0];
    for (int i = 1; i < n; ++i) for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], orig[i][d]); mx[d] = std::max(mx[d], orig[i][d]); }
    const float ex = std::max(1e-6f, mx.x - mn.x), ey = std::max(1e-6f, mx.y - mn.y), ez = std::max(1e-6f, mx.z - mn.z);
    const float vol = ex * ey * ez;
    float leaf = (cfg.max_points == 0) ? 0.0f : cbrt_(vol / std::max<float>(1.0, (float)(size_t)cfg.max_points));

    // leaf-adaptation loop (:120-148)
    const int max_iters = 8; const float tol = 0.05f, leaf_min = 1e-4f, leaf_max = 1e4f;
    std::vector<V3> dsp, dsn;
    size_t N = 0; bool converged = false;
    for (int iter = 0; iter < max_iters; ++iter) {
        leaf_used = leaf;
        N = (size_t)voxel_grid(leaf, true, dsp, dsn);
        leaf_hist.push_back(leaf); nvox_hist.push_back((int)N);
        if (N == 0) { leaf *= 0.5f; if (leaf < leaf_min) { leaf = leaf_min; break; } continue; }
        const float ratio_out = (float)N / (float)cfg.max_points;
        if (std::fabs(ratio_out - 1.0f) <= tol) { converged = true; break; }
        leaf *= cbrt_(ratio_out);
        if (!std::isfinite(leaf)) leaf = 0.05f;
        if (leaf < leaf_min) leaf = leaf_min;
        if (leaf > leaf_max) leaf = leaf_max;
    }
    if (!converged) flags |= FLAG_LEAF_NOT_CONVERGED;
    leaf_size = leaf;                                                    // :148 (quirk: may differ from leaf_used)

    // unpack + renormalise (:150-165)
    const size_t M = dsp.size();
    pts = dsp; nrms.resize(M);
    for (size_t i = 0; i < M; ++i) {
        V3 nn = dsn[i];
        float L = norm3(nn);
        if (L > 1e-6f) nn = div3(nn, L); else nn = {0, 0, 0};
        nrms[i] = nn;
    }
    pcd_size = M;
    return true;
}

// src/rosa.cpp:183-199 + create_orthonormal_frame include/rosa.hpp:112-128 (column 0 only)
bool RosaOracle::rosa_init() {
    pset.resize(pcd_size); vset.resize(pcd_size); vvar.assign(pcd_size, 0.0f);
    for (size_t i = 0; i < pcd_size; ++i) {
        pset[i] = pts[i];
        const V3 v = nrms[i];
        const float n = norm3(v);
        if (n == 0.0) { vset[i] = {1, 0, 0}; continue; }
        const V3 z = div3(v, n);
        const bool use_x = (double)std::fabs(z.x) < 0.9;
        const V3 a = use_x ? V3{1, 0, 0} : V3{0, 1, 0};
        const float az = dot3(a, z);
        vset[i] = normalize3(sub3(a, mul3(z, az)));
    }
    return true;
}

// src/rosa.cpp:625-651
float RosaOracle::similarity_metric(const V3& p1, const V3& v1, const V3& p2, const V3& v2, float range_r, float scale) const {
    const V3 d = sub3(p1, p2);
    const float dot12_raw = dot3(v1, v2);
    const float dot12 = std::max(0.0f, std::min(1.0f, dot12_raw));
    if (dot12 == 0.0f) return 0.0f;
    const float proj1 = dot3(d, v1);
    const float proj2 = -dot3(d, v2);
    const float s1 = scale * proj1, s2 = scale * proj2;
    float dist1 = norm3(add3(d, mul3(v1, s1))) / range_r;
    float dist2_ = norm3(sub3(d, mul3(v2, s2))) / range_r;
    auto bump = [](float t) {
        if (t > 1.0f) return 0.0f;
        const float t2 = t * t; const float t3 = t2 * t;
        return 2.0f * t3 - 3.0f * t2 + 1.0f;
    };
    const float k1 = bump(dist1); if (k1 == 0.0f) return 0.0f;
    const float k2 = bump(dist2_); if (k2 == 0.0f) return 0.0f;
    const float align2 = dot12 * dot12;
    return std::min(k1, k2) * align2;
}

// src/rosa.cpp:201-256
bool RosaOracle::similarity_neighbor_extraction() {
    simi_nbs.assign(pcd_size, {});
    if (pts.empty()) return false;
    kd_.build(pts.data(), (int)pts.size());
    const float radius_r = 10.0f * leaf_size;
    const float th_sim = 0.1f * leaf_size;
    const float r2 = radius_r * radius_r;     // KdTreeFLANN::radiusSearch squares in double and narrows: == float product
    std::vector<int> ni; std::vector<float> nd;
    for (size_t i = 0; i < pcd_size; ++i) {
        const V3& P1 = pts[i]; const V3& N1 = nrms[i];
        if (!finite3(N1)) continue;
        const V3 v1 = normalize3(N1);
        kd_.radius(P1, r2, ni, nd);
        std::vector<int> keep; keep.reserve(ni.size());
        for (int idx : ni) {
            if (idx == (int)i) continue;
            const V3& N2 = nrms[idx];
            if (!finite3(N2)) continue;
            const V3 v2 = normalize3(N2);
            const float w = similarity_metric(P1, v1, pts[idx], v2, radius_r);
            if (w > th_sim) keep.push_back(idx);
        }
        simi_nbs[i].swap(keep);
    }
    return true;
}

// src/rosa.cpp:653-683.  The reachable SET is order independent; faithful mode returns
// the reference's DFS visit order, canonical mode (C4) returns it sorted ascending.
void RosaOracle::compute_active_samples(int seed, const V3& p, const V3& n, std::vector<int>& out) {
    out.clear();
    state_.assign(pcd_size, 0);
    stack_.clear();
    auto in_slab = [&](int i) -> bool {
        const V3& q = pts[i];
        float d = n.x * (p.x - q.x) + n.y * (p.y - q.y) + n.z * (p.z - q.z);
        return std::fabs(d) < leaf_size;
    };
    if (!in_slab(seed)) return;
    state_[seed] = 1; stack_.push_back(seed);
    while (!stack_.empty()) {
        int cur = stack_.back(); stack_.pop_back();
        state_[cur] = 2; out.push_back(cur);
        for (int nb : simi_nbs[cur]) if (state_[nb] == 0 && in_slab(nb)) { state_[nb] = 1; stack_.push_back(nb); }
    }
    if (!opt.faithful_order) std::sort(out.begin(), out.end());
}

// Canonical choice C4: float sums over an active set.  The reference adds in DFS visit
// order (faithful mode).  Canonical: "lane" l = (index >> 5) & 31 owns the members of
// its 32-index words and adds them in ascending index; the 32 partials are combined by
// an xor butterfly (offsets 16,8,4,2,1) -- exactly what one warp does on the GPU.
template <int K, class F>
void RosaOracle::reduce_active(const std::vector<int>& idxs, F contrib, float out[K]) {
    float c[K];
    if (opt.faithful_order) {
        for (int k = 0; k < K; ++k) out[k] = 0.0f;
        for (int i : idxs) { contrib(i, c); for (int k = 0; k < K; ++k) out[k] += c[k]; }
        return;
    }
    float part[32][K];
    for (int l = 0; l < 32; ++l) for (int k = 0; k < K; ++k) part[l][k] = 0.0f;
    for (int i : idxs) { contrib(i, c); const int l = (i >> 5) & 31; for (int k = 0; k < K; ++k) part[l][k] += c[k]; }
    for (int off = 16; off >= 1; off >>= 1) {
        float nxt[32][K];
        for (int l = 0; l < 32; ++l) for (int k = 0; k < K; ++k) nxt[l][k] = part[l][k] + part[l ^ off][k];
        std::memcpy(part, nxt, sizeof(part));
    }
    for (int k = 0; k < K; ++k) out[k] = part[0][k];
}

// unit normal as written at src/rosa.cpp:693-696,739-742 (double literal 1e-20) and
// :716-718,:340-341 (float literal 1e-20f)
static inline V3 unit_normal_d(const V3& v) {
    float s2 = sqn3(v);
    const float inv_len = ((double)s2 > 1e-20) ? 1.0f / std::sqrt(s2) : 0.0f;
    return mul3(v, inv_len);
}
static inline V3 unit_normal_f(const V3& v) {
    float s2 = sqn3(v);
    const float inv_len = (s2 > 1e-20f) ? 1.0f / std::sqrt(s2) : 0.0f;
    return mul3(v, inv_len);
}

// normal covariance shared by src/rosa.cpp:685-706 and :734-753
static inline M3 cov_from_sums(const float s[9], float inv_m) {
    // s: sum_v[3], sum_outer lower (xx, yx, yy, zx, zy, zz)
    const V3 mu{s[0] * inv_m, s[1] * inv_m, s[2] * inv_m};
    M3 C;
    C(0,0) = s[3] * inv_m - mu.x * mu.x;
    C(1,0) = s[4] * inv_m - mu.y * mu.x;
    C(1,1) = s[5] * inv_m - mu.y * mu.y;
    C(2,0) = s[6] * inv_m - mu.z * mu.x;
    C(2,1) = s[7] * inv_m - mu.z * mu.y;
    C(2,2) = s[8] * inv_m - mu.z * mu.z;
    C(0,1) = C(1,0); C(0,2) = C(2,0); C(1,2) = C(2,1);
    return C;
}

// src/rosa.cpp:685-706
V3 RosaOracle::compute_symmetrynormal(const std::vector<int>& idxs) {
    const int m = (int)idxs.size();
    if (m == 0) return {1, 0, 0};
    float s[9];
    reduce_active<9>(idxs, [&](int i, float* c) {
        const V3 vn = unit_normal_d(nrms[i]);
        c[0] = vn.x; c[1] = vn.y; c[2] = vn.z;
        c[3] = vn.x * vn.x; c[4] = vn.y * vn.x; c[5] = vn.y * vn.y;
        c[6] = vn.z * vn.x; c[7] = vn.z * vn.y; c[8] = vn.z * vn.z;
    }, s);
    const float inv_m = 1.0f / (float)m;
    M3 C = cov_from_sums(s, inv_m);
    float ev[3]; M3 Q; eig3_sym(C, ev, Q);
    return col3(Q, 0);
}

// src/rosa.cpp:708-732
float RosaOracle::symmnormal_variance(const V3& symm, const std::vector<int>& idxs) {
    const int m = (int)idxs.size();
    if (m == 0) return 0.0f;
    float s[2];
    reduce_active<2>(idxs, [&](int i, float* c) {
        const V3& v = nrms[i];
        const float s2 = sqn3(v);
        const float inv_len = (s2 > 1e-20f) ? 1.0f / std::sqrt(s2) : 0.0f;
        const float a = dot3(v, symm) * inv_len;
        c[0] = a; c[1] = a * a;
    }, s);
    const float inv_m = 1.0f / (float)m;
    const float mu = s[0] * inv_m;
    float var = s[1] * inv_m - mu * mu;
    if (m > 1) var *= (float)m / (float)(m - 1);
    return var;
}

// src/rosa.cpp:734-753
V3 RosaOracle::cov_eigs_from_normals(const std::vector<int>& idxs) {
    float s[9];
    reduce_active<9>(idxs, [&](int i, float* c) {
        const V3 vn = unit_normal_d(nrms[i]);
        c[0] = vn.x; c[1] = vn.y; c[2] = vn.z;
        c[3] = vn.x * vn.x; c[4] = vn.y * vn.x; c[5] = vn.y * vn.y;
        c[6] = vn.z * vn.x; c[7] = vn.z * vn.y; c[8] = vn.z * vn.z;
    }, s);
    const float inv_m = 1.0f / (float)std::max<size_t>(1, idxs.size());
    M3 C = cov_from_sums(s, inv_m);
    float ev[3]; M3 Q; eig3_sym(C, ev, Q);
    return {ev[0], ev[1], ev[2]};
}

// src/rosa.cpp:755-791
V3 RosaOracle::closest_projection_point(const std::vector<int>& idxs) {
    if (idxs.empty()) return {1e8f, 1e8f, 1e8f};
    float s[12];   // M row-major (9) then B (3)
    reduce_active<12>(idxs, [&](int i, float* c) {
        const V3& p = pts[i]; const V3& v = nrms[i];
        const float s2 = sqn3(v);
        const float inv_len = (s2 > 1e-20f) ? (float)(1.0 / (double)std::sqrt(s2)) : 0.0f;   // :765 double literal
        const V3 vn = mul3(v, inv_len);
        float P[9];
        for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) P[r * 3 + cc] = (r == cc ? 1.0f : 0.0f) - vn[r] * vn[cc];
        for (int t = 0; t < 9; ++t) c[t] = P[t];
        for (int r = 0; r < 3; ++r) c[9 + r] = sum3(P[r * 3] * p.x, P[r * 3 + 1] * p.y, P[r * 3 + 2] * p.z);
    }, s);
    M3 M; for (int t = 0; t < 9; ++t) M.m[t] = s[t];
    V3 B{s[9], s[10], s[11]};
    const float tr = sum3(M(0,0), M(1,1), M(2,2));
    const float reg = 1e-6f * std::max(tr, 1e-6f);
    M(0,0) += reg; M(1,1) += reg; M(2,2) += reg;
    V3 X;
    if (llt3_solve(M, B, X)) return X;
    if (ldlt3_solve(M, B, X)) return X;
    return {1e8f, 1e8f, 1e8f};
}

// src/rosa.cpp:258-380
bool RosaOracle::drosa() {
    const int M = (int)pcd_size;
    surf_nbs.assign(M, {});
    kd_.build(pts.data(), M);
    { std::vector<int> ki; std::vector<float> kd;
      for (int i = 0; i < M; ++i) { kd_.knn(pts[i], cfg.nb_knn, ki, kd); surf_nbs[i] = ki; } }    // :261-267

    std::vector<V3> vnew(M, V3{0, 0, 0});
    std::vector<int> active;
    for (int it = 0; it < cfg.niter_drosa; ++it) {                                                // :274-309
        for (int p = 0; p < M; ++p) {
            compute_active_samples(p, pset[p], vset[p], active);
            if (!active.empty()) {
                V3 nv = compute_symmetrynormal(active);
                vnew[p] = nv;
                vvar[p] = symmnormal_variance(nv, active);
            } else vvar[p] = 0.0f;
        }
        vset = vnew;
        const float eps = 1e-5;
        for (int p = 0; p < M; ++p) { float a = vvar[p] * vvar[p]; a = a * a; vvar[p] = 1.0f / (a + eps); }   // :291-292
        for (int p = 0; p < M; ++p) {                                                             // :295-308 (in place)
            const auto& snb = surf_nbs[p];
            if (snb.empty()) continue;
            M3 Mm; for (int t = 0; t < 9; ++t) Mm.m[t] = 0.0f;
            for (int j : snb) {
                const V3 v = vset[j]; const float w = vvar[j];
                const V3 wv = mul3(v, w);      // Eigen binds the scalar to the lhs of the outer product: (w*v) * v^T
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Mm(r, c) += wv[r] * v[c];
            }
            float ev[3]; M3 Q; eig3_sym(Mm, ev, Q);
            vset[p] = normalize3_dynrow(col3(Q, 2));      // :306-307 (row of a MatrixXf: sequential squaredNorm)
        }
        vset_iters.insert(vset_iters.end(), vset.begin(), vset.end());
    }

    // position step (:311-365)
    const float PLANAR_TH = 0.1f;
    std::vector<V3> goodCenters;
    std::vector<V3> tmp_pt(M, V3{0, 0, 0});       // tmp_pt_->resize(M): default PointXYZ = (0,0,0)
    std::vector<int> poorIdx;
    active_size_last.assign(M, 0);
    for (int p = 0; p < M; ++p) {
        const V3 pp = pset[p];
        compute_active_samples(p, pp, vset[p], active);
        active_size_last[p] = (int)active.size();
        if (active.empty()) { poorIdx.push_back(p); continue; }
        const V3 evals = cov_eigs_from_normals(active);
        const float eps = 1e-6f;
        const float planar_score = (evals.y - evals.x) / std::max(evals.z, eps);
        V3 center;
        if (planar_score < PLANAR_TH) {
            float s[6];
            reduce_active<6>(active, [&](int i, float* c) {
                const V3 vn = unit_normal_f(nrms[i]);
                c[0] = pts[i].x; c[1] = pts[i].y; c[2] = pts[i].z; c[3] = vn.x; c[4] = vn.y; c[5] = vn.z;
            }, s);
            const float inv_m = 1.0f / (float)active.size();
            const V3 mean_p{s[0] * inv_m, s[1] * inv_m, s[2] * inv_m};
            const V3 mean_n{s[3] * inv_m, s[4] * inv_m, s[5] * inv_m};
            center = sub3(mean_p, mul3(mean_n, 0.5f * cfg.max_proj_range));
        } else center = closest_projection_point(active);
        const bool finite = finite3(center);
        const bool in_range = norm3(sub3(center, pp)) < cfg.max_proj_range;
        if (finite && in_range) { pset[p] = center; tmp_pt[p] = pts[p]; goodCenters.push_back(center); }
        else poorIdx.push_back(p);
    }
    // poor-point fix-up (:367-376).  BUG in the reference, restated as written: gid is a slot
    // index of tmp_pt (size M) but indexes the COMPACTED goodCenters.  Out-of-bounds (UB in the
    // reference) is defined here as "leave pset unchanged and raise FLAG_POOR_FIXUP_OOB".
    kd_.build(tmp_pt.data(), M);
    { std::vector<int> ki; std::vector<float> kd;
      for (int poor : poorIdx) {
          if (kd_.knn(pts[poor], 1, ki, kd) > 0) {
              const int gid = ki[0];
              if (gid < (int)goodCenters.size()) pset[poor] = goodCenters[gid];
              else flags |= FLAG_POOR_FIXUP_OOB;
          }
      } }
    poor_idx = poorIdx;
    pset_drosa = pset;
    return true;
}

// src/rosa.cpp:793-841
bool RosaOracle::local_line_fit(const std::vector<int>& nb, V3& mean_out, V3& dir_out, float& conf_out, int min_nb) {
    const int m = (int)nb.size();
    const float EPS = 1e-6f;
    if (m < min_nb) { mean_out = {0, 0, 0}; dir_out = {1, 0, 0}; conf_out = 0.0f; return false; }
    V3 mean{0, 0, 0};
    for (int j : nb) mean = add3(mean, pset[j]);
    mean = div3(mean, (float)m);
    M3 C; for (int t = 0; t < 9; ++t) C.m[t] = 0.0f;
    for (int j : nb) {
        const V3 q = sub3(pset[j], mean);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C(r, c) += q[r] * q[c];
    }
    for (int t = 0; t < 9; ++t) C.m[t] /= (float)m;
    float ev[3]; M3 Q;
    if (!eig3_sym(C, ev, Q)) { mean_out = mean; dir_out = {1, 0, 0}; conf_out = 0.0f; return false; }
    const float denom = std::max(EPS, sum3(ev[0], ev[1], ev[2]));
    mean_out = mean;
    dir_out = col3(Q, 2);
    const float nrm = norm3(dir_out);
    if (nrm > EPS) dir_out = div3(dir_out, nrm); else dir_out = {1, 0, 0};
    conf_out = ev[2] / denom;
    return true;
}

// src/rosa.cpp:382-422
bool RosaOracle::dcrosa() {
    const float CONF_TH = 0.5f; const int MIN_NB = 3;
    const int M = (int)pcd_size;
    std::vector<V3> newp(M);
    for (int it = 0; it < cfg.niter_dcrosa; ++it) {
        for (int i = 0; i < M; ++i) {
            const auto& nb = simi_nbs[i];
            if (nb.empty()) { newp[i] = pset[i]; continue; }
            V3 acc{0, 0, 0};
            for (int j : nb) acc = add3(acc, pset[j]);
            newp[i] = div3(acc, (float)nb.size());
        }
        pset.swap(newp);
        for (int i = 0; i < M; ++i) {
            const auto& nb = simi_nbs[i];
            newp[i] = pset[i];
            if ((int)nb.size() < MIN_NB) continue;
            V3 mean, dir; float conf;
            if (local_line_fit(nb, mean, dir, conf, MIN_NB) && conf > CONF_TH) {
                const V3 x = pset[i];
                const float t = dot3(dir, sub3(x, mean));
                newp[i] = add3(mean, mul3(dir, t));
            }
        }
        pset.swap(newp);
    }
    return true;
}

// src/rosa.cpp:424-532
bool RosaOracle::vertex_sampling() {
    if (pcd_size == 0) { skelver.clear(); return false; }
    std::vector<V3> cloud;
    for (size_t i = 0; i < pcd_size; ++i) if (finite3(pset[i])) cloud.push_back(pset[i]);
    if (cloud.size() != pcd_size) flags |= FLAG_NONFINITE_PSET;
    pcd_size = cloud.size();                                             // :448 (shrinks; later rows use compacted indices)
    const int Mc = (int)pcd_size;
    const float R = leaf_size;
    std::vector<char> covered(Mc, 0);
    corresp.assign(Mc, -1);
    std::vector<float> mind2(Mc, std::numeric_limits<float>::infinity());
    skelver.clear(); seeds.clear();
    kd_.build(cloud.data(), Mc);
    std::vector<int> ix; std::vector<float> d2s;
    int num_covered = Mc;
    const float R2 = R * R;
    for (size_t start = 0; start < (size_t)Mc && num_covered > 0;) {
        int seed = -1;
        for (int i = 0; i < Mc; ++i) if (!covered[i]) { seed = i; start = i + 1; break; }
        if (seed == -1) break;
        const int vid = (int)skelver.size();
        skelver.push_back(pset[seed]);                                   // :475 (uncompacted pset, compacted index)
        seeds.push_back(seed);
        const int found = kd_.radius(cloud[seed], R2, ix, d2s);
        if (found <= 0) { covered[seed] = 1; num_covered -= 1; mind2[seed] = 0.0f; corresp[seed] = vid; continue; }
        for (int j = 0; j < found; ++j) {
            const int idx = ix[j]; const float d2 = d2s[j];
            if (d2 < mind2[idx]) { mind2[idx] = d2; corresp[idx] = vid; }
            if (!covered[idx]) { covered[idx] = 1; num_covered -= 1; }
        }
    }
    // recentering (:506-530)
    const int V = (int)skelver.size();
    std::vector<V3> sum(V, V3{0, 0, 0}); std::vector<int> cnt(V, 0);
    for (int i = 0; i < Mc; ++i) { int vid = corresp[i]; if (vid >= 0) { sum[vid] = add3(sum[vid], pts[i]); cnt[vid]++; } }
    const float a = cfg.alpha_recenter;
    const float b = (float)(1.0 - (double)a);            // :529 double scalar narrowed before the float multiply
    for (int v = 0; v < V; ++v) {
        if (cnt[v] == 0) continue;
        const V3 ec = div3(sum[v], (float)cnt[v]);
        skelver[v] = add3(mul3(skelver[v], a), mul3(ec, b));
    }
    skelver_sampled = skelver;
    return true;
}

// src/rosa.cpp:534-620
bool RosaOracle::vertex_smooth() {
    int V = (int)skelver.size();
    if (V == 0) return false;
    std::vector<V3> tmp;
    for (int i = 0; i < V; ++i) if (finite3(skelver[i])) tmp.push_back(skelver[i]);
    V = (int)tmp.size();
    kd_.build(tmp.data(), V);
    std::vector<std::vector<int>> neighbors(V);
    { std::vector<int> ix; std::vector<float> d2;
      const float r2 = cfg.radius_smooth * cfg.radius_smooth;
      for (int i = 0; i < V; ++i) {
          if (kd_.radius(tmp[i], r2, ix, d2) > 0) {
              if (ix.empty() || ix[0] != i) ix.insert(ix.begin(), i);
              neighbors[i] = ix;
          } else neighbors[i] = {i};
      } }
    std::vector<V3> cur = skelver, next = skelver;
    const float a = cfg.alpha_recenter;
    for (int it = 0; it < cfg.niter_smooth; ++it) {
        for (int i = 0; i < V; ++i) {
            const auto& nbr = neighbors[i];
            const int n = (int)nbr.size();
            if (n <= 1) { next[i] = cur[i]; continue; }
            V3 mean{0, 0, 0};
            for (int j : nbr) mean = add3(mean, cur[j]);
            mean = div3(mean, (float)n);
            M3 C; for (int t = 0; t < 9; ++t) C.m[t] = 0.0f;
            for (int j : nbr) {
                const V3 d = sub3(cur[j], mean);
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C(r, c) += d[r] * d[c];
            }
            for (int t = 0; t < 9; ++t) C.m[t] /= (float)n;
            float ev[3]; M3 Q; eig3_sym(C, ev, Q);
            const V3 dir = col3(Q, 2);
            const V3 p = cur[i];
            const V3 delta = sub3(p, mean);
            const V3 proj = add3(mean, mul3(dir, dot3(dir, delta)));
            next[i] = add3(mul3(p, a), mul3(proj, 1.0f - a));
        }
        cur.swap(next);
    }
    skelver = cur;
    return true;
}

}  // namespace orc
