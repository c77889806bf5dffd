// oracle/oracle_capi.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// extern "C" surface of the CPU oracle for ctypes (tests/, __graft_entry__.smoke(),
// bench.py cpu_baseline / --impl reference only).
#include "oracle_rosa.hpp"
#include "oracle_gskel.hpp"
#include <string>
#include <cstring>

using namespace orc;

namespace {
template <class T> long put(const std::vector<T>& v, void* dst, long cap_bytes) {
    const long need = (long)(v.size() * sizeof(T));
    if (dst && cap_bytes >= need && need > 0) std::memcpy(dst, v.data(), need);
    return (long)v.size();
}
long put_csr_off(const std::vector<std::vector<int>>& a, void* dst, long cap) {
    std::vector<int> off(a.size() + 1, 0);
    for (size_t i = 0; i < a.size(); ++i) off[i + 1] = off[i] + (int)a[i].size();
    return put(off, dst, cap);
}
long put_csr_idx(const std::vector<std::vector<int>>& a, void* dst, long cap) {
    std::vector<int> flat;
    for (auto& r : a) flat.insert(flat.end(), r.begin(), r.end());
    return put(flat, dst, cap);
}
}  // namespace

extern "C" {

struct orc_rosa_config { int max_points, min_points; float pts_dist_lim; int ne_knn, nb_knn; float max_proj_range;
                         int niter_drosa, niter_dcrosa, niter_smooth; float alpha_recenter, radius_smooth; };
struct orc_rosa_opts { int faithful_normals, faithful_order, libm_cbrt, stop_after_stage; };

void* orc_rosa_create(const orc_rosa_config* c, const orc_rosa_opts* o) {
    auto* r = new RosaOracle();
    if (c) {
        r->cfg.max_points = c->max_points; r->cfg.min_points = c->min_points; r->cfg.pts_dist_lim = c->pts_dist_lim;
        r->cfg.ne_knn = c->ne_knn; r->cfg.nb_knn = c->nb_knn; r->cfg.max_proj_range = c->max_proj_range;
        r->cfg.niter_drosa = c->niter_drosa; r->cfg.niter_dcrosa = c->niter_dcrosa; r->cfg.niter_smooth = c->niter_smooth;
        r->cfg.alpha_recenter = c->alpha_recenter; r->cfg.radius_smooth = c->radius_smooth;
    }
    if (o) { r->opt.faithful_normals = o->faithful_normals; r->opt.faithful_order = o->faithful_order;
             r->opt.libm_cbrt = o->libm_cbrt; r->opt.stop_after_stage = o->stop_after_stage > 0 ? o->stop_after_stage : 7; }
    return r;
}
void orc_rosa_destroy(void* h) { delete (RosaOracle*)h; }
// returns 0 on success, else the failing stage 1..7
int orc_rosa_run(void* h, const float* xyz, int n, int stride_floats) {
    auto* r = (RosaOracle*)h;
    r->run(xyz, n, stride_floats);
    return r->failed_stage;
}
float orc_rosa_leaf_size(void* h) { return ((RosaOracle*)h)->leaf_size; }
float orc_rosa_leaf_used(void* h) { return ((RosaOracle*)h)->leaf_used; }
unsigned orc_rosa_flags(void* h) { return ((RosaOracle*)h)->flags; }
double orc_rosa_stage_ms(void* h, int stage) { return (stage >= 1 && stage <= 7) ? ((RosaOracle*)h)->stage_ms[stage] : 0.0; }

// Tap fetch: returns the element count (floats for V3 taps count 3 per row); copies when dst != NULL and fits.
long orc_rosa_get(void* h, const char* name, void* dst, long cap_bytes) {
    auto* r = (RosaOracle*)h;
    const std::string s(name);
    if (s == "filtered") return put(r->orig, dst, cap_bytes);
    if (s == "filt_idx") return put(r->filt_idx, dst, cap_bytes);
    if (s == "normals_full") return put(r->normals_full, dst, cap_bytes);
    if (s == "knn_full") return put(r->knn_full, dst, cap_bytes);
    if (s == "leaf_hist") return put(r->leaf_hist, dst, cap_bytes);
    if (s == "nvox_hist") return put(r->nvox_hist, dst, cap_bytes);
    if (s == "vox_keys") return put(r->vox_keys, dst, cap_bytes);
    if (s == "vox_count") return put(r->vox_count, dst, cap_bytes);
    if (s == "pts") return put(r->pts, dst, cap_bytes);
    if (s == "nrms") return put(r->nrms, dst, cap_bytes);
    if (s == "surf_off") return put_csr_off(r->surf_nbs, dst, cap_bytes);
    if (s == "surf_idx") return put_csr_idx(r->surf_nbs, dst, cap_bytes);
    if (s == "simi_off") return put_csr_off(r->simi_nbs, dst, cap_bytes);
    if (s == "simi_idx") return put_csr_idx(r->simi_nbs, dst, cap_bytes);
    if (s == "vset") return put(r->vset, dst, cap_bytes);
    if (s == "vset_iters") return put(r->vset_iters, dst, cap_bytes);
    if (s == "vvar") return put(r->vvar, dst, cap_bytes);
    if (s == "pset") return put(r->pset, dst, cap_bytes);
    if (s == "pset_drosa") return put(r->pset_drosa, dst, cap_bytes);
    if (s == "poor_idx") return put(r->poor_idx, dst, cap_bytes);
    if (s == "active_size") return put(r->active_size_last, dst, cap_bytes);
    if (s == "corresp") return put(r->corresp, dst, cap_bytes);
    if (s == "seeds") return put(r->seeds, dst, cap_bytes);
    if (s == "skelver_sampled") return put(r->skelver_sampled, dst, cap_bytes);
    if (s == "skelver") return put(r->skelver, dst, cap_bytes);
    return -1;
}

// ---- stateless graph (G3-G5)
// Outputs: rng CSR (graph_adj adjacency with duplicates), MST edges (u,v) in Kruskal order, MST CSR, type.
void* orc_graph_build(const float* xyz, int G, int stride, float fuse_dist_th) {
    std::vector<V3> pos(G);
    for (int i = 0; i < G; ++i) pos[i] = {xyz[(size_t)i * stride], xyz[(size_t)i * stride + 1], xyz[(size_t)i * stride + 2]};
    return new GraphResult(build_graph(pos, fuse_dist_th));
}
void orc_graph_destroy(void* g) { delete (GraphResult*)g; }
long orc_graph_get(void* g, const char* name, void* dst, long cap_bytes) {
    auto* r = (GraphResult*)g;
    const std::string s(name);
    if (s == "rng_off") return put_csr_off(r->adj_rng, dst, cap_bytes);
    if (s == "rng_idx") return put_csr_idx(r->adj_rng, dst, cap_bytes);
    if (s == "adj_off") return put_csr_off(r->adj, dst, cap_bytes);
    if (s == "adj_idx") return put_csr_idx(r->adj, dst, cap_bytes);
    if (s == "joints") return put(r->joints, dst, cap_bytes);
    if (s == "leafs") return put(r->leafs, dst, cap_bytes);
    if (s == "type") return put(r->type, dst, cap_bytes);
    if (s == "mst_edges") {
        std::vector<int> e; for (auto& x : r->mst_edges) { e.push_back(x.u); e.push_back(x.v); }
        return put(e, dst, cap_bytes);
    }
    if (s == "mst_w") { std::vector<float> w; for (auto& x : r->mst_edges) w.push_back(x.w); return put(w, dst, cap_bytes); }
    return -1;
}

// ---- stateful GSkel (G1-G10)
struct orc_gskel_config { float gnd_th, fuse_dist_th, fuse_conf_th, lkf_pn, lkf_mn; int max_obs_wo_conf, niter_smooth_vertex;
                          float vertex_smooth_coef; int min_branch_length; };
void* orc_gskel_create(const orc_gskel_config* c) {
    auto* g = new GSkelOracle();
    if (c) { g->cfg.gnd_th = c->gnd_th; g->cfg.fuse_dist_th = c->fuse_dist_th; g->cfg.fuse_conf_th = c->fuse_conf_th;
             g->cfg.lkf_pn = c->lkf_pn; g->cfg.lkf_mn = c->lkf_mn; g->cfg.max_obs_wo_conf = c->max_obs_wo_conf;
             g->cfg.niter_smooth_vertex = c->niter_smooth_vertex; g->cfg.vertex_smooth_coef = c->vertex_smooth_coef;
             g->cfg.min_branch_length = c->min_branch_length; }
    return g;
}
void orc_gskel_destroy(void* h) { delete (GSkelOracle*)h; }
int orc_gskel_run(void* h, const float* xyz, int n, int stride) { auto* g = (GSkelOracle*)h; g->run(xyz, n, stride); return g->failed_stage; }
long orc_gskel_get(void* h, const char* name, void* dst, long cap_bytes) {
    auto* g = (GSkelOracle*)h;
    const std::string s(name);
    if (s == "global") return put(g->global_cloud, dst, cap_bytes);
    if (s == "prelim") { std::vector<V3> p; for (auto& v : g->prelim_vers) p.push_back(v.position); return put(p, dst, cap_bytes); }
    if (s == "prelim_frozen") { std::vector<int> p; for (auto& v : g->prelim_vers) p.push_back(v.frozen ? 1 : 0); return put(p, dst, cap_bytes); }
    if (s == "adj_off") return put_csr_off(g->global_adj, dst, cap_bytes);
    if (s == "adj_idx") return put_csr_idx(g->global_adj, dst, cap_bytes);
    if (s == "joints") return put(g->joints, dst, cap_bytes);
    if (s == "leafs") return put(g->leafs, dst, cap_bytes);
    if (s == "new_idx") return put(g->new_vers_indxs, dst, cap_bytes);
    if (s == "branch_off") return put_csr_off(g->branches, dst, cap_bytes);
    if (s == "branch_idx") return put_csr_idx(g->branches, dst, cap_bytes);
    if (s == "type") { std::vector<int> t; for (auto& v : g->global_vers) t.push_back(v.type); return put(t, dst, cap_bytes); }
    if (s == "obs") { std::vector<int> t; for (auto& v : g->global_vers) t.push_back(v.obs_count); return put(t, dst, cap_bytes); }
    return -1;
}

// ---- primitives, exposed for unit tests
int orc_eig3(const float* A9, float* ev3, float* Q9) {
    M3 A, Q; std::memcpy(A.m, A9, 36);
    bool ok = eig3_sym(A, ev3, Q);
    std::memcpy(Q9, Q.m, 36);
    return ok ? 1 : 0;
}
float orc_cbrtf(float x) { return canon_cbrtf(x); }
void orc_pcl_eigen33(const float* A9, float* ev, float* v3) {
    M3 A; std::memcpy(A.m, A9, 36);
    V3 v = pcl_eigen33_smallest(A, *ev);
    v3[0] = v.x; v3[1] = v.y; v3[2] = v.z;
}
int orc_llt3(const float* A9, const float* b3, float* x3) {
    M3 A; std::memcpy(A.m, A9, 36); V3 x;
    bool ok = llt3_solve(A, {b3[0], b3[1], b3[2]}, x);
    x3[0] = x.x; x3[1] = x.y; x3[2] = x.z; return ok;
}
int orc_ldlt3(const float* A9, const float* b3, float* x3) {
    M3 A; std::memcpy(A.m, A9, 36); V3 x;
    bool ok = ldlt3_solve(A, {b3[0], b3[1], b3[2]}, x);
    x3[0] = x.x; x3[1] = x.y; x3[2] = x.z; return ok;
}
// gskel_node.cpp:137-138 tf2::doTransform(PointCloud2): float32 affine = Translation3f * Quaternionf (Eigen 3.4
// Quaternion::toRotationMatrix), point = linear * p + translation with Eigen's 3-term reduction a + (b + c).
void orc_tf_points(const float* xyz, int n, int stride, const double* trans, const double* rot_xyzw, float* out3) {
    const float x = (float)rot_xyzw[0], y = (float)rot_xyzw[1], z = (float)rot_xyzw[2], w = (float)rot_xyzw[3];
    const float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    const float R[9] = {1.0f - (tyy + tzz), txy - twz, txz + twy,
                        txy + twz, 1.0f - (txx + tzz), tyz - twx,
                        txz - twy, tyz + twx, 1.0f - (txx + tyy)};
    const float T[3] = {(float)trans[0], (float)trans[1], (float)trans[2]};
    for (int i = 0; i < n; ++i) {
        const float* p = xyz + (size_t)i * stride;
        for (int r = 0; r < 3; ++r) {
            const float a = R[3 * r] * p[0], b = R[3 * r + 1] * p[1], c = R[3 * r + 2] * p[2];
            out3[3 * (size_t)i + r] = (a + (b + c)) + T[r];
        }
    }
}
// exact kNN / radius over a cloud (kd-tree), canonical (d2, idx) order
void orc_knn(const float* xyz, int n, int stride, int k, int* out_idx /* n*k */) {
    std::vector<V3> p(n);
    for (int i = 0; i < n; ++i) p[i] = {xyz[(size_t)i * stride], xyz[(size_t)i * stride + 1], xyz[(size_t)i * stride + 2]};
    KdTree kd; kd.build(p.data(), n);
    std::vector<int> ki; std::vector<float> kdist;
    for (int i = 0; i < n; ++i) {
        int f = kd.knn(p[i], k, ki, kdist);
        for (int t = 0; t < k; ++t) out_idx[(size_t)i * k + t] = t < f ? ki[t] : -1;
    }
}

}  // extern "C"
