// oracle/oracle_gskel.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See oracle_gskel.hpp.
// Citations are into /root/reference/osep_skeleton_decomp/.
#include "oracle_gskel.hpp"

namespace orc {

// ---- Eigen fixed 3x3 helpers (coefficient-based lazy products; 3-term sums are a0+(a1+a2))
static inline void mat3_mul(const float* A, const float* B, float* C) {
    float t[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
        t[r * 3 + c] = sum3(A[r * 3] * B[c], A[r * 3 + 1] * B[3 + c], A[r * 3 + 2] * B[6 + c]);
    for (int i = 0; i < 9; ++i) C[i] = t[i];
}
// Eigen Matrix3f::inverse(): cofactor/determinant (Eigen/src/LU/InverseImpl.h, compute_inverse<.,.,3>)
static inline float cof3(const float* m, int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}
static inline void mat3_inv(const float* m, float* out) {
    const float c0 = cof3(m, 0, 0), c1 = cof3(m, 1, 0), c2 = cof3(m, 2, 0);
    const float det = sum3(c0 * m[0], c1 * m[3], c2 * m[6]);
    const float invdet = 1.0f / det;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
        if (r == 0) continue;
        out[r * 3 + c] = cof3(m, c, r) * invdet;
    }
    out[0] = c0 * invdet; out[1] = c1 * invdet; out[2] = c2 * invdet;
}

// include/lkf_vertex_fuse.hpp:16-28
static void lkf_update(GVertex& g, const V3& z, float pn, float mn) {
    float Pp[9], S[9], Si[9], K[9];
    for (int i = 0; i < 9; ++i) { const float d = (i % 4 == 0) ? 1.0f : 0.0f; Pp[i] = g.kP[i] + d * pn; }
    for (int i = 0; i < 9; ++i) { const float d = (i % 4 == 0) ? 1.0f : 0.0f; S[i] = Pp[i] + d * mn; }
    mat3_inv(S, Si);
    mat3_mul(Pp, Si, K);
    const V3 inn = sub3(z, g.kx);
    const V3 Kd{sum3(K[0] * inn.x, K[1] * inn.y, K[2] * inn.z), sum3(K[3] * inn.x, K[4] * inn.y, K[5] * inn.z),
                sum3(K[6] * inn.x, K[7] * inn.y, K[8] * inn.z)};
    g.kx = add3(g.kx, Kd);
    float IK[9];
    for (int i = 0; i < 9; ++i) { const float d = (i % 4 == 0) ? 1.0f : 0.0f; IK[i] = d - K[i]; }
    mat3_mul(IK, Pp, g.kP);
}

bool GSkelOracle::run(const float* xyz, int n, int stride) {     // gskel.cpp:21-38
    new_cands.resize(n);
    for (int i = 0; i < n; ++i) new_cands[i] = {xyz[(size_t)i * stride], xyz[(size_t)i * stride + 1], xyz[(size_t)i * stride + 2]};
    failed_stage = 0;
    typedef bool (GSkelOracle::*Fn)();
    Fn steps[7] = {&GSkelOracle::increment_skeleton, &GSkelOracle::graph_adj_stage, &GSkelOracle::mst_stage, &GSkelOracle::vertex_merge,
                   &GSkelOracle::prune, &GSkelOracle::smooth_vertex_positions, &GSkelOracle::extract_branches};
    for (int s = 0; s < 7; ++s) if (!(this->*steps[s])()) { failed_stage = s + 1; return false; }
    return true;
}

// gskel.cpp:40-199
bool GSkelOracle::increment_skeleton() {
    if (new_cands.empty()) return false;
    for (auto& v : prelim_vers) v.just_approved = false;
    std::vector<V3> prelim_cloud; prelim_cloud.reserve(prelim_vers.size());
    for (const auto& v : prelim_vers) prelim_cloud.push_back(v.position);
    if (!prelim_cloud.empty()) kd_.build(prelim_cloud.data(), (int)prelim_cloud.size());
    const float fuse_r2 = cfg.fuse_dist_th * cfg.fuse_dist_th;
    std::vector<int> ib; std::vector<float> db;
    std::vector<V3> spawned;
    for (const V3& ver : new_cands) {
        if (!finite3(ver)) continue;
        int chosen = -1; float chosen_d2 = std::numeric_limits<float>::max();
        bool any_in_radius = false, frozen_in_radius = false;
        if (!prelim_cloud.empty()) {
            if (kd_.radius(ver, fuse_r2, ib, db) > 0) {
                any_in_radius = true;
                for (size_t k = 0; k < ib.size(); ++k) {
                    const int i = ib[k];
                    if (prelim_vers[i].frozen) { frozen_in_radius = true; continue; }
                    if (db[k] < chosen_d2) { chosen_d2 = db[k]; chosen = i; }
                }
            }
        }
        auto near_spawned = [&]() {
            for (const V3& s : spawned) if (sqn3(sub3(s, ver)) <= fuse_r2) return true;
            return false;
        };
        if (chosen >= 0 && chosen_d2 <= fuse_r2) {
            GVertex& g = prelim_vers[chosen];
            lkf_update(g, ver, cfg.lkf_pn, cfg.lkf_mn);
            if (!finite3(g.kx)) { g.marked_for_deletion = true; continue; }
            g.position = g.kx;
            g.obs_count++;
            const float trace = sum3(g.kP[0], g.kP[4], g.kP[8]);
            if (!g.conf_check && trace < cfg.fuse_conf_th) { g.conf_check = true; g.just_approved = true; g.frozen = true; g.unconf_check = 0; }
            else g.unconf_check++;
            if (g.unconf_check > cfg.max_obs_wo_conf) g.marked_for_deletion = true;
        } else if (any_in_radius || frozen_in_radius) {
            continue;
        } else if (!near_spawned()) {
            GVertex nv;
            nv.kx = ver; for (int i = 0; i < 9; ++i) nv.kP[i] = (i % 4 == 0) ? 1.0f : 0.0f;
            nv.position = ver; nv.obs_count++; nv.smooth_iters = cfg.niter_smooth_vertex;
            prelim_vers.push_back(nv);
            spawned.push_back(ver);
        }
    }
    { std::vector<GVertex> k; for (auto& v : prelim_vers) if (!v.marked_for_deletion) k.push_back(v); prelim_vers.swap(k); }
    new_vers_indxs.clear();
    for (auto& v : prelim_vers) {
        if (v.just_approved && v.position.z > cfg.gnd_th) {
            global_vers.push_back(v);
            global_cloud.push_back(v.position);
            new_vers_indxs.push_back((int)global_vers.size() - 1);
            v.just_approved = false; v.updated = true;
        }
    }
    { std::vector<GVertex> kept; std::vector<V3> clean;
      for (auto& v : global_vers) { if (!finite3(v.position)) continue; kept.push_back(v); clean.push_back(v.position); }
      global_vers.swap(kept); global_cloud.swap(clean); gskel_size = global_vers.size(); }
    return !new_vers_indxs.empty();
}

// gskel.cpp:201-246 (stateless core)
void graph_adj(const std::vector<V3>& pos, float fuse_dist_th, std::vector<std::vector<int>>& adj) {
    const int G = (int)pos.size();
    adj.assign(G, {});
    KdTree kd; kd.build(pos.data(), G);
    const int K = 10;
    const float max_dist_th = 2.5f * fuse_dist_th;
    std::vector<int> ind; std::vector<float> d2;
    for (int i = 0; i < G; ++i) {
        const int n_nbs = kd.knn(pos[i], K, ind, d2);
        for (int j = 1; j < n_nbs; ++j) {
            const int nb = ind[j];
            const float dist_to_nb = norm3(sub3(pos[i], pos[nb]));
            if (dist_to_nb > max_dist_th) continue;
            bool good = true;
            for (int k = 1; k < n_nbs; ++k) {
                if (k == j) continue;
                const int o = ind[k];
                const float d_nb_o = norm3(sub3(pos[nb], pos[o]));
                const float d_i_o = norm3(sub3(pos[i], pos[o]));
                if (d_nb_o < dist_to_nb && d_i_o < dist_to_nb) { good = false; break; }
            }
            if (good) { adj[i].push_back(nb); adj[nb].push_back(i); }
        }
    }
}

// gskel.cpp:539-567 (incl. the case-2 fall-through: every non-leaf ends with type 0)
void graph_decomp(const std::vector<std::vector<int>>& adj, std::vector<int>& joints, std::vector<int>& leafs, std::vector<int>& type) {
    joints.clear(); leafs.clear(); type.assign(adj.size(), 0);
    for (int i = 0; i < (int)adj.size(); ++i) {
        const int degree = (int)adj[i].size();
        if (degree == 1) { leafs.push_back(i); type[i] = 1; }
        else { if (degree > 2) joints.push_back(i); type[i] = 0; }
    }
}

// gskel.cpp:248-281 + gskel.hpp:35-68.  Canonical choice C6: edges ordered by (w, u, v).
void mst_from_adj(const std::vector<V3>& pos, const std::vector<std::vector<int>>& adj_in, GraphResult& out) {
    const int G = (int)pos.size();
    std::vector<GEdge> edges;
    for (int i = 0; i < G; ++i) for (int nb : adj_in[i]) {
        if (nb <= i) continue;
        edges.push_back({i, nb, norm3(sub3(pos[i], pos[nb]))});
    }
    std::stable_sort(edges.begin(), edges.end(), [](const GEdge& a, const GEdge& b) {
        if (a.w != b.w) return a.w < b.w;
        if (a.u != b.u) return a.u < b.u;
        return a.v < b.v;
    });
    std::vector<int> parent(G);
    for (int i = 0; i < G; ++i) parent[i] = i;
    auto find = [&](int x) { int r = x; while (parent[r] != r) r = parent[r]; while (parent[x] != r) { int nx = parent[x]; parent[x] = r; x = nx; } return r; };
    out.adj.assign(G, {}); out.mst_edges.clear();
    for (const auto& e : edges) {
        int rx = find(e.u), ry = find(e.v);
        if (rx == ry) continue;
        parent[ry] = rx;
        out.adj[e.u].push_back(e.v); out.adj[e.v].push_back(e.u);
        out.mst_edges.push_back(e);
    }
    graph_decomp(out.adj, out.joints, out.leafs, out.type);
}

GraphResult build_graph(const std::vector<V3>& pos, float fuse_dist_th) {
    GraphResult r;
    if (pos.empty()) return r;
    graph_adj(pos, fuse_dist_th, r.adj_rng);
    mst_from_adj(pos, r.adj_rng, r);
    return r;
}

bool GSkelOracle::size_assert() const {                      // gskel.cpp:591-599
    return global_vers.size() == global_cloud.size() && global_cloud.size() == global_adj.size() && global_adj.size() == gskel_size;
}
void GSkelOracle::decomp() {
    std::vector<int> type;
    graph_decomp(global_adj, joints, leafs, type);
    for (size_t i = 0; i < global_vers.size() && i < type.size(); ++i) global_vers[i].type = type[i];
}
bool GSkelOracle::graph_adj_stage() {
    if (global_cloud.empty()) return false;
    std::vector<V3> pos(global_cloud.begin(), global_cloud.begin() + gskel_size);
    graph_adj(pos, cfg.fuse_dist_th, global_adj);
    return size_assert();
}
bool GSkelOracle::mst_stage() {
    if (gskel_size == 0 || global_adj.empty()) return false;
    std::vector<V3> pos; for (auto& v : global_vers) pos.push_back(v.position);
    GraphResult r; mst_from_adj(pos, global_adj, r);
    global_adj = r.adj;
    decomp();
    return size_assert();
}

// gskel.cpp:569-589
void GSkelOracle::merge_into(int keep, int del) {
    GVertex& Vi = global_vers[keep]; GVertex& Vj = global_vers[del];
    const int tot = Vi.obs_count + Vj.obs_count;
    Vi.position = div3(add3(mul3(Vi.position, (float)Vi.obs_count), mul3(Vj.position, (float)Vj.obs_count)), (float)tot);
    Vi.obs_count = tot;
    for (int nb : global_adj[del]) {
        if (nb == keep) continue;
        auto& kn = global_adj[keep];
        if (std::find(kn.begin(), kn.end(), nb) == kn.end()) kn.push_back(nb);
        auto& nn = global_adj[nb];
        std::replace(nn.begin(), nn.end(), del, keep);
    }
}

// gskel.cpp:283-353
bool GSkelOracle::vertex_merge() {
    const int N_new = (int)new_vers_indxs.size();
    if (gskel_size == 0 || N_new == 0) return false;
    if ((float)N_new / (float)gskel_size > 0.5) return false;
    auto is_joint = [&](int idx) { return std::find(joints.begin(), joints.end(), idx) != joints.end(); };
    std::set<int> to_delete;
    for (int new_id : new_vers_indxs) {
        if (new_id < 0 || new_id >= (int)gskel_size) continue;
        if (to_delete.count(new_id)) continue;
        const std::vector<int> nbrs = global_adj[new_id];
        for (int nb_id : nbrs) {
            if (nb_id < 0 || nb_id >= (int)gskel_size) continue;
            if (new_id == nb_id || to_delete.count(nb_id)) continue;
            bool do_merge = false;
            if (is_joint(new_id) && is_joint(nb_id)) {
                do_merge = true;
                joints.erase(std::remove(joints.begin(), joints.end(), nb_id), joints.end());
            }
            const float dist = norm3(sub3(global_vers[new_id].position, global_vers[nb_id].position));
            if (!do_merge && dist < 0.5f * cfg.fuse_dist_th) do_merge = true;
            if (!do_merge) continue;
            merge_into(nb_id, new_id);
            to_delete.insert(new_id);
            break;
        }
    }
    if (to_delete.empty()) return true;
    for (auto it = to_delete.rbegin(); it != to_delete.rend(); ++it) {
        const int del = *it;
        if (del < 0 || del > (int)global_vers.size()) continue;
        global_vers.erase(global_vers.begin() + del);
        global_cloud.erase(global_cloud.begin() + del);
        global_adj.erase(global_adj.begin() + del);
        for (auto& nbrs : global_adj) {
            nbrs.erase(std::remove(nbrs.begin(), nbrs.end(), del), nbrs.end());
            for (auto& v : nbrs) if (v > del) --v;
            std::sort(nbrs.begin(), nbrs.end());
            nbrs.erase(std::unique(nbrs.begin(), nbrs.end()), nbrs.end());
        }
        new_vers_indxs.erase(std::remove(new_vers_indxs.begin(), new_vers_indxs.end(), del), new_vers_indxs.end());
        for (auto& id : new_vers_indxs) if (id > del) --id;
    }
    gskel_size = global_vers.size();
    decomp();
    return size_assert();
}

// gskel.cpp:355-394
bool GSkelOracle::prune() {
    const int N_new = (int)new_vers_indxs.size();
    if (gskel_size == 0 || N_new == 0) return false;
    if ((float)N_new / (float)gskel_size > 0.5) return false;
    std::set<int> joint_set(joints.begin(), joints.end()), new_set(new_vers_indxs.begin(), new_vers_indxs.end());
    std::vector<int> to_delete;
    for (int leaf : leafs) {
        if (!new_set.count(leaf)) continue;
        if (global_adj[leaf].size() != 1) continue;
        if (joint_set.count(global_adj[leaf][0])) to_delete.push_back(leaf);
    }
    std::sort(to_delete.rbegin(), to_delete.rend());
    for (int idx : to_delete) {
        global_vers.erase(global_vers.begin() + idx);
        global_cloud.erase(global_cloud.begin() + idx);
        global_adj.erase(global_adj.begin() + idx);
        for (auto& nbrs : global_adj) {
            nbrs.erase(std::remove(nbrs.begin(), nbrs.end(), idx), nbrs.end());
            for (auto& v : nbrs) if (v > idx) --v;
        }
        new_vers_indxs.erase(std::remove(new_vers_indxs.begin(), new_vers_indxs.end(), idx), new_vers_indxs.end());
        for (auto& id : new_vers_indxs) if (id > idx) --id;
    }
    gskel_size = global_vers.size();
    decomp();
    return true;
}

// gskel.cpp:396-434
bool GSkelOracle::smooth_vertex_positions() {
    if (gskel_size == 0 || global_adj.empty()) return false;
    std::vector<V3> new_pos(gskel_size, V3{0, 0, 0});
    for (size_t i = 0; i < gskel_size; ++i) {
        GVertex& v = global_vers[i];
        const auto& nbrs = global_adj[i];
        if (v.smooth_iters > 0) {
            V3 avg{0, 0, 0};
            for (int j : nbrs) avg = add3(avg, global_vers[j].position);
            avg = div3(avg, (float)nbrs.size());
            new_pos[i] = add3(mul3(v.position, 1.0f - cfg.vertex_smooth_coef), mul3(avg, cfg.vertex_smooth_coef));
            --v.smooth_iters;
        } else new_pos[i] = v.position;
    }
    for (size_t i = 0; i < gskel_size; ++i) global_vers[i].position = new_pos[i];
    global_cloud.clear();
    for (const auto& v : global_vers) global_cloud.push_back(v.position);
    return true;
}

// gskel.cpp:436-536, restated with its quirks: std::swap on the loop copies (:448,:485), the
// stray ';' that disables the visited check in the joints loop (:486), unsorted pairs inserted
// into visited_edges (:474-477).  Branch sort is made stable (ties unordered in the reference).
bool GSkelOracle::extract_branches() {
    if (gskel_size == 0) return false;
    branches.clear();
    std::set<std::pair<int, int>> visited;
    auto is_endpoint = [&](int v) { return global_vers[v].type == 1 || global_vers[v].type == 3; };
    auto walk = [&](int endp, int nb) {
        std::vector<int> branch; branch.push_back(endp);
        int prev = endp, curr = nb;
        size_t guard = 0;
        while (guard++ <= gskel_size + 1) {
            branch.push_back(curr);
            if (is_endpoint(curr) && curr != endp) break;
            int next = -1;
            for (int nn : global_adj[curr]) if (nn != prev) { next = nn; break; }
            if (next == -1) break;
            prev = curr; curr = next;
        }
        for (size_t i = 1; i < branch.size(); ++i) visited.insert({branch[i - 1], branch[i]});
        branches.push_back(branch);
    };
    for (int endp0 : leafs) {
        int endp = endp0;
        const std::vector<int>& lst = global_adj[endp0];
        for (int nb0 : lst) {
            int nb = nb0;
            if (endp > nb) std::swap(endp, nb);
            if (visited.count({endp, nb})) continue;
            walk(endp, nb);
        }
    }
    for (int endp0 : joints) {
        int endp = endp0;
        const std::vector<int>& lst = global_adj[endp0];
        for (int nb0 : lst) {
            int nb = nb0;
            if (endp > nb) std::swap(endp, nb);
            walk(endp, nb);
        }
    }
    std::stable_sort(branches.begin(), branches.end(), [](const std::vector<int>& a, const std::vector<int>& b) { return a.size() < b.size(); });
    const size_t min_bl = (size_t)cfg.min_branch_length;
    branches.erase(std::remove_if(branches.begin(), branches.end(), [min_bl](const std::vector<int>& b) { return b.size() < min_bl; }), branches.end());
    return true;
}

}  // namespace orc
