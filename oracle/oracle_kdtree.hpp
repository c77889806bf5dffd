// oracle/oracle_kdtree.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Exact kd-tree standing in for pcl::search::KdTree -> pcl::KdTreeFLANN ->
// FLANN KDTreeSingleIndex (max leaf 15, L2_Simple, eps = 0, sorted = true), the
// neighbour search behind rosa.cpp:99,207,231,263,265,367,372,456,480,553,566
// and gskel.cpp:55,77,205,216.  FLANN is not vendored under /root/reference
// (PARITY UNPINNED); published semantics restated:
//   * squared distances, float, accumulated x,y,z in order (dist2()).
//   * nearestKSearch: k clamped to the cloud size, ascending distance.
//   * radiusSearch: strict d2 < r*r, all neighbours, ascending distance.
//   * non-finite points are not inserted; original indices are reported.
// Canonical choice C1: ties are ordered by (d2, index) -- FLANN's own tie order
// depends on its tree layout and is not reproducible across builds.
#pragma once
#include "oracle_math.hpp"
#include <vector>
#include <numeric>

namespace orc {

class KdTree {
public:
    void build(const V3* pts, int n) {
        pts_ = pts;
        idx_.clear();
        for (int i = 0; i < n; ++i) if (finite3(pts[i])) idx_.push_back(i);
        nodes_.clear();
        if (!idx_.empty()) { nodes_.reserve(idx_.size() / 4 + 8); build_rec(0, (int)idx_.size()); }
    }
    int size() const { return (int)idx_.size(); }

    // k nearest by (d2, idx); returns count (<= k), outputs ascending.
    int knn(const V3& q, int k, std::vector<int>& out_i, std::vector<float>& out_d) const {
        out_i.clear(); out_d.clear();
        if (idx_.empty() || k <= 0) return 0;
        k = std::min(k, (int)idx_.size());
        best_i_.assign(k, -1); best_d_.assign(k, std::numeric_limits<float>::infinity());
        cnt_ = 0; kk_ = k;
        knn_rec(0, q);
        out_i.assign(best_i_.begin(), best_i_.begin() + cnt_);
        out_d.assign(best_d_.begin(), best_d_.begin() + cnt_);
        return cnt_;
    }

    // all with d2 < r2, ascending (d2, idx); returns count.
    int radius(const V3& q, float r2, std::vector<int>& out_i, std::vector<float>& out_d) const {
        out_i.clear(); out_d.clear();
        if (idx_.empty()) return 0;
        tmp_.clear();
        radius_rec(0, q, r2);
        std::sort(tmp_.begin(), tmp_.end());
        for (auto& e : tmp_) { out_d.push_back(e.first); out_i.push_back(e.second); }
        return (int)tmp_.size();
    }

private:
    struct Node { int lo, hi, dim, left, right; float split_lo, split_hi; };
    const V3* pts_ = nullptr;
    std::vector<int> idx_;
    std::vector<Node> nodes_;
    mutable std::vector<int> best_i_;
    mutable std::vector<float> best_d_;
    mutable int cnt_ = 0, kk_ = 0;
    mutable std::vector<std::pair<float, int>> tmp_;

    int build_rec(int lo, int hi) {
        int id = (int)nodes_.size();
        nodes_.push_back({lo, hi, -1, -1, -1, 0.f, 0.f});
        if (hi - lo <= 15) return id;
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = lo; i < hi; ++i) for (int d = 0; d < 3; ++d) {
            float v = pts_[idx_[i]][d]; mn[d] = std::min(mn[d], v); mx[d] = std::max(mx[d], v);
        }
        int dim = 0;
        if (mx[1] - mn[1] > mx[dim] - mn[dim]) dim = 1;
        if (mx[2] - mn[2] > mx[dim] - mn[dim]) dim = 2;
        if (!(mx[dim] > mn[dim])) return id;   // all identical: keep as a leaf
        int mid = (lo + hi) / 2;
        std::nth_element(idx_.begin() + lo, idx_.begin() + mid, idx_.begin() + hi,
                         [&](int a, int b) { return pts_[a][dim] < pts_[b][dim]; });
        float slo = -INFINITY, shi = INFINITY;
        for (int i = lo; i < mid; ++i) slo = std::max(slo, pts_[idx_[i]][dim]);   // max of left
        shi = pts_[idx_[mid]][dim];
        for (int i = mid; i < hi; ++i) shi = std::min(shi, pts_[idx_[i]][dim]);    // min of right
        int l = build_rec(lo, mid);
        int r = build_rec(mid, hi);
        nodes_[id].dim = dim; nodes_[id].left = l; nodes_[id].right = r;
        nodes_[id].split_lo = slo; nodes_[id].split_hi = shi;
        return id;
    }

    static bool less_pair(float d, int i, float d2, int i2) { return d < d2 || (d == d2 && i < i2); }

    void knn_insert(float d, int i) const {
        if (cnt_ == kk_ && !less_pair(d, i, best_d_[kk_ - 1], best_i_[kk_ - 1])) return;
        int pos = (cnt_ < kk_) ? cnt_++ : kk_ - 1;
        while (pos > 0 && less_pair(d, i, best_d_[pos - 1], best_i_[pos - 1])) {
            best_d_[pos] = best_d_[pos - 1]; best_i_[pos] = best_i_[pos - 1]; --pos;
        }
        best_d_[pos] = d; best_i_[pos] = i;
    }

    // Monotonicity of IEEE rounding makes (q[dim]-plane)^2 a true lower bound of the
    // float-evaluated dist2() of any point beyond that plane, so pruning on a strict
    // '>' never drops a candidate that ties on distance.
    void knn_rec(int id, const V3& q) const {
        const Node& nd = nodes_[id];
        if (nd.dim < 0) {
            for (int i = nd.lo; i < nd.hi; ++i) knn_insert(dist2(q, pts_[idx_[i]]), idx_[i]);
            return;
        }
        float v = q[nd.dim];
        int first = nd.left, second = nd.right;
        float gap;   // distance from q to the far child's slab
        if (v <= nd.split_lo) { gap = nd.split_hi - v; }
        else if (v >= nd.split_hi) { first = nd.right; second = nd.left; gap = v - nd.split_lo; }
        else if (v - nd.split_lo < nd.split_hi - v) { gap = nd.split_hi - v; }
        else { first = nd.right; second = nd.left; gap = v - nd.split_lo; }
        knn_rec(first, q);
        float worst = (cnt_ == kk_) ? best_d_[kk_ - 1] : std::numeric_limits<float>::infinity();
        if (gap < 0.f) gap = 0.f;
        if (!(gap * gap > worst)) knn_rec(second, q);
    }

    void radius_rec(int id, const V3& q, float r2) const {
        const Node& nd = nodes_[id];
        if (nd.dim < 0) {
            for (int i = nd.lo; i < nd.hi; ++i) {
                float d = dist2(q, pts_[idx_[i]]);
                if (d < r2) tmp_.push_back({d, idx_[i]});
            }
            return;
        }
        float v = q[nd.dim];
        float gl = v - nd.split_lo; if (gl < 0.f) gl = 0.f;   // distance to the left slab
        float gr = nd.split_hi - v; if (gr < 0.f) gr = 0.f;   // distance to the right slab
        if (gl * gl < r2) radius_rec(nd.left, q, r2);
        if (gr * gr < r2) radius_rec(nd.right, q, r2);
    }
};

}  // namespace orc
