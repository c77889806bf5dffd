// host/rosa.hpp -- C++ drop-in for the reference's `Rosa` class
// (osep_skeleton_decomp/include/rosa.hpp:20-96) over the C ABI of libosep_skel.so.
//
// Same type and member names, same meaning:
//   struct RosaConfig { ... }                      rosa.hpp:20-33
//   explicit Rosa(const RosaConfig&)               rosa.hpp:50, rosa.cpp:15-34
//   bool rosa_run()                                rosa.hpp:51, rosa.cpp:36-55 (prints the same stdout lines)
//   PointCloudXYZ& input_cloud()                   rosa.hpp:52  (mutable buffer the node fills: rosa_node.cpp:129)
//   MatrixXf& output_vertices()                    rosa.hpp:53  (V x 3, column-major like Eigen::MatrixXf)
// Without PCL / Eigen this header supplies two minimal look-alikes (osep::PointXYZ with pcl::PointXYZ's
// 16-byte layout, osep::MatrixXf with rows()/cols()/operator()(r,c)/data()); when built inside the
// reference's workspace define OSEP_WITH_PCL_EIGEN and the real pcl::PointCloud<pcl::PointXYZ> /
// Eigen::MatrixXf are used instead, so RosaNode (rosa_node.cpp) compiles unchanged.
#pragma once
#include "../../include/osep_skel.h"
#include <chrono>
#include <cstdio>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef OSEP_WITH_PCL_EIGEN
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <Eigen/Core>
#endif

struct RosaConfig {          // field-for-field the reference's struct (rosa.hpp:20-33)
    int max_points;
    int min_points;
    float pts_dist_lim;
    int ne_knn;
    int nb_knn;
    float max_proj_range;
    int niter_drosa;
    int niter_dcrosa;
    int niter_smooth;
    float alpha_recenter;
    float radius_smooth;
};

namespace osep {
#ifndef OSEP_WITH_PCL_EIGEN
struct alignas(16) PointXYZ { float x = 0.f, y = 0.f, z = 0.f, pad = 1.f; };     // pcl::PointXYZ: {x,y,z,1.0f}, 16 B
struct PointCloudXYZ {
    std::vector<PointXYZ> points;
    unsigned width = 0, height = 1; bool is_dense = true;
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = 0; }
    void resize(size_t n) { points.resize(n); width = (unsigned)n; }
};
struct MatrixXf {                       // column-major, like Eigen::MatrixXf
    std::vector<float> d; long r = 0, c = 0;
    long rows() const { return r; } long cols() const { return c; }
    void resize(long rr, long cc) { r = rr; c = cc; d.resize((size_t)rr * cc); }
    float& operator()(long i, long j) { return d[(size_t)j * r + i]; }
    float operator()(long i, long j) const { return d[(size_t)j * r + i]; }
    float* data() { return d.data(); } const float* data() const { return d.data(); }
};
#else
using PointXYZ = pcl::PointXYZ;
using PointCloudXYZ = pcl::PointCloud<pcl::PointXYZ>;
using MatrixXf = Eigen::MatrixXf;
#endif
}  // namespace osep

class Rosa {
public:
    explicit Rosa(const RosaConfig& cfg, int device = 0, int capacity_points = 400000) : cfg_(cfg) {
        osep_rosa_config c;
        c.max_points = cfg.max_points; c.min_points = cfg.min_points; c.pts_dist_lim = cfg.pts_dist_lim; c.ne_knn = cfg.ne_knn;
        c.nb_knn = cfg.nb_knn; c.max_proj_range = cfg.max_proj_range; c.niter_drosa = cfg.niter_drosa; c.niter_dcrosa = cfg.niter_dcrosa;
        c.niter_smooth = cfg.niter_smooth; c.alpha_recenter = cfg.alpha_recenter; c.radius_smooth = cfg.radius_smooth;
        const int rc = osep_rosa_create(&c, device, capacity_points, &h_);
        if (rc != OSEP_OK) throw std::runtime_error(std::string("osep_rosa_create failed: ") + osep_last_error());   // no CPU fallback
        running = true;
    }
    ~Rosa() { if (h_) osep_rosa_destroy(h_); }
    Rosa(const Rosa&) = delete; Rosa& operator=(const Rosa&) = delete;

    bool rosa_run() { return run_impl(nullptr, 0, 0, 0); }            // rosa.cpp:36-55: same stdout protocol, RUN_STEP early exit

    // addition: rosa_run() straight from a sensor_msgs/PointCloud2 data block, replacing
    // pcl::fromROSMsg + rosa_run (rosa_node.cpp:129,136).  x/y/z must be consecutive float32 fields at byte offset off_x.
    bool rosa_run_pointcloud2(const void* data, int n_points, int point_step, int off_x = 0) { return run_impl(data, n_points, point_step, off_x); }

private:
    bool run_impl(const void* raw, int raw_n, int point_step, int off_x) {
        static const char* names[7] = {"preprocess", "rosa_init", "similarity_neighbor_extraction", "drosa", "dcrosa", "vertex_sampling", "vertex_smooth"};
        auto ts = std::chrono::high_resolution_clock::now();
        const int n = raw ? raw_n : (int)orig_.points.size();
        std::cout << "PointCloud size before downsampling: " << n << std::endl;
        osep_rosa_result res;
        const int rc = raw ? osep_rosa_run_pointcloud2(h_, raw, raw_n, point_step, off_x, off_x + 4, off_x + 8, &res)
                           : osep_rosa_run(h_, n ? reinterpret_cast<const float*>(orig_.points.data()) : nullptr, n, 4, &res);
        last_ = res;
        if (mirror_filter_ && !raw && rc >= 0 && res.n_filtered >= 0 && res.n_filtered != n) {
            std::vector<float> f((size_t)3 * res.n_filtered);
            if (osep_rosa_get(h_, "filtered", f.data(), (long)(f.size() * sizeof(float))) == (long)f.size()) {
                orig_.points.resize((size_t)res.n_filtered);
                for (int i = 0; i < res.n_filtered; ++i) { auto& p = orig_.points[(size_t)i]; p.x = f[3 * (size_t)i]; p.y = f[3 * (size_t)i + 1]; p.z = f[3 * (size_t)i + 2]; }
            }
        }
        if (rc < 0) { std::cerr << "[ROSA] device error: " << osep_last_error() << std::endl; running = false; skelver_.resize(0, 3); return false; }
        for (int s = 1; s <= 7; ++s) {
            const bool ok = (rc == 0 || s < rc);
            running = ok;
            std::cout << (ok ? "[SUCCESS] " : "[FAILED] ") << names[s - 1] << std::endl;
            if (!ok) { skelver_.resize(0, 3); return false; }
            if (s == 1) std::cout << "PointCloud size after downsampling: " << res.n_downsampled << std::endl;
        }
        const float* v = nullptr; int V = 0;
        osep_rosa_vertices(h_, &v, &V);
        skelver_.resize(V, 3);
        for (long j = 0; j < 3; ++j) for (long i = 0; i < V; ++i) skelver_(i, j) = v[(size_t)j * V + i];
        std::cout << "Number of vertices: " << skelver_.rows() << std::endl;
        auto te = std::chrono::high_resolution_clock::now();
        std::cout << "[ROSA] Time Elapsed: " << std::chrono::duration_cast<std::chrono::milliseconds>(te - ts).count() << " ms" << std::endl;
        return running;
    }
public:
    osep::PointCloudXYZ& input_cloud() { return orig_; }
    // The reference filters input_cloud() IN PLACE (rosa.cpp:84: orig_pcd_.swap(filtered)): after rosa_run() the caller's cloud
    // holds only the finite points within pts_dist_lim.  Off by default (one extra device-to-host copy of N' points per frame);
    // enable it when a caller reads input_cloud() after the run.
    void set_filter_input_in_place(bool on) { mirror_filter_ = on; }
    osep::MatrixXf& output_vertices() { return skelver_; }

    // additions (not in the reference): summary of the last frame and the device handle
    const osep_rosa_result& last_result() const { return last_; }
    osep_rosa* handle() { return h_; }

private:
    RosaConfig cfg_;
    bool running = false;
    osep_rosa* h_ = nullptr;
    osep::PointCloudXYZ orig_;
    osep::MatrixXf skelver_;
    osep_rosa_result last_{};
    bool mirror_filter_ = false;
};
