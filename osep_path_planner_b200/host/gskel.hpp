// host/gskel.hpp -- C++ drop-in for the reference's `GSkel` class
// (osep_skeleton_decomp/include/gskel.hpp:23-142) over the C ABI of libosep_skel.so:
//   struct GSkelConfig                              gskel.hpp:23-33
//   explicit GSkel(const GSkelConfig&)              gskel.hpp:108, gskel.cpp:13-19
//   bool gskel_run()                                gskel.hpp:109, gskel.cpp:21-38
//   PointCloudXYZ& input_vertices()                 gskel.hpp:111 (candidates, already in the global frame: gskel_node.cpp:138-139)
//   PointCloudXYZ::Ptr& output_gskel()              gskel.hpp:112
// plus graph() -- adjacency / joints / leafs the reference keeps internal ("TODO: Publish edges", gskel_node.cpp:5).
#pragma once
#include "rosa.hpp"
#include <memory>

struct GSkelConfig {         // field-for-field the reference's struct (gskel.hpp:23-33)
    float gnd_th;
    float fuse_dist_th;
    float fuse_conf_th;
    float lkf_pn;
    float lkf_mn;
    int max_obs_wo_conf;
    int niter_smooth_vertex;
    float vertex_smooth_coef;
    int min_branch_length;
};

class GSkel {
public:
#ifdef OSEP_WITH_PCL_EIGEN
    using CloudPtr = osep::PointCloudXYZ::Ptr;
#else
    using CloudPtr = std::shared_ptr<osep::PointCloudXYZ>;
#endif
    explicit GSkel(const GSkelConfig& cfg, int device = 0, int capacity_vertices = 16384) : cfg_(cfg) {
        osep_gskel_config c;
        c.gnd_th = cfg.gnd_th; c.fuse_dist_th = cfg.fuse_dist_th; c.fuse_conf_th = cfg.fuse_conf_th; c.lkf_pn = cfg.lkf_pn; c.lkf_mn = cfg.lkf_mn;
        c.max_obs_wo_conf = cfg.max_obs_wo_conf; c.niter_smooth_vertex = cfg.niter_smooth_vertex; c.vertex_smooth_coef = cfg.vertex_smooth_coef;
        c.min_branch_length = cfg.min_branch_length;
        if (osep_gskel_create(&c, device, capacity_vertices, &h_) != OSEP_OK) throw std::runtime_error(std::string("osep_gskel_create failed: ") + osep_last_error());
        new_cands_.reset(new osep::PointCloudXYZ); global_vers_cloud_.reset(new osep::PointCloudXYZ);
        running = true;
    }
    ~GSkel() { if (h_) osep_gskel_destroy(h_); }
    GSkel(const GSkel&) = delete; GSkel& operator=(const GSkel&) = delete;

    bool gskel_run() { return run_impl(nullptr, nullptr); }           // gskel.cpp:21-38

    // addition: tf2::doTransform + fromROSMsg + gskel_run (gskel_node.cpp:137-142) in one call; the
    // candidates in input_vertices() are in the sensor frame, translation / rotation_xyzw come from lookupTransform.
    bool gskel_run_tf(const double translation[3], const double rotation_xyzw[4]) { return run_impl(translation, rotation_xyzw); }
    // same, with the candidates taken from a Rosa handle's DEVICE vertices (the frame that just ran): the transform runs on
    // the device and the sensor-frame vertices never cross to the host (osep_gskel_run_tf_device)
    bool gskel_run_tf_device(osep_rosa* rosa, const double translation[3], const double rotation_xyzw[4]) { return run_impl(translation, rotation_xyzw, rosa); }

private:
    bool run_impl(const double* translation, const double* rotation_xyzw, osep_rosa* rosa = nullptr) {
        static const char* names[7] = {"increment_skeleton", "graph_adj", "mst", "vertex_merge", "prune", "smooth_vertex_positions", "extract_branches"};
        auto ts = std::chrono::high_resolution_clock::now();
        const int n = (int)new_cands_->points.size();
        const float* cands = n ? reinterpret_cast<const float*>(new_cands_->points.data()) : nullptr;
        const int rc = rosa ? osep_gskel_run_tf_device(h_, rosa, translation, rotation_xyzw)
                            : (translation ? osep_gskel_run_tf(h_, cands, n, 4, translation, rotation_xyzw) : osep_gskel_run(h_, cands, n, 4));
        if (rc < 0) { std::cerr << "[GSKEL] device error: " << osep_last_error() << std::endl; running = false; return false; }
        // the reference mutates the output cloud in place stage by stage; mirror the final state of this tick
        const float* xyz4 = nullptr; int G = 0;
        osep_gskel_vertices(h_, &xyz4, &G);
        global_vers_cloud_->points.resize(G);
        for (int i = 0; i < G; ++i) { auto& p = global_vers_cloud_->points[i]; p.x = xyz4[4 * i]; p.y = xyz4[4 * i + 1]; p.z = xyz4[4 * i + 2]; }
        global_vers_cloud_->width = (unsigned)G; global_vers_cloud_->height = 1; global_vers_cloud_->is_dense = true;
        for (int s = 1; s <= 7; ++s) {
            const bool ok = (rc == 0 || s < rc);
            running = ok;
            std::cout << (ok ? "[SUCCESS] " : "[FAILED] ") << names[s - 1] << std::endl;
            if (!ok) return false;
        }
        auto te = std::chrono::high_resolution_clock::now();
        std::cout << "[GSKEL] Time Elapsed: " << std::chrono::duration_cast<std::chrono::milliseconds>(te - ts).count() << " ms" << std::endl;
        return running;
    }
public:
    osep::PointCloudXYZ& input_vertices() { return *new_cands_; }
    CloudPtr& output_gskel() { return global_vers_cloud_; }

    // addition: the skeleton graph (MST adjacency, joints, leafs) of the current global skeleton
    int graph(osep_graph_view* out) { return osep_gskel_graph(h_, out); }
    osep_gskel* handle() { return h_; }

private:
    GSkelConfig cfg_;
    bool running = false;
    osep_gskel* h_ = nullptr;
    CloudPtr new_cands_, global_vers_cloud_;
};
