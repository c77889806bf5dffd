// host/demo_node.cpp -- what RosaNode::process_tick + GSkelNode::process_tick do (rosa_node.cpp:120-140,
// gskel_node.cpp:114-147) with the ROS plumbing replaced by a synthetic frame source, to show the shim
// classes are used exactly like the reference's.  Build: make -C osep_path_planner_b200/host.
// Usage: demo_node [n_points] [n_frames]
#include "gskel.hpp"
#include <cmath>
#include <cstdlib>
#include <random>

int main(int argc, char** argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 20000;
    const int frames = argc > 2 ? std::atoi(argv[2]) : 3;
    RosaConfig rc{1000, 100, 50.0f, 20, 10, 10.0f, 6, 5, 3, 0.3f, 5.0f};                 // rosa_node.cpp:50-60
    GSkelConfig gc{-100.0f, 3.0f, 0.1f, 1e-4f, 0.1f, 5, 3, 0.3f, 5};                       // gskel_node.cpp:62-70 (gnd_th lowered: synthetic scene)
    Rosa rosa(rc, 0, n);
    GSkel gskel(gc);
    std::mt19937 rng(0);
    std::uniform_real_distribution<float> uth(1.5707963f, 4.7123890f), uz(-25.f, 25.f);
    std::normal_distribution<float> noise(0.f, 0.02f);
    for (int f = 0; f < frames; ++f) {
        auto& cloud = rosa.input_cloud();                                                  // pcl::fromROSMsg(*msg, rosa_->input_cloud())
        cloud.resize(n);
        for (int i = 0; i < n; ++i) {
            const float th = uth(rng), z = uz(rng);
            cloud.points[i].x = 15.f + 2.f * std::cos(th) + noise(rng);
            cloud.points[i].y = 2.f * std::sin(th) + noise(rng);
            cloud.points[i].z = z + noise(rng);
        }
        if (rosa.rosa_run()) {                                                             // rosa_node.cpp:137-140
            auto& v = rosa.output_vertices();
            auto& cand = gskel.input_vertices();                                           // pcl::fromROSMsg(cloud_tf, gskel_->input_vertices())
            cand.resize(v.rows());
            for (long i = 0; i < v.rows(); ++i) { cand.points[i].x = v(i, 0); cand.points[i].y = v(i, 1); cand.points[i].z = v(i, 2); }
            if (gskel.gskel_run()) std::printf("global skeleton: %zu vertices\n", gskel.output_gskel()->points.size());
        }
    }
    return 0;
}
