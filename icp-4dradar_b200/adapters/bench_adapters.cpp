// bench_adapters.cpp — end-to-end timings THROUGH the header-only adapters, i.e. at the boundary a maintainer of the
// reference would actually call: pageable pcl::PointCloud<pcl::PointXYZI> clouds (32-byte rows) in, PCL-shaped results
// out, every host<->device copy and every per-call allocation inside the timed region.
//   C1 shape  pcl::IterativeClosestPoint::align        (/root/reference/src/iterative_closest_point.cpp:510-521)
//   map shape KD_TREE::Build / Add_Points / Nearest_Search (/root/reference/src/radar_odometry.cpp:347,390,396)
//   GICP      fast_gicp::FastGICPSingleThread::align    (/root/reference/src/radar_odometry.cpp:399-411)
// Prints one JSON object; bench.py runs it and attaches the numbers as `configs.adapters`.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "icp4r/kd_tree.hpp"
#include "icp4r/registration.hpp"
#include "icp4r/voxel_grid.hpp"

using PointType = pcl::PointXYZI;
using Cloud = pcl::PointCloud<PointType>;

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
template <typename F>
static double median_ms(int reps, int warm, F f) {
    std::vector<double> t;
    for (int i = 0; i < warm + reps; ++i) {
        const double t0 = now_ms();
        f(i);
        const double t1 = now_ms();
        if (i >= warm) t.push_back(t1 - t0);
    }
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

// points on the floor and the walls of a 2*ext x 2*ext x 4 m hall, 2 cm noise
static void hall(std::mt19937& rng, int n, float ext, std::vector<PointType, Eigen::aligned_allocator<PointType>>& out) {
    std::uniform_real_distribution<float> U(-ext, ext), Z(0.f, 4.f);
    std::normal_distribution<float> N(0.f, 0.02f);
    std::uniform_int_distribution<int> W(0, 5);
    out.resize(n);
    for (int i = 0; i < n; ++i) {
        PointType p;
        const int w = W(rng);
        if (w < 2) { p.x = U(rng); p.y = U(rng); p.z = N(rng); }
        else if (w == 2) { p.x = -ext + N(rng); p.y = U(rng); p.z = Z(rng); }
        else if (w == 3) { p.x = ext + N(rng); p.y = U(rng); p.z = Z(rng); }
        else if (w == 4) { p.x = U(rng); p.y = -ext + N(rng); p.z = Z(rng); }
        else { p.x = U(rng); p.y = ext + N(rng); p.z = Z(rng); }
        p.intensity = (float)i;
        out[i] = p;
    }
}
static PointType moved(const PointType& p, float yaw, float tx, float ty) {
    PointType q = p;
    q.x = std::cos(yaw) * p.x - std::sin(yaw) * p.y + tx;
    q.y = std::sin(yaw) * p.x + std::cos(yaw) * p.y + ty;
    return q;
}

int main() {
    std::mt19937 rng(11);
    // ---- C1 shape: 1,024 + 1,024 points, 30 iterations
    std::shared_ptr<Cloud> src(new Cloud), tgt(new Cloud);
    hall(rng, 1024, 20.f, tgt->points);
    src->points.resize(tgt->points.size());
    for (std::size_t i = 0; i < tgt->points.size(); ++i) src->points[i] = moved(tgt->points[i], 0.01f, 0.05f, -0.03f);
    Cloud aligned;
    int iters_c1 = 0;
    const double c1_ms = median_ms(200, 10, [&](int) {
        icp4r::IterativeClosestPoint<PointType, PointType> icp;  // constructed per frame, like the reference
        icp.setMaximumIterations(30);
        icp.setInputSource(src);
        icp.setInputTarget(tgt);
        icp.align(aligned);
        iters_c1 = icp.getIterations();
        if (!icp.hasConverged()) std::fprintf(stderr, "C1 align did not converge\n");
    });
    // ---- map shape: Build 200 k, Add_Points 4,096, batched 5-NN of 4,096 queries, Sector_Search
    icp4r::KD_TREE<PointType> tree(0.3f, 0.6f, 0.5f);
    icp4r::KD_TREE<PointType>::PointVector map_pts, scan;
    hall(rng, 200000, 100.f, map_pts);
    hall(rng, 4096, 30.f, scan);
    const double build_ms = median_ms(10, 2, [&](int) { tree.Build(map_pts); });
    std::vector<int32_t> idx, found;
    std::vector<float> d2;
    const double knn_ms = median_ms(50, 5, [&](int) { tree.Nearest_Search_Batch(scan, 5, idx, d2, found, 2.0); });
    icp4r::KD_TREE<PointType>::PointVector one_nn;
    std::vector<float> one_d;
    const double knn1_ms = median_ms(200, 10, [&](int i) { tree.Nearest_Search(scan[i % scan.size()], 5, one_nn, one_d, 2.0); });
    icp4r::KD_TREE<PointType>::PointVector sub;
    PointType centre;
    centre.x = centre.y = centre.z = 0.f;
    const double sector_ms = median_ms(20, 3, [&](int) { tree.Sector_Search(centre, 80.f, 0.f, sub); });
    const double add_ms = median_ms(20, 3, [&](int) { tree.Add_Points(scan, false); });
    // ---- GICP shape: scan vs the sector sub-map, fast_gicp defaults (what the scan-to-map node runs per frame)
    std::shared_ptr<Cloud> gs(new Cloud), gt(new Cloud);
    gs->points.resize(scan.size());
    for (std::size_t i = 0; i < scan.size(); ++i) gs->points[i] = moved(scan[i], 0.005f, 0.04f, 0.02f);
    gt->points.assign(sub.begin(), sub.end());
    int iters_g = 0;
    const double gicp_ms = median_ms(30, 3, [&](int) {
        icp4r::FastGICPSingleThread<PointType, PointType> reg;
        reg.setCorrespondenceRandomness(5);
        reg.setInputSource(gs);
        reg.setInputTarget(gt);
        reg.align(aligned);
        iters_g = reg.getIterations();
    });
    // ---- VoxelGrid over the map
    std::shared_ptr<Cloud> mc(new Cloud);
    mc->points.assign(map_pts.begin(), map_pts.end());
    Cloud ds;
    const double vg_ms = median_ms(10, 2, [&](int) {
        icp4r::VoxelGrid<PointType> sor;
        sor.setInputCloud(mc);
        sor.setLeafSize(0.5f, 0.5f, 0.5f);
        sor.filter(ds);
    });
    std::printf(
        "{\"clouds\": \"pageable pcl::PointCloud<pcl::PointXYZI> (32-byte rows) passed as they lie in memory (icp4r_set_point_layout)\", "
        "\"c1_align_ms\": %.4f, \"c1_iterations\": %d, \"c1_registrations_per_s\": %.1f, "
        "\"build_200k_ms\": %.4f, \"nearest_search_batch_4096x5_ms\": %.4f, \"nearest_search_single_ms\": %.4f, "
        "\"sector_search_ms\": %.4f, \"sector_points\": %d, \"add_points_4096_ms\": %.4f, "
        "\"gicp_align_ms\": %.4f, \"gicp_iterations\": %d, \"gicp_target_points\": %d, "
        "\"voxel_grid_200k_ms\": %.4f, \"voxel_grid_leaves\": %d}\n",
        c1_ms, iters_c1, 1e3 / c1_ms, build_ms, knn_ms, knn1_ms, sector_ms, (int)sub.size(), add_ms, gicp_ms, iters_g, (int)gt->points.size(), vg_ms,
        (int)ds.points.size());
    return 0;
}
