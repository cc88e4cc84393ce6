// test_adapters.cpp — exercises the two reference call shapes through the header-only adapters.
// Built by adapters/Makefile (compile check on CPU); run on the GPU box by tests/test_gpu_adapters.py.
// Self-checking against an in-file exhaustive search: exits 0 on success.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <random>

#include "icp4r/kd_tree.hpp"
#include "icp4r/registration.hpp"
#include "icp4r/voxel_grid.hpp"

using PointType = pcl::PointXYZI;

static float d2f(const PointType& a, const PointType& b) {
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return (dx * dx + dy * dy) + dz * dz;
}

#define REQUIRE(c)                                                     \
    do {                                                               \
        if (!(c)) {                                                    \
            std::fprintf(stderr, "FAILED %s:%d %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                  \
        }                                                              \
    } while (0)

int main() {
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> U(-20.f, 20.f), Z(-2.f, 2.f);
    auto rnd = [&](float i) {
        PointType p;
        p.x = U(rng);
        p.y = U(rng);
        p.z = Z(rng);
        p.intensity = i;
        return p;
    };
    // ---- the ikd-Tree shape, as radar_odometry.cpp:92,347-348,390 uses it
    icp4r::KD_TREE<PointType> ikd_Tree(0.3f, 0.6f, 0.5f);
    icp4r::KD_TREE<PointType>::PointVector first, more;
    for (int i = 0; i < 3000; ++i) first.push_back(rnd((float)i));
    for (int i = 0; i < 2000; ++i) more.push_back(rnd((float)(3000 + i)));
    ikd_Tree.Build(first);
    ikd_Tree.set_downsample_param(0.5f);
    REQUIRE(ikd_Tree.Add_Points(more, false) == 0);
    REQUIRE(ikd_Tree.size() == 5000 && ikd_Tree.validnum() == 5000);
    std::vector<PointType> all(first.begin(), first.end());
    all.insert(all.end(), more.begin(), more.end());
    for (int t = 0; t < 50; ++t) {
        PointType q = rnd(0.f);
        icp4r::KD_TREE<PointType>::PointVector nn;
        std::vector<float> dd;
        ikd_Tree.Nearest_Search(q, 5, nn, dd, t % 2 ? 3.0 : INFINITY);
        // exhaustive check
        std::vector<std::pair<float, int>> ref;
        for (int i = 0; i < (int)all.size(); ++i) {
            const float d = d2f(q, all[i]);
            if (t % 2 == 0 || (double)d <= 9.0) ref.push_back({d, i});
        }
        std::sort(ref.begin(), ref.end());
        const int want = std::min<int>(5, (int)ref.size());
        REQUIRE((int)nn.size() == want && (int)dd.size() == want);
        for (int i = 0; i < want; ++i) {
            REQUIRE(dd[i] == ref[i].first);
            REQUIRE((int)nn[i].intensity == ref[i].second);  // the copy carries the original point's payload
        }
    }
    BoxPointType r = ikd_Tree.tree_range();
    REQUIRE(r.vertex_min[0] >= -20.f && r.vertex_max[0] <= 20.f && r.vertex_min[0] < r.vertex_max[0]);

    // ---- the PCL registration shape, as iterative_closest_point.cpp:510-521 uses it
    pcl::PointCloud<PointType>::Ptr cloud_src_in(new pcl::PointCloud<PointType>), cloud_tar_in(new pcl::PointCloud<PointType>);
    pcl::PointCloud<PointType>::Ptr Final(new pcl::PointCloud<PointType>);
    const float yaw = 0.02f, tx = 0.15f, ty = -0.1f;
    for (int i = 0; i < 2000; ++i) {
        PointType p = rnd(1.f);
        cloud_tar_in->push_back(p);
        PointType s = p;  // source = target moved by the inverse of (yaw, t): ICP must recover (yaw, t)
        const float x = p.x - tx, y = p.y - ty;
        s.x = std::cos(yaw) * x + std::sin(yaw) * y;
        s.y = -std::sin(yaw) * x + std::cos(yaw) * y;
        cloud_src_in->push_back(s);
    }
    icp4r::IterativeClosestPoint<PointType, PointType> icp;
    icp.setInputSource(cloud_src_in);
    icp.setInputTarget(cloud_tar_in);
    icp.align(*Final);
    REQUIRE(icp.hasConverged());
    const icp4r::Matrix4f T = icp.getFinalTransformation();
    std::printf("icp: score %.3g, T(0,3)=%.4f T(1,3)=%.4f T(1,0)=%.5f\n", icp.getFitnessScore(), T(0, 3), T(1, 3), T(1, 0));
    REQUIRE(std::fabs(T(0, 3) - tx) < 2e-3f && std::fabs(T(1, 3) - ty) < 2e-3f && std::fabs(T(1, 0) - std::sin(yaw)) < 2e-4f);
    REQUIRE(icp.getFitnessScore() < 1e-4);
    REQUIRE(Final->size() == cloud_src_in->size());
    REQUIRE(std::fabs(Final->points[0].x - cloud_tar_in->points[0].x) < 1e-2f);

    // ---- the fast_gicp shape, as radar_odometry.cpp:399-411 uses it
    icp4r::FastGICPSingleThread<PointType, PointType> fgicp_st;
    fgicp_st.clearTarget();
    fgicp_st.clearSource();
    fgicp_st.setInputTarget(cloud_tar_in);
    fgicp_st.setInputSource(cloud_src_in);
    fgicp_st.setCorrespondenceRandomness(5);
    fgicp_st.align(*Final);
    std::printf("gicp-shape: converged %d, score %.3g, iterations %d\n", (int)fgicp_st.hasConverged(), fgicp_st.getFitnessScore(),
                fgicp_st.getIterations());
    // an empty target must not throw: PCL prints an error and hasConverged() stays false
    pcl::PointCloud<PointType>::Ptr empty(new pcl::PointCloud<PointType>);
    icp4r::IterativeClosestPoint<PointType, PointType> icp2;
    icp2.setInputSource(cloud_src_in);
    icp2.setInputTarget(empty);
    icp2.align(*Final);
    REQUIRE(!icp2.hasConverged());
    // ---- the rest of the KD_TREE surface (ikd_Tree.h:243-249)
    {
        BoxPointType box;
        for (int a = 0; a < 3; ++a) {
            box.vertex_min[a] = -5.f;
            box.vertex_max[a] = 5.f;
        }
        icp4r::KD_TREE<PointType>::PointVector in_box, in_ball;
        ikd_Tree.Box_Search(box, in_box);
        size_t want = 0;
        for (const auto& p : all) want += (p.x >= -5.f && p.x < 5.f && p.y >= -5.f && p.y < 5.f && p.z >= -5.f && p.z < 5.f);
        REQUIRE(in_box.size() == want && want > 0);
        PointType c0 = rnd(0.f);
        ikd_Tree.Radius_Search(c0, 4.f, in_ball);
        want = 0;
        for (const auto& p : all) want += d2f(c0, p) <= 16.f;
        REQUIRE(in_ball.size() == want);
        std::vector<BoxPointType> boxes(1, box);
        const int before = ikd_Tree.validnum();
        const int deleted = ikd_Tree.Delete_Point_Boxes(boxes);
        REQUIRE(deleted == (int)in_box.size() && ikd_Tree.validnum() == before - deleted && ikd_Tree.size() == before);
        ikd_Tree.Box_Search(box, in_ball);
        REQUIRE(in_ball.empty());
        ikd_Tree.Add_Point_Boxes(boxes);
        REQUIRE(ikd_Tree.validnum() == before);
        icp4r::KD_TREE<PointType>::PointVector removed, flat;
        ikd_Tree.acquire_removed_points(removed);  // ikd_Tree.h:247: the box's points came back before anybody asked
        REQUIRE(removed.empty());
        icp4r::KD_TREE<PointType>::PointVector victims(all.begin(), all.begin() + 7);
        ikd_Tree.Delete_Points(victims);
        REQUIRE(ikd_Tree.validnum() == before - 7);
        ikd_Tree.acquire_removed_points(removed);
        REQUIRE(removed.size() == 7);
        for (int i = 0; i < 7; ++i) REQUIRE(removed[i].x == all[i].x && removed[i].intensity == all[i].intensity);
        ikd_Tree.acquire_removed_points(removed);  // appends, and hands every point out once
        REQUIRE(removed.size() == 7);
        ikd_Tree.flatten(flat);                    // ikd_Tree.h:246: every non-deleted point
        REQUIRE((int)flat.size() == before - 7 && flat[0].intensity == all[7].intensity);
        std::printf("box/radius/delete shapes: %zu in box, %d deleted and restored, 7 points deleted\n", in_box.size(), deleted);
    }
    // ---- the pcl::VoxelGrid shape, as radar_odometry.cpp:426-429 uses it
    {
        icp4r::VoxelGrid<PointType> sor;
        pcl::PointCloud<PointType>::Ptr downSizeFilterMap(new pcl::PointCloud<PointType>);
        sor.setInputCloud(cloud_tar_in);
        sor.setLeafSize(0.5f, 0.5f, 0.5f);
        sor.filter(*downSizeFilterMap);
        REQUIRE(!downSizeFilterMap->empty() && downSizeFilterMap->size() <= cloud_tar_in->size());
        // every input point's leaf holds exactly one output point: count-weighted means reproduce the cloud's mean
        std::vector<long long> keys_in, keys_out;
        auto leaf_key = [](const PointType& p) {
            return ((long long)std::floor(p.x * 2.f) + 100000) + ((long long)std::floor(p.y * 2.f) + 100000) * 1000000LL +
                   ((long long)std::floor(p.z * 2.f) + 100000) * 1000000000000LL;
        };
        for (const auto& p : cloud_tar_in->points) keys_in.push_back(leaf_key(p));
        std::sort(keys_in.begin(), keys_in.end());
        keys_in.erase(std::unique(keys_in.begin(), keys_in.end()), keys_in.end());
        REQUIRE(keys_in.size() == downSizeFilterMap->size());
        std::printf("voxel-grid-shape: %zu -> %zu points\n", cloud_tar_in->size(), downSizeFilterMap->size());
    }
    std::printf("adapters ok\n");
    return 0;
}
