// voxel_grid.hpp — icp4r::VoxelGrid<PointT>, the call shape of pcl::VoxelGrid as the scan-to-map node uses it
// (/root/reference/src/radar_odometry.cpp:426-429):
//     pcl::VoxelGrid<pcl::PointXYZI> sor;  sor.setInputCloud(RadarCloudMap);
//     sor.setLeafSize(0.5f, 0.5f, 0.5f);    sor.filter(*downSizeFilterMap);
// One output point per occupied leaf: the float mean of x, y, z (and intensity), leaves in ascending index order
// (x fastest), header/sensor pose copied from the input like pcl::Filter does. Only cubic leaves are supported
// (the reference never uses another shape); a leaf size that would overflow the leaf index leaves `output` equal to
// the input, as PCL does.
#pragma once
#include <memory>
#include <vector>

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "common.hpp"

namespace icp4r {

template <typename PointT>
class VoxelGrid {
   public:
    using PointCloud = pcl::PointCloud<PointT>;
    using PointCloudConstPtr = typename PointCloud::ConstPtr;

    explicit VoxelGrid(int device = 0) : h_(shared_handle(device)) {}

    void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }
    void setLeafSize(float lx, float ly, float lz) {
        if (lx != ly || ly != lz) throw std::runtime_error("icp4r::VoxelGrid: only cubic leaves are supported");
        leaf_ = lx;
    }

    void filter(PointCloud& output) {
        output.points.clear();
        if (!input_ || input_->points.empty()) {
            output.width = output.height = 0;
            return;
        }
        const int32_t n = (int32_t)input_->points.size();
        std::vector<float> out(4 * (std::size_t)n);
        int32_t cnt = 0;
        int rc;
        {
            RowLayout<PointT> lay(h_);
            rc = icp4r_voxel_grid(h_, rows(input_->points.data()), n, ICP4R_HOST, leaf_, out.data(), n, &cnt);
        }
        if (rc == ICP4R_ERR_INVALID) {  // PCL: "Leaf size is too small for the input dataset" -> output = input
            output = *input_;
            return;
        }
        check(h_, rc, "VoxelGrid::filter");
        output.points.resize((std::size_t)cnt);
        for (int32_t i = 0; i < cnt; ++i) {
            PointT p{};
            p.x = out[4 * i];
            p.y = out[4 * i + 1];
            p.z = out[4 * i + 2];
            set_intensity(p, out[4 * i + 3], has_intensity<PointT>());
            output.points[(std::size_t)i] = p;
        }
        output.width = (uint32_t)cnt;
        output.height = 1;
        output.is_dense = true;
    }

   private:
    template <typename P>
    static void set_intensity(P& p, float v, std::true_type) { p.intensity = v; }
    template <typename P>
    static void set_intensity(P&, float, std::false_type) {}

    icp4r_handle h_;
    PointCloudConstPtr input_;
    float leaf_ = 0.5f;
};

}  // namespace icp4r
