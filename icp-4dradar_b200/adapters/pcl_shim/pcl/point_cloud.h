// Minimal pcl::PointCloud stand-in (see point_types.h in this directory for when it is used).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#include "point_types.h"

namespace pcl {
template <typename PointT>
class PointCloud {
   public:
    using Ptr = std::shared_ptr<PointCloud<PointT>>;
    using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
    std::vector<PointT, Eigen::aligned_allocator<PointT>> points;
    std::uint32_t width = 0, height = 0;
    bool is_dense = true;
    std::size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); }
    void push_back(const PointT& p) { points.push_back(p); }
    void resize(std::size_t n) { points.resize(n); }
    PointT& operator[](std::size_t i) { return points[i]; }
    const PointT& operator[](std::size_t i) const { return points[i]; }
};
}  // namespace pcl
