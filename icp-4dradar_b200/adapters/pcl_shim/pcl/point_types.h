// Minimal stand-ins for the PCL / Eigen names the adapters touch, used ONLY when the real headers are not
// installed (this build image has neither PCL nor Eigen). With PCL present, do not add this directory to the
// include path: the adapters then compile against the real <pcl/point_types.h> / <pcl/point_cloud.h>.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

namespace pcl {
struct alignas(16) PointXYZ {
    float x = 0.f, y = 0.f, z = 0.f, _pad = 1.f;
};
struct alignas(16) PointXYZI {
    float x = 0.f, y = 0.f, z = 0.f, _pad = 1.f;
    float intensity = 0.f;
    float _pad2[3] = {0.f, 0.f, 0.f};
};
}  // namespace pcl

namespace Eigen {
template <typename T>
using aligned_allocator = std::allocator<T>;
}
