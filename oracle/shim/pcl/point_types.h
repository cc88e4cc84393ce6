/* TEST INFRASTRUCTURE ONLY — minimal stand-in for <pcl/point_types.h>.
 *
 * PCL is not installed in this image.  The reference's vendored ikd-Tree
 * (/root/reference/third_party/ikd-Tree/ikd_Tree.h:11,62) needs exactly four names from PCL/Eigen:
 * pcl::PointXYZ, pcl::PointXYZI, pcl::PointXYZINormal and Eigen::aligned_allocator.  This header
 * provides them with the same field names and the same 16-byte-aligned layouts PCL uses
 * (PointXYZ 16 B, PointXYZI 32 B, PointXYZINormal 48 B) so the reference sources compile
 * unmodified from where they lie (see oracle/Makefile, target _ref/libikd_ref.so).
 * Nothing in the product (icp-4dradar_b200/) includes this file.
 */
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring> /* the real PCL header pulls these in transitively; ikd_Tree.cpp relies on it */
#include <vector>
#include <memory>

namespace pcl {

struct alignas(16) PointXYZ {
    float x, y, z, _pad;
    PointXYZ() : x(0.f), y(0.f), z(0.f), _pad(1.f) {}
    PointXYZ(float px, float py, float pz) : x(px), y(py), z(pz), _pad(1.f) {}
};

struct alignas(16) PointXYZI {
    float x, y, z, _pad;
    float intensity;
    float _pad2[3];
    PointXYZI() : x(0.f), y(0.f), z(0.f), _pad(1.f), intensity(0.f), _pad2{0.f, 0.f, 0.f} {}
};

struct alignas(16) PointXYZINormal {
    float x, y, z, _pad;
    float normal_x, normal_y, normal_z, _padn;
    float intensity, curvature;
    float _pad2[2];
    PointXYZINormal()
        : x(0.f), y(0.f), z(0.f), _pad(1.f), normal_x(0.f), normal_y(0.f), normal_z(0.f), _padn(0.f),
          intensity(0.f), curvature(0.f), _pad2{0.f, 0.f} {}
};

}  // namespace pcl

namespace Eigen {
/* std::allocator honours alignas(16) through aligned operator new (C++17) and glibc malloc is
 * 16-byte aligned anyway, which is all Eigen::aligned_allocator guarantees. */
template <typename T>
using aligned_allocator = std::allocator<T>;
}  // namespace Eigen
