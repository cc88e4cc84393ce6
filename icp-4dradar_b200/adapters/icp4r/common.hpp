// common.hpp — shared pieces of the header-only adapters over include/icp4r.h.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "icp4r.h"

namespace icp4r {

// intensity is carried through the map when the point type has one (pcl::PointXYZI), else 0
template <typename P, typename = void>
struct has_intensity : std::false_type {};
template <typename P>
struct has_intensity<P, decltype(void(std::declval<P>().intensity))> : std::true_type {};

template <typename P>
inline float intensity_of(const P& p, std::true_type) { return p.intensity; }
template <typename P>
inline float intensity_of(const P&, std::false_type) { return 0.f; }

// The reference's clouds go to the library AS THEY LIE IN MEMORY: a std::vector / pcl::PointCloud of PointType is an
// array of sizeof(PointType)-byte rows with x, y, z in the first 12 bytes (every PCL point type starts with
// PCL_ADD_POINT4D); RowLayout tells the handle that stride (and where the intensity sits) for the calls inside its
// scope and restores the packed default afterwards. No host-side pack loop, no temporary vector: the rows are copied
// to the device unmodified and repacked there (icp4r_set_point_layout).
template <typename P>
inline int32_t w_offset(std::true_type) {
    static_assert(std::is_standard_layout<P>::value, "point type must be standard-layout");
    return (int32_t)offsetof(P, intensity);
}
template <typename P>
inline int32_t w_offset(std::false_type) { return -1; }

template <typename P>
class RowLayout {
   public:
    explicit RowLayout(icp4r_handle h) : h_(h) {
        static_assert(sizeof(P) % 4 == 0 && sizeof(P) >= 12, "point rows must be whole floats");
        static_assert(offsetof(P, x) == 0 && offsetof(P, y) == 4 && offsetof(P, z) == 8, "x, y, z must lead the point type");
        icp4r_set_point_layout(h_, (int32_t)sizeof(P), w_offset<P>(has_intensity<P>()));
    }
    ~RowLayout() { icp4r_set_point_layout(h_, 16, 12); }
    RowLayout(const RowLayout&) = delete;
    RowLayout& operator=(const RowLayout&) = delete;

   private:
    icp4r_handle h_;
};
template <typename P>
inline const float* rows(const P* first) { return reinterpret_cast<const float*>(first); }

// One handle per THREAD and device for objects that the reference constructs per frame on the stack
// (pcl::IterativeClosestPoint at iterative_closest_point.cpp:510, FastGICP at radar_odometry.cpp:399): creating a CUDA
// stream and buffers per frame would be wasted work. An icp4r handle is stateful (staging buffers, transient target map,
// error text, captured launch graphs) and serves one caller thread at a time (include/icp4r.h), so the shared one is
// thread_local: registration objects used from different threads (a ROS callback thread and a processing thread) get
// different handles and are as independent as the reference's objects; objects of one thread share one.
// The handles live until the thread ends (they are deliberately not destroyed during static destruction, when the CUDA
// runtime may already be gone).
inline icp4r_handle shared_handle(int device = 0) {
    static thread_local icp4r_handle h[16] = {nullptr};
    if (device < 0 || device >= 16) throw std::runtime_error("icp4r: bad device");
    if (!h[device]) {
        if (icp4r_create(device, &h[device]) != ICP4R_OK)
            throw std::runtime_error(std::string("icp4r_create failed: ") + icp4r_last_error(nullptr));
    }
    return h[device];
}

inline void check(icp4r_handle h, int rc, const char* what) {
    if (rc != ICP4R_OK) throw std::runtime_error(std::string(what) + ": " + icp4r_last_error(h));
}

}  // namespace icp4r
