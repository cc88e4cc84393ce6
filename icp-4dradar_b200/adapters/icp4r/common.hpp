// common.hpp — shared pieces of the header-only adapters over include/icp4r.h.
#pragma once
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

template <typename It>
inline std::vector<float> pack_xyzw(It first, It last) {
    std::vector<float> out;
    out.reserve(4 * static_cast<std::size_t>(last - first));
    for (; first != last; ++first) {
        using P = typename std::decay<decltype(*first)>::type;
        out.push_back(first->x);
        out.push_back(first->y);
        out.push_back(first->z);
        out.push_back(intensity_of(*first, has_intensity<P>()));
    }
    return out;
}

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
