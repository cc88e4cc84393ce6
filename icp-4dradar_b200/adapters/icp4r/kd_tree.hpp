// kd_tree.hpp — drop-in for the reference's KD_TREE<PointType> call shape on top of libicp4r_cuda.
//
// Mirrors the public interface of /root/reference/third_party/ikd-Tree/ikd_Tree.h:227-251 as used at
// /root/reference/src/radar_odometry.cpp:92 (ctor), :347 (Build), :348 (set_downsample_param), :390 (Add_Points),
// :396 (Sector_Search): same names, argument meaning, output-vector conventions (the callee clears them,
// ikd_Tree.cpp:371,390-391) and return values.  The k-d tree itself is gone: points live in a device voxel
// grid; there is no rebuild thread, so delete_param / balance_param are accepted and ignored.
//
// Differences a caller can observe:
//   * exact ties in distance go to the lowest insertion index (the reference: traversal order);
//   * Nearest_Search ranks at most ICP4R_MAX_K (16) neighbours per query: a larger k_nearest is clamped to 16
//     (the reference accepts any k; its call sites use 5);
//   * failures of the CUDA runtime surface as std::runtime_error (the reference never throws from these calls);
//     a query before Build returns empty vectors like the reference does;
//   * Nearest_Search is also offered for a whole batch of queries (one launch instead of N calls).
#pragma once
#include <cmath>
#include <memory>
#include <vector>

#include <pcl/point_types.h>

#include "common.hpp"

struct BoxPointType {  // same layout as ikd_Tree.h:32-35
    float vertex_min[3];
    float vertex_max[3];
};
static_assert(sizeof(BoxPointType) == 6 * sizeof(float), "a vector<BoxPointType> is passed to the C ABI as n x 6 floats");

namespace icp4r {

template <typename PointType>
class KD_TREE {
   public:
    using PointVector = std::vector<PointType, Eigen::aligned_allocator<PointType>>;
    using Ptr = std::shared_ptr<KD_TREE<PointType>>;

    explicit KD_TREE(float delete_param = 0.5f, float balance_param = 0.6f, float box_length = 0.2f, int device = 0)
        : downsample_size_(box_length) {
        (void)delete_param;
        (void)balance_param;
        check(nullptr, icp4r_create(device, &h_), "icp4r_create");
    }
    ~KD_TREE() {
        if (h_) icp4r_destroy(h_);
    }
    KD_TREE(const KD_TREE&) = delete;
    KD_TREE& operator=(const KD_TREE&) = delete;

    void Set_delete_criterion_param(float) {}
    void Set_balance_criterion_param(float) {}
    void set_downsample_param(float box_length) {
        downsample_size_ = box_length;
        check(h_, icp4r_map_set_downsample(h_, box_length), "set_downsample_param");
    }
    void InitializeKDTree(float = 0.5f, float = 0.7f, float box_length = 0.2f) { set_downsample_param(box_length); }

    int size() {
        int32_t s = 0, v = 0;
        check(h_, icp4r_map_size(h_, &s, &v), "size");
        return s;
    }
    int validnum() {
        int32_t s = 0, v = 0;
        check(h_, icp4r_map_size(h_, &s, &v), "validnum");
        return v;
    }

    void Build(PointVector point_cloud) {  // by value, like the reference
        mirror_.assign(point_cloud.begin(), point_cloud.end());
        {
            RowLayout<PointType> lay(h_);
            check(h_, icp4r_map_build(h_, rows(point_cloud.data()), (int32_t)point_cloud.size(), ICP4R_HOST, 0.f), "Build");
        }
        check(h_, icp4r_map_set_downsample(h_, downsample_size_), "Build");
    }

    int Add_Points(PointVector& PointToAdd, bool downsample_on) {
        int32_t replaced = 0;
        {
            RowLayout<PointType> lay(h_);
            check(h_, icp4r_map_add_points(h_, rows(PointToAdd.data()), (int32_t)PointToAdd.size(), ICP4R_HOST, downsample_on ? 1 : 0, &replaced),
                  "Add_Points");
        }
        mirror_.insert(mirror_.end(), PointToAdd.begin(), PointToAdd.end());
        return replaced;
    }

    void Nearest_Search(PointType point, int k_nearest, PointVector& Nearest_Points, std::vector<float>& Point_Distance,
                        double max_dist = INFINITY) {
        PointVector().swap(Nearest_Points);
        std::vector<float>().swap(Point_Distance);
        if (k_nearest <= 0 || mirror_.empty()) return;  // before Build: empty result, like the reference (ikd_Tree.cpp:370-371)
        // the reference accepts any k; the library ranks up to ICP4R_MAX_K per query: larger requests are clamped (a radar
        // map query for more than 16 neighbours does not occur in the reference, radar_odometry.cpp uses 5)
        const int k = k_nearest < ICP4R_MAX_K ? k_nearest : ICP4R_MAX_K;
        // max_dist: +inf = ungated. The library encodes "ungated" as <= 0, the reference treats 0 as "exact hits only"
        // (dist <= 0): map it to the smallest positive gate so that only exact hits pass
        const double gate = std::isinf(max_dist) ? 0.0 : (max_dist > 0.0 ? max_dist : 1e-30);
        const float q[4] = {point.x, point.y, point.z, 0.f};
        std::vector<int32_t> idx(k);
        std::vector<float> d2(k);
        int32_t found = 0;
        check(h_, icp4r_map_knn(h_, q, 1, ICP4R_HOST, k, gate, idx.data(), d2.data(), &found), "Nearest_Search");
        for (int i = 0; i < found; ++i) {  // ascending, like ikd_Tree.cpp:392-396
            Nearest_Points.push_back(mirror_[idx[i]]);
            Point_Distance.push_back(d2[i]);
        }
    }

    // batch form: rows of k indices into insertion order (-1 = none) and squared distances
    void Nearest_Search_Batch(const PointVector& queries, int k_nearest, std::vector<int32_t>& indices, std::vector<float>& sq_dist,
                              std::vector<int32_t>& found, double max_dist = INFINITY) {
        RowLayout<PointType> lay(h_);
        indices.assign(queries.size() * k_nearest, -1);
        sq_dist.assign(queries.size() * k_nearest, INFINITY);
        found.assign(queries.size(), 0);
        check(h_, icp4r_map_knn(h_, rows(queries.data()), (int32_t)queries.size(), ICP4R_HOST, k_nearest, std::isinf(max_dist) ? 0.0 : max_dist,
                               indices.data(), sq_dist.data(), found.data()),
              "Nearest_Search_Batch");
    }

    void Sector_Search(PointType point, const float radius, const float heading, PointVector& Storage) {
        Storage.clear();
        const float c[3] = {point.x, point.y, point.z};
        std::vector<int32_t> idx(mirror_.size() ? mirror_.size() : 1);
        int32_t n = 0;
        check(h_, icp4r_map_sector(h_, c, radius, heading, ICP4R_HOST, idx.data(), (int32_t)idx.size(), &n), "Sector_Search");
        for (int i = 0; i < n && i < (int)idx.size(); ++i) Storage.push_back(mirror_[idx[i]]);
    }

    // ikd_Tree.h:243-249 — the searches return point copies like the reference, in unspecified order
    void Box_Search(const BoxPointType& Box_of_Point, PointVector& Storage) {
        Storage.clear();
        std::vector<int32_t> idx(mirror_.size() ? mirror_.size() : 1);
        int32_t n = 0;
        check(h_, icp4r_map_box_search(h_, Box_of_Point.vertex_min, Box_of_Point.vertex_max, ICP4R_HOST, idx.data(), (int32_t)idx.size(), &n),
              "Box_Search");
        for (int i = 0; i < n && i < (int)idx.size(); ++i) Storage.push_back(mirror_[idx[i]]);
    }

    void Radius_Search(PointType point, const float radius, PointVector& Storage) {
        Storage.clear();
        const float c[3] = {point.x, point.y, point.z};
        std::vector<int32_t> idx(mirror_.size() ? mirror_.size() : 1);
        int32_t n = 0;
        check(h_, icp4r_map_radius_search(h_, c, radius, ICP4R_HOST, idx.data(), (int32_t)idx.size(), &n), "Radius_Search");
        for (int i = 0; i < n && i < (int)idx.size(); ++i) Storage.push_back(mirror_[idx[i]]);
    }

    int Delete_Point_Boxes(std::vector<BoxPointType>& BoxPoints) {
        const std::vector<uint8_t> before = valid_mask();
        int32_t n = 0;
        check(h_, icp4r_map_delete_boxes(h_, BoxPoints.empty() ? nullptr : BoxPoints[0].vertex_min, (int32_t)BoxPoints.size(), &n),
              "Delete_Point_Boxes");
        if (n > 0) note_removed(before);
        return n;
    }

    void Add_Point_Boxes(std::vector<BoxPointType>& BoxPoints) {
        check(h_, icp4r_map_add_boxes(h_, BoxPoints.empty() ? nullptr : BoxPoints[0].vertex_min, (int32_t)BoxPoints.size(), nullptr),
              "Add_Point_Boxes");
        if (!removed_pending_.empty()) {  // a revived point was never "removed"
            const std::vector<uint8_t> now = valid_mask();
            std::vector<int32_t> keep;
            for (int32_t i : removed_pending_)
                if (!now[(std::size_t)i]) keep.push_back(i);
            removed_pending_.swap(keep);
        }
    }

    void Delete_Points(PointVector& PointToDel) {
        if (PointToDel.empty()) return;
        const std::vector<uint8_t> before = valid_mask();
        int32_t n = 0;
        {
            RowLayout<PointType> lay(h_);
            check(h_, icp4r_map_delete_points(h_, rows(PointToDel.data()), (int32_t)PointToDel.size(), ICP4R_HOST, &n), "Delete_Points");
        }
        if (n > 0) note_removed(before);
    }

    // ikd_Tree.h:246 — flatten(root, Storage, NOT_RECORD) appends every non-deleted point of the subtree; there are no tree
    // nodes here, so the adapter's form takes no node and appends every valid point of the map (insertion order; the
    // reference: traversal order).
    void flatten(PointVector& Storage) {
        const std::vector<uint8_t> v = valid_mask();
        for (std::size_t i = 0; i < v.size(); ++i)
            if (v[i]) Storage.push_back(mirror_[i]);
    }

    // ikd_Tree.h:247, ikd_Tree.cpp:567-579 — appends the points that Delete_Points / Delete_Point_Boxes removed since the
    // last call (not the ones down-sampling replaced, ikd_Tree.cpp:1393) and forgets them. The reference reports a deleted
    // point only once a rebuild has purged its node (timing dependent); here it is reported right after the delete call —
    // the same set once the reference's rebuilds have caught up.
    void acquire_removed_points(PointVector& removed_points) {
        for (int32_t i : removed_pending_) removed_points.push_back(mirror_[(std::size_t)i]);
        removed_pending_.clear();
    }

    BoxPointType tree_range() {
        float r[6];
        check(h_, icp4r_map_range(h_, r), "tree_range");
        BoxPointType b;
        for (int a = 0; a < 3; ++a) {
            b.vertex_min[a] = r[a];
            b.vertex_max[a] = r[3 + a];
        }
        return b;
    }

    const PointType& point_at(int index) const { return mirror_[index]; }
    icp4r_handle native_handle() { return h_; }

   private:
    std::vector<uint8_t> valid_mask() {
        std::vector<uint8_t> v(mirror_.size());
        if (!v.empty()) check(h_, icp4r_map_points(h_, ICP4R_HOST, nullptr, v.data(), (int32_t)v.size()), "valid mask");
        return v;
    }
    void note_removed(const std::vector<uint8_t>& before) {
        const std::vector<uint8_t> now = valid_mask();
        for (std::size_t i = 0; i < now.size() && i < before.size(); ++i)
            if (before[i] && !now[i]) removed_pending_.push_back((int32_t)i);
    }

    icp4r_handle h_ = nullptr;
    float downsample_size_;
    std::vector<int32_t> removed_pending_;  // removed by a delete call, not yet handed out by acquire_removed_points
    std::vector<PointType, Eigen::aligned_allocator<PointType>> mirror_;  // host copies, returned by value like the reference
};

}  // namespace icp4r
