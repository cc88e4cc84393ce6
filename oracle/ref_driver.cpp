/* ref_driver.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * A thin extern "C" wrapper around the REFERENCE's own KD_TREE<pcl::PointXYZI>
 * (/root/reference/third_party/ikd-Tree/ikd_Tree.h:227-251), compiled unmodified from where it
 * lies together with this file into oracle/_ref/libikd_ref.so (oracle/Makefile).  It pins the
 * oracle's kNN / Add_Points / Sector_Search restatements and serves as the `kind: "reference"`
 * CPU baseline.  No reference source is copied into this repository.
 *
 * Index recovery: the tree copies whole points into its nodes (ikd_Tree.cpp:622,830) and out again
 * (:393), so the point index is carried bit-cast in PointXYZI::intensity (valid for
 * idx < 0x7F800000: the bit patterns are finite floats/denormals that are only ever moved).
 */
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "ikd_Tree.h"

using Tree = KD_TREE<pcl::PointXYZI>;
using PV = Tree::PointVector;

static inline pcl::PointXYZI mk(const float* xyzw, int32_t index) {
    pcl::PointXYZI p;
    p.x = xyzw[0];
    p.y = xyzw[1];
    p.z = xyzw[2];
    std::memcpy(&p.intensity, &index, 4);
    return p;
}
static inline int32_t idx_of(const pcl::PointXYZI& p) {
    int32_t i;
    std::memcpy(&i, &p.intensity, 4);
    return i;
}

extern "C" {

void* ikdref_create(float delete_param, float balance_param, float box_length) {
    return new Tree(delete_param, balance_param, box_length); /* ~64 MB object: heap only */
}
void ikdref_destroy(void* h) { delete static_cast<Tree*>(h); }

void ikdref_build(void* h, const float* xyzw, int n, int index_base) {
    PV v;
    v.reserve(n);
    for (int i = 0; i < n; ++i) v.push_back(mk(xyzw + 4 * (size_t)i, index_base + i));
    static_cast<Tree*>(h)->Build(v);
}

int ikdref_add_points(void* h, const float* xyzw, int n, int index_base, int downsample_on) {
    PV v;
    v.reserve(n);
    for (int i = 0; i < n; ++i) v.push_back(mk(xyzw + 4 * (size_t)i, index_base + i));
    return static_cast<Tree*>(h)->Add_Points(v, downsample_on != 0);
}

void ikdref_set_downsample(void* h, float box) { static_cast<Tree*>(h)->set_downsample_param(box); }
int ikdref_size(void* h) { return static_cast<Tree*>(h)->size(); }
int ikdref_validnum(void* h) { return static_cast<Tree*>(h)->validnum(); }

/* batch of Nearest_Search calls; nthreads > 1 fans the queries out (Nearest_Search is reader-safe,
 * ikd_Tree.cpp:372-388). Rows are [nq,k]; unfilled slots idx=-1, d2=inf. */
void ikdref_knn(void* h, const float* q, int nq, int k, double max_dist, int32_t* idx, float* d2,
                int32_t* found, int nthreads) {
    Tree* t = static_cast<Tree*>(h);
    const double md = (max_dist > 0.0) ? max_dist : INFINITY;
    auto work = [&](int lo, int hi) {
        PV pts;
        std::vector<float> dd;
        for (int i = lo; i < hi; ++i) {
            pcl::PointXYZI p = mk(q + 4 * (size_t)i, 0);
            t->Nearest_Search(p, k, pts, dd, md);
            const int f = (int)pts.size();
            for (int s = 0; s < k; ++s) {
                idx[(size_t)i * k + s] = s < f ? idx_of(pts[s]) : -1;
                d2[(size_t)i * k + s] = s < f ? dd[s] : INFINITY;
            }
            if (found) found[i] = f;
        }
    };
    if (nthreads <= 1 || nq < 64) {
        work(0, nq);
        return;
    }
    std::vector<std::thread> th;
    const int chunk = (nq + nthreads - 1) / nthreads;
    for (int w = 0; w < nthreads; ++w) {
        const int lo = w * chunk, hi = lo + chunk < nq ? lo + chunk : nq;
        if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto& x : th) x.join();
}

/* orc_knn_fn-compatible callback (oracle.h): ctx = {tree handle, nthreads} */
struct ikdref_cb_ctx {
    void* tree;
    int nthreads;
};
void ikdref_knn_cb(void* ctx, const float* q, int nq, int k, double max_dist, int32_t* idx, float* d2,
                   int32_t* found) {
    ikdref_cb_ctx* c = static_cast<ikdref_cb_ctx*>(ctx);
    ikdref_knn(c->tree, q, nq, k, max_dist, idx, d2, found, c->nthreads);
}

int ikdref_sector(void* h, const float centre[3], float radius, float heading, int32_t* idx_out, int cap) {
    pcl::PointXYZI c = mk(centre, 0);
    PV out;
    static_cast<Tree*>(h)->Sector_Search(c, radius, heading, out);
    const int n = (int)out.size();
    for (int i = 0; i < n && i < cap; ++i) idx_out[i] = idx_of(out[i]);
    return n;
}

int ikdref_radius(void* h, const float centre[3], float radius, int32_t* idx_out, int cap) {
    pcl::PointXYZI c = mk(centre, 0);
    PV out;
    static_cast<Tree*>(h)->Radius_Search(c, radius, out);
    const int n = (int)out.size();
    for (int i = 0; i < n && i < cap; ++i) idx_out[i] = idx_of(out[i]);
    return n;
}

int ikdref_box(void* h, const float bmin[3], const float bmax[3], int32_t* idx_out, int cap) {
    BoxPointType b;
    for (int a = 0; a < 3; ++a) {
        b.vertex_min[a] = bmin[a];
        b.vertex_max[a] = bmax[a];
    }
    PV out;
    static_cast<Tree*>(h)->Box_Search(b, out);
    const int n = (int)out.size();
    for (int i = 0; i < n && i < cap; ++i) idx_out[i] = idx_of(out[i]);
    return n;
}

int ikdref_delete_boxes(void* h, const float* boxes6, int nb) {
    std::vector<BoxPointType> v(nb);
    for (int i = 0; i < nb; ++i)
        for (int a = 0; a < 3; ++a) {
            v[i].vertex_min[a] = boxes6[6 * i + a];
            v[i].vertex_max[a] = boxes6[6 * i + 3 + a];
        }
    return static_cast<Tree*>(h)->Delete_Point_Boxes(v);
}

void ikdref_add_boxes(void* h, const float* boxes6, int nb) {
    std::vector<BoxPointType> v(nb);
    for (int i = 0; i < nb; ++i)
        for (int a = 0; a < 3; ++a) {
            v[i].vertex_min[a] = boxes6[6 * i + a];
            v[i].vertex_max[a] = boxes6[6 * i + 3 + a];
        }
    static_cast<Tree*>(h)->Add_Point_Boxes(v);
}

void ikdref_delete_points(void* h, const float* xyzw, int n) {
    PV v;
    v.reserve(n);
    for (int i = 0; i < n; ++i) v.push_back(mk(xyzw + 4 * (size_t)i, 0));
    static_cast<Tree*>(h)->Delete_Points(v);
}

/* indices of all valid points currently in the tree */
int ikdref_flatten(void* h, int32_t* idx_out, int cap) {
    Tree* t = static_cast<Tree*>(h);
    PV out;
    if (t->Root_Node == nullptr) return 0;
    t->flatten(t->Root_Node, out, NOT_RECORD);
    const int n = (int)out.size();
    for (int i = 0; i < n && i < cap; ++i) idx_out[i] = idx_of(out[i]);
    return n;
}

void ikdref_range(void* h, float out6[6]) {
    BoxPointType b = static_cast<Tree*>(h)->tree_range();
    for (int a = 0; a < 3; ++a) {
        out6[a] = b.vertex_min[a];
        out6[3 + a] = b.vertex_max[a];
    }
}

int ikdref_hw_threads(void) {
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

} /* extern "C" */
