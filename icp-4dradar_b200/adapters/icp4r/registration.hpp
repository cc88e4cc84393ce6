// registration.hpp — drop-ins for the two registration objects the reference constructs per frame:
//
//   pcl::IterativeClosestPoint<PointXYZI, PointXYZI>     /root/reference/src/iterative_closest_point.cpp:510-521
//   fast_gicp::FastGICPSingleThread<PointXYZI, PointXYZI> /root/reference/src/radar_odometry.cpp:399-411
//
// Same method names and meaning (setInputSource / setInputTarget / align / hasConverged / getFitnessScore /
// getFinalTransformation, plus clearSource / clearTarget / setCorrespondenceRandomness for the GICP shape).
// Errors follow the reference's behaviour: nothing throws out of align(); a failed call leaves
// hasConverged() == false and the final transformation at the initial guess.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstring>
#include <memory>

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#if __has_include(<Eigen/Core>)
#include <Eigen/Core>
#define ICP4R_HAVE_EIGEN 1
#endif

#include "common.hpp"

namespace icp4r {

#ifdef ICP4R_HAVE_EIGEN
using Matrix4f = Eigen::Matrix4f;
inline Matrix4f to_matrix4f(const double T[16]) {
    Matrix4f m;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) m(i, j) = (float)T[4 * i + j];
    return m;
}
inline void from_matrix4f(const Matrix4f& m, double T[16]) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) T[4 * i + j] = (double)m(i, j);
}
#else
struct Matrix4f {  // stand-in when Eigen is not installed (this image); same element access as Eigen
    float v[16];
    float& operator()(int i, int j) { return v[4 * i + j]; }
    float operator()(int i, int j) const { return v[4 * i + j]; }
    static Matrix4f Identity() {
        Matrix4f m;
        for (int i = 0; i < 16; ++i) m.v[i] = (i % 5 == 0) ? 1.f : 0.f;
        return m;
    }
};
inline Matrix4f to_matrix4f(const double T[16]) {
    Matrix4f m;
    for (int i = 0; i < 16; ++i) m.v[i] = (float)T[i];
    return m;
}
inline void from_matrix4f(const Matrix4f& m, double T[16]) {
    for (int i = 0; i < 16; ++i) T[i] = (double)m.v[i];
}
#endif

template <typename PointSource, typename PointTarget>
class RegistrationBase {
   public:
    using PointCloudSource = pcl::PointCloud<PointSource>;
    using PointCloudTarget = pcl::PointCloud<PointTarget>;
    using PointCloudSourceConstPtr = std::shared_ptr<const PointCloudSource>;
    using PointCloudTargetConstPtr = std::shared_ptr<const PointCloudTarget>;

    explicit RegistrationBase(int device = 0) : h_(shared_handle(device)) {
        icp4r_default_opts(&opts_);
        std::memset(&res_, 0, sizeof(res_));
        for (int i = 0; i < 16; ++i) T_[i] = (i % 5 == 0) ? 1.0 : 0.0;
    }

    // clouds are retained, not copied, like PCL
    void setInputSource(const PointCloudSourceConstPtr& cloud) { src_ = cloud; }
    void setInputTarget(const PointCloudTargetConstPtr& cloud) { tgt_ = cloud; }
    void setMaximumIterations(int n) { opts_.max_iterations = n; }
    void setMaxCorrespondenceDistance(double d) { opts_.max_corr_dist = (d >= 1e150) ? 0.0 : d; }
    void setTransformationEpsilon(double e) { opts_.trans_eps = e; }
    void setEuclideanFitnessEpsilon(double e) { opts_.mse_abs_eps = e; }
    int getMaximumIterations() const { return opts_.max_iterations; }

    void align(PointCloudSource& output) { align(output, Matrix4f::Identity()); }
    void align(PointCloudSource& output, const Matrix4f& guess) {
        converged_ = false;
        from_matrix4f(guess, opts_.T0);
        std::memcpy(T_, opts_.T0, sizeof(T_));
        if (!src_ || !tgt_ || src_->points.empty() || tgt_->points.empty()) return;  // PCL prints and returns
        static_assert(sizeof(PointSource) == sizeof(PointTarget), "source and target rows share one layout per call");
        const int32_t n = (int32_t)src_->points.size();
        std::vector<float> out(4 * (std::size_t)n);
        int rc, rc_out;
        {
            RowLayout<PointSource> lay(h_);  // the clouds are read as they lie in memory
            rc = icp4r_register(h_, rows(src_->points.data()), n, rows(tgt_->points.data()), (int32_t)tgt_->points.size(), ICP4R_HOST, &opts_, T_,
                                &res_, nullptr);
            // output = source transformed by the final pose (pcl::transformPointCloud at the end of align)
            rc_out = rc == ICP4R_OK ? icp4r_transform_points(h_, T_, rows(src_->points.data()), n, ICP4R_HOST, out.data()) : rc;
        }
        if (rc != ICP4R_OK) {
            last_error_ = icp4r_last_error(h_);
            return;
        }
        converged_ = res_.converged != 0;
        if (rc_out == ICP4R_OK) {
            output.points.assign(src_->points.begin(), src_->points.end());
            for (std::size_t i = 0; i < output.points.size(); ++i) {
                output.points[i].x = out[4 * i];
                output.points[i].y = out[4 * i + 1];
                output.points[i].z = out[4 * i + 2];
            }
        }
    }

    bool hasConverged() const { return converged_; }
    // mean squared nearest-neighbour distance after the final transform; the library computes it inside
    // align() (PCL runs one more full 1-NN pass per call, twice per frame at iterative_closest_point.cpp:516,520)
    double getFitnessScore(double /*max_range*/ = DBL_MAX) const { return res_.n_fitness > 0 ? res_.fitness : DBL_MAX; }
    Matrix4f getFinalTransformation() const { return to_matrix4f(T_); }
    const double* getFinalTransformationDouble() const { return T_; }
    int getIterations() const { return res_.iterations; }
    const std::string& lastError() const { return last_error_; }

   protected:
    icp4r_handle h_;
    icp4r_opts opts_;
    icp4r_result res_;
    double T_[16];
    bool converged_ = false;
    PointCloudSourceConstPtr src_;
    PointCloudTargetConstPtr tgt_;
    std::string last_error_;
};

// pcl::IterativeClosestPoint shape: 1-NN, closed-form SVD update, PCL defaults (10 iterations, ungated,
// |dMSE| < 1e-12 exit) — iterative_closest_point.cpp:510-514 sets nothing else.
template <typename PointSource, typename PointTarget>
class IterativeClosestPoint : public RegistrationBase<PointSource, PointTarget> {
   public:
    explicit IterativeClosestPoint(int device = 0) : RegistrationBase<PointSource, PointTarget>(device) {
        this->opts_.residual = ICP4R_P2P_SVD;
        this->opts_.max_iterations = 10;
        this->opts_.early_exit = 1;
        this->opts_.mse_abs_eps = 1e-12;
        this->opts_.max_corr_dist = 0.0;
    }
};

// fast_gicp::FastGICPSingleThread shape on the library's GICP cost (plane-regularised k-NN covariances of both
// clouds, 1-NN Mahalanobis residual, Levenberg-Marquardt), with fast_gicp's defaults: k = 20 (capped at
// ICP4R_MAX_K = 16), 64 iterations, rotation / translation epsilons 2e-3 / 5e-4, ungated.
template <typename PointSource, typename PointTarget>
class FastGICPSingleThread : public RegistrationBase<PointSource, PointTarget> {
   public:
    explicit FastGICPSingleThread(int device = 0) : RegistrationBase<PointSource, PointTarget>(device) {
        this->opts_.residual = ICP4R_GICP;
        this->opts_.k = 20;  // fast_gicp default k_correspondences
        if (this->opts_.k > ICP4R_MAX_K) this->opts_.k = ICP4R_MAX_K;
        this->opts_.max_iterations = 64;
        this->opts_.early_exit = 1;
        this->opts_.rot_eps = 2e-3;
        this->opts_.trans_eps = 5e-4;
        this->opts_.max_corr_dist = 0.0;
    }
    void clearSource() { this->src_.reset(); }
    void clearTarget() { this->tgt_.reset(); }
    void setCorrespondenceRandomness(int k) { this->opts_.k = k < 3 ? 3 : (k > ICP4R_MAX_K ? ICP4R_MAX_K : k); }
};

}  // namespace icp4r
