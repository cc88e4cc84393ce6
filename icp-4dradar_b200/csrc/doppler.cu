// doppler.cu — Doppler static-point filter and ego-velocity estimate for 4D-radar frames: the step right before
// registration in the reference's scan-to-scan node (/root/reference/src/iterative_closest_point.cpp:85-128 fitSineRansac,
// :354-386 per-point angles, :387-407 static/dynamic split, :412-431 least squares), its second CPU hot loop
// (0.2 N hypotheses x N points of double-precision trigonometry per frame).
//
// One warp scores one hypothesis (lanes stride the points); the best hypothesis (highest score, first on ties, as
// the reference's strict `>`), the split and the 3x3 normal equations follow in one single-block kernel.
// Deliberate deviations from the reference: sample indices come from a counter-based generator
// with a caller-supplied seed (reference: unseeded std::random_device and an inclusive [0, n] index range that can
// read past the end); alpha / beta are double atan2 / asin rounded to float.
#include <cmath>

#include "ctx.h"
#include "device_math.cuh"

namespace icp4r {

#define ICP4R_DEG2RAD(x) ((x) * 0.017453293)  // pcl_macros.h DEG2RAD, as used at iterative_closest_point.cpp:106-108

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct DopplerPoint {  // per point, prepared once
    double alpha;      // DEG2RAD(arfa)
    double cbv;        // cos(DEG2RAD(beta)) * v_r
    double cb;         // cos(DEG2RAD(beta))
    double sb;         // sin(DEG2RAD(beta))
};

struct DopplerOut {
    double A, b, score;
    double v[3];
    int n_static;
    int best_iteration;
};

__global__ void __launch_bounds__(256) doppler_prep_kernel(const float* __restrict__ rec, int n, DopplerPoint* __restrict__ dp, int iterations,
                                                           int* __restrict__ scores) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = i; it < iterations; it += gridDim.x * blockDim.x) scores[it] = 0;  // the scoring kernel adds partial counts
    if (i >= n) return;
    const float x = rec[5 * (size_t)i], y = rec[5 * (size_t)i + 1], z = rec[5 * (size_t)i + 2], vr = rec[5 * (size_t)i + 4];
    const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    const float arfa = (float)((double)__fmul_rn((float)atan2((double)y, (double)x), 180.f) / M_PI);
    const float beta = (float)((double)__fmul_rn((float)asin((double)__fdiv_rn(z, dist)), 180.f) / M_PI);
    DopplerPoint d;
    d.alpha = ICP4R_DEG2RAD(arfa);
    d.cb = cos(ICP4R_DEG2RAD(beta));
    d.sb = sin(ICP4R_DEG2RAD(beta));
    d.cbv = d.cb * (double)vr;
    dp[i] = d;
}

// one warp per (hypothesis, slice of the points): blockIdx.y selects one of gridDim.y slices. A frame has 0.2 N hypotheses
// (800 for a 4,000-point frame = 100 blocks of 8 warps: half the SMs idle with one warp per hypothesis); the counts are
// integers, so the partition does not change the result.
__global__ void __launch_bounds__(256) doppler_score_kernel(const float* __restrict__ rec, const DopplerPoint* __restrict__ dp, int n, int iterations,
                                                            uint64_t seed, double sigma, int* __restrict__ scores, double* __restrict__ As,
                                                            double* __restrict__ bs) {
    const int lane = threadIdx.x & 31;
    const int it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (it >= iterations) return;
    const int i1 = (int)(splitmix64(seed + 2ull * (uint64_t)it) % (uint64_t)n);
    const int i2 = (int)(splitmix64(seed + 2ull * (uint64_t)it + 1ull) % (uint64_t)n);
    const DopplerPoint p1 = dp[i1], p2 = dp[i2];
    const double v1 = (double)rec[5 * (size_t)i1 + 4], v2 = (double)rec[5 * (size_t)i2 + 4];
    const double k = (v1 * p1.cb) / (v2 * p2.cb);
    const double b = atan((cos(p1.alpha) - k * cos(p2.alpha)) / (sin(p1.alpha) - k * sin(p2.alpha)));
    const double A = p1.cb * v1 / cos(p1.alpha + b);
    int sc = 0;
    const int per = (n + (int)gridDim.y - 1) / (int)gridDim.y;
    const int j0 = (int)blockIdx.y * per, j1 = min(n, j0 + per);
    for (int j = j0 + lane; j < j1; j += 32) {
        const DopplerPoint pj = dp[j];
        const double delta = pj.cbv - (A * cos(pj.alpha + b));
        if (fabs(delta) < sigma) ++sc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sc += __shfl_xor_sync(FULL, sc, o);
    if (lane == 0) {
        if (sc) atomicAdd(scores + it, sc);
        if (blockIdx.y == 0) {
            As[it] = A;
            bs[it] = b;
        }
    }
}

// single block: argmax (first maximum), split, normal equations in a fixed summation order, 3x3 solve
__global__ void __launch_bounds__(1024) doppler_final_kernel(const float* __restrict__ rec, const DopplerPoint* __restrict__ dp, int n, int iterations,
                                                             const int* __restrict__ scores, const double* __restrict__ As,
                                                             const double* __restrict__ bs, double split, uint8_t* __restrict__ mask,
                                                             DopplerOut* __restrict__ out) {
    __shared__ unsigned long long s_best[32];
    __shared__ double s_sum[32][10];
    __shared__ double s_model[2];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // best = max score, lowest iteration: key = (score << 32) | (0xffffffff - it), maximised
    unsigned long long best = 0ull;
    for (int it = tid; it < iterations; it += 1024) {
        const int sc = scores[it];
        if (sc > 0) {
            const unsigned long long key = ((unsigned long long)(unsigned)sc << 32) | (unsigned long long)(0xffffffffu - (unsigned)it);
            best = key > best ? key : best;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ob = __shfl_xor_sync(FULL, best, o);
        best = ob > best ? ob : best;
    }
    if (lane == 0) s_best[w] = best;
    __syncthreads();
    if (tid == 0) {
        unsigned long long bb = 0ull;
        for (int k = 0; k < 32; ++k) bb = s_best[k] > bb ? s_best[k] : bb;
        const int sc = (int)(bb >> 32);
        const int it = sc > 0 ? (int)(0xffffffffu - (unsigned)(bb & 0xffffffffull)) : -1;
        s_model[0] = it >= 0 ? As[it] : 0.0;  // the reference leaves A = b = 0 when no hypothesis scores
        s_model[1] = it >= 0 ? bs[it] : 0.0;
        out->A = s_model[0];
        out->b = s_model[1];
        out->score = (double)sc;
        out->best_iteration = it;
    }
    __syncthreads();
    const double A = s_model[0], b = s_model[1];
    double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // K^T K (6), K^T v (3), count
    for (int j = tid; j < n; j += 1024) {
        const DopplerPoint pj = dp[j];
        const double delta = pj.cbv - (A * cos(pj.alpha + b));
        const bool is_static = !(delta > split);  // the SIGNED test of :394
        if (mask) mask[j] = is_static ? 1 : 0;
        if (!is_static) continue;
        const double vr = (double)rec[5 * (size_t)j + 4];
        const double k0 = cos(pj.alpha) * pj.cb, k1 = sin(pj.alpha) * pj.cb, k2 = pj.sb;
        acc[0] += k0 * k0; acc[1] += k0 * k1; acc[2] += k0 * k2; acc[3] += k1 * k1; acc[4] += k1 * k2; acc[5] += k2 * k2;
        acc[6] += k0 * vr; acc[7] += k1 * vr; acc[8] += k2 * vr;
        acc[9] += 1.0;
    }
#pragma unroll
    for (int v = 0; v < 10; ++v) {
        double x = acc[v];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
        if (lane == 0) s_sum[w][v] = x;
    }
    __syncthreads();
    if (tid == 0) {
        double t[10];
        for (int v = 0; v < 10; ++v) {
            double x = 0.0;
            for (int k = 0; k < 32; ++k) x += s_sum[k][v];
            t[v] = x;
        }
        out->n_static = (int)t[9];
        out->v[0] = out->v[1] = out->v[2] = 0.0;
        const double M[9] = {t[0], t[1], t[2], t[1], t[3], t[4], t[2], t[4], t[5]};
        const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
        const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
        if (t[9] >= 3.0 && fabs(det) > 0.0) {
            const double Mi[9] = {c00 / det, (M[2] * M[7] - M[1] * M[8]) / det, (M[1] * M[5] - M[2] * M[4]) / det,
                                  c01 / det, (M[0] * M[8] - M[2] * M[6]) / det, (M[2] * M[3] - M[0] * M[5]) / det,
                                  c02 / det, (M[1] * M[6] - M[0] * M[7]) / det, (M[0] * M[4] - M[1] * M[3]) / det};
            for (int a = 0; a < 3; ++a) out->v[a] = Mi[3 * a] * t[6] + Mi[3 * a + 1] * t[7] + Mi[3 * a + 2] * t[8];
        }
    }
}

int doppler_filter(Ctx* c, const float* d_rec, int n, int iterations, uint64_t seed, double sigma, double split, uint8_t* d_mask,
                   void* out_host /* DopplerOut, pinned */, bool sync_now) {
    if (iterations <= 0) iterations = (int)(n * 0.2);  // fitSineRansac(..., PointsNum * 0.2), :389
    CKS(reserve_grow(c, c->d_q, (size_t)std::max(n, 1) * sizeof(DopplerPoint)));
    CKS(reserve_grow(c, c->d_partials, std::max((size_t)std::max(iterations, 1) * 24 + 256, (size_t)c->sm_count * 4 * ICP4R_ACC_LEN * sizeof(double) + 1024)));
    CKS(reserve(c, c->d_res, 256));
    DopplerPoint* dp = c->d_q.as<DopplerPoint>();
    double* As = c->d_partials.as<double>();
    double* bs = As + std::max(iterations, 1);
    int* scores = reinterpret_cast<int*>(bs + std::max(iterations, 1));
    DopplerOut* d_out = c->d_res.as<DopplerOut>();
    if (n > 0) {
        doppler_prep_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_rec, n, dp, iterations, scores);
        if (iterations > 0) {
            const int hb = (iterations + 7) / 8;
            const int slices = std::max(1, std::min(8, (c->sm_count * 4) / std::max(hb, 1)));  // ~4 blocks per SM
            doppler_score_kernel<<<dim3(hb, slices), 256, 0, c->stream>>>(d_rec, dp, n, iterations, seed, sigma, scores, As, bs);
        }
        c->launches += 2;
    }
    doppler_final_kernel<<<1, 1024, 0, c->stream>>>(d_rec, dp, n, n > 0 ? iterations : 0, scores, As, bs, split, d_mask, d_out);
    c->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_host, d_out, sizeof(DopplerOut), cudaMemcpyDeviceToHost, c->stream));
    if (sync_now) CK(cudaStreamSynchronize(c->stream));  // (a caller with more work to enqueue synchronises once, later)
    return ICP4R_OK;
}

}  // namespace icp4r
