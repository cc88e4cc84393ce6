// device_math.cuh — scalar device helpers shared by every kernel of libicp4r_cuda (sm_100a).
//
// Everything that decides a correspondence is written with explicit round-to-nearest intrinsics so
// nvcc cannot contract it into FMAs: the reference evaluates these expressions with separate roundings
// (it is built with `-g` only, /root/reference/CMakeLists.txt:5-6) and the neighbour indices must be
// bit-exact.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace icp4r {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint64_t KEY_EMPTY = ~0ull;

// float squared distance, (dx*dx + dy*dy) + dz*dz — calc_dist, ikd_Tree.cpp:1427-1431
__device__ __forceinline__ float dist2_exact(float qx, float qy, float qz, float px, float py, float pz) {
    const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// one row of p' = R p + t in double, left to right, unfused — pointAssociateToMap,
// /root/reference/src/radar_odometry.cpp:137-145 (Eigen evaluates the 3x3 * 3x1 product coefficient-wise)
__device__ __forceinline__ double xform_row(double a, double b, double c, double t, double x, double y, double z) {
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a, x), __dmul_rn(b, y)), __dmul_rn(c, z)), t);
}

__device__ __forceinline__ void xform_point(const double* __restrict__ T, float px, float py, float pz, double pw[3]) {
    const double x = (double)px, y = (double)py, z = (double)pz;
    pw[0] = xform_row(T[0], T[1], T[2], T[3], x, y, z);
    pw[1] = xform_row(T[4], T[5], T[6], T[7], x, y, z);
    pw[2] = xform_row(T[8], T[9], T[10], T[11], x, y, z);
}

// (d2, index) packed so that unsigned comparison orders by distance first, lowest index on ties.
// d2 >= 0 always, so its IEEE bit pattern is monotone as an unsigned integer.
__device__ __forceinline__ uint64_t pack_key(float d2, int idx) {
    return ((uint64_t)__float_as_uint(d2) << 32) | (uint32_t)idx;
}
__device__ __forceinline__ float key_d2(uint64_t k) { return __uint_as_float((uint32_t)(k >> 32)); }
__device__ __forceinline__ int key_idx(uint64_t k) { return (int)(uint32_t)k; }

// sorted (ascending) top-K of packed keys held in registers
template <int K>
struct TopK {
    uint64_t key[K];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < K; ++i) key[i] = KEY_EMPTY;
    }
    __device__ __forceinline__ uint64_t worst() const { return key[K - 1]; }
    __device__ __forceinline__ void insert(uint64_t k) {
        if (k < key[K - 1]) {
            key[K - 1] = k;
#pragma unroll
            for (int j = K - 1; j > 0; --j) {
                const uint64_t a = key[j - 1], b = key[j];
                const bool sw = b < a;
                key[j - 1] = sw ? b : a;
                key[j] = sw ? a : b;
            }
        }
    }
    __device__ __forceinline__ void pop_front() {
#pragma unroll
        for (int j = 0; j < K - 1; ++j) key[j] = key[j + 1];
        key[K - 1] = KEY_EMPTY;
    }
};

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
    const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
    const uint32_t mh = __reduce_min_sync(FULL, hi);
    const uint32_t ml = __reduce_min_sync(FULL, hi == mh ? lo : 0xffffffffu);
    return ((uint64_t)mh << 32) | ml;
}

// Merge the 32 per-lane sorted lists: afterwards lane r (< K) returns the r-th smallest key of the union,
// other lanes return KEY_EMPTY. Destroys the lists.
template <int K>
__device__ __forceinline__ uint64_t warp_merge_topk(TopK<K>& t, int lane) {
    uint64_t mine = KEY_EMPTY;
#pragma unroll
    for (int r = 0; r < K; ++r) {
        const uint64_t m = warp_min_u64(t.key[0]);
        if (m != KEY_EMPTY && t.key[0] == m) t.pop_front();  // keys are unique (index in the low word)
        if (lane == r) mine = m;
    }
    return mine;
}

// ------------------------------------------------------------------------------------------------
// small dense algebra (fp64, one thread)

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

__device__ inline void mat4_mul(const double A[16], const double B[16], double C[16]) {
    double R[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
            R[4 * i + j] = s;
        }
#pragma unroll
    for (int i = 0; i < 16; ++i) C[i] = R[i];
}

__device__ __forceinline__ int tri6(int i, int j) { return i * 6 - i * (i - 1) / 2 + (j - i); }

// H x = -g, H symmetric positive definite given as its 21-entry upper triangle; returns 0 on success
static __device__ __noinline__ int chol6_solve(const double* H21, const double* g, double x[6]) {
    double L[36];
    for (int i = 0; i < 36; ++i) L[i] = 0.0;
    for (int j = 0; j < 6; ++j) {
        double s = H21[tri6(j, j)];
        for (int k = 0; k < j; ++k) s -= L[6 * j + k] * L[6 * j + k];
        if (!(s > 0.0)) return 1;
        L[6 * j + j] = sqrt(s);
        for (int i = j + 1; i < 6; ++i) {
            double v = H21[tri6(j, i)];
            for (int k = 0; k < j; ++k) v -= L[6 * i + k] * L[6 * j + k];
            L[6 * i + j] = v / L[6 * j + j];
        }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) {
        double s = -g[i];
        for (int k = 0; k < i; ++k) s -= L[6 * i + k] * y[k];
        y[i] = s / L[6 * i + i];
    }
    for (int i = 5; i >= 0; --i) {
        double s = y[i];
        for (int k = i + 1; k < 6; ++k) s -= L[6 * k + i] * x[k];
        x[i] = s / L[6 * i + i];
    }
    return 0;
}

// exp of xi = (omega, v) in SE(3), row-major 4x4
static __device__ __noinline__ void se3_exp(const double xi[6], double T[16]) {
    const double wx = xi[0], wy = xi[1], wz = xi[2];
    const double th2 = wx * wx + wy * wy + wz * wz, th = sqrt(th2);
    double A, B, C;
    if (th < 1e-5) {
        A = 1.0 - th2 / 6.0;
        B = 0.5 - th2 / 24.0;
        C = 1.0 / 6.0 - th2 / 120.0;
    } else {
        A = sin(th) / th;
        B = (1.0 - cos(th)) / th2;
        C = (th - sin(th)) / (th2 * th);
    }
    const double W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double W2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += W[3 * i + k] * W[3 * k + j];
            W2[3 * i + j] = s;
        }
    for (int i = 0; i < 3; ++i) {
        double Vr[3];
        for (int j = 0; j < 3; ++j) {
            const double I = (i == j) ? 1.0 : 0.0;
            T[4 * i + j] = I + A * W[3 * i + j] + B * W2[3 * i + j];
            Vr[j] = I + B * W[3 * i + j] + C * W2[3 * i + j];
        }
        T[4 * i + 3] = Vr[0] * xi[3] + Vr[1] * xi[4] + Vr[2] * xi[5];
    }
    T[12] = T[13] = T[14] = 0.0;
    T[15] = 1.0;
}

// ------------------------------------------------------------------------------------------------
// per-correspondence contributions. acc layout (ICP4R_ACC_LEN = 32 doubles):
//   P2P_SVD : [0] n, [1..3] sum p', [4..6] sum q, [7..15] sum p' q^T, [16] sum d2
//   GN kinds: [0..20] H upper triangle, [21..26] g = J^T r, [27] cost, [28] n

__device__ __forceinline__ void acc_gn(double* acc, const double J[6], double r) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i; j < 6; ++j) acc[t++] += J[i] * J[j];
#pragma unroll
    for (int i = 0; i < 6; ++i) acc[21 + i] += J[i] * r;
    acc[27] += r * r;
}

__device__ __forceinline__ void contrib_p2p_svd(double* acc, const double pw[3], float qx, float qy, float qz, float d2) {
    const double q[3] = {(double)qx, (double)qy, (double)qz};
    acc[0] += 1.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        acc[1 + i] += pw[i];
        acc[4 + i] += q[i];
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[7 + 3 * i + j] += pw[i] * q[j];
    }
    acc[16] += (double)d2;
}

// LidarDistanceFactor (radarFactor.hpp:156-158): r = p' - c, J = [-[p']x | I]
__device__ __forceinline__ void contrib_p2p_gn(double* acc, const double pw[3], float cx, float cy, float cz) {
    const double r0 = pw[0] - (double)cx, r1 = pw[1] - (double)cy, r2 = pw[2] - (double)cz;
    const double J0[6] = {0, pw[2], -pw[1], 1, 0, 0};
    const double J1[6] = {-pw[2], 0, pw[0], 0, 1, 0};
    const double J2[6] = {pw[1], -pw[0], 0, 0, 0, 1};
    acc_gn(acc, J0, r0);
    acc_gn(acc, J1, r1);
    acc_gn(acc, J2, r2);
    acc[28] += 1.0;
}

// LOAM plane through k points: A n = -1 by normal equations (adjugate), d = 1/|n|, n /= |n|
template <int K, typename F>
__device__ __forceinline__ bool plane_fit(const F (&P)[K][3], int k, double n[3], double& d) {
    double m00 = 0, m01 = 0, m02 = 0, m11 = 0, m12 = 0, m22 = 0, v0 = 0, v1 = 0, v2 = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (j < k) {
            const double x = (double)P[j][0], y = (double)P[j][1], z = (double)P[j][2];
            m00 += x * x;
            m01 += x * y;
            m02 += x * z;
            m11 += y * y;
            m12 += y * z;
            m22 += z * z;
            v0 -= x;
            v1 -= y;
            v2 -= z;
        }
    }
    const double c00 = m11 * m22 - m12 * m12, c01 = m02 * m12 - m01 * m22, c02 = m01 * m12 - m02 * m11;
    const double c11 = m00 * m22 - m02 * m02, c12 = m01 * m02 - m00 * m12, c22 = m00 * m11 - m01 * m01;
    const double det = (m00 * c00 + m01 * c01) + m02 * c02;
    if (!(fabs(det) > 0.0) || !isfinite(det)) return false;
    // n = adj(M) v / det, then n /= |n|, d = 1 / |n|. The common factor 1/det cancels in the direction, so the
    // three divisions by det and the three by |n| collapse into one reciprocal square root and one reciprocal
    // (this runs once per source point per iteration, redundantly on every lane of the warp).
    const double ux = (c00 * v0 + c01 * v1) + c02 * v2;
    const double uy = (c01 * v0 + c11 * v1) + c12 * v2;
    const double uz = (c02 * v0 + c12 * v1) + c22 * v2;
    const double uu = (ux * ux + uy * uy) + uz * uz;
    if (!(uu > 0.0) || !isfinite(uu)) return false;
    const double iu = rsqrt(uu);                 // 1 / |u|
    const double sgn = det > 0.0 ? 1.0 : -1.0;   // n = u / det / |u / det| = sign(det) u / |u|
    n[0] = sgn * ux * iu;
    n[1] = sgn * uy * iu;
    n[2] = sgn * uz * iu;
    d = fabs(det) * iu;                          // 1 / |u / det|
    return true;
}

}  // namespace icp4r
