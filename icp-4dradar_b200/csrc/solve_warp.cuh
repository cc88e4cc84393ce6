// solve_warp.cuh — the per-iteration pose update done by ONE WARP (north_star item 4): 6x6 Cholesky solve with
// one matrix row per lane, SE(3) exponential and pose composition with one matrix entry per lane, and a one-sided
// Jacobi SVD of the 3x3 cross-covariance with one matrix row per lane.  This code sits on the serial tail of every
// iteration (it runs after the last block has reduced the partial sums), so it is written for latency: registers
// only, constant indices, shuffles instead of memory, rsqrt instead of sqrt + divide.
//
// All functions must be called by a full, converged warp.
#pragma once
#include "device_math.cuh"

namespace icp4r {

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL, v, src); }

// H x = -g for the symmetric positive definite 6x6 H given as its 21-entry upper triangle (row-major).
// Lane i < 6 works on row i. Returns false (warp-uniform) when a pivot is not positive. x_i is returned in lane i.
__device__ __forceinline__ bool warp_chol6_solve(const double* __restrict__ H21, const double* __restrict__ g, int lane, double& x_out) {
    const int i = lane < 6 ? lane : 5;
    double a[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const int lo = i < j ? i : j, hi = i < j ? j : i;
        a[j] = H21[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)];
    }
    double b = -g[i];
    double col[6];  // col[k] = L[k][i] for k > i (filled at step j == i)
    double myinv = 0.0;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const double d = shfl_d(a[j], j);  // current H[j][j]
        if (!(d > 0.0)) ok = false;
        const double inv = rsqrt(d);       // 1 / L[j][j]
        const double lij = a[j] * inv;     // L[i][j] for i >= j
        if (i == j) myinv = inv;
        const double yj = shfl_d(b, j) * inv;  // forward substitution fused in: y_j
        if (i == j) b = yj;
        else if (i > j) b -= lij * yj;
#pragma unroll
        for (int k = j + 1; k < 6; ++k) {
            const double lkj = shfl_d(lij, k);  // L[k][j]
            if (i == j) col[k] = lkj;
            if (i > j) a[k] -= lij * lkj;
        }
    }
    // back substitution: x_j = (y_j - sum_{k>j} L[k][j] x_k) / L[j][j]
#pragma unroll
    for (int j = 5; j >= 0; --j) {
        const double xj = shfl_d(b * myinv, j);
        if (i == j) b = xj;
        else if (i < j) b -= col[j] * xj;
    }
    x_out = b;
    return ok;
}

// D = exp(xi^) (3x4 part) with lane e < 12 computing entry (e / 4, e % 4); xi in shared memory.
__device__ __forceinline__ double warp_se3_exp_entry(const double* __restrict__ xi, int lane) {
    const double wx = xi[0], wy = xi[1], wz = xi[2];
    const double th2 = wx * wx + wy * wy + wz * wz, th = sqrt(th2);
    double A, B, C;  // sin th / th, (1 - cos th) / th^2, (th - sin th) / th^3
    if (th < 1e-5) {
        A = 1.0 - th2 / 6.0;
        B = 0.5 - th2 / 24.0;
        C = 1.0 / 6.0 - th2 / 120.0;
    } else {
        double sn, cs;
        sincos(th, &sn, &cs);
        A = sn / th;
        B = (1.0 - cs) / th2;
        C = (th - sn) / (th2 * th);
    }
    const int e = lane < 12 ? lane : 0;
    const int r = e >> 2, c = e & 3;
    // W = [w]x : W[r][c]; W2 = w w^T - th2 I
    auto Wrc = [&](int rr, int cc) -> double {
        if (rr == cc) return 0.0;
        const int k = 3 - rr - cc;  // the remaining axis
        const double wk = xi[k];
        // sign: W[0][1] = -wz, W[0][2] = +wy, W[1][0] = +wz, W[1][2] = -wx, W[2][0] = -wy, W[2][1] = +wx
        const bool pos = ((cc - rr + 3) % 3) == 2;
        return pos ? wk : -wk;
    };
    if (c < 3) {
        const double I = r == c ? 1.0 : 0.0;
        return I + A * Wrc(r, c) + B * (xi[r] * xi[c] - I * th2);
    }
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double I = r == k ? 1.0 : 0.0;
        const double V = I + B * Wrc(r, k) + C * (xi[r] * xi[k] - I * th2);
        t += V * xi[3 + k];
    }
    return t;
}

// T <- D * T with D (3x4, row-major in shared memory Ds[12]) and T (4x4 in shared memory Ts[16]).
// Lane e < 12 returns the new T entry (e / 4, e % 4).
__device__ __forceinline__ double warp_compose_entry(const double* __restrict__ Ds, const double* __restrict__ Ts, int lane) {
    const int e = lane < 12 ? lane : 0;
    const int r = e >> 2, c = e & 3;
    double s = 0.0;
    s += Ds[4 * r + 0] * Ts[0 + c];
    s += Ds[4 * r + 1] * Ts[4 + c];
    s += Ds[4 * r + 2] * Ts[8 + c];
    s += Ds[4 * r + 3] * Ts[12 + c];
    return s;
}

__device__ __forceinline__ double sum3(double v) {  // sum over lanes 0..3 (lane 3 must hold 0), result in all four
    v += __shfl_xor_sync(FULL, v, 1);
    v += __shfl_xor_sync(FULL, v, 2);
    return v;
}

// Kabsch rotation from H = sum (p - pm)(q - qm)^T (row-major in shared memory Hs[9]) by one-sided Jacobi SVD,
// lane i < 3 holding row i of the working matrix A and of V (H V = U S). R = V U^T with the column of the smallest
// singular value rebuilt by cross products so det R = +1 (Umeyama's reflection fix). Row i of R is returned in
// lane i (R0, R1, R2).
//
// Vw (optional, 9 doubles in shared memory, row-major): warm start. The batched-pairs loop solves 30 nearly identical
// problems in a row; starting from the V of the previous one (A = H V, already almost column-orthogonal) the sweeps end
// after one or two instead of six to eight. V is written back. The rotation returned is the same (V U^T does not
// depend on the order or sign of the singular vectors).
__device__ __forceinline__ void warp_kabsch(const double* __restrict__ Hs, int lane, double& R0, double& R1, double& R2,
                                            double* __restrict__ Vw = nullptr, int* sweeps_out = nullptr) {
    // every group of 4 lanes holds the same 3 rows (+ a zero lane), so the whole warp stays converged
    const int i = lane & 3;
    const bool rowlane = i < 3;
    double a[3], v[3];
    if (Vw == nullptr) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            a[j] = rowlane ? Hs[3 * (rowlane ? i : 0) + j] : 0.0;
            v[j] = (rowlane && i == j) ? 1.0 : 0.0;
        }
    } else {
        const int r = rowlane ? i : 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            a[j] = rowlane ? (Hs[3 * r + 0] * Vw[j] + Hs[3 * r + 1] * Vw[3 + j]) + Hs[3 * r + 2] * Vw[6 + j] : 0.0;
            v[j] = rowlane ? Vw[3 * r + j] : 0.0;
        }
        __syncwarp();
    }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;  // (0,1), (0,2), (1,2)
            const double al = sum3(a[p] * a[p]), be = sum3(a[q] * a[q]), ga = sum3(a[p] * a[q]);
            if (fabs(ga) <= 1e-300 || fabs(ga) <= 1e-16 * sqrt(al * be)) continue;  // warp-uniform
            off += fabs(ga);
            // rotation that zeroes a_p . a_q: tan 2t = 2 ga / (be - al), |t| <= pi/4, written with two rsqrt
            // instead of a divide and two square roots (this loop is on the serial tail of every iteration)
            const double dd = be - al, g2 = 2.0 * ga;
            const double invh = rsqrt(dd * dd + g2 * g2);      // 1 / hypot(dd, g2)
            const double c2 = 0.5 + 0.5 * fabs(dd) * invh;     // cos^2 t
            const double invc = rsqrt(c2);
            const double c = c2 * invc;
            const double s = (dd >= 0 ? 0.5 : -0.5) * g2 * invh * invc;
            const double ap = a[p], aq = a[q], vp = v[p], vq = v[q];
            a[p] = c * ap - s * aq;
            a[q] = s * ap + c * aq;
            v[p] = c * vp - s * vq;
            v[q] = s * vp + c * vq;
        }
        if (sweeps_out) *sweeps_out = sweep + 1;
        if (off == 0.0) break;
    }
    if (Vw != nullptr && lane < 3) {
#pragma unroll
        for (int j = 0; j < 3; ++j) Vw[3 * lane + j] = v[j];
    }
    double sg[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) sg[j] = sqrt(sum3(a[j] * a[j]));
    int cc = 0;
    if (sg[1] < sg[cc]) cc = 1;
    if (sg[2] < sg[cc]) cc = 2;
    const int ia = (cc + 1) % 3, ib = (cc + 2) % 3;
    const double sga = ia == 0 ? sg[0] : (ia == 1 ? sg[1] : sg[2]);
    const double sgb = ib == 0 ? sg[0] : (ib == 1 ? sg[1] : sg[2]);
    if (!(sga > 0.0) || !(sgb > 0.0)) {  // rank < 2: rotation undefined -> identity
        R0 = i == 0 ? 1.0 : 0.0;
        R1 = i == 1 ? 1.0 : 0.0;
        R2 = i == 2 ? 1.0 : 0.0;
        return;
    }
    // this lane's components (index i) of ua, ub, va, vb
    const double aa = ia == 0 ? a[0] : (ia == 1 ? a[1] : a[2]);
    const double ab = ib == 0 ? a[0] : (ib == 1 ? a[1] : a[2]);
    const double ua_i = aa / sga;
    double ub_i = ab / sgb;
    const double va_i = ia == 0 ? v[0] : (ia == 1 ? v[1] : v[2]);
    const double vb_i = ib == 0 ? v[0] : (ib == 1 ? v[1] : v[2]);
    const double dab = sum3(rowlane ? ua_i * ub_i : 0.0);
    ub_i -= dab * ua_i;
    const double nb = sqrt(sum3(rowlane ? ub_i * ub_i : 0.0));
    ub_i /= nb;
    // full vectors in every lane
    double ua[3], ub[3], va[3], vb[3], uc[3], vc[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        ua[j] = shfl_d(ua_i, j);
        ub[j] = shfl_d(ub_i, j);
        va[j] = shfl_d(va_i, j);
        vb[j] = shfl_d(vb_i, j);
    }
    cross3(ua, ub, uc);
    cross3(va, vb, vc);
    const double vai = i == 0 ? va[0] : (i == 1 ? va[1] : va[2]);
    const double vbi = i == 0 ? vb[0] : (i == 1 ? vb[1] : vb[2]);
    const double vci = i == 0 ? vc[0] : (i == 1 ? vc[1] : vc[2]);
    R0 = vai * ua[0] + vbi * ub[0] + vci * uc[0];
    R1 = vai * ua[1] + vbi * ub[1] + vci * uc[1];
    R2 = vai * ua[2] + vbi * ub[2] + vci * uc[2];
}

// The same Kabsch rotation computed by ONE thread with the whole 3x3 problem in registers: no shuffles on the critical
// path (each of the ~30 lane-to-lane exchanges of the warp version costs more than the arithmetic between them), no
// divisions or square roots (reciprocal square roots only), the convergence test without a square root.
// H row-major; R row-major. Every thread of a warp may call it redundantly (warp-uniform control flow).
__device__ __forceinline__ void thread_kabsch(const double* __restrict__ H, double R[9]) {
    double a[3][3], v[3][3];  // columns: a[j] = j-th column of A = H V, v[j] = j-th column of V
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            a[j][i] = H[3 * i + j];
            v[j][i] = i == j ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;  // (0,1), (0,2), (1,2)
            const double al = (a[p][0] * a[p][0] + a[p][1] * a[p][1]) + a[p][2] * a[p][2];
            const double be = (a[q][0] * a[q][0] + a[q][1] * a[q][1]) + a[q][2] * a[q][2];
            const double ga = (a[p][0] * a[q][0] + a[p][1] * a[q][1]) + a[p][2] * a[q][2];
            if (fabs(ga) <= 1e-300 || ga * ga <= 1e-32 * (al * be)) continue;
            rotated = true;
            // rotation that zeroes a_p . a_q: tan 2t = 2 ga / (be - al), |t| <= pi/4
            const double dd = be - al, g2 = 2.0 * ga;
            const double invh = rsqrt(dd * dd + g2 * g2);      // 1 / hypot(dd, g2)
            const double c2 = 0.5 + 0.5 * fabs(dd) * invh;     // cos^2 t
            const double invc = rsqrt(c2);
            const double c = c2 * invc;
            const double s = (dd >= 0 ? 0.5 : -0.5) * g2 * invh * invc;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double ap = a[p][i], aq = a[q][i], vp = v[p][i], vq = v[q][i];
                a[p][i] = c * ap - s * aq;
                a[q][i] = s * ap + c * aq;
                v[p][i] = c * vp - s * vq;
                v[q][i] = s * vp + c * vq;
            }
        }
        if (!rotated) break;
    }
    double n2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) n2[j] = (a[j][0] * a[j][0] + a[j][1] * a[j][1]) + a[j][2] * a[j][2];
    int cc = 0;
    if (n2[1] < n2[cc]) cc = 1;
    if (n2[2] < n2[cc]) cc = 2;
    // the two dominant columns, selected without dynamic register indexing
    double ua[3], ub[3], va[3], vb[3], na, nb;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ua[i] = cc == 0 ? a[1][i] : (cc == 1 ? a[2][i] : a[0][i]);
        ub[i] = cc == 0 ? a[2][i] : (cc == 1 ? a[0][i] : a[1][i]);
        va[i] = cc == 0 ? v[1][i] : (cc == 1 ? v[2][i] : v[0][i]);
        vb[i] = cc == 0 ? v[2][i] : (cc == 1 ? v[0][i] : v[1][i]);
    }
    na = cc == 0 ? n2[1] : (cc == 1 ? n2[2] : n2[0]);
    nb = cc == 0 ? n2[2] : (cc == 1 ? n2[0] : n2[1]);
    if (!(na > 0.0) || !(nb > 0.0)) {  // rank < 2: rotation undefined -> identity
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    const double ia = rsqrt(na), ib = rsqrt(nb);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ua[i] *= ia;
        ub[i] *= ib;
    }
    const double dab = (ua[0] * ub[0] + ua[1] * ub[1]) + ua[2] * ub[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) ub[i] -= dab * ua[i];
    const double inb = rsqrt((ub[0] * ub[0] + ub[1] * ub[1]) + ub[2] * ub[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) ub[i] *= inb;
    double uc[3], vc[3];
    cross3(ua, ub, uc);  // the third pair rebuilt by cross products: det R = +1 (Umeyama's reflection fix)
    cross3(va, vb, vc);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R[3 * i + j] = (va[i] * ua[j] + vb[i] * ub[j]) + vc[i] * uc[j];
}

}  // namespace icp4r
