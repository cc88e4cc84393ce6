/* oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE). See oracle.h for scope and pin status.
 *
 * Build: gcc -O3 -fopenmp -ffp-contract=off  (contraction MUST stay off: the reference is built with
 * `-g` only, /root/reference/CMakeLists.txt:5-6, so its float/double expressions are evaluated with
 * separate roundings; the CUDA library uses __fmul_rn/__fadd_rn/__dmul_rn/__dadd_rn to match).
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_default_opts(orc_opts* o) {
    memset(o, 0, sizeof(*o));
    o->residual = ORC_P2P_SVD;
    o->k = 1;
    o->max_iterations = 10; /* PCL default; iterative_closest_point.cpp:513 leaves it */
    o->early_exit = 0;
    o->max_corr_dist = 0.0;
    o->rot_eps = 2e-3;   /* fast_gicp defaults quoted in SURVEY.md §8 a11 */
    o->trans_eps = 5e-4;
    o->mse_abs_eps = 1e-12;
    o->plane_thresh = 0.2;
    for (int i = 0; i < 16; ++i) o->T0[i] = (i % 5 == 0) ? 1.0 : 0.0;
    o->interp_s = 1.0;
}

/* ------------------------------------------------------------------ kNN ---------------------------- */

/* calc_dist, ikd_Tree.cpp:1427-1431: left-to-right float sum of squares */
static inline float dist2f(const float* a, const float* b) {
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return (dx * dx + dy * dy) + dz * dz;
}

static inline double gate2(double max_dist) {
    if (!(max_dist > 0.0) || isinf(max_dist)) return INFINITY;
    return max_dist * max_dist; /* ikd_Tree.cpp:880 */
}

void orc_knn(const float* tgt, const uint8_t* valid, int m, const float* q, int nq, int k, double max_dist,
             int32_t* idx, float* d2, int32_t* found) {
    const double g2 = gate2(max_dist);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nq; ++i) {
        int32_t* bi = idx + (size_t)i * k;
        float* bd = d2 + (size_t)i * k;
        int cnt = 0;
        for (int s = 0; s < k; ++s) {
            bi[s] = -1;
            bd[s] = INFINITY;
        }
        const float* qi = q + 4 * (size_t)i;
        for (int j = 0; j < m; ++j) {
            if (valid && !valid[j]) continue;
            float d = dist2f(qi, tgt + 4 * (size_t)j);
            if (!((double)d <= g2)) continue; /* ikd_Tree.cpp:895 gate (inclusive) */
            /* ascending j => an equal distance never displaces an earlier (lower) index */
            if (cnt == k && !(d < bd[k - 1])) continue;
            int pos = (cnt < k) ? cnt : k - 1;
            while (pos > 0 && d < bd[pos - 1]) {
                bd[pos] = bd[pos - 1];
                bi[pos] = bi[pos - 1];
                --pos;
            }
            bd[pos] = d;
            bi[pos] = j;
            if (cnt < k) ++cnt;
        }
        if (found) found[i] = cnt;
    }
}

void orc_knn_brute_cb(void* cloud, const float* q, int nq, int k, double max_dist, int32_t* idx, float* d2,
                      int32_t* found) {
    const orc_cloud* c = (const orc_cloud*)cloud;
    orc_knn(c->xyzw, c->valid, c->m, q, nq, k, max_dist, idx, d2, found);
}

/* ------------------------------------------------------------------ transforms --------------------- */

static inline void xform1(const double T[16], const float* p, double o[3]) {
    const double x = p[0], y = p[1], z = p[2];
    o[0] = ((T[0] * x + T[1] * y) + T[2] * z) + T[3];
    o[1] = ((T[4] * x + T[5] * y) + T[6] * z) + T[7];
    o[2] = ((T[8] * x + T[9] * y) + T[10] * z) + T[11];
}

void orc_transform(const double T[16], const float* xyzw, int n, float* out_f32, double* out_f64) {
    for (int i = 0; i < n; ++i) {
        double o[3];
        xform1(T, xyzw + 4 * (size_t)i, o);
        if (out_f32) {
            out_f32[4 * (size_t)i + 0] = (float)o[0];
            out_f32[4 * (size_t)i + 1] = (float)o[1];
            out_f32[4 * (size_t)i + 2] = (float)o[2];
            out_f32[4 * (size_t)i + 3] = xyzw[4 * (size_t)i + 3];
        }
        if (out_f64) {
            out_f64[3 * (size_t)i + 0] = o[0];
            out_f64[3 * (size_t)i + 1] = o[1];
            out_f64[3 * (size_t)i + 2] = o[2];
        }
    }
}

void orc_mat4_mul(const double A[16], const double B[16], double C[16]) {
    double R[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
            R[4 * i + j] = s;
        }
    memcpy(C, R, sizeof(R));
}

/* ------------------------------------------------------------------ residual functors -------------- */

static inline void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

/* v' = q * v for a unit quaternion (x,y,z,w): v + w*(2 u x v) + u x (2 u x v) */
static void quat_rotate(const double q[4], const double v[3], double o[3]) {
    double uv[3], uuv[3];
    cross3(q, v, uv);
    uv[0] += uv[0];
    uv[1] += uv[1];
    uv[2] += uv[2];
    cross3(q, uv, uuv);
    o[0] = v[0] + q[3] * uv[0] + uuv[0];
    o[1] = v[1] + q[3] * uv[1] + uuv[1];
    o[2] = v[2] + q[3] * uv[2] + uuv[2];
}

/* identity.slerp(s, q) as the functors at radarFactor.hpp:26-28,78-80 use it */
static void slerp_from_identity(double s, const double q[4], double o[4]) {
    const double d = q[3]; /* dot(identity, q) */
    const double ad = fabs(d);
    double s0, s1;
    if (ad >= 1.0 - DBL_EPSILON) {
        s0 = 1.0 - s;
        s1 = s;
    } else {
        const double th = acos(ad), sn = sin(th);
        s0 = sin((1.0 - s) * th) / sn;
        s1 = sin(s * th) / sn;
    }
    if (d < 0.0) s1 = -s1;
    o[0] = s1 * q[0];
    o[1] = s1 * q[1];
    o[2] = s1 * q[2];
    o[3] = s0 + s1 * q[3];
}

/* LidarDistanceFactor, radarFactor.hpp:147-160 */
void orc_res_distance(const double q[4], const double t[3], const double p[3], const double c[3], double r[3]) {
    double w[3];
    quat_rotate(q, p, w);
    r[0] = (w[0] + t[0]) - c[0];
    r[1] = (w[1] + t[1]) - c[1];
    r[2] = (w[2] + t[2]) - c[2];
}

/* LidarPlaneNormFactor, radarFactor.hpp:113-124 */
void orc_res_plane_norm(const double q[4], const double t[3], const double p[3], const double n[3], double d,
                        double r[1]) {
    double w[3];
    quat_rotate(q, p, w);
    w[0] += t[0];
    w[1] += t[1];
    w[2] += t[2];
    r[0] = (n[0] * w[0] + n[1] * w[1] + n[2] * w[2]) + d;
}

/* LidarPlaneFactor, radarFactor.hpp:63-64 (ctor normal) and :68-89 */
void orc_res_plane(const double q[4], const double t[3], const double p[3], const double j[3],
                   const double l[3], const double m[3], double s, double r[1]) {
    double jl[3] = {j[0] - l[0], j[1] - l[1], j[2] - l[2]};
    double jm[3] = {j[0] - m[0], j[1] - m[1], j[2] - m[2]};
    double nrm[3];
    cross3(jl, jm, nrm);
    double len = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    if (len > 0.0) {
        nrm[0] /= len;
        nrm[1] /= len;
        nrm[2] /= len;
    }
    double qs[4], lp[3];
    slerp_from_identity(s, q, qs);
    quat_rotate(qs, p, lp);
    lp[0] += s * t[0];
    lp[1] += s * t[1];
    lp[2] += s * t[2];
    r[0] = (lp[0] - j[0]) * nrm[0] + (lp[1] - j[1]) * nrm[1] + (lp[2] - j[2]) * nrm[2];
}

/* RadarEdgeFactor, radarFactor.hpp:18-42 */
void orc_res_edge(const double q[4], const double t[3], const double p[3], const double a[3],
                  const double b[3], double s, double r[3]) {
    double qs[4], lp[3];
    slerp_from_identity(s, q, qs);
    quat_rotate(qs, p, lp);
    lp[0] += s * t[0];
    lp[1] += s * t[1];
    lp[2] += s * t[2];
    double u[3] = {lp[0] - a[0], lp[1] - a[1], lp[2] - a[2]};
    double v[3] = {lp[0] - b[0], lp[1] - b[1], lp[2] - b[2]};
    double nu[3];
    cross3(u, v, nu);
    double de[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    double dn = sqrt(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]);
    r[0] = nu[0] / dn;
    r[1] = nu[1] / dn;
    r[2] = nu[2] / dn;
}

/* ------------------------------------------------------------------ small dense algebra ------------ */

/* One-sided (Hestenes) Jacobi SVD of a 3x3: A V = U S. Kabsch rotation R = V' U'^T with the column of
 * the smallest singular value rebuilt by cross products so det(R) = +1 (Umeyama's reflection fix;
 * PCL TransformationEstimationSVD -> pcl::umeyama). H = sum (p-pm)(q-qm)^T maps
 * source to target: R = argmin sum |R p - q|^2. */
void orc_svd3_rotation(const double H[9], double R[9]) {
    double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(A, H, sizeof(A));
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double al = 0, be = 0, ga = 0;
                for (int i = 0; i < 3; ++i) {
                    al += A[3 * i + p] * A[3 * i + p];
                    be += A[3 * i + q] * A[3 * i + q];
                    ga += A[3 * i + p] * A[3 * i + q];
                }
                if (ga == 0.0 || fabs(ga) <= 1e-300) continue;
                double lim = 1e-16 * sqrt(al * be);
                if (fabs(ga) <= lim) continue;
                off += fabs(ga);
                double zeta = (be - al) / (2.0 * ga);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < 3; ++i) {
                    double ap = A[3 * i + p], aq = A[3 * i + q];
                    A[3 * i + p] = c * ap - s * aq;
                    A[3 * i + q] = s * ap + c * aq;
                    double vp = V[3 * i + p], vq = V[3 * i + q];
                    V[3 * i + p] = c * vp - s * vq;
                    V[3 * i + q] = s * vp + c * vq;
                }
            }
        if (off == 0.0) break;
    }
    double sg[3];
    for (int j = 0; j < 3; ++j)
        sg[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
    /* indices of the two largest singular values (a,b) and the smallest (c), cyclic order a->b->c */
    int c = 0;
    if (sg[1] < sg[c]) c = 1;
    if (sg[2] < sg[c]) c = 2;
    int a = (c + 1) % 3, b = (c + 2) % 3;
    double ua[3], ub[3], uc[3], va[3], vb[3], vc[3];
    if (!(sg[a] > 0.0) || !(sg[b] > 0.0)) { /* rank < 2: rotation undefined, return identity */
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    for (int i = 0; i < 3; ++i) {
        ua[i] = A[3 * i + a] / sg[a];
        ub[i] = A[3 * i + b] / sg[b];
        va[i] = V[3 * i + a];
        vb[i] = V[3 * i + b];
    }
    /* re-orthogonalise ub against ua (guards the nearly rank-1 case) */
    double dab = ua[0] * ub[0] + ua[1] * ub[1] + ua[2] * ub[2];
    double nb = 0;
    for (int i = 0; i < 3; ++i) {
        ub[i] -= dab * ua[i];
        nb += ub[i] * ub[i];
    }
    nb = sqrt(nb);
    for (int i = 0; i < 3; ++i) ub[i] /= nb;
    cross3(ua, ub, uc);
    cross3(va, vb, vc);
    /* H = U S V^T (H V = U S); minimiser of sum |R p - q|^2 with H = sum p q^T is R = V U^T */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = va[i] * ua[j] + vb[i] * ub[j] + vc[i] * uc[j];
}

static inline int tri(int i, int j) { /* upper-triangular row-major index, i<=j, 6x6 */
    return i * 6 - i * (i - 1) / 2 + (j - i);
}

int orc_chol6_solve(const double H21[21], const double g[6], double x[6]) {
    double L[36];
    memset(L, 0, sizeof(L));
    for (int j = 0; j < 6; ++j) {
        double s = H21[tri(j, j)];
        for (int k = 0; k < j; ++k) s -= L[6 * j + k] * L[6 * j + k];
        if (!(s > 0.0)) return 1;
        L[6 * j + j] = sqrt(s);
        for (int i = j + 1; i < 6; ++i) {
            double v = H21[tri(j, i)];
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

void orc_se3_exp(const double xi[6], double T[16]) {
    const double wx = xi[0], wy = xi[1], wz = xi[2];
    const double th2 = wx * wx + wy * wy + wz * wz, th = sqrt(th2);
    double A, B, C; /* sin th/th, (1-cos th)/th^2, (th - sin th)/th^3 */
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
    double R[9], Vm[9];
    for (int i = 0; i < 9; ++i) {
        const double I = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = I + A * W[i] + B * W2[i];
        Vm[i] = I + B * W[i] + C * W2[i];
    }
    for (int i = 0; i < 3; ++i) {
        T[4 * i + 0] = R[3 * i + 0];
        T[4 * i + 1] = R[3 * i + 1];
        T[4 * i + 2] = R[3 * i + 2];
        T[4 * i + 3] = Vm[3 * i + 0] * xi[3] + Vm[3 * i + 1] * xi[4] + Vm[3 * i + 2] * xi[5];
    }
    T[12] = T[13] = T[14] = 0.0;
    T[15] = 1.0;
}

int orc_plane_fit(const double* P, int k, double n[3], double* d) {
    double m00 = 0, m01 = 0, m02 = 0, m11 = 0, m12 = 0, m22 = 0, v0 = 0, v1 = 0, v2 = 0;
    for (int j = 0; j < k; ++j) {
        const double x = P[3 * j], y = P[3 * j + 1], z = P[3 * j + 2];
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
    const double c00 = m11 * m22 - m12 * m12, c01 = m02 * m12 - m01 * m22, c02 = m01 * m12 - m02 * m11;
    const double c11 = m00 * m22 - m02 * m02, c12 = m01 * m02 - m00 * m12, c22 = m00 * m11 - m01 * m01;
    const double det = (m00 * c00 + m01 * c01) + m02 * c02;
    if (!(fabs(det) > 0.0) || !isfinite(det)) return 0;
    double nx = ((c00 * v0 + c01 * v1) + c02 * v2) / det;
    double ny = ((c01 * v0 + c11 * v1) + c12 * v2) / det;
    double nz = ((c02 * v0 + c12 * v1) + c22 * v2) / det;
    const double nn = sqrt((nx * nx + ny * ny) + nz * nz);
    if (!(nn > 0.0) || !isfinite(nn)) return 0;
    *d = 1.0 / nn;
    n[0] = nx / nn;
    n[1] = ny / nn;
    n[2] = nz / nn;
    return 1;
}

/* ------------------------------------------------------------------ accumulation ------------------- */

static inline void acc_gn(double* acc, const double J[6], double r) {
    int t = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) acc[t++] += J[i] * J[j];
    for (int i = 0; i < 6; ++i) acc[21 + i] += J[i] * r;
    acc[27] += r * r;
}

/* Accumulate one source point's contribution. pw = transformed point (double). Returns 1 if used. */
static int contribute(int residual, int k, double plane_thresh, const double pw[3], const float* tgt,
                      const int32_t* nn, const float* nd2, int found, double* acc) {
    if (residual == ORC_P2P_SVD) {
        if (found < 1) return 0;
        const float* c = tgt + 4 * (size_t)nn[0];
        const double q[3] = {c[0], c[1], c[2]};
        acc[0] += 1.0;
        for (int i = 0; i < 3; ++i) {
            acc[1 + i] += pw[i];
            acc[4 + i] += q[i];
            for (int j = 0; j < 3; ++j) acc[7 + 3 * i + j] += pw[i] * q[j];
        }
        acc[16] += (double)nd2[0];
        return 1;
    }
    if (residual == ORC_P2P_GN) {
        if (found < 1) return 0;
        const float* c = tgt + 4 * (size_t)nn[0];
        /* r = p' - c, J = [-[p']x | I]  (LidarDistanceFactor, radarFactor.hpp:156-158) */
        const double r[3] = {pw[0] - c[0], pw[1] - c[1], pw[2] - c[2]};
        const double J0[6] = {0, pw[2], -pw[1], 1, 0, 0};
        const double J1[6] = {-pw[2], 0, pw[0], 0, 1, 0};
        const double J2[6] = {pw[1], -pw[0], 0, 0, 0, 1};
        acc_gn(acc, J0, r[0]);
        acc_gn(acc, J1, r[1]);
        acc_gn(acc, J2, r[2]);
        acc[28] += 1.0;
        return 1;
    }
    if (residual == ORC_P2PLANE_KNN) {
        if (found < k || k < 3) return 0;
        double P[3 * 16];
        for (int j = 0; j < k; ++j) {
            const float* c = tgt + 4 * (size_t)nn[j];
            P[3 * j] = c[0];
            P[3 * j + 1] = c[1];
            P[3 * j + 2] = c[2];
        }
        double n[3], d;
        if (!orc_plane_fit(P, k, n, &d)) return 0;
        for (int j = 0; j < k; ++j) {
            double e = ((n[0] * P[3 * j] + n[1] * P[3 * j + 1]) + n[2] * P[3 * j + 2]) + d;
            if (!(fabs(e) <= plane_thresh)) return 0;
        }
        /* r = n.p' + d (LidarPlaneNormFactor, radarFactor.hpp:122), J = [(p' x n)^T | n^T] */
        const double r = ((n[0] * pw[0] + n[1] * pw[1]) + n[2] * pw[2]) + d;
        double pxn[3];
        cross3(pw, n, pxn);
        const double J[6] = {pxn[0], pxn[1], pxn[2], n[0], n[1], n[2]};
        acc_gn(acc, J, r);
        acc[28] += 1.0;
        return 1;
    }
    if (residual == ORC_P2PLANE_3PT) {
        if (found < 3) return 0;
        /* LidarPlaneFactor (radarFactor.hpp:63-64,86), s = 1: plane through the three nearest points j, l, m with
         * unit normal (j-l) x (j-m); r = (p' - j) . n ; J = [(p' x n)^T | n^T] */
        const float* fj = tgt + 4 * (size_t)nn[0];
        const float* fl = tgt + 4 * (size_t)nn[1];
        const float* fm = tgt + 4 * (size_t)nn[2];
        const double jl[3] = {(double)fj[0] - fl[0], (double)fj[1] - fl[1], (double)fj[2] - fl[2]};
        const double jm[3] = {(double)fj[0] - fm[0], (double)fj[1] - fm[1], (double)fj[2] - fm[2]};
        double n[3];
        cross3(jl, jm, n);
        const double len = sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
        if (!(len > 0.0)) return 0;
        n[0] /= len;
        n[1] /= len;
        n[2] /= len;
        const double r = ((pw[0] - fj[0]) * n[0] + (pw[1] - fj[1]) * n[1]) + (pw[2] - fj[2]) * n[2];
        double pxn[3];
        cross3(pw, n, pxn);
        const double J[6] = {pxn[0], pxn[1], pxn[2], n[0], n[1], n[2]};
        acc_gn(acc, J, r);
        acc[28] += 1.0;
        return 1;
    }
    if (residual == ORC_P2LINE) {
        if (found < 2) return 0;
        const float* fa = tgt + 4 * (size_t)nn[0];
        const float* fb = tgt + 4 * (size_t)nn[1];
        const double a[3] = {fa[0], fa[1], fa[2]}, b[3] = {fb[0], fb[1], fb[2]};
        const double ba[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
        const double L = sqrt((ba[0] * ba[0] + ba[1] * ba[1]) + ba[2] * ba[2]);
        if (!(L > 0.0)) return 0;
        /* r = ((p'-a) x (p'-b)) / |a-b|  (RadarEdgeFactor, radarFactor.hpp:34-39) */
        const double u[3] = {pw[0] - a[0], pw[1] - a[1], pw[2] - a[2]};
        const double v[3] = {pw[0] - b[0], pw[1] - b[1], pw[2] - b[2]};
        double nu[3];
        cross3(u, v, nu);
        /* dr/dp' = [b-a]x / L ; dp'/dxi = [-[p']x | I] */
        const double e[3] = {ba[0] / L, ba[1] / L, ba[2] / L};
        const double D[9] = {0, -e[2], e[1], e[2], 0, -e[0], -e[1], e[0], 0};
        const double Px[9] = {0, pw[2], -pw[1], -pw[2], 0, pw[0], pw[1], -pw[0], 0}; /* -[p']x */
        for (int i = 0; i < 3; ++i) {
            double J[6];
            for (int j = 0; j < 3; ++j) {
                J[j] = (D[3 * i] * Px[j] + D[3 * i + 1] * Px[3 + j]) + D[3 * i + 2] * Px[6 + j];
                J[3 + j] = D[3 * i + j];
            }
            acc_gn(acc, J, nu[i] / L);
        }
        acc[28] += 1.0;
        return 1;
    }
    return 0;
}

/* ---- RadarEdgeFactor / LidarPlaneFactor with an interpolation ratio s != 1 (radarFactor.hpp:26-32,78-84) ------------
 * The functors place the point with slerp(I, q, s) p + s t. The loop here follows them literally: the pose is converted to
 * the functor's (q = x,y,z,w ; t) parameters, the RESIDUAL is the functor restatement above (orc_res_edge /
 * orc_res_plane), and the JACOBIAN with respect to the left perturbation T <- exp(xi^) T is taken by central differences
 * of that functor (h = 1e-6: error ~1e-10 relative) — deliberately not the closed form the device kernel uses. */
static void pose_to_qt(const double T[16], double q[4], double t[3]) {
    const double m00 = T[0], m01 = T[1], m02 = T[2], m10 = T[4], m11 = T[5], m12 = T[6], m20 = T[8], m21 = T[9], m22 = T[10];
    const double tr = m00 + m11 + m22;
    double x, y, z, w;
    if (tr > 0.0) {
        const double S = sqrt(tr + 1.0) * 2.0;
        w = 0.25 * S;
        x = (m21 - m12) / S;
        y = (m02 - m20) / S;
        z = (m10 - m01) / S;
    } else if (m00 > m11 && m00 > m22) {
        const double S = sqrt(1.0 + m00 - m11 - m22) * 2.0;
        w = (m21 - m12) / S;
        x = 0.25 * S;
        y = (m01 + m10) / S;
        z = (m02 + m20) / S;
    } else if (m11 > m22) {
        const double S = sqrt(1.0 + m11 - m00 - m22) * 2.0;
        w = (m02 - m20) / S;
        x = (m01 + m10) / S;
        y = 0.25 * S;
        z = (m12 + m21) / S;
    } else {
        const double S = sqrt(1.0 + m22 - m00 - m11) * 2.0;
        w = (m10 - m01) / S;
        x = (m02 + m20) / S;
        y = (m12 + m21) / S;
        z = 0.25 * S;
    }
    const double nq = sqrt(x * x + y * y + z * z + w * w);
    q[0] = x / nq;
    q[1] = y / nq;
    q[2] = z / nq;
    q[3] = w / nq;
    t[0] = T[3];
    t[1] = T[7];
    t[2] = T[11];
}

/* the pose the functors place a point with: (slerp(I, q, s), s t) as a 4x4 matrix */
void orc_interp_pose(const double T[16], double s, double Ts[16]) {
    double q[4], t[3], qs[4];
    pose_to_qt(T, q, t);
    slerp_from_identity(s, q, qs);
    const double e[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int c = 0; c < 3; ++c) {
        double col[3];
        quat_rotate(qs, e[c], col);
        Ts[0 + c] = col[0];
        Ts[4 + c] = col[1];
        Ts[8 + c] = col[2];
    }
    Ts[3] = s * t[0];
    Ts[7] = s * t[1];
    Ts[11] = s * t[2];
    Ts[12] = Ts[13] = Ts[14] = 0.0;
    Ts[15] = 1.0;
}

/* residual rows (1 for the plane, 3 for the edge) of source point p under pose T */
static int interp_residual(int residual, const double T[16], double s, const double p[3], const float* tgt, const int32_t* nn, double r[3]) {
    double q[4], t[3];
    pose_to_qt(T, q, t);
    if (residual == ORC_P2PLANE_3PT) {
        const float *fj = tgt + 4 * (size_t)nn[0], *fl = tgt + 4 * (size_t)nn[1], *fm = tgt + 4 * (size_t)nn[2];
        const double j[3] = {fj[0], fj[1], fj[2]}, l[3] = {fl[0], fl[1], fl[2]}, m[3] = {fm[0], fm[1], fm[2]};
        orc_res_plane(q, t, p, j, l, m, s, r);
        return 1;
    }
    const float *fa = tgt + 4 * (size_t)nn[0], *fb = tgt + 4 * (size_t)nn[1];
    const double a[3] = {fa[0], fa[1], fa[2]}, b[3] = {fb[0], fb[1], fb[2]};
    orc_res_edge(q, t, p, a, b, s, r);
    return 3;
}

static int contribute_interp(int residual, const double T[16], double s, const float* psrc, const float* tgt, const int32_t* nn, int found,
                             double* acc) {
    const int need = residual == ORC_P2PLANE_3PT ? 3 : 2;
    if (found < need) return 0;
    if (residual == ORC_P2PLANE_3PT) { /* degenerate plane: the three points are collinear */
        const float *fj = tgt + 4 * (size_t)nn[0], *fl = tgt + 4 * (size_t)nn[1], *fm = tgt + 4 * (size_t)nn[2];
        const double jl[3] = {(double)fj[0] - fl[0], (double)fj[1] - fl[1], (double)fj[2] - fl[2]};
        const double jm[3] = {(double)fj[0] - fm[0], (double)fj[1] - fm[1], (double)fj[2] - fm[2]};
        double n[3];
        cross3(jl, jm, n);
        if (!(sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]) > 0.0)) return 0;
    } else {
        const float *fa = tgt + 4 * (size_t)nn[0], *fb = tgt + 4 * (size_t)nn[1];
        const double ba[3] = {(double)fb[0] - fa[0], (double)fb[1] - fa[1], (double)fb[2] - fa[2]};
        if (!(sqrt((ba[0] * ba[0] + ba[1] * ba[1]) + ba[2] * ba[2]) > 0.0)) return 0;
    }
    const double p[3] = {psrc[0], psrc[1], psrc[2]};
    double r0[3];
    const int rows = interp_residual(residual, T, s, p, tgt, nn, r0);
    double J[3][6];
    const double h = 1e-6;
    for (int i = 0; i < 6; ++i) {
        double xi[6] = {0, 0, 0, 0, 0, 0}, D[16], Tp[16], Tm[16], rp[3], rm[3];
        xi[i] = h;
        orc_se3_exp(xi, D);
        orc_mat4_mul(D, T, Tp);
        xi[i] = -h;
        orc_se3_exp(xi, D);
        orc_mat4_mul(D, T, Tm);
        interp_residual(residual, Tp, s, p, tgt, nn, rp);
        interp_residual(residual, Tm, s, p, tgt, nn, rm);
        for (int k = 0; k < rows; ++k) J[k][i] = (rp[k] - rm[k]) / (2.0 * h);
    }
    for (int k = 0; k < rows; ++k) acc_gn(acc, J[k], r0[k]);
    acc[28] += 1.0;
    return 1;
}

static int knn_k_for(const orc_opts* o) {
    switch (o->residual) {
        case ORC_P2P_SVD:
        case ORC_P2P_GN:
            return 1;
        case ORC_P2LINE:
            return 2;
        case ORC_P2PLANE_3PT:
            return 3;
        default:
            return o->k > 0 ? o->k : 5;
    }
}

int orc_accumulate(const float* src, int n, const float* tgt, int m, orc_knn_fn knn, void* ctx,
                   const orc_opts* o, const double T[16], double acc[ORC_ACC_LEN], int32_t* idx_out) {
    (void)m;
    const int k = knn_k_for(o);
    if (k > 16) return -1;
    float* q = (float*)malloc(sizeof(float) * 4 * (size_t)n);
    double* pw = (double*)malloc(sizeof(double) * 3 * (size_t)n);
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)n * k);
    float* d2 = (float*)malloc(sizeof(float) * (size_t)n * k);
    int32_t* fnd = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    const int interp = (o->residual == ORC_P2LINE || o->residual == ORC_P2PLANE_3PT) && o->interp_s > 0.0 && o->interp_s != 1.0;
    double Tplace[16];
    if (interp) orc_interp_pose(T, o->interp_s, Tplace);
    else memcpy(Tplace, T, sizeof(Tplace));
    orc_transform(Tplace, src, n, q, pw);
    knn(ctx, q, n, k, o->max_corr_dist, idx, d2, fnd);
    memset(acc, 0, sizeof(double) * ORC_ACC_LEN);
    int used = 0;
    for (int i = 0; i < n; ++i) {
        if (interp)
            used += contribute_interp(o->residual, T, o->interp_s, src + 4 * (size_t)i, tgt, idx + (size_t)i * k, fnd[i], acc);
        else
            used += contribute(o->residual, k, o->plane_thresh, pw + 3 * (size_t)i, tgt, idx + (size_t)i * k,
                               d2 + (size_t)i * k, fnd[i], acc);
    }
    if (idx_out) memcpy(idx_out, idx, sizeof(int32_t) * (size_t)n * k);
    free(q);
    free(pw);
    free(idx);
    free(d2);
    free(fnd);
    return used;
}

static void fitness_pass(const float* src, int n, orc_knn_fn knn, void* ctx, double max_dist, const double T[16],
                         orc_result* res) {
    float* q = (float*)malloc(sizeof(float) * 4 * (size_t)n);
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    float* d2 = (float*)malloc(sizeof(float) * (size_t)n);
    int32_t* fnd = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    orc_transform(T, src, n, q, NULL);
    knn(ctx, q, n, 1, max_dist, idx, d2, fnd);
    double s = 0;
    int c = 0;
    for (int i = 0; i < n; ++i)
        if (fnd[i] > 0) {
            s += (double)d2[i];
            ++c;
        }
    res->n_fitness = c;
    res->fitness = c ? s / c : INFINITY;
    free(q);
    free(idx);
    free(d2);
    free(fnd);
}

int orc_register(const float* src, int n, const float* tgt, int m, orc_knn_fn knn, void* ctx, const orc_opts* o,
                 double T_out[16], orc_result* res, double* dump_pose, double* dump_acc, int32_t* dump_idx) {
    const int k = knn_k_for(o);
    double T[16];
    memcpy(T, o->T0, sizeof(T));
    orc_result r;
    memset(&r, 0, sizeof(r));
    double mse_prev = INFINITY;
    int it = 0;
    r.converged = 0;
    for (; it < o->max_iterations; ++it) {
        double acc[ORC_ACC_LEN];
        if (dump_pose) memcpy(dump_pose + 16 * (size_t)it, T, sizeof(T));
        int used = orc_accumulate(src, n, tgt, m, knn, ctx, o, T, acc,
                                  dump_idx ? dump_idx + (size_t)it * n * k : NULL);
        if (used < 0) return -1;
        if (dump_acc) memcpy(dump_acc + ORC_ACC_LEN * (size_t)it, acc, sizeof(acc));
        r.n_corr = used;
        double D[16];
        if (o->residual == ORC_P2P_SVD) {
            if (used < 3) break; /* PCL: fewer than 3 correspondences -> not converged */
            const double cnt = acc[0];
            double pm[3], qm[3], H[9], R[9];
            for (int i = 0; i < 3; ++i) {
                pm[i] = acc[1 + i] / cnt;
                qm[i] = acc[4 + i] / cnt;
            }
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) H[3 * i + j] = acc[7 + 3 * i + j] / cnt - pm[i] * qm[j];
            orc_svd3_rotation(H, R);
            for (int i = 0; i < 3; ++i) {
                D[4 * i + 0] = R[3 * i + 0];
                D[4 * i + 1] = R[3 * i + 1];
                D[4 * i + 2] = R[3 * i + 2];
                D[4 * i + 3] = qm[i] - ((R[3 * i] * pm[0] + R[3 * i + 1] * pm[1]) + R[3 * i + 2] * pm[2]);
            }
            D[12] = D[13] = D[14] = 0;
            D[15] = 1;
            r.last_cost = acc[16] / cnt;
            orc_mat4_mul(D, T, T);
            if (o->early_exit) {
                if (fabs(r.last_cost - mse_prev) < o->mse_abs_eps) {
                    r.converged = 1;
                    ++it;
                    break;
                }
                mse_prev = r.last_cost;
            }
        } else {
            if (used < 6) break;
            double xi[6];
            if (orc_chol6_solve(acc, acc + 21, xi)) break;
            orc_se3_exp(xi, D);
            r.last_cost = acc[27];
            orc_mat4_mul(D, T, T);
            if (o->early_exit) {
                const double wn = sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]);
                const double vn = sqrt(xi[3] * xi[3] + xi[4] * xi[4] + xi[5] * xi[5]);
                if (wn < o->rot_eps && vn < o->trans_eps) {
                    r.converged = 1;
                    ++it;
                    break;
                }
            }
        }
    }
    if (it >= o->max_iterations) r.converged = 1; /* PCL: reaching max_iterations counts as converged */
    r.iterations = it;
    fitness_pass(src, n, knn, ctx, o->max_corr_dist, T, &r);
    memcpy(T_out, T, sizeof(T));
    if (res) *res = r;
    return 0;
}

/* ------------------------------------------------------------------ fp32 PCL mirror ---------------- */

int orc_icp_p2p_f32(const float* src, int n, const float* tgt, int m, orc_knn_fn knn, void* ctx,
                    int max_iterations, float T_out[16]) {
    (void)m;
    float T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    float* cur = (float*)malloc(sizeof(float) * 4 * (size_t)n);
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    float* d2 = (float*)malloc(sizeof(float) * (size_t)n);
    int32_t* fnd = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    memcpy(cur, src, sizeof(float) * 4 * (size_t)n);
    for (int it = 0; it < max_iterations; ++it) {
        knn(ctx, cur, n, 1, 0.0, idx, d2, fnd);
        float pm[3] = {0, 0, 0}, qm[3] = {0, 0, 0};
        int c = 0;
        for (int i = 0; i < n; ++i)
            if (fnd[i] > 0) {
                for (int a = 0; a < 3; ++a) {
                    pm[a] += cur[4 * i + a];
                    qm[a] += tgt[4 * (size_t)idx[i] + a];
                }
                ++c;
            }
        if (c < 3) break;
        for (int a = 0; a < 3; ++a) {
            pm[a] /= (float)c;
            qm[a] /= (float)c;
        }
        float Hf[9] = {0};
        for (int i = 0; i < n; ++i)
            if (fnd[i] > 0)
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b)
                        Hf[3 * a + b] += (cur[4 * i + a] - pm[a]) * (tgt[4 * (size_t)idx[i] + b] - qm[b]);
        double H[9], R[9];
        for (int a = 0; a < 9; ++a) H[a] = (double)(Hf[a] / (float)c);
        orc_svd3_rotation(H, R); /* the 3x3 factorisation itself is not the float-sensitive part */
        float D[16] = {0};
        for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b) D[4 * a + b] = (float)R[3 * a + b];
            D[4 * a + 3] = qm[a] - (D[4 * a] * pm[0] + D[4 * a + 1] * pm[1] + D[4 * a + 2] * pm[2]);
        }
        D[15] = 1.f;
        float Tn[16];
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) {
                float s = 0;
                for (int e = 0; e < 4; ++e) s += D[4 * a + e] * T[4 * e + b];
                Tn[4 * a + b] = s;
            }
        memcpy(T, Tn, sizeof(T));
        for (int i = 0; i < n; ++i) { /* PCL transforms the already-transformed cloud by the increment */
            float x = cur[4 * i], y = cur[4 * i + 1], z = cur[4 * i + 2];
            for (int a = 0; a < 3; ++a) cur[4 * i + a] = D[4 * a] * x + D[4 * a + 1] * y + D[4 * a + 2] * z + D[4 * a + 3];
        }
    }
    memcpy(T_out, T, sizeof(T));
    free(cur);
    free(idx);
    free(d2);
    free(fnd);
    return 0;
}

/* ------------------------------------------------------------------ map maintenance ---------------- */

static inline int in_box(const float* p, const float bmin[3], const float bmax[3]) {
    /* Search_by_range / Delete_by_range point test, ikd_Tree.cpp:1034,678 */
    return bmin[0] <= p[0] && bmax[0] > p[0] && bmin[1] <= p[1] && bmax[1] > p[1] && bmin[2] <= p[2] &&
           bmax[2] > p[2];
}

int orc_map_add_points(float* pts, uint8_t* valid, int* m_io, const float* add, int n, int downsample_on,
                       float ds) {
    int m = *m_io;
    int counter = 0;
    for (int i = 0; i < n; ++i) {
        const float* p = add + 4 * (size_t)i;
        float* slot = pts + 4 * (size_t)m; /* every offered point takes the next index */
        memcpy(slot, p, 4 * sizeof(float));
        if (!downsample_on) {
            valid[m++] = 1;
            continue;
        }
        float bmin[3], bmax[3], mid[3];
        for (int a = 0; a < 3; ++a) {
            bmin[a] = floorf(p[a] / ds) * ds;                                    /* ikd_Tree.cpp:432-437 */
            bmax[a] = bmin[a] + ds;
            mid[a] = (float)((double)bmin[a] + (double)(bmax[a] - bmin[a]) / 2.0); /* :438-440 */
        }
        float min_dist = dist2f(p, mid);
        int winner = -1; /* -1: the new point */
        int in_cnt = 0;
        for (int j = 0; j < m; ++j) {
            if (!valid[j] || !in_box(pts + 4 * (size_t)j, bmin, bmax)) continue;
            ++in_cnt;
            float d = dist2f(pts + 4 * (size_t)j, mid);
            if (d < min_dist) { /* strict: the new point wins ties, :447 */
                min_dist = d;
                winner = j;
            }
        }
        int same = (winner < 0);
        if (!same) {
            const float* w = pts + 4 * (size_t)winner;
            same = fabs(p[0] - w[0]) < 1e-6 && fabs(p[1] - w[1]) < 1e-6 && fabs(p[2] - w[2]) < 1e-6; /* :1422 */
        }
        valid[m] = 0;
        if (in_cnt > 1 || same) { /* :453-457 */
            for (int j = 0; j < m; ++j)
                if (valid[j] && in_box(pts + 4 * (size_t)j, bmin, bmax)) valid[j] = 0;
            if (winner < 0)
                valid[m] = 1;
            else
                valid[winner] = 1;
            ++counter;
        }
        ++m;
    }
    *m_io = m;
    return counter;
}

static float heading_of(const float* a, const float* b) { /* calc_heading, ikd_Tree.cpp:1434-1448 */
    float h;
    const float s = sqrtf(dist2f(a, b));
    const float as = asinf((a[0] - b[0]) / s);
    if (a[1] - b[1] < 0)
        h = (float)(180 + (double)(as * 180) / M_PI);
    else
        h = (float)((double)(-as * 180) / M_PI);
    if (h > 180 && h < 360) h = h - 360;
    return h;
}

int orc_map_sector(const float* pts, const uint8_t* valid, int m, const float c[3], float radius, float heading,
                   int32_t* out, int cap) {
    int cnt = 0;
    for (int j = 0; j < m; ++j) {
        const float* p = pts + 4 * (size_t)j;
        const int alive = !valid || valid[j];
        const float dh = fabsf(heading_of(p, c) - heading);
        /* ikd_Tree.cpp:1114-1116: (alive && in radius && dh < 60) || dh > 300. The reference lets
         * lazily-deleted nodes through the `dh > 300` arm until the next rebuild physically drops them
         * (timing dependent); the restatement never returns a deleted point. */
        if (alive && ((dist2f(p, c) <= radius * radius && dh < 60) || dh > 300)) {
            if (cnt < cap) out[cnt] = j;
            ++cnt;
        }
    }
    return cnt;
}

/* ------------------------------------------------------------------ box / radius search, deletes --- */
/* KD_TREE::Box_Search -> Search_by_range (ikd_Tree.cpp:401-405,1024-1051): non-deleted points with
 * min <= p < max on every axis (half-open, float compares). Output order in the reference is tree order;
 * here ascending index — compare as sets. */
static int in_box6(const float* p, const float* b6) { /* half-open box, b6 = min xyz, max xyz */
    return b6[0] <= p[0] && b6[3] > p[0] && b6[1] <= p[1] && b6[4] > p[1] && b6[2] <= p[2] && b6[5] > p[2];
}
int orc_map_box_search(const float* pts, const uint8_t* valid, int m, const float bmin[3], const float bmax[3], int32_t* out, int cap) {
    const float b6[6] = {bmin[0], bmin[1], bmin[2], bmax[0], bmax[1], bmax[2]};
    int cnt = 0;
    for (int j = 0; j < m; ++j)
        if ((!valid || valid[j]) && in_box6(pts + 4 * (size_t)j, b6)) {
            if (cnt < cap) out[cnt] = j;
            ++cnt;
        }
    return cnt;
}

/* KD_TREE::Radius_Search -> Search_by_radius (ikd_Tree.cpp:408-412,1054-1095): non-deleted points with
 * calc_dist(p, centre) <= radius * radius (float). The reference also takes whole subtrees whose bounding
 * sphere lies inside the query sphere, which is the same set up to float rounding exactly on the boundary. */
int orc_map_radius_search(const float* pts, const uint8_t* valid, int m, const float c[3], float radius, int32_t* out, int cap) {
    int cnt = 0;
    const float r2 = radius * radius;
    for (int j = 0; j < m; ++j)
        if ((!valid || valid[j]) && dist2f(pts + 4 * (size_t)j, c) <= r2) {
            if (cnt < cap) out[cnt] = j;
            ++cnt;
        }
    return cnt;
}

/* KD_TREE::Delete_Point_Boxes -> Delete_by_range(.., is_downsample = false) (ikd_Tree.cpp:544-565,656-719):
 * every non-deleted point inside a box is deleted; returns how many. `userdel` remembers that the deletion can be
 * undone by Add_Point_Boxes (points removed by down-sampling can not: point_downsample_deleted, :782,789). */
int orc_map_delete_boxes(const float* pts, uint8_t* valid, uint8_t* userdel, int m, const float* boxes6, int nb) {
    int cnt = 0;
    for (int b = 0; b < nb; ++b)
        for (int j = 0; j < m; ++j)
            if (valid[j] && in_box6(pts + 4 * (size_t)j, boxes6 + 6 * (size_t)b)) {
                valid[j] = 0;
                userdel[j] = 1;
                ++cnt;
            }
    return cnt;
}

/* KD_TREE::Add_Point_Boxes -> Add_by_range (ikd_Tree.cpp:500-519,771-824): points inside a box that were deleted by
 * Delete_Points / Delete_Point_Boxes come back. Returns how many (the reference returns nothing). */
int orc_map_add_boxes(const float* pts, uint8_t* valid, uint8_t* userdel, int m, const float* boxes6, int nb) {
    int cnt = 0;
    for (int b = 0; b < nb; ++b)
        for (int j = 0; j < m; ++j)
            if (!valid[j] && userdel[j] && in_box6(pts + 4 * (size_t)j, boxes6 + 6 * (size_t)b)) {
                valid[j] = 1;
                userdel[j] = 0;
                ++cnt;
            }
    return cnt;
}

/* KD_TREE::Delete_Points -> Delete_by_point (ikd_Tree.cpp:522-541,721-768): for every requested point, in order,
 * ONE non-deleted point with |dx|, |dy|, |dz| < EPSS = 1e-6 (same_point, :1422-1424) is deleted. The reference
 * finds it by descending the tree along the split comparisons (with coordinates equal on the split axis it can
 * miss an existing copy); the restatement deletes the lowest-index match. Returns how many were deleted. */
int orc_map_delete_points(const float* pts, uint8_t* valid, uint8_t* userdel, int m, const float* targets, int n) {
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        const float* t = targets + 4 * (size_t)i;
        for (int j = 0; j < m; ++j) {
            const float* p = pts + 4 * (size_t)j;
            if (valid[j] && (double)fabsf(p[0] - t[0]) < 1e-6 && (double)fabsf(p[1] - t[1]) < 1e-6 && (double)fabsf(p[2] - t[2]) < 1e-6) {
                valid[j] = 0;
                userdel[j] = 1;
                ++cnt;
                break;
            }
        }
    }
    return cnt;
}

/* ------------------------------------------------------------------ VoxelGrid (PCL 1.8 restatement) --- */
/* pcl::VoxelGrid<PointXYZI>::applyFilter with the defaults the reference uses (src/radar_odometry.cpp:426-429:
 * setLeafSize(0.5, 0.5, 0.5), downsample_all_data, min_points_per_voxel 0, no field filter). PCL (1.8, the ROS
 * Melodic version the reference builds against) is not part of /root/reference; this restates its published
 * algorithm: min/max over finite points, min_b = floor(min * inv_leaf), leaf index
 * (floor(x*inv) - min_b.x) + (floor(y*inv) - min_b.y) * div.x + (floor(z*inv) - min_b.z) * div.x * div.y in float
 * arithmetic, points grouped by leaf index ascending, centroid = float sums / count (CentroidPoint accumulators).
 * PCL orders the points of a leaf with an unstable std::sort; this restatement fixes the order to ascending input
 * index. Parity unpinned: the reference has no test vector for this step and PCL cannot be built here. */
typedef struct { uint32_t key; int32_t idx; } vg_pair;
static int vg_cmp(const void* a, const void* b) {
    const vg_pair *x = (const vg_pair*)a, *y = (const vg_pair*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

int orc_voxel_grid(const float* pts, const uint8_t* valid, int n, float leaf, float* out, int cap) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    int nf = 0;
    for (int i = 0; i < n; ++i) {
        const float* p = pts + 4 * (size_t)i;
        if ((valid && !valid[i]) || !isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
        for (int a = 0; a < 3; ++a) {
            if (p[a] < mn[a]) mn[a] = p[a];
            if (p[a] > mx[a]) mx[a] = p[a];
        }
        ++nf;
    }
    if (nf == 0) return 0;
    const float inv = 1.0f / leaf;
    int min_b[3];
    long long div[3];
    for (int a = 0; a < 3; ++a) {
        min_b[a] = (int)floorf(mn[a] * inv);
        div[a] = (long long)(int)floorf(mx[a] * inv) - min_b[a] + 1;
    }
    if ((double)div[0] * (double)div[1] * (double)div[2] > 2147483647.0) return -1;
    const int mul1 = (int)div[0], mul2 = (int)(div[0] * div[1]);
    vg_pair* pr = (vg_pair*)malloc(sizeof(vg_pair) * (size_t)nf);
    int k = 0;
    for (int i = 0; i < n; ++i) {
        const float* p = pts + 4 * (size_t)i;
        if ((valid && !valid[i]) || !isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
        const int i0 = (int)(floorf(p[0] * inv) - (float)min_b[0]);
        const int i1 = (int)(floorf(p[1] * inv) - (float)min_b[1]);
        const int i2 = (int)(floorf(p[2] * inv) - (float)min_b[2]);
        pr[k].key = (uint32_t)(i0 + i1 * mul1 + i2 * mul2);
        pr[k].idx = i;
        ++k;
    }
    qsort(pr, (size_t)nf, sizeof(vg_pair), vg_cmp);
    int cnt = 0;
    for (int s = 0; s < nf;) {
        int e = s;
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
        while (e < nf && pr[e].key == pr[s].key) {
            const float* p = pts + 4 * (size_t)pr[e].idx;
            for (int a = 0; a < 4; ++a) sum[a] = sum[a] + p[a];
            ++e;
        }
        if (cnt < cap) {
            const float fn = (float)(e - s);
            for (int a = 0; a < 4; ++a) out[4 * (size_t)cnt + a] = sum[a] / fn;
        }
        ++cnt;
        s = e;
    }
    free(pr);
    return cnt;
}

/* ------------------------------------------------------------------ GICP (fast_gicp restatement) --- */

/* eigenvector of the smallest eigenvalue of a symmetric 3x3 (cyclic Jacobi) */
static void smallest_eigvec3(const double C[9], double n[3]) {
    double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(A, C, sizeof(A));
    for (int sweep = 0; sweep < 50; ++sweep) {
        double off = fabs(A[1]) + fabs(A[2]) + fabs(A[5]);
        if (off <= 1e-20 * (fabs(A[0]) + fabs(A[4]) + fabs(A[8])) || off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double apq = A[3 * p + q];
                if (fabs(apq) < 1e-300) continue;
                const double theta = (A[3 * q + q] - A[3 * p + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) { /* A <- A J */
                    const double akp = A[3 * k + p], akq = A[3 * k + q];
                    A[3 * k + p] = c * akp - s * akq;
                    A[3 * k + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) { /* A <- J^T A */
                    const double apk = A[3 * p + k], aqk = A[3 * q + k];
                    A[3 * p + k] = c * apk - s * aqk;
                    A[3 * q + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[3 * k + p], vkq = V[3 * k + q];
                    V[3 * k + p] = c * vkp - s * vkq;
                    V[3 * k + q] = s * vkp + c * vkq;
                }
            }
    }
    int m = 0;
    if (A[4] < A[4 * m]) m = 1;
    if (A[8] < A[4 * m]) m = 2;
    const double len = sqrt(V[m] * V[m] + V[3 + m] * V[3 + m] + V[6 + m] * V[6 + m]);
    n[0] = V[m] / len;
    n[1] = V[3 + m] / len;
    n[2] = V[6 + m] / len;
}

void orc_gicp_normals(const float* pts, int n, orc_knn_fn knn, void* ctx, int k, double* out) {
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)n * k);
    float* d2 = (float*)malloc(sizeof(float) * (size_t)n * k);
    int32_t* fnd = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    knn(ctx, pts, n, k, 0.0, idx, d2, fnd);
    for (int i = 0; i < n; ++i) {
        const int f = fnd[i];
        double mean[3] = {0, 0, 0}, C[9] = {0};
        for (int j = 0; j < f; ++j)
            for (int a = 0; a < 3; ++a) mean[a] += pts[4 * (size_t)idx[(size_t)i * k + j] + a];
        for (int a = 0; a < 3; ++a) mean[a] /= (f > 0 ? f : 1);
        for (int j = 0; j < f; ++j) {
            double d[3];
            for (int a = 0; a < 3; ++a) d[a] = (double)pts[4 * (size_t)idx[(size_t)i * k + j] + a] - mean[a];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) C[3 * a + b] += d[a] * d[b];
        }
        for (int a = 0; a < 9; ++a) C[a] /= k; /* fast_gicp divides by k_correspondences_ */
        smallest_eigvec3(C, out + 3 * (size_t)i);
    }
    free(idx);
    free(d2);
    free(fnd);
}

static int inv3(const double A[9], double Ai[9]) {
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    if (!(fabs(det) > 0.0)) return 0;
    Ai[0] = c00 / det;
    Ai[1] = (A[2] * A[7] - A[1] * A[8]) / det;
    Ai[2] = (A[1] * A[5] - A[2] * A[4]) / det;
    Ai[3] = c01 / det;
    Ai[4] = (A[0] * A[8] - A[2] * A[6]) / det;
    Ai[5] = (A[2] * A[3] - A[0] * A[5]) / det;
    Ai[6] = c02 / det;
    Ai[7] = (A[1] * A[6] - A[0] * A[7]) / det;
    Ai[8] = (A[0] * A[4] - A[1] * A[3]) / det;
    return 1;
}

#define GICP_ALPHA 0.999 /* 1 - 1e-3: C = I - alpha n n^T */

/* Mahalanobis matrix of one correspondence under rotation R (row-major 3x3 inside T) */
static int gicp_mahalanobis(const double T[16], const double na[3], const double nb[3], double M[9]) {
    double ra[3];
    for (int i = 0; i < 3; ++i) ra[i] = T[4 * i] * na[0] + T[4 * i + 1] * na[1] + T[4 * i + 2] * na[2];
    double RCR[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) RCR[3 * i + j] = (i == j ? 2.0 : 0.0) - GICP_ALPHA * (nb[i] * nb[j] + ra[i] * ra[j]);
    return inv3(RCR, M);
}

int orc_gicp_linearize(const float* src, const double* sn, int n, const float* tgt, const double* tn, orc_knn_fn knn, void* ctx,
                       const orc_opts* o, const double T[16], double acc[ORC_ACC_LEN], int32_t* idx_out) {
    float* q = (float*)malloc(sizeof(float) * 4 * (size_t)n);
    double* pw = (double*)malloc(sizeof(double) * 3 * (size_t)n);
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    float* d2 = (float*)malloc(sizeof(float) * (size_t)n);
    int32_t* fnd = (int32_t*)malloc(sizeof(int32_t) * (size_t)n);
    orc_transform(T, src, n, q, pw);
    knn(ctx, q, n, 1, o->max_corr_dist, idx, d2, fnd);
    memset(acc, 0, sizeof(double) * ORC_ACC_LEN);
    int used = 0;
    for (int i = 0; i < n; ++i) {
        if (idx_out) idx_out[i] = fnd[i] > 0 ? idx[i] : -1;
        if (fnd[i] < 1) continue;
        const int j = idx[i];
        double M[9];
        if (!gicp_mahalanobis(T, sn + 3 * (size_t)i, tn + 3 * (size_t)j, M)) continue;
        const double* a = pw + 3 * (size_t)i;
        const double e[3] = {(double)tgt[4 * (size_t)j] - a[0], (double)tgt[4 * (size_t)j + 1] - a[1], (double)tgt[4 * (size_t)j + 2] - a[2]};
        /* J = [skew(Ta) | -I] */
        const double J[18] = {0, -a[2], a[1], -1, 0, 0, a[2], 0, -a[0], 0, -1, 0, -a[1], a[0], 0, 0, 0, -1};
        double MJ[18], Me[3];
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 6; ++c) MJ[6 * r + c] = M[3 * r] * J[c] + M[3 * r + 1] * J[6 + c] + M[3 * r + 2] * J[12 + c];
            Me[r] = M[3 * r] * e[0] + M[3 * r + 1] * e[1] + M[3 * r + 2] * e[2];
        }
        int t = 0;
        for (int r = 0; r < 6; ++r)
            for (int c = r; c < 6; ++c) acc[t++] += J[r] * MJ[c] + J[6 + r] * MJ[6 + c] + J[12 + r] * MJ[12 + c];
        for (int r = 0; r < 6; ++r) acc[21 + r] += J[r] * Me[0] + J[6 + r] * Me[1] + J[12 + r] * Me[2];
        acc[27] += e[0] * Me[0] + e[1] * Me[1] + e[2] * Me[2];
        acc[28] += 1.0;
        ++used;
    }
    free(q);
    free(pw);
    free(idx);
    free(d2);
    free(fnd);
    return used;
}

/* sum e^T M e with the correspondences and Mahalanobis matrices of the linearisation pose T_lin, at pose T */
static double gicp_error(const float* src, const double* sn, int n, const float* tgt, const double* tn, const int32_t* idx,
                         const double T_lin[16], const double T[16]) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        if (idx[i] < 0) continue;
        const int j = idx[i];
        double M[9], a[3];
        if (!gicp_mahalanobis(T_lin, sn + 3 * (size_t)i, tn + 3 * (size_t)j, M)) continue;
        xform1(T, src + 4 * (size_t)i, a);
        const double e[3] = {(double)tgt[4 * (size_t)j] - a[0], (double)tgt[4 * (size_t)j + 1] - a[1], (double)tgt[4 * (size_t)j + 2] - a[2]};
        for (int r = 0; r < 3; ++r) s += e[r] * (M[3 * r] * e[0] + M[3 * r + 1] * e[1] + M[3 * r + 2] * e[2]);
    }
    return s;
}

static int gicp_delta_converged(const double D[16], double rot_eps, double trans_eps) {
    double m = 0.0;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            const double v = fabs(D[4 * i + j] - (i == j ? 1.0 : 0.0)) / rot_eps;
            if (v > m) m = v;
        }
        const double v = fabs(D[4 * i + 3]) / trans_eps;
        if (v > m) m = v;
    }
    return m < 1.0;
}

int orc_gicp_register(const float* src, const double* sn, int n, const float* tgt, const double* tn, int m, orc_knn_fn knn, void* ctx,
                      const orc_opts* o, double T_out[16], orc_result* res, double* dump_pose, double* dump_acc) {
    (void)m;
    double x0[16];
    memcpy(x0, o->T0, sizeof(x0));
    orc_result r;
    memset(&r, 0, sizeof(r));
    double lambda = -1.0;
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    int it = 0, converged = 0, last_conv = 0;
    for (; it < o->max_iterations && !converged; ++it) {
        double acc[ORC_ACC_LEN];
        if (dump_pose) memcpy(dump_pose + 16 * (size_t)it, x0, sizeof(x0));
        r.n_corr = orc_gicp_linearize(src, sn, n, tgt, tn, knn, ctx, o, x0, acc, idx);
        if (dump_acc) memcpy(dump_acc + ORC_ACC_LEN * (size_t)it, acc, sizeof(acc));
        if (r.n_corr < 6) break;
        const double y0 = acc[27];
        r.last_cost = y0;
        if (lambda < 0.0) {
            double mx = 0.0;
            for (int i = 0; i < 6; ++i) mx = fmax(mx, fabs(acc[tri(i, i)]));
            lambda = 1e-9 * mx;
        }
        double nu = 2.0, D[16];
        int stepped = 0;
        for (int t = 0; t < 10; ++t) {
            double Hl[21], d[6];
            memcpy(Hl, acc, sizeof(Hl));
            for (int i = 0; i < 6; ++i) Hl[tri(i, i)] += lambda;
            if (orc_chol6_solve(Hl, acc + 21, d)) break;
            orc_se3_exp(d, D);
            double xi[16];
            orc_mat4_mul(D, x0, xi);
            const double yi = gicp_error(src, sn, n, tgt, tn, idx, x0, xi);
            double den = 0.0;
            for (int i = 0; i < 6; ++i) den += d[i] * (lambda * d[i] - acc[21 + i]);
            const double rho = (y0 - yi) / den;
            if (rho < 0) {
                if (gicp_delta_converged(D, o->rot_eps, o->trans_eps)) {
                    stepped = 1;
                    break;
                }
                lambda = nu * lambda;
                nu = 2 * nu;
                continue;
            }
            memcpy(x0, xi, sizeof(x0));
            const double f = 2 * rho - 1;
            lambda = lambda * fmax(1.0 / 3.0, 1 - f * f * f);
            stepped = 1;
            break;
        }
        if (!stepped) break; /* "lm not converged!!" */
        last_conv = gicp_delta_converged(D, o->rot_eps, o->trans_eps);
        if (o->early_exit) converged = last_conv; /* early_exit = 0: run every iteration, report the last verdict */
    }
    r.converged = o->early_exit ? converged : (it >= o->max_iterations ? last_conv : 0);
    r.iterations = it;
    fitness_pass(src, n, knn, ctx, o->max_corr_dist, x0, &r);
    memcpy(T_out, x0, sizeof(x0));
    if (res) *res = r;
    free(idx);
    return 0;
}

/* ------------------------------------------------------------------ Doppler filter ----------------- */

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

#define ORC_DEG2RAD(x) ((x) * 0.017453293) /* pcl_macros.h, as used at iterative_closest_point.cpp:106-108 */

int orc_doppler_filter(const float* rec, int n, int iterations, uint64_t seed, double sigma, double split, uint8_t* mask,
                       orc_doppler_out* out) {
    memset(out, 0, sizeof(*out));
    out->best_iteration = -1;
    if (n <= 0) return 0;
    if (iterations <= 0) iterations = (int)(n * 0.2); /* fitSineRansac(..., PointsNum * 0.2), :389 */
    float* arfa = (float*)malloc(sizeof(float) * (size_t)n);
    float* beta = (float*)malloc(sizeof(float) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        const float x = rec[5 * (size_t)i], y = rec[5 * (size_t)i + 1], z = rec[5 * (size_t)i + 2];
        const float dist = sqrtf(x * x + y * y + z * z);                              /* :378-379 */
        arfa[i] = (float)((double)((float)atan2((double)y, (double)x) * 180) / M_PI);  /* :382 */
        beta[i] = (float)((double)((float)asin((double)(z / dist)) * 180) / M_PI);     /* :383 */
    }
    int* scores = (int*)malloc(sizeof(int) * (size_t)iterations);
    double* As = (double*)malloc(sizeof(double) * (size_t)iterations);
    double* bs = (double*)malloc(sizeof(double) * (size_t)iterations);
#pragma omp parallel for schedule(dynamic, 8)
    for (int it = 0; it < iterations; ++it) {
        const int i1 = (int)(splitmix64(seed + 2ull * (uint64_t)it) % (uint64_t)n);
        const int i2 = (int)(splitmix64(seed + 2ull * (uint64_t)it + 1ull) % (uint64_t)n);
        const double v1 = rec[5 * (size_t)i1 + 4], v2 = rec[5 * (size_t)i2 + 4];
        const double k = (v1 * cos(ORC_DEG2RAD(beta[i1]))) / (v2 * cos(ORC_DEG2RAD(beta[i2])));
        const double b = atan((cos(ORC_DEG2RAD(arfa[i1])) - k * cos(ORC_DEG2RAD(arfa[i2]))) /
                              (sin(ORC_DEG2RAD(arfa[i1])) - k * sin(ORC_DEG2RAD(arfa[i2]))));
        const double A = cos(ORC_DEG2RAD(beta[i1])) * v1 / cos((ORC_DEG2RAD(arfa[i1])) + b);
        int sc = 0;
        for (int j = 0; j < n; ++j) {
            const double delta = (cos(ORC_DEG2RAD(beta[j])) * (double)rec[5 * (size_t)j + 4]) - (A * cos(ORC_DEG2RAD(arfa[j]) + b));
            if (fabs(delta) < sigma) ++sc;
        }
        scores[it] = sc;
        As[it] = A;
        bs[it] = b;
    }
    int best = -1, best_sc = 0;
    for (int it = 0; it < iterations; ++it)
        if (scores[it] > best_sc) { /* strict: the first maximum wins, :120 */
            best_sc = scores[it];
            best = it;
        }
    const double A = best >= 0 ? As[best] : 0.0, b = best >= 0 ? bs[best] : 0.0; /* reference leaves A = b = 0 when nothing scores */
    out->A = A;
    out->b = b;
    out->score = best_sc;
    out->best_iteration = best;
    double KK[6] = {0, 0, 0, 0, 0, 0}, Kv[3] = {0, 0, 0};
    int ns = 0;
    for (int j = 0; j < n; ++j) {
        const double vr = rec[5 * (size_t)j + 4];
        const double delta = (cos(ORC_DEG2RAD(beta[j])) * vr) - (A * cos(ORC_DEG2RAD(arfa[j]) + b));
        const int is_static = !(delta > split); /* :394-403 */
        if (mask) mask[j] = (uint8_t)is_static;
        if (!is_static) continue;
        ++ns;
        const double k0 = cos(ORC_DEG2RAD(arfa[j])) * cos(ORC_DEG2RAD(beta[j]));
        const double k1 = sin(ORC_DEG2RAD(arfa[j])) * cos(ORC_DEG2RAD(beta[j]));
        const double k2 = sin(ORC_DEG2RAD(beta[j]));
        KK[0] += k0 * k0; KK[1] += k0 * k1; KK[2] += k0 * k2; KK[3] += k1 * k1; KK[4] += k1 * k2; KK[5] += k2 * k2;
        Kv[0] += k0 * vr; Kv[1] += k1 * vr; Kv[2] += k2 * vr;
    }
    out->n_static = ns;
    const double M[9] = {KK[0], KK[1], KK[2], KK[1], KK[3], KK[4], KK[2], KK[4], KK[5]};
    double Mi[9];
    if (ns >= 3 && inv3(M, Mi))
        for (int a = 0; a < 3; ++a) out->v[a] = Mi[3 * a] * Kv[0] + Mi[3 * a + 1] * Kv[1] + Mi[3 * a + 2] * Kv[2];
    free(arfa);
    free(beta);
    free(scores);
    free(As);
    free(bs);
    return 0;
}
