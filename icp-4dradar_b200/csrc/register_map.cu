// register_map.cu — scan-to-map registration: the per-iteration hot path
//   transform -> exact kNN over the voxel grid -> distance gate -> residual + Jacobian -> fp64 accumulators
//   -> SE(3) solve -> pose update
// as ONE kernel per iteration (the last block to finish reduces the block partials in a fixed order and does
// the 3x3 Jacobi-SVD / 6x6 Cholesky update), the whole loop captured in a CUDA graph.
//
// Replaces pcl::IterativeClosestPoint::align (/root/reference/src/iterative_closest_point.cpp:510-514) and the
// FastGICP align call (/root/reference/src/radar_odometry.cpp:399-405) on top of the map that replaces the
// ikd-Tree (radar_odometry.cpp:390); residuals are those of /root/reference/include/radarFactor.hpp.
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>

#include "ctx.h"
#include "device_math.cuh"
#include "grid_knn.cuh"
#include "solve_warp.cuh"

namespace icp4r {

constexpr int RM_WARPS = 28;
constexpr int RM_BLOCKS_PER_SM = 1;  // measured: two 14-warp blocks per SM are no faster in batch mode (1.88 vs 1.85 ms) and slower for one scan (0.54 vs 0.46 ms)
constexpr int RM_THREADS = RM_WARPS * 32;
constexpr int RM_PARK = 32;    // plane fits parked per warp (P2PLANE_KNN)

enum { MODE_ITER = 0, MODE_ITER_NOSOLVE = 1, MODE_FITNESS = 2, MODE_FITNESS_NOFINAL = 3 };

struct ResultBlock {  // what travels back to the host in one copy
    double T[16];
    icp4r_result res;
    int xch_timeout;
    int pad;
};

// The pose increment D (3x4 in shared memory) as the map q -> q - D^-1 q: how far a point at q moved with this update.
__device__ __forceinline__ void publish_increment(RegState* st, const double* Ds, int iter, int lane) {
    if (lane < 12) {
        const int r = lane >> 2, cc = lane & 3;
        st->dA[lane] = cc < 3 ? (float)((r == cc ? 1.0 : 0.0) - Ds[4 * cc + r]) : (float)((Ds[r] * Ds[3] + Ds[4 + r] * Ds[7]) + Ds[8 + r] * Ds[11]);
    }
    if (lane == 0) st->last_pass = iter;
}

// T_s = (exp(s log R), s t) and M = s J_l(s phi) J_l(phi)^-1 for the interpolated functors (see reg_iter_kernel)
__device__ __noinline__ void interp_pose(const double* T, double s, double* Tss, double* M) {
    const double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
    // phi = log R
    const double w[3] = {0.5 * (R[7] - R[5]), 0.5 * (R[2] - R[6]), 0.5 * (R[3] - R[1])};  // sin(theta) * axis
    const double sn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const double cs = 0.5 * ((R[0] + R[4] + R[8]) - 1.0);
    const double th = atan2(sn, cs);
    const double f = th < 1e-8 ? 1.0 + th * th / 6.0 : th / sn;
    const double phi[3] = {f * w[0], f * w[1], f * w[2]};
    auto skew = [](const double* v, double* K) {
        K[0] = 0; K[1] = -v[2]; K[2] = v[1];
        K[3] = v[2]; K[4] = 0; K[5] = -v[0];
        K[6] = -v[1]; K[7] = v[0]; K[8] = 0;
    };
    auto mul3 = [](const double* A, const double* B, double* C) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) C[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
    };
    double K[9], K2[9];
    skew(phi, K);
    mul3(K, K, K2);
    const double th2 = th * th;
    // J_l(phi)^-1 = I - K/2 + (1/th^2 - (1 + cos)/(2 th sin)) K^2
    const double ci = th < 1e-4 ? 1.0 / 12.0 + th2 / 720.0 : 1.0 / th2 - (1.0 + cos(th)) / (2.0 * th * sin(th));
    double Jinv[9];
    for (int e = 0; e < 9; ++e) Jinv[e] = (e % 4 == 0 ? 1.0 : 0.0) - 0.5 * K[e] + ci * K2[e];
    // exp(s phi) and J_l(s phi): angle a = s th, generator Ks = s K
    const double a = s * th, a2 = a * a;
    const double A1 = a < 1e-4 ? 1.0 - a2 / 6.0 : sin(a) / a;
    const double B1 = a < 1e-4 ? 0.5 - a2 / 24.0 : (1.0 - cos(a)) / a2;
    const double C1 = a < 1e-4 ? 1.0 / 6.0 - a2 / 120.0 : (a - sin(a)) / (a2 * a);
    double Rs[9], Jl[9];
    for (int e = 0; e < 9; ++e) {
        const double I = e % 4 == 0 ? 1.0 : 0.0;
        Rs[e] = I + A1 * (s * K[e]) + B1 * (s * s * K2[e]);
        Jl[e] = I + B1 * (s * K[e]) + C1 * (s * s * K2[e]);
    }
    double JJ[9];
    mul3(Jl, Jinv, JJ);
    for (int e = 0; e < 9; ++e) M[e] = s * JJ[e];
    for (int r = 0; r < 3; ++r) {
        Tss[4 * r + 0] = Rs[3 * r + 0];
        Tss[4 * r + 1] = Rs[3 * r + 1];
        Tss[4 * r + 2] = Rs[3 * r + 2];
        Tss[4 * r + 3] = s * T[4 * r + 3];
    }
    Tss[12] = Tss[13] = Tss[14] = 0.0;
    Tss[15] = 1.0;
}

// Jacobian row of a residual with gradient g at the placed point lp (see reg_iter_kernel): out[0..2] = d/d omega, out[3..5] = d/d v
__device__ __forceinline__ void interp_row(const double g[3], const double lp[3], const double* T, const double* Tss, const double* M, double s,
                                           double* out) {
    const double u[3] = {lp[0] - Tss[3], lp[1] - Tss[7], lp[2] - Tss[11]};  // R_s p
    const double t[3] = {T[3], T[7], T[11]};
    double a[3], b[3];
    cross3(u, g, a);
    cross3(t, g, b);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        out[j] = ((a[0] * M[j] + a[1] * M[3 + j]) + a[2] * M[6 + j]) + s * b[j];
        out[3 + j] = s * g[j];
    }
}

// Pose update from the reduced accumulators, run by one full warp (see solve_warp.cuh). `tot`, `Ts` and `ws`
// (>= 32 doubles of scratch) are shared memory; every branch is warp-uniform.
__device__ __forceinline__ void warp_solve_and_update(int residual, const RegParams& P, RegState* st, const double* tot, const double* Ts,
                                                      double* ws, int iter, int lane) {
    const bool last = (iter == P.max_iterations - 1);
    double* Ds = ws;        // [12] increment, 3x4 row-major
    double* aux = ws + 16;  // [9] cross-covariance or [6] xi
    bool stop = false, conv = false;
    int its = iter + 1;
    if (residual == ICP4R_P2P_SVD) {
        const double cnt = tot[0];
        if (lane == 0) st->n_corr = (int)cnt;
        if (cnt < 3.0) {  // PCL: fewer than 3 correspondences -> not converged
            if (lane == 0) {
                st->done = 1;
                st->converged = 0;
                st->iterations = iter;
            }
            return;
        }
        // every lane solves the 3x3 problem redundantly in its own registers (thread_kabsch): no lane-to-lane exchange
        // on this serial tail of the iteration
        const double inv = 1.0 / cnt;
        const double pm[3] = {tot[1] * inv, tot[2] * inv, tot[3] * inv};
        const double qm[3] = {tot[4] * inv, tot[5] * inv, tot[6] * inv};
        double Hc[9], Rk[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) Hc[e] = tot[7 + e] * inv - pm[e / 3] * qm[e % 3];
        thread_kabsch(Hc, Rk);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                Ds[4 * r + 0] = Rk[3 * r + 0];
                Ds[4 * r + 1] = Rk[3 * r + 1];
                Ds[4 * r + 2] = Rk[3 * r + 2];
                Ds[4 * r + 3] = qm[r] - ((Rk[3 * r + 0] * pm[0] + Rk[3 * r + 1] * pm[1]) + Rk[3 * r + 2] * pm[2]);
            }
        }
        __syncwarp();
        const double tn = warp_compose_entry(Ds, Ts, lane);
        if (lane < 12) st->T[lane] = tn;
        publish_increment(st, Ds, iter, lane);
        const double mse = tot[16] / cnt;
        if (lane == 0) st->last_cost = mse;
        if (P.early_exit) {
            if (fabs(mse - __ldcg(&st->mse_prev)) < P.mse_abs_eps) {
                stop = true;
                conv = true;
            } else if (lane == 0) {
                st->mse_prev = mse;
            }
        }
    } else {
        const double cnt = tot[28];
        if (lane == 0) st->n_corr = (int)cnt;
        double x = 0.0;
        bool ok = cnt >= 6.0;
        if (ok) ok = warp_chol6_solve(tot, tot + 21, lane, x);
        if (!ok) {
            if (lane == 0) {
                st->done = 1;
                st->converged = 0;
                st->iterations = iter;
            }
            return;
        }
        if (lane < 6) aux[lane] = x;
        __syncwarp();
        const double de = warp_se3_exp_entry(aux, lane);
        if (lane < 12) Ds[lane] = de;
        __syncwarp();
        const double tn = warp_compose_entry(Ds, Ts, lane);
        if (lane < 12) st->T[lane] = tn;
        publish_increment(st, Ds, iter, lane);
        if (lane == 0) st->last_cost = tot[27];
        if (P.early_exit) {
            const double wn = sqrt(aux[0] * aux[0] + aux[1] * aux[1] + aux[2] * aux[2]);
            const double vn = sqrt(aux[3] * aux[3] + aux[4] * aux[4] + aux[5] * aux[5]);
            if (wn < P.rot_eps && vn < P.trans_eps) {
                stop = true;
                conv = true;
            }
        }
    }
    if (!stop && last) {
        stop = true;
        conv = true;  // PCL: reaching max_iterations counts as converged
        its = P.max_iterations;
    }
    if (stop && lane == 0) {
        st->done = 1;
        st->converged = conv ? 1 : 0;
        st->iterations = its;
    }
}

__device__ void write_result(const RegState* st, ResultBlock* out) {
    for (int i = 0; i < 16; ++i) out->T[i] = st->T[i];
    out->res.converged = st->converged;
    out->res.iterations = st->iterations;
    out->res.n_corr = st->n_corr;
    out->res.n_fitness = st->fit_cnt;
    out->res.fitness = st->fit_cnt > 0 ? st->fit_sum / (double)st->fit_cnt : INFINITY;
    out->res.last_cost = st->last_cost;
    out->xch_timeout = st->xch_timeout;
}

// One iteration (or the fitness pass). One warp owns one source point at a time, start to finish:
//   transform -> grid kNN (lanes stride the candidates) -> the k neighbours broadcast by shuffle ->
//   residual and Jacobian row(s) (every lane computes the same few fp64 values) ->
//   lane t adds ITS product J_i*J_j (or J_i*r, r*r, 1) to ITS accumulator: the 29 (or 17) running sums of
//   the normal equations live one per lane, in one register, for the whole kernel.
// No block barrier on the per-point path. At the end the 8 warps' lanes are summed in a fixed order into the
// block partial, and the last block to finish sums the partials in a fixed order and solves.
// LB = false compiles the kernel without the keep-the-neighbours proof (exactly k neighbours ranked, no per-point state):
// the flavour for a single small scan, where the proof cannot shorten the iteration's latency chain (one point per warp:
// the slowest point sets the time) and its extra registers only cost.
// FLAVOUR 0: lean, 1: with the keep-the-neighbours proof (LB), 2: block-chunked work distribution without the proof
// (slab-sharded maps: the ownership test needs the per-block lists, the proof does not pay on the maps that get sharded)
// Returns true in the block that finished last (the one that reduced and solved). `load_p`: (re)load the per-call
// parameter block into shared memory — the persistent loop below does it once, not every iteration.
template <int KIND, int K, int MODE, int FLAVOUR>
__device__ __forceinline__ bool reg_iter_body(const GridDesc& g, const RegParams* __restrict__ prm, RegState* __restrict__ st,
                                              double* __restrict__ partials, ResultBlock* __restrict__ out, const int iter, const bool load_p,
                                              const bool check_done = true) {
    constexpr bool FIT = (MODE == MODE_FITNESS || MODE == MODE_FITNESS_NOFINAL);
    constexpr bool LB = FLAVOUR == 1;
    constexpr bool CHUNKED = FLAVOUR != 0;

    constexpr int RK = FIT ? ICP4R_P2P_SVD : KIND;  // residual actually accumulated
    // blockIdx.y selects one of the independent scans of a batched call (each has its own parameters, state,
    // partials and result); single registrations launch with gridDim.y == 1
    prm += blockIdx.y;
    st += blockIdx.y;
    partials += (size_t)blockIdx.y * gridDim.x * ICP4R_ACC_LEN;
    out += blockIdx.y;
    // (state another block wrote during an earlier iteration of the SAME launch — the persistent loop — must not come
    // out of this SM's L1: volatile / ld.cg)
    if (!FIT && check_done && *reinterpret_cast<volatile int*>(&st->done)) return false;  // (the persistent loop looks once per block)
    if (FIT && st->xch_timeout) {  // a sharded loop that lost a peer: nothing to measure, hand the flag to the host
        if (MODE == MODE_FITNESS && blockIdx.x == 0 && threadIdx.x == 0) write_result(st, out);
        return false;
    }

    __shared__ WarpSegs segs[RM_WARPS];
    __shared__ double scr[RM_WARPS][3][12];  // per-warp operands of the lane products (up to 3 residual rows)
    constexpr bool PARKED = (!FIT && KIND == ICP4R_P2PLANE_KNN && K <= 8);  // plane fits done up to 32 points at a time, one per lane
    // per warp: PARK rows of 8 doubles (the residual row of a parked point) + PARK x (K + 1) ints (its neighbours + itself).
    // The flavour with the lane-parallel phase parks up to 32 points per warp in dynamic shared memory; the lean flavour
    // parks 8 in static arrays (constant addresses: the two base pointers cost it four registers it does not have).
    constexpr int PARK = LB ? RM_PARK : 8;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ double scrq_s[(PARKED && !LB) ? RM_WARPS : 1][8][8];
    __shared__ int nbq_s[(PARKED && !LB) ? RM_WARPS : 1][8][K + 1];
    double (*scrq)[PARK][8] = LB ? reinterpret_cast<double (*)[PARK][8]>(dyn_smem) : reinterpret_cast<double (*)[PARK][8]>(&scrq_s[0][0][0]);
    int (*nbq)[PARK][K + 1] = LB ? reinterpret_cast<int (*)[PARK][K + 1]>(dyn_smem + (size_t)(blockDim.x >> 5) * PARK * 8 * sizeof(double))
                                 : reinterpret_cast<int (*)[PARK][K + 1]>(&nbq_s[0][0][0]);
    __shared__ double red[RM_WARPS][ICP4R_ACC_LEN];
    __shared__ double tot[ICP4R_ACC_LEN];
    __shared__ double Ts[16];
    __shared__ bool is_last;
    __shared__ int own_list[RM_THREADS], own_cnt[RM_WARPS], own_n;  // sharded maps: this block's owned source points

    __shared__ RegParams P;  // per-call parameters: one copy per block instead of ~25 registers per thread
    __shared__ float s_dA[12];  // displacement map of the last pose increment and the pass it belongs to (publish_increment)
    __shared__ int s_last_pass;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid < 16) Ts[tid] = __ldcg(&st->T[tid]);
    if (tid >= 16 && tid < 28) s_dA[tid - 16] = __ldcg(&st->dA[tid - 16]);
    if (tid == 28) s_last_pass = __ldcg(&st->last_pass);
    if (load_p && tid >= 32 && tid < 32 + (int)(sizeof(RegParams) / 4)) reinterpret_cast<uint32_t*>(&P)[tid - 32] = reinterpret_cast<const uint32_t*>(prm)[tid - 32];
    __syncthreads();
    // RadarEdgeFactor / LidarPlaneFactor with an interpolation ratio s != 1 (radarFactor.hpp:26-32,78-84): the point is
    // placed with T_s = (slerp(I, q, s), s t) instead of T, for the search and for the residual alike. Under the left
    // perturbation T <- exp(xi^) T the placed point moves as
    //     d lp = -[R_s p]x M d omega - s [t]x d omega + s d v,      M = s J_l(s phi) J_l(phi)^-1,  phi = log R
    // (J_l = left Jacobian of SO(3)), so a residual row with gradient g has J = [ (R_s p x g)^T M + s (t x g)^T | s g^T ].
    constexpr bool CAN_INTERP = !FIT && (KIND == ICP4R_P2LINE || KIND == ICP4R_P2PLANE_3PT);
    __shared__ double Tss[CAN_INTERP ? 16 : 1], Ms[CAN_INTERP ? 9 : 1];
    const bool interp = CAN_INTERP && P.interp_s != 1.0;
    if (CAN_INTERP && interp) {
        if (tid == 0) interp_pose(Ts, P.interp_s, Tss, Ms);
        __syncthreads();
    }
    const double* Tx = (CAN_INTERP && interp) ? Tss : Ts;  // the pose source points are placed with

    const float4* __restrict__ pts = P.map_pts;  // map points by insertion index
    const SegAddr sgaddr = seg_addr(&segs[w]);
    // which two operands this lane multiplies (indices into scr[w][row][*])
    int ia = 8, ib = 8;  // scr[..][8] == 0
    if (RK == ICP4R_P2P_SVD) {  // scr = {1, p'x,p'y,p'z, qx,qy,qz, d2, 0}
        if (lane == 0) ia = 0, ib = 0;
        else if (lane < 7) ia = 0, ib = lane;
        else if (lane < 16) ia = 1 + (lane - 7) / 3, ib = 4 + (lane - 7) % 3;
        else if (lane == 16) ia = 0, ib = 7;
    } else {  // scr = {J0..J5, r, 1, 0}
        if (lane < 21) {
            int t = lane, i = 0;
            while (t >= 6 - i) {
                t -= 6 - i;
                ++i;
            }
            ia = i;
            ib = i + t;
        } else if (lane < 27) ia = lane - 21, ib = 6;
        else if (lane == 27) ia = 6, ib = 6;
        else if (lane == 28) ia = 7, ib = 7;
    }
    double acc = 0.0;  // this lane's running sum
    unsigned st_cand = 0, st_q = 0, st_settled = 0;  // work counters of this warp (icp4r_set_stats)
#ifdef ICP4R_PHASE_TIMING
    const long long tp0 = clock64();
#endif

    // P2PLANE_KNN: a warp finds the neighbours of its points one after the other but fits the planes of up to PARK
    // points at once, one point per lane (the fit is ~400 fp64 instructions that every lane would otherwise execute
    // redundantly for every single point); lane q keeps point q's neighbour indices until the flush
    int parked = 0;
    auto flush_parked = [&]() {
        if (PARKED) {
            __syncwarp();  // the parked indices were written by other lanes
            if (lane < parked) {
                const int* my_nb = nbq[w][lane];
                const float4 p = __ldg(P.src + my_nb[K]);
                double pw[3];
                xform_point(Ts, p.x, p.y, p.z, pw);
                float Pn[K][3];
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    if (j < P.k) {
                        const float4 cpt = __ldg(pts + my_nb[j]);
                        Pn[j][0] = cpt.x;
                        Pn[j][1] = cpt.y;
                        Pn[j][2] = cpt.z;
                    } else {
                        Pn[j][0] = Pn[j][1] = Pn[j][2] = 0.f;
                    }
                }
                double nrm[3], d;
                bool ok = plane_fit<K, float>(Pn, P.k, nrm, d);
                if (ok) {
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        if (j < P.k) {
                            const double e = ((nrm[0] * (double)Pn[j][0] + nrm[1] * (double)Pn[j][1]) + nrm[2] * (double)Pn[j][2]) + d;
                            if (!(fabs(e) <= P.plane_thresh)) ok = false;
                        }
                    }
                }
                double* s = scrq[w][lane];
                if (ok) {  // LidarPlaneNormFactor (radarFactor.hpp:122): r = n.p' + d, J = [(p' x n)^T | n^T]
                    double pxn[3];
                    cross3(pw, nrm, pxn);
                    s[0] = pxn[0]; s[1] = pxn[1]; s[2] = pxn[2];
                    s[3] = nrm[0]; s[4] = nrm[1]; s[5] = nrm[2];
                    s[6] = ((nrm[0] * pw[0] + nrm[1] * pw[1]) + nrm[2] * pw[2]) + d;
                    s[7] = 1.0;
                } else {
#pragma unroll
                    for (int t8 = 0; t8 < 8; ++t8) s[t8] = 0.0;
                }
            }
            __syncwarp();
            if (ia < 8) {
                for (int q = 0; q < parked; ++q) acc += scrq[w][q][ia] * scrq[w][q][ib];  // fixed order
            }
            __syncwarp();
            parked = 0;
        }
    };

    const int n = P.n;
    const int kq = FIT ? 1 : P.k;
    const int nwb = blockDim.x >> 5;  // warps in this block (<= RM_WARPS)

    // Sharded map: the rank whose slab holds the transformed point owns it. Ownership changes with the pose, and a
    // warp that strides over ALL points would find a random share of its points owned (the slowest warp sets the
    // kernel time), so each block first compacts the owned points of its contiguous chunk into an ordered list
    // (ballot + prefix: deterministic) and its warps stride over that list.
    // with only a point or two per warp the lane-parallel phase would be one more serial step on the iteration's latency
    // chain: the per-point path does the same test inline then
    const bool lane_phase = LB && PARKED && P.nb_state != nullptr && iter > 0 && (n + (int)gridDim.x - 1) / (int)gridDim.x >= 2 * nwb;
    // ---- one source point, start to finish, by one warp ---------------------------------------------------------------
    auto per_point = [&](const int i) {
        const float4 p = __ldg(P.src + i);
        double pw[3];
        xform_point(Tx, p.x, p.y, p.z, pw);
        const float qx = (float)pw[0], qy = (float)pw[1], qz = (float)pw[2];
        // What is remembered per source point from the previous pass: its k neighbours (nb_prev) and a lower bound LB on the
        // distance to every OTHER map point (nb_state). The pose update moved the point by delta, so the others are now at
        // least LB - delta away: if the farthest remembered neighbour is closer than that, the k nearest neighbours are
        // exactly the remembered ones — no search, only a re-ranking by (d2, index). Otherwise the remembered neighbours
        // still bound the k-th distance (hint): one pass over the ball of that radius, widened a little so that the
        // search also yields a useful new LB (the (k+1)-th distance, or the radius up to which it ranked every point).
        constexpr int KS = FIT ? 1 : (LB ? K + 1 : K);
        float hint = -1.0f;
        int32_t* nbp = P.nb_prev ? P.nb_prev + (size_t)i * K : nullptr;
        NbState* nbs = (LB && P.nb_state) ? P.nb_state + i : nullptr;
        const float margin_q = fmaxf(g.margin, 9.5367431640625e-7f * fmaxf(fabsf(qx), fmaxf(fabsf(qy), fabsf(qz))));
        uint64_t mine = KEY_EMPTY;
        bool settled = false;
        float lb_now = 0.0f;
        if (nbp != nullptr && (FIT ? P.max_iterations > 0 : iter > 0)) {
            const int kp = P.k;
            const int pj = lane < kp ? __ldcg(nbp + lane) : 0;
            float dh = 0.0f;
            if (lane < kp && pj >= 0) {
                const float4 c = __ldg(pts + pj);
                dh = dist2_exact(qx, qy, qz, c.x, c.y, c.z);
            }
            const bool all_valid = !__any_sync(FULL, pj < 0);
            const unsigned hb = __reduce_max_sync(FULL, __float_as_uint(dh));  // d2 >= 0: bit order == value order; NaN sorts last
            const float hbf = __uint_as_float(hb);
            // how far this point moved with the last pose increment (see publish_increment)
            const float ex = __fmaf_rn(s_dA[0], qx, __fmaf_rn(s_dA[1], qy, __fmaf_rn(s_dA[2], qz, s_dA[3])));
            const float ey = __fmaf_rn(s_dA[4], qx, __fmaf_rn(s_dA[5], qy, __fmaf_rn(s_dA[6], qz, s_dA[7])));
            const float ez = __fmaf_rn(s_dA[8], qx, __fmaf_rn(s_dA[9], qy, __fmaf_rn(s_dA[10], qz, s_dA[11])));
            const float delta = sqrtf(__fmaf_rn(ex, ex, __fmaf_rn(ey, ey, ez * ez))) * 1.0001f + 2.0f * margin_q;
            if (!lane_phase && all_valid && nbs != nullptr && hb < 0x7f800000u && (FIT || hbf <= P.gate_f)) {
                const float2 sv = __ldcg(reinterpret_cast<const float2*>(nbs));
                if (__float_as_int(sv.y) == P.epoch * 4096 + s_last_pass) {
                    lb_now = sv.x - delta;
                    settled = sqrtf(hbf) * 1.000001f + margin_q < lb_now;  // false for NaN
                }
            }
            if (settled) {
                const uint64_t key = (lane < kp) ? pack_key(dh, pj) : KEY_EMPTY;
                if (FIT) {
                    const uint64_t m1 = warp_min_u64(key);
                    if (lane == 0 && key_d2(m1) <= P.gate_f) mine = m1;
                } else {
                    int rank = 0;  // keys are unique (index in the low word): the ranks are a permutation
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        const uint64_t kj = __shfl_sync(FULL, key, j);
                        if (j < kp && kj < key) ++rank;
                    }
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        const uint64_t kj = __shfl_sync(FULL, key, j);
                        const int rj = __shfl_sync(FULL, rank, j);
                        if (j < kp && rj == lane) mine = kj;
                    }
                }
                ++st_settled;
            } else if (FIT) {
                // any remembered neighbour bounds the nearest distance: take the closest of them
                const unsigned hm = __reduce_min_sync(FULL, (lane < kp && pj >= 0) ? __float_as_uint(dh) : 0x7f800000u);
                if (hm < 0x7f800000u) hint = __uint_as_float(hm);
            } else if (all_valid) {
                if (hb < 0x7f800000u) {
                    hint = hbf;
                    if (nbs != nullptr) {
                        // widen the ball by a few times the last displacement (what is left of a converging loop's motion)
                        const float r = sqrtf(hbf);
                        // ... unless that displacement is still a good part of r: the wider ball would buy a bound that the next
                        // increment eats again (iterations 2-4 of the C2 batch looked at 176 candidates per search with the
                        // widening capped at r/2, against 109 for the unbounded first search)
                        const float slack = P.slack_a * delta > P.slack_b * r ? 0.0f : fmaxf(P.slack_a * delta, 0.05f * r);
                        const float wide = (r + slack) * (r + slack);
                        hint = (hbf <= P.gate_f) ? fminf(wide, fmaxf(P.gate_f, hbf)) : wide;
                    }
                }
            } else if (P.gate_r < 3.0e38f) {
                // fewer than k neighbours inside the gate last time: most likely still so. The gate itself is the
                // bound then: one pass over the gate ball instead of growing shell by shell up to it.
                hint = P.gate_f;
            }
        }
#ifdef ICP4R_KNN_TIMING
        const long long tq0 = clock64();
#endif
        if (!settled) {
            float cover2 = 0.0f;
            mine = warp_grid_knn<KS, false>(g, P.map_sorted, P.map_cell_start, nullptr, P.map_m, sgaddr, qx, qy, qz, P.gate_f, P.gate_r, lane, hint, &st_cand,
                                            &cover2);
            ++st_q;
            if (!FIT && nbs != nullptr) {
                // every map point that is not one of the kq nearest is at least this far away
                const uint64_t nxt = __shfl_sync(FULL, mine, kq < KS ? kq : KS - 1);
                const bool full = __shfl_sync(FULL, mine, kq - 1) != KEY_EMPTY;
                const float d_next = (kq < KS && nxt != KEY_EMPTY) ? key_d2(nxt) : INFINITY;
                const float lb2 = fminf(fminf(d_next, cover2), P.gate_f);
                lb_now = full ? fmaxf(sqrtf(lb2) * 0.999999f - margin_q, 0.0f) : 0.0f;
            }
        }
        if (!FIT && nbs != nullptr && lane == 0) __stcg(reinterpret_cast<float2*>(nbs), make_float2(lb_now, __int_as_float(P.epoch * 4096 + iter)));
#ifdef ICP4R_KNN_TIMING
        const int tq_cycles = (int)(clock64() - tq0);
#endif
        const bool have = (lane < kq) && (mine != KEY_EMPTY);
        if (!FIT && nbp != nullptr && lane < kq) nbp[lane] = have ? key_idx(mine) : -1;
        if (!FIT && P.dump_idx && lane < kq) P.dump_idx[((size_t)iter * n + i) * kq + lane] = have ? key_idx(mine) : -1;
#ifdef ICP4R_KNN_TIMING
        // dev probe: the last two slots of the index dump carry the search's cycle count and the kind of bound it had
        if (!FIT && P.dump_idx && kq >= 3 && lane == 0) {
            P.dump_idx[((size_t)iter * n + i) * kq + kq - 1] = tq_cycles;
            P.dump_idx[((size_t)iter * n + i) * kq + kq - 2] = hint < 0.0f ? 0 : (hint == P.gate_f ? 2 : 1);
        }
#endif
        const int found = __popc(__ballot_sync(FULL, have));
        float4 nb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (have) nb = __ldg(pts + key_idx(mine));  // lane j holds neighbour j

        int rows = 0;  // residual rows written to scr[w]
        if (RK == ICP4R_P2P_SVD) {
            if (found >= 1) {
                const float cx = __shfl_sync(FULL, nb.x, 0), cy = __shfl_sync(FULL, nb.y, 0), cz = __shfl_sync(FULL, nb.z, 0);
                const float d2 = key_d2(__shfl_sync(FULL, mine, 0));
                if (lane == 0) {
                    double* s = scr[w][0];
                    s[0] = 1.0; s[1] = pw[0]; s[2] = pw[1]; s[3] = pw[2];
                    s[4] = (double)cx; s[5] = (double)cy; s[6] = (double)cz; s[7] = (double)d2; s[8] = 0.0;
                }
                rows = 1;
            }
        } else if (RK == ICP4R_P2P_GN) {
            if (found >= 1) {
                const float cx = __shfl_sync(FULL, nb.x, 0), cy = __shfl_sync(FULL, nb.y, 0), cz = __shfl_sync(FULL, nb.z, 0);
                if (lane < 3) {  // LidarDistanceFactor (radarFactor.hpp:156-158): r = p' - c, J = [-[p']x | I]
                    double* s = scr[w][lane];
                    const double c3[3] = {(double)cx, (double)cy, (double)cz};
                    s[0] = lane == 0 ? 0.0 : (lane == 1 ? -pw[2] : pw[1]);
                    s[1] = lane == 0 ? pw[2] : (lane == 1 ? 0.0 : -pw[0]);
                    s[2] = lane == 0 ? -pw[1] : (lane == 1 ? pw[0] : 0.0);
                    s[3] = lane == 0 ? 1.0 : 0.0;
                    s[4] = lane == 1 ? 1.0 : 0.0;
                    s[5] = lane == 2 ? 1.0 : 0.0;
                    s[6] = pw[lane] - c3[lane];
                    s[7] = lane == 0 ? 1.0 : 0.0;  // the correspondence is counted once
                    s[8] = 0.0;
                }
                rows = 3;
            }
        } else if (RK == ICP4R_P2PLANE_KNN && PARKED) {
            if (found == kq && kq >= 3) {  // park the neighbours in lane `parked`; the fit happens in flush_parked()
                if (lane < kq) nbq[w][parked][lane] = key_idx(mine);
                if (lane == K) nbq[w][parked][K] = i;
                if (++parked == PARK) flush_parked();
            }
        } else if (RK == ICP4R_P2PLANE_KNN) {
            if (found == kq && kq >= 3) {
                float Pn[K][3];  // kept as float (converted at use) to hold register pressure down
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    Pn[j][0] = __shfl_sync(FULL, nb.x, j);
                    Pn[j][1] = __shfl_sync(FULL, nb.y, j);
                    Pn[j][2] = __shfl_sync(FULL, nb.z, j);
                }
                double nrm[3], d;
                bool ok = plane_fit<K, float>(Pn, kq, nrm, d);
                if (ok) {
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        if (j < kq) {
                            const double e = ((nrm[0] * (double)Pn[j][0] + nrm[1] * (double)Pn[j][1]) + nrm[2] * (double)Pn[j][2]) + d;
                            if (!(fabs(e) <= P.plane_thresh)) ok = false;
                        }
                    }
                }
                if (ok) {  // LidarPlaneNormFactor (radarFactor.hpp:122): r = n.p' + d, J = [(p' x n)^T | n^T]
                    if (lane == 0) {
                        double* s = scr[w][0];
                        double pxn[3];
                        cross3(pw, nrm, pxn);
                        s[0] = pxn[0]; s[1] = pxn[1]; s[2] = pxn[2];
                        s[3] = nrm[0]; s[4] = nrm[1]; s[5] = nrm[2];
                        s[6] = ((nrm[0] * pw[0] + nrm[1] * pw[1]) + nrm[2] * pw[2]) + d;
                        s[7] = 1.0; s[8] = 0.0;
                    }
                    rows = 1;
                }
            }
        } else if (RK == ICP4R_GICP) {
            // fast_gicp cost: e = b - T a, J = [skew(T a) | -I], weight M = (C_B + R C_A R^T)^-1 with C = I - 0.999 n n^T.
            // With (C_B + R C_A R^T) = L L^T the three rows L^-1 J, L^-1 e feed the same lane-owned accumulators.
            bool okc = false;
            double li[6] = {0, 0, 0, 0, 0, 0};
            const int j0 = found >= 1 ? key_idx(__shfl_sync(FULL, mine, 0)) : -1;
            if (found >= 1) {
                const float cx = __shfl_sync(FULL, nb.x, 0), cy = __shfl_sync(FULL, nb.y, 0), cz = __shfl_sync(FULL, nb.z, 0);
                const double* na = P.src_normals + 3 * (size_t)i;
                const double* nbn = P.tgt_normals + 3 * (size_t)j0;
                const double a0 = na[0], a1 = na[1], a2 = na[2], b0 = nbn[0], b1 = nbn[1], b2 = nbn[2];
                const double r0 = Ts[0] * a0 + Ts[1] * a1 + Ts[2] * a2, r1 = Ts[4] * a0 + Ts[5] * a1 + Ts[6] * a2,
                             r2 = Ts[8] * a0 + Ts[9] * a1 + Ts[10] * a2;
                const double al = 0.999;
                const double s00 = 2.0 - al * (b0 * b0 + r0 * r0), s10 = -al * (b1 * b0 + r1 * r0), s11 = 2.0 - al * (b1 * b1 + r1 * r1);
                const double s20 = -al * (b2 * b0 + r2 * r0), s21 = -al * (b2 * b1 + r2 * r1), s22 = 2.0 - al * (b2 * b2 + r2 * r2);
                const double l00 = sqrt(s00), l10 = s10 / l00, l20 = s20 / l00;
                const double d11 = s11 - l10 * l10;
                const double l11 = sqrt(d11), l21 = (s21 - l20 * l10) / l11;
                const double d22 = s22 - l20 * l20 - l21 * l21;
                const double l22 = sqrt(d22);
                okc = (s00 > 0.0) && (d11 > 0.0) && (d22 > 0.0);
                if (okc) {
                    li[0] = 1.0 / l00;
                    li[2] = 1.0 / l11;
                    li[5] = 1.0 / l22;
                    li[1] = -l10 * li[0] * li[2];
                    li[4] = -l21 * li[2] * li[5];
                    li[3] = -(l20 * li[0] + l21 * li[1]) * li[5];
                    if (lane < 3) {
                        const double e[3] = {(double)cx - pw[0], (double)cy - pw[1], (double)cz - pw[2]};
                        const double J[3][6] = {{0, -pw[2], pw[1], -1, 0, 0}, {pw[2], 0, -pw[0], 0, -1, 0}, {-pw[1], pw[0], 0, 0, 0, -1}};
                        const double w0 = lane == 0 ? li[0] : (lane == 1 ? li[1] : li[3]);
                        const double w1 = lane == 0 ? 0.0 : (lane == 1 ? li[2] : li[4]);
                        const double w2 = lane == 2 ? li[5] : 0.0;
                        double* s = scr[w][lane];
#pragma unroll
                        for (int cix = 0; cix < 6; ++cix) s[cix] = (w0 * J[0][cix] + w1 * J[1][cix]) + w2 * J[2][cix];
                        s[6] = (w0 * e[0] + w1 * e[1]) + w2 * e[2];
                        s[7] = lane == 0 ? 1.0 : 0.0;
                        s[8] = 0.0;
                    }
                    rows = 3;
                }
            }
            if (!FIT && lane == 0 && P.corr) {
                GicpCorr cr;
                cr.idx = okc ? j0 : -1;
                cr.pad = 0;
#pragma unroll
                for (int t6 = 0; t6 < 6; ++t6) cr.li[t6] = li[t6];
                P.corr[i] = cr;
            }
        } else if (RK == ICP4R_P2PLANE_3PT) {
            if (found >= 3) {  // LidarPlaneFactor (radarFactor.hpp:63-64,86), s = 1: plane through the 3 nearest points
                const double j3[3] = {(double)__shfl_sync(FULL, nb.x, 0), (double)__shfl_sync(FULL, nb.y, 0), (double)__shfl_sync(FULL, nb.z, 0)};
                const double l3[3] = {(double)__shfl_sync(FULL, nb.x, 1), (double)__shfl_sync(FULL, nb.y, 1), (double)__shfl_sync(FULL, nb.z, 1)};
                const double m3[3] = {(double)__shfl_sync(FULL, nb.x, 2), (double)__shfl_sync(FULL, nb.y, 2), (double)__shfl_sync(FULL, nb.z, 2)};
                const double jl[3] = {j3[0] - l3[0], j3[1] - l3[1], j3[2] - l3[2]};
                const double jm[3] = {j3[0] - m3[0], j3[1] - m3[1], j3[2] - m3[2]};
                double nrm[3];
                cross3(jl, jm, nrm);
                const double len = sqrt((nrm[0] * nrm[0] + nrm[1] * nrm[1]) + nrm[2] * nrm[2]);
                if (len > 0.0) {
                    nrm[0] /= len;
                    nrm[1] /= len;
                    nrm[2] /= len;
                    if (lane == 0) {
                        double* s = scr[w][0];
                        if (CAN_INTERP && interp) {
                            interp_row(nrm, pw, Ts, Tss, Ms, P.interp_s, s);
                        } else {
                            double pxn[3];
                            cross3(pw, nrm, pxn);
                            s[0] = pxn[0]; s[1] = pxn[1]; s[2] = pxn[2];
                            s[3] = nrm[0]; s[4] = nrm[1]; s[5] = nrm[2];
                        }
                        s[6] = ((pw[0] - j3[0]) * nrm[0] + (pw[1] - j3[1]) * nrm[1]) + (pw[2] - j3[2]) * nrm[2];
                        s[7] = 1.0; s[8] = 0.0;
                    }
                    rows = 1;
                }
            }
        } else if (RK == ICP4R_P2LINE) {
            if (found >= 2) {
                const double a[3] = {(double)__shfl_sync(FULL, nb.x, 0), (double)__shfl_sync(FULL, nb.y, 0), (double)__shfl_sync(FULL, nb.z, 0)};
                const double b[3] = {(double)__shfl_sync(FULL, nb.x, 1), (double)__shfl_sync(FULL, nb.y, 1), (double)__shfl_sync(FULL, nb.z, 1)};
                const double ba[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
                const double L = sqrt((ba[0] * ba[0] + ba[1] * ba[1]) + ba[2] * ba[2]);
                if (L > 0.0) {  // RadarEdgeFactor (radarFactor.hpp:34-39), s = 1
                    if (lane < 3) {
                        const double u[3] = {pw[0] - a[0], pw[1] - a[1], pw[2] - a[2]};
                        const double v[3] = {pw[0] - b[0], pw[1] - b[1], pw[2] - b[2]};
                        double nu[3];
                        cross3(u, v, nu);
                        const double e[3] = {ba[0] / L, ba[1] / L, ba[2] / L};
                        const double D[9] = {0, -e[2], e[1], e[2], 0, -e[0], -e[1], e[0], 0};
                        const double Px[9] = {0, pw[2], -pw[1], -pw[2], 0, pw[0], pw[1], -pw[0], 0};
                        double* s = scr[w][lane];
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            double Dr[3], nuj = 0.0;
#pragma unroll
                            for (int r3 = 0; r3 < 3; ++r3) {  // select row `lane` without dynamic register indexing
                                Dr[r3] = lane == 0 ? D[r3] : (lane == 1 ? D[3 + r3] : D[6 + r3]);
                            }
                            nuj = lane == 0 ? nu[0] : (lane == 1 ? nu[1] : nu[2]);
                            s[j] = (Dr[0] * Px[j] + Dr[1] * Px[3 + j]) + Dr[2] * Px[6 + j];
                            s[3 + j] = Dr[j];
                            s[6] = nuj / L;
                            if (CAN_INTERP && interp && j == 2) interp_row(Dr, pw, Ts, Tss, Ms, P.interp_s, s);  // overwrites s[0..5]
                        }
                        s[7] = lane == 0 ? 1.0 : 0.0;
                        s[8] = 0.0;
                    }
                    rows = 3;
                }
            }
        }
        if (rows) {
            __syncwarp();
            for (int r3 = 0; r3 < rows; ++r3) acc += scr[w][r3][ia] * scr[w][r3][ib];
            __syncwarp();
        }
    };

    if constexpr (!CHUNKED) {
        // lean flavour (a single small scan, never sharded): the warps stride over the points, nothing else
        for (int i = blockIdx.x * nwb + w; i < n; i += (int)gridDim.x * nwb) per_point(i);
    } else {
    // Work distribution. Every block owns a contiguous chunk of the source points and walks it blockDim.x points at a time:
    //   phase A  one point per THREAD: is the point this block's business at all (slab-sharded maps: the rank whose slab
    //            holds the transformed point owns it), and — P2PLANE_KNN after the first iteration — are its remembered
    //            neighbours provably still the k nearest (see the per-point path below)? Such points skip the search;
    //            their planes are fitted up to 32 at a time, one per lane, by their own warp.
    //   phase B  the remaining points are compacted, in order (ballot + prefix: deterministic), into a block-wide list
    //            and the warps stride over it one point per WARP: search, plane fit (parked), accumulation.
    // A static split of ALL points over the warps made the slowest warp (the one with the most searches) set the kernel
    // time: 30 % of the warp time of the batched C2 launch was spent at the block barrier (profiles/).
    const bool sharded = P.shard_axis >= 0;
    const int per = (n + (int)gridDim.x - 1) / (int)gridDim.x;
    const int cbeg = min(n, (int)blockIdx.x * per), cend = min(n, cbeg + per);
    for (int c0 = cbeg; c0 < cend; c0 += (int)blockDim.x) {
        const int my_i = (c0 + tid < cend) ? c0 + tid : -1;
        bool need = my_i >= 0;
        if (lane_phase || sharded) {
            const int kp = P.k;
            bool ok = false;
            int nbv[K];
            float lb_new = 0.0f;
            if (need) {
                const float4 p = __ldg(P.src + my_i);
                double pw[3];
                xform_point(Tx, p.x, p.y, p.z, pw);
                const float qx = (float)pw[0], qy = (float)pw[1], qz = (float)pw[2];
                if (sharded) {
                    const float v = P.shard_axis == 0 ? qx : (P.shard_axis == 1 ? qy : qz);
                    need = (v >= P.slab_lo) && (v < P.slab_hi);
                }
                if (need && lane_phase) {
                    const float margin_q = fmaxf(g.margin, 9.5367431640625e-7f * fmaxf(fabsf(qx), fmaxf(fabsf(qy), fabsf(qz))));
                    const int32_t* nbp = P.nb_prev + (size_t)my_i * K;
                    uint64_t key[K];
                    bool all_valid = true;
                    float hmax = 0.0f;
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        key[j] = KEY_EMPTY;
                        if (j < kp) {
                            const int pj = __ldcg(nbp + j);
                            if (pj < 0) {
                                all_valid = false;
                            } else {
                                const float4 c = __ldg(pts + pj);
                                const float d = dist2_exact(qx, qy, qz, c.x, c.y, c.z);
                                hmax = fmaxf(hmax, d);
                                if (!(d <= P.gate_f)) all_valid = false;  // also catches NaN
                                key[j] = pack_key(d, pj);
                            }
                        }
                    }
                    if (all_valid) {
                        const float2 sv = __ldcg(reinterpret_cast<const float2*>(P.nb_state + my_i));
                        if (__float_as_int(sv.y) == P.epoch * 4096 + s_last_pass) {
                            const float ex = __fmaf_rn(s_dA[0], qx, __fmaf_rn(s_dA[1], qy, __fmaf_rn(s_dA[2], qz, s_dA[3])));
                            const float ey = __fmaf_rn(s_dA[4], qx, __fmaf_rn(s_dA[5], qy, __fmaf_rn(s_dA[6], qz, s_dA[7])));
                            const float ez = __fmaf_rn(s_dA[8], qx, __fmaf_rn(s_dA[9], qy, __fmaf_rn(s_dA[10], qz, s_dA[11])));
                            const float delta = sqrtf(__fmaf_rn(ex, ex, __fmaf_rn(ey, ey, ez * ez))) * 1.0001f + 2.0f * margin_q;
                            lb_new = sv.x - delta;
                            ok = sqrtf(hmax) * 1.000001f + margin_q < lb_new;
                        }
                    }
                    if (ok) {
                        // ascending (d2, index): insertion network over the K packed keys in registers
#pragma unroll
                        for (int a2 = 1; a2 < K; ++a2)
#pragma unroll
                            for (int b2 = a2; b2 > 0; --b2) {
                                const uint64_t lo = key[b2 - 1] < key[b2] ? key[b2 - 1] : key[b2];
                                const uint64_t hi = key[b2 - 1] < key[b2] ? key[b2] : key[b2 - 1];
                                key[b2 - 1] = lo;
                                key[b2] = hi;
                            }
#pragma unroll
                        for (int j = 0; j < K; ++j) nbv[j] = j < kp ? key_idx(key[j]) : -1;
                        need = false;
                    }
                }
            }
            if (PARKED) {
                const unsigned okm = __ballot_sync(FULL, ok);
                if (okm != 0u) {
                    if (ok) {
                        const int slot = __popc(okm & ((1u << lane) - 1u));  // parked == 0 here: phase B flushes before it ends
                        int32_t* nbp = P.nb_prev + (size_t)my_i * K;
#pragma unroll
                        for (int j = 0; j < K; ++j)
                            if (j < kp) {
                                nbq[w][slot][j] = nbv[j];
                                nbp[j] = nbv[j];
                                if (P.dump_idx) P.dump_idx[((size_t)iter * n + my_i) * kp + j] = nbv[j];
                            }
                        nbq[w][slot][K] = my_i;
                        __stcg(reinterpret_cast<float2*>(P.nb_state + my_i), make_float2(lb_new, __int_as_float(P.epoch * 4096 + iter)));
                    }
                    parked = __popc(okm);
                    st_settled += (lane == 0) ? __popc(okm) : 0;
                    flush_parked();
                }
            }
        }
        const bool direct = !(lane_phase || sharded);  // every point of the round takes the per-point path: no list needed
        if (!direct) {   // the block's list of points that need the per-point path, in point order
            const unsigned bal = __ballot_sync(FULL, need);
            __syncthreads();  // the previous round's list is no longer read
            if (lane == 0) own_cnt[w] = __popc(bal);
            __syncthreads();
            int before = 0;
            for (int j = 0; j < w; ++j) before += own_cnt[j];
            if (need) own_list[before + __popc(bal & ((1u << lane) - 1u))] = my_i;
            if (tid == 0) {
                int t = 0;
                for (int j = 0; j < nwb; ++j) t += own_cnt[j];
                own_n = t;
            }
            __syncthreads();
        }
    const int list_n = direct ? min((int)blockDim.x, cend - c0) : own_n;
    for (int sl = w; sl < list_n; sl += nwb) per_point(direct ? c0 + sl : own_list[sl]);
    flush_parked();  // phase A of the next round parks from slot 0
    }
    }

    flush_parked();
    if (P.stats != nullptr && lane == 0) {
        atomicAdd(P.stats + 0, (unsigned long long)st_q);
        atomicAdd(P.stats + 1, (unsigned long long)st_cand);
        atomicAdd(P.stats + 2, (unsigned long long)st_settled);
    }
    // block partial in a fixed order: value v = sum over warps 0..7 of lane v's accumulator
#ifdef ICP4R_PHASE_TIMING
    const long long tp1 = clock64();
#endif
    red[w][lane] = acc;
    __syncthreads();
#ifdef ICP4R_PHASE_TIMING
    const long long tp2 = clock64();
    if (lane == 0) atomicMax((unsigned long long*)(partials + 160 * ICP4R_ACC_LEN) + blockIdx.x * 4 + 0, (unsigned long long)(tp1 - tp0));
    if (tid == 0) ((unsigned long long*)(partials + 160 * ICP4R_ACC_LEN))[blockIdx.x * 4 + 1] = (unsigned long long)(tp2 - tp0);
#endif
    // warp 0 publishes the block partial and takes the ticket: the fence that orders the two is needed only in the
    // lanes that wrote (a block-wide fence made all 28 warps wait for it), the acquire side only in the lane that
    // read the ticket — the barrier below hands the ordering on to the other threads, which read with ld.cg
    if (w == 0) {
        double x = 0.0;
#pragma unroll
        for (int j = 0; j < nwb; ++j) x += red[j][lane];
        partials[(size_t)blockIdx.x * ICP4R_ACC_LEN + lane] = x;
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            unsigned* tk = FIT ? &st->ticket_fit : &st->ticket;
            const unsigned t = atomicAdd(tk, 1u);
            is_last = (t == gridDim.x - 1);
            if (is_last) *tk = 0;
            __threadfence();
        }
    }
    __syncthreads();
    if (!is_last) return false;
#ifdef ICP4R_PHASE_TIMING
    const long long tp3 = clock64();
#endif

    // last block: reduce the block partials in a fixed order (deterministic for a given grid size)
    {
        const int v = tid & 31, grp8 = tid >> 5;
        // 8 independent chains keep 8 loads in flight per thread; the summation order is still fixed
        double s8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int nb_ = (int)gridDim.x;
        for (int b0 = grp8; b0 < nb_; b0 += nwb * 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int b = b0 + u * nwb;
                if (b < nb_) s8[u] += __ldcg(partials + (size_t)b * ICP4R_ACC_LEN + v);
            }
        }
        red[grp8][v] = ((s8[0] + s8[1]) + (s8[2] + s8[3])) + ((s8[4] + s8[5]) + (s8[6] + s8[7]));
    }
    __syncthreads();
    if (tid < ICP4R_ACC_LEN) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < nwb; ++j) s += red[j][tid];
        tot[tid] = s;
    }
    __syncthreads();
    if ((MODE == MODE_ITER || MODE == MODE_FITNESS) && P.xt != nullptr) {
        // Slab-sharded map, fused flavour: the sum over ranks happens HERE, over NVLink peer memory, instead of a
        // separate collective launch. Warp 0 of every rank's last block stores its 32 partial sums into every peer's
        // exchange buffer and collects the peers' sums from its own, adding them in rank order — so all ranks hold
        // bit-identical totals, take the same decisions and therefore run the same number of exchanges.
        // Transport: every double travels as two 8-byte words {32 data bits, 32-bit epoch}; an aligned 8-byte store is
        // one transaction, so a word seen with the expected epoch is complete — no flag word, no system-scope fence
        // (two block-wide fences cost 9 us per iteration before). The epoch is the count of exchanges (kept on the
        // device), consecutive exchanges alternate between two buffers: a rank can only overwrite buffer p after every
        // peer has published the epoch in between, i.e. after every peer finished reading p.
        if (w == 0) {
            const XchTable xt = *P.xt;
            Xch* mine = xt.peer[xt.rank];
            const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(&mine->seq) + 1ull;
            const int par = (int)(epoch & 1ull);
            const unsigned long long tag = (epoch & 0xFFFFFFFFull) << 32;
            const unsigned long long bits = (unsigned long long)__double_as_longlong(tot[lane]);
            const unsigned long long w0 = tag | (bits & 0xFFFFFFFFull), w1 = tag | (bits >> 32);
            for (int r = 0; r < xt.world; ++r) {
                volatile unsigned long long* dst = &xt.peer[r]->ll[par][xt.rank][2 * lane];
                dst[0] = w0;
                dst[1] = w1;
            }
            unsigned long long t0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            double sum = 0.0;
            bool late = false;
            for (int r = 0; r < xt.world; ++r) {
                volatile unsigned long long* src = &mine->ll[par][r][2 * lane];
                unsigned long long a, b;
                for (;;) {
                    a = src[0];
                    b = src[1];
                    if ((a & 0xFFFFFFFF00000000ull) == tag && (b & 0xFFFFFFFF00000000ull) == tag) break;
                    unsigned long long t1;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t1 - t0 > 10000000000ull) {  // 10 s: a peer never arrived; report instead of hanging the GPU
                        late = true;
                        break;
                    }
                }
                sum += __longlong_as_double((long long)((a & 0xFFFFFFFFull) | (b << 32)));
            }
            // a peer that never published: the sums are incomplete. Stop the loop right here — no solve on garbage, the
            // remaining iteration launches and the fitness pass become no-ops instead of spinning 10 s each — and report
            late = __any_sync(FULL, late);
            if (late && lane == 0) {
                st->xch_timeout = 1;
                st->done = 1;
                st->converged = 0;
                st->iterations = iter;
            }
            tot[lane] = sum;
            if (lane == 0) mine->seq = epoch;
        }
        __syncthreads();
        if (*reinterpret_cast<volatile int*>(&st->xch_timeout)) {
            if (MODE == MODE_FITNESS && tid == 0) write_result(st, out);
            return true;
        }
    }
    if (FIT) {
        if (tid == 0) {
            if (MODE == MODE_FITNESS) {
                st->fit_sum = tot[16];
                st->fit_cnt = (int)tot[0];
                write_result(st, out);
            } else {
                st->acc[0] = tot[0];
                st->acc[1] = tot[16];
            }
        }
        return true;
    }
    if (tid < ICP4R_ACC_LEN) {
        if (P.dump_acc) P.dump_acc[(size_t)iter * ICP4R_ACC_LEN + tid] = tot[tid];
        if (MODE == MODE_ITER_NOSOLVE) st->acc[tid] = tot[tid];
    }
    if (tid < 16 && P.dump_pose) P.dump_pose[(size_t)iter * 16 + tid] = Ts[tid];
#ifdef ICP4R_PHASE_TIMING
    const long long tp4 = clock64();
#endif
    if (MODE == MODE_ITER && w == 0) warp_solve_and_update(KIND, P, st, tot, Ts, &red[0][0], iter, lane);
#ifdef ICP4R_PHASE_TIMING
    if (tid == 0) {
        unsigned long long* dbg = (unsigned long long*)(partials + 160 * ICP4R_ACC_LEN) + 4 * 160;
        dbg[0] = (unsigned long long)(tp3 - tp0);   // last block: start of main loop -> start of final phase
        dbg[1] = (unsigned long long)(tp4 - tp3);   // final reduce
        dbg[2] = (unsigned long long)(clock64() - tp4);  // solve
    }
#endif
    return true;
}

template <int KIND, int K, int MODE, int FLAVOUR>
__global__ void __launch_bounds__(RM_THREADS, RM_BLOCKS_PER_SM)
    reg_iter_kernel(GridDesc g, const RegParams* __restrict__ prm, RegState* __restrict__ st,
                    double* __restrict__ partials, ResultBlock* __restrict__ out, int iter) {
    reg_iter_body<KIND, K, MODE, FLAVOUR>(g, prm, st, partials, out, iter, true);
}

// The whole iteration loop of ONE scan as ONE persistent launch (cooperative: every block is resident, one per SM).
// Iterations [it0, it1) of the lean flavour run back to back inside the kernel; what separates two iterations is no
// longer a kernel boundary (launch gap, parameter reload, cold L1) but a grid-wide hand-over built on the ticket the
// iteration already takes: the block that finishes last reduces, solves, updates the pose and then publishes
// st->loop_epoch = iteration + 1 (release); thread 0 of every other block spins on that word (acquire) and the block
// moves on. Early exit leaves the loop on the device (st->done), so there is no look at a flag from the host either.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <int KIND, int K>
__global__ void __launch_bounds__(RM_THREADS, RM_BLOCKS_PER_SM)
    reg_loop_kernel(GridDesc g, const RegParams* __restrict__ prm, RegState* __restrict__ st, double* __restrict__ partials,
                    ResultBlock* __restrict__ out, int it0, int it1) {
    __shared__ int s_done;  // one thread per block looks at the global flag: 130 k threads polling one L2 line would serialise
    for (int it = it0; it < it1; ++it) {
        const bool last = reg_iter_body<KIND, K, MODE_ITER, 0>(g, prm, st, partials, out, it, it == it0, false);
        __syncthreads();  // the last block's warp 0 has written the new pose; everybody is done with the shared scratch
        if (threadIdx.x == 0) {
            if (last) {
                st_release_gpu(&st->loop_epoch, it + 1);
            } else {
                while (ld_acquire_gpu(&st->loop_epoch) < it + 1) {
                }
            }
            s_done = *reinterpret_cast<volatile int*>(&st->done);
        }
        __syncthreads();
        if (s_done) break;
    }
}

// sharded path: solve after the cross-rank sum of st->acc
__global__ void __launch_bounds__(32) solve_kernel(int residual, const RegParams* __restrict__ prm, RegState* __restrict__ st, int iter) {
    if (st->done) return;
    __shared__ double tot[ICP4R_ACC_LEN], Ts[16], ws[32];
    const int lane = threadIdx.x;
    const RegParams P = *prm;
    tot[lane] = st->acc[lane];
    if (lane < 16) Ts[lane] = st->T[lane];
    __syncwarp();
    warp_solve_and_update(residual, P, st, tot, Ts, ws, iter, lane);
}
__global__ void fitness_final_kernel(RegState* __restrict__ st, ResultBlock* __restrict__ out) {
    if (threadIdx.x == 0) {
        st->fit_cnt = (int)st->acc[0];
        st->fit_sum = st->acc[1];
        write_result(st, out);
    }
}

__global__ void init_state_kernel(RegState* st, const double* T0) {
    st += blockIdx.x;   // one block per scan of a batched call
    T0 += 16 * blockIdx.x;
    const int t = threadIdx.x;
    if (t < 16) st->T[t] = T0[t];
    if (t < ICP4R_ACC_LEN) st->acc[t] = 0.0;
    if (t == 0) {
        st->mse_prev = INFINITY;
        st->last_cost = 0.0;
        st->lm_lambda = -1.0;
        st->lm_last_conv = 0;
        st->fit_sum = 0.0;
        st->fit_cnt = 0;
        st->xch_timeout = 0;
        st->done = 0;
        st->converged = 0;
        st->iterations = 0;
        st->n_corr = 0;
        st->ticket = 0;
        st->ticket_fit = 0;
        st->loop_epoch = 0;
        st->last_pass = -1;
    }
}

template <int KIND, int K, int MODE, int FLAVOUR>
static void launch_iter_mode(Ctx* c, int blocks, int threads, int nscan, const GridDesc& g, const RegParams* prm, RegState* st, double* partials,
                             ResultBlock* out, int iter) {
    constexpr bool FIT = (MODE == MODE_FITNESS || MODE == MODE_FITNESS_NOFINAL);
    constexpr bool PARKED = (!FIT && KIND == ICP4R_P2PLANE_KNN && K <= 8);
    // dynamic shared memory: the parked plane-fit rows (see reg_iter_kernel); sized for the largest block
    constexpr bool LB = FLAVOUR == 1;
    constexpr size_t per_warp = (PARKED && LB) ? (size_t)RM_PARK * (8 * sizeof(double) + (K + 1) * sizeof(int)) : 0;
    static bool attr_set = false;
    if (!attr_set && per_warp > 0) {
        cudaFuncSetAttribute(reg_iter_kernel<KIND, K, MODE, FLAVOUR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(per_warp * RM_WARPS));
        attr_set = true;
    }
    reg_iter_kernel<KIND, K, MODE, FLAVOUR><<<dim3(blocks, nscan), threads, per_warp * (threads / 32), c->stream>>>(g, prm, st, partials, out, iter);
}

template <int KIND, int K, int MODE>
static void launch_iter_flavour(Ctx* c, int flavour, int blocks, int threads, int nscan, const GridDesc& g, const RegParams* prm, RegState* st,
                                double* partials, ResultBlock* out, int iter) {
    // GICP keeps no per-point state between iterations (its LM step moves the pose outside the kernel): no proof flavour
    if (flavour == 1 && KIND != ICP4R_GICP) launch_iter_mode<KIND, K, MODE, (KIND != ICP4R_GICP ? 1 : 0)>(c, blocks, threads, nscan, g, prm, st, partials, out, iter);
    else if (flavour == 2) launch_iter_mode<KIND, K, MODE, 2>(c, blocks, threads, nscan, g, prm, st, partials, out, iter);
    else launch_iter_mode<KIND, K, MODE, 0>(c, blocks, threads, nscan, g, prm, st, partials, out, iter);
}

template <int KIND, int K>
static void launch_iter(Ctx* c, int mode, int blocks, int threads, int nscan, const GridDesc& g, const RegParams* prm, RegState* st,
                        double* partials, ResultBlock* out, int iter, int flavour) {
    switch (mode) {
        case MODE_ITER:
            launch_iter_flavour<KIND, K, MODE_ITER>(c, flavour, blocks, threads, nscan, g, prm, st, partials, out, iter);
            break;
        case MODE_ITER_NOSOLVE:
            launch_iter_flavour<KIND, K, MODE_ITER_NOSOLVE>(c, flavour, blocks, threads, nscan, g, prm, st, partials, out, iter);
            break;
        case MODE_FITNESS:
            launch_iter_flavour<KIND, K, MODE_FITNESS>(c, flavour, blocks, threads, nscan, g, prm, st, partials, out, iter);
            break;
        default:
            launch_iter_flavour<KIND, K, MODE_FITNESS_NOFINAL>(c, flavour, blocks, threads, nscan, g, prm, st, partials, out, iter);
            break;
    }
    c->launches += 1;
}

static void dispatch_iter(Ctx* c, int kind, int K, int mode, int blocks, int threads, int nscan, const GridDesc& g,
                          const RegParams* prm, RegState* st, double* partials, ResultBlock* out, int iter, int lb = 0) {
    switch (kind) {
        case ICP4R_P2P_SVD:
            launch_iter<ICP4R_P2P_SVD, 1>(c, mode, blocks, threads, nscan, g, prm, st, partials, out, iter, lb);
            break;
        case ICP4R_P2P_GN:
            launch_iter<ICP4R_P2P_GN, 1>(c, mode, blocks, threads, nscan, g, prm, st, partials, out, iter, lb);
            break;
        case ICP4R_P2LINE:
            launch_iter<ICP4R_P2LINE, 2>(c, mode, blocks, threads, nscan, g, prm, st, partials, out, iter, lb);
            break;
        case ICP4R_GICP:
            launch_iter<ICP4R_GICP, 1>(c, mode, blocks, threads, nscan, g, prm, st, partials, out, iter, lb);
            break;
        case ICP4R_P2PLANE_3PT:
            launch_iter<ICP4R_P2PLANE_3PT, 5>(c, mode, blocks, threads, nscan, g, prm, st, partials, out, iter, lb);
            break;
        default:
            if (K <= 5) launch_iter<ICP4R_P2PLANE_KNN, 5>(c, mode, blocks, threads, nscan, g, prm, st, partials, out, iter, lb);
            else if (K <= 8) launch_iter<ICP4R_P2PLANE_KNN, 8>(c, mode, blocks, threads, nscan, g, prm, st, partials, out, iter, lb);
            else launch_iter<ICP4R_P2PLANE_KNN, 16>(c, mode, blocks, threads, nscan, g, prm, st, partials, out, iter, lb);
            break;
    }
}

// persistent loop (reg_loop_kernel): cooperative launch, lean flavour, one scan
template <int KIND, int K>
static cudaError_t launch_loop_kind(Ctx* c, int blocks, int threads, GridDesc g, const RegParams* prm, RegState* st, double* partials,
                                    ResultBlock* out, int it0, int it1) {
    void* args[] = {&g, &prm, &st, &partials, &out, &it0, &it1};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&reg_loop_kernel<KIND, K>), dim3(blocks), dim3(threads), args, 0, c->stream);
}
static cudaError_t launch_loop(Ctx* c, int kind, int K, int blocks, int threads, const GridDesc& g, const RegParams* prm, RegState* st,
                               double* partials, ResultBlock* out, int it0, int it1) {
    c->launches += 1;
    switch (kind) {
        case ICP4R_P2P_SVD: return launch_loop_kind<ICP4R_P2P_SVD, 1>(c, blocks, threads, g, prm, st, partials, out, it0, it1);
        case ICP4R_P2P_GN: return launch_loop_kind<ICP4R_P2P_GN, 1>(c, blocks, threads, g, prm, st, partials, out, it0, it1);
        case ICP4R_P2LINE: return launch_loop_kind<ICP4R_P2LINE, 2>(c, blocks, threads, g, prm, st, partials, out, it0, it1);
        case ICP4R_P2PLANE_3PT: return launch_loop_kind<ICP4R_P2PLANE_3PT, 5>(c, blocks, threads, g, prm, st, partials, out, it0, it1);
        default:
            if (K <= 5) return launch_loop_kind<ICP4R_P2PLANE_KNN, 5>(c, blocks, threads, g, prm, st, partials, out, it0, it1);
            if (K <= 8) return launch_loop_kind<ICP4R_P2PLANE_KNN, 8>(c, blocks, threads, g, prm, st, partials, out, it0, it1);
            return launch_loop_kind<ICP4R_P2PLANE_KNN, 16>(c, blocks, threads, g, prm, st, partials, out, it0, it1);
    }
}

static int knn_k_for(const icp4r_opts* o) {
    switch (o->residual) {
        case ICP4R_P2P_SVD:
        case ICP4R_P2P_GN:
        case ICP4R_GICP:
            return 1;
        case ICP4R_P2LINE:
            return 2;
        case ICP4R_P2PLANE_3PT:
            return 3;
        default:
            return o->k > 0 ? o->k : 5;
    }
}

int register_against_map(Ctx* c, Map& mp, const float4* d_src, int n, const icp4r_opts* o, int shard_axis, float slab_lo,
                         float slab_hi, double* T_out_host, icp4r_result* res_host, const icp4r_dump* dump) {
    if (!mp.built) return fail(c, ICP4R_ERR_STATE, "registration target has no built map");
    if (o->residual < 0 || o->residual > ICP4R_P2PLANE_3PT) return fail(c, ICP4R_ERR_INVALID, "bad residual kind %d", o->residual);
    const int k = knn_k_for(o);
    if (k > ICP4R_MAX_K) return fail(c, ICP4R_ERR_INVALID, "k=%d exceeds ICP4R_MAX_K", k);
    if (o->residual == ICP4R_P2PLANE_KNN && k < 3) return fail(c, ICP4R_ERR_INVALID, "P2PLANE_KNN needs k >= 3");
    if (o->max_iterations < 0 || n < 0) return fail(c, ICP4R_ERR_INVALID, "negative size");
    const bool sharded = shard_axis >= 0;
    if (sharded && !(o->max_corr_dist > 0.0 && std::isfinite(o->max_corr_dist)))
        return fail(c, ICP4R_ERR_INVALID, "sharded registration needs a finite max_corr_dist (halo guarantee)");

    // per-call parameters -> pinned staging -> device
    CKS(reserve(c, c->d_params, sizeof(RegParams)));
    CKS(reserve(c, c->d_state, sizeof(RegState)));
    CKS(reserve(c, c->d_T, 16 * sizeof(double)));
    CKS(reserve(c, c->d_res, sizeof(ResultBlock)));
    // one source point per warp at a time; warps per block chosen so that the points spread over all SMs
    // (one block per SM keeps the number of partials the last block has to sum at <= sm_count)
    const int max_blocks = c->sm_count * RM_BLOCKS_PER_SM;
    int wpb = (n + max_blocks - 1) / max_blocks;
    wpb = std::min(std::max(wpb, 4), RM_WARPS);
    const int threads = wpb * 32;
    int blocks = std::min(std::max(1, (n + wpb - 1) / wpb), max_blocks);
    // sized once for the largest grid so the pointer baked into captured graphs never moves
    CKS(reserve(c, c->d_partials, (size_t)c->sm_count * 4 * ICP4R_ACC_LEN * sizeof(double) + 1024));

    struct Stage {
        RegParams prm;
        double T0[16];
        ResultBlock out;
        int done_flag;
    };
    Stage* hs = static_cast<Stage*>(c->h_pinned);
    RegParams& P = hs->prm;
    std::memset(&P, 0, sizeof(P));
    P.src = d_src;
    P.n = n;
    P.residual = o->residual;
    P.k = k;
    P.max_iterations = o->max_iterations;
    P.early_exit = o->early_exit;
    gate_params(o->max_corr_dist, &P.gate_f, &P.gate_r);
    P.rot_eps = o->rot_eps;
    P.trans_eps = o->trans_eps;
    P.mse_abs_eps = o->mse_abs_eps;
    P.plane_thresh = o->plane_thresh;
    P.slack_a = c->slack_a;
    P.slack_b = c->slack_b;
    P.interp_s = (o->interp_s > 0.0 && (o->residual == ICP4R_P2LINE || o->residual == ICP4R_P2PLANE_3PT)) ? o->interp_s : 1.0;
    P.dump_pose = dump ? dump->pose : nullptr;
    P.dump_acc = dump ? dump->acc : nullptr;
    P.dump_idx = dump ? dump->idx : nullptr;
    P.map_sorted = mp.grid.sorted;
    P.map_cell_start = mp.grid.cell_start;
    P.map_coarse = mp.grid.coarse;
    P.map_pts = mp.pts.as<float4>();
    P.map_m = mp.grid.m;
    P.shard_axis = sharded ? shard_axis : -1;
    P.slab_lo = slab_lo;
    P.slab_hi = slab_hi;
    const bool fused_shard = sharded && c->xch_ready && c->world >= 1;
    if (fused_shard) {
        P.xt = c->d_xt.as<XchTable>();
    }
    P.stats = c->stats ? c->d_stats.as<unsigned long long>() : nullptr;
    const bool gicp = o->residual == ICP4R_GICP;
    if (c->use_hints) {
        CKS(reserve_grow(c, c->d_nbprev, (size_t)n * ICP4R_MAX_K * sizeof(int32_t)));
        P.nb_prev = c->d_nbprev.as<int32_t>();
        if (c->use_lb && !gicp && !sharded && o->max_iterations < 4096 && P.interp_s == 1.0) {
            const void* before = c->d_nbstate.p;
            CKS(reserve_grow(c, c->d_nbstate, (size_t)std::max(n, 1) * sizeof(NbState)));
            if (c->d_nbstate.p != before) CK(cudaMemsetAsync(c->d_nbstate.p, 0xFF, c->d_nbstate.cap, c->stream));
            P.nb_state = c->d_nbstate.as<NbState>();
            P.epoch = (++c->reg_epoch) & 0x3FFFF;
        }
        // worth it only with many points per warp (see reg_iter_kernel), measured: 16 per warp (16 C2 scans per launch)
        // 2.14 -> 1.91 ms, 4 per warp (C5, dense 3-D map) 0.705 -> 0.95 ms (the wider search balls cost more than the
        // skipped searches save), 1 per warp (C2 single) no difference: the lean flavour below 8 per warp
        if ((n + blocks - 1) / blocks < 8 * wpb) P.nb_state = nullptr;
        // Sharded maps: a rank only refreshes the entries of the points it owns, so a point that changes owner finds
        // an older entry — still k valid points of this rank's slab, hence still a bound — or none (-1). Entries of
        // an earlier CALL must not survive (the map may have changed since): clear them.
        if (sharded && n > 0) CK(cudaMemsetAsync(c->d_nbprev.p, 0xFF, (size_t)n * ICP4R_MAX_K * sizeof(int32_t), c->stream));
    }
    if (gicp) {
        if (sharded) return fail(c, ICP4R_ERR_UNSUPPORTED, "GICP is not available for sharded maps yet");
        // covariances (normals) of both clouds from their own k nearest neighbours; the map's are cached per k
        const int kc = std::min(std::max(o->k > 0 ? o->k : 20, 3), ICP4R_MAX_K);
        CKS(gicp_normals(c, mp, kc));
        Map& sm = c->srcmap;
        if (n <= 8192) {
            // a scan: exhaustive k-NN (n^2 distances, tens of microseconds) beats building a grid that is searched once
            CKS(gicp_normals_small(c, d_src, n, kc, sm.normals));
        } else {
            sm.m = 0;
            sm.built = false;
            sm.quick_build = true;  // only its own 5-NN is searched, once: a volume-estimated cell is good enough
            sm.user_cell = 0.f;
            sm.hint_cell = 0.f;
            CKS(map_reserve(c, sm, n));
            CK(cudaMemcpyAsync(sm.pts.p, d_src, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
            CK(cudaMemsetAsync(sm.valid.p, 1, (size_t)n, c->stream));
            sm.m = n;
            CKS(map_rebuild_grid(c, sm));
            CKS(gicp_normals(c, sm, kc));
        }
        CKS(reserve_grow(c, c->d_gicp_corr, (size_t)std::max(n, 1) * sizeof(GicpCorr)));
        P.src_normals = sm.normals.as<double>();
        P.tgt_normals = mp.normals.as<double>();
        P.tgt_pts = mp.pts.as<float4>();
        P.corr = c->d_gicp_corr.as<GicpCorr>();
    }
    std::memcpy(hs->T0, o->T0, sizeof(hs->T0));
    const int lb = sharded ? 2 : (P.nb_state != nullptr ? 1 : 0);  // flavour of the iteration kernel (see reg_iter_kernel)

    RegParams* d_prm = c->d_params.as<RegParams>();
    RegState* d_st = c->d_state.as<RegState>();
    ResultBlock* d_out = c->d_res.as<ResultBlock>();
    double* d_part = c->d_partials.as<double>();
    CK(cudaMemcpyAsync(d_prm, &hs->prm, sizeof(RegParams), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_T.p, hs->T0, sizeof(hs->T0), cudaMemcpyHostToDevice, c->stream));
    init_state_kernel<<<1, 32, 0, c->stream>>>(d_st, c->d_T.as<double>());
    c->launches += 1;

    // geometry only by value: buffers and their length travel in RegParams (see ctx.h), so Add_Points between two
    // registrations does not invalidate the captured loop as long as the cell geometry stays
    GridDesc g = mp.grid;
    g.sorted = nullptr;
    g.cell_start = nullptr;
    g.coarse = nullptr;
    g.m = 0;
    const int iters = o->max_iterations;

    // The loop is a fixed sequence of launches whose arguments are all stable device pointers (per-call values
    // travel through RegParams in device memory), so it is captured once per (kind, k, grid size, iterations,
    // flavour, map identity) and replayed as a CUDA graph. Early exit is a device flag that turns the remaining
    // launches into no-ops. The NCCL all-reduce of the sharded flavour is captured like any other stream operation.
    const bool prof = c->profiling && !sharded && !gicp;
    int enqueue_status = ICP4R_OK;
    // iterations [it0, it1) and, if `fit`, the fitness pass
    auto enqueue_loop = [&](int it0, int it1, bool fit) {
        double* acc_ptr = reinterpret_cast<double*>(reinterpret_cast<char*>(d_st) + offsetof(RegState, acc));
        if (gicp) {
            // linearise (grid-wide, accumulators left in st->acc) then one single-block Levenberg-Marquardt step
            for (int it = it0; it < it1; ++it) {
                dispatch_iter(c, o->residual, k, MODE_ITER_NOSOLVE, blocks, threads, 1, g, d_prm, d_st, d_part, d_out, it, lb);
                gicp_lm_step(c, d_prm, d_st, it);
            }
            if (fit) dispatch_iter(c, o->residual, k, MODE_FITNESS, blocks, threads, 1, g, d_prm, d_st, d_part, d_out, 0, lb);
        } else if (sharded && !fused_shard) {
            for (int it = 0; it < iters; ++it) {
                dispatch_iter(c, o->residual, k, MODE_ITER_NOSOLVE, blocks, threads, 1, g, d_prm, d_st, d_part, d_out, it, lb);
                if (shard_allreduce(c, acc_ptr, ICP4R_ACC_LEN) != ICP4R_OK) enqueue_status = ICP4R_ERR_NCCL;
                solve_kernel<<<1, 32, 0, c->stream>>>(o->residual, d_prm, d_st, it);
                c->launches += 1;
            }
            dispatch_iter(c, o->residual, k, MODE_FITNESS_NOFINAL, blocks, threads, 1, g, d_prm, d_st, d_part, d_out, 0, lb);
            if (shard_allreduce(c, acc_ptr, 2) != ICP4R_OK) enqueue_status = ICP4R_ERR_NCCL;
            fitness_final_kernel<<<1, 32, 0, c->stream>>>(d_st, d_out);
            c->launches += 1;
        } else {
            if (prof) cudaEventRecord(c->prof_events[0], c->stream);
            for (int it = it0; it < it1; ++it) {
                dispatch_iter(c, o->residual, k, MODE_ITER, blocks, threads, 1, g, d_prm, d_st, d_part, d_out, it, lb);
                if (prof) cudaEventRecord(c->prof_events[it + 1], c->stream);
            }
            if (fit) dispatch_iter(c, o->residual, k, MODE_FITNESS, blocks, threads, 1, g, d_prm, d_st, d_part, d_out, 0, lb);
            if (prof) cudaEventRecord(c->prof_events[iters + 1], c->stream);
        }
    };
    if (prof) {
        while ((int)c->prof_events.size() < iters + 2) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            c->prof_events.push_back(e);
        }
    }
    // One scan, lean flavour, not sharded, not GICP: the whole loop is ONE cooperative launch (reg_loop_kernel) followed by
    // the fitness launch — no graph, no chunks, no look at the done flag from the host. The grid is one block per SM
    // at most, so all of it is resident, which is what the in-kernel hand-over between iterations needs.
    if (c->use_persist && c->coop_ok && !sharded && !gicp && !prof && lb == 0 && n > 0 && iters > 0 && blocks <= c->sm_count * RM_BLOCKS_PER_SM) {
        const cudaError_t le = launch_loop(c, o->residual, k, blocks, threads, g, d_prm, d_st, d_part, d_out, 0, iters);
        if (le == cudaSuccess) {
            dispatch_iter(c, o->residual, k, MODE_FITNESS, blocks, threads, 1, g, d_prm, d_st, d_part, d_out, 0, lb);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(&hs->out, d_out, sizeof(ResultBlock), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            c->prof_ms.clear();
            if (T_out_host) std::memcpy(T_out_host, hs->out.T, sizeof(hs->out.T));
            if (res_host) *res_host = hs->out.res;
            return ICP4R_OK;
        }
        cudaGetLastError();  // a device / driver that refuses the cooperative launch: the per-iteration path below
        c->coop_ok = false;
        c->launches -= 1;
    }
    const bool want_graph = c->use_graph && n > 0 && !c->profiling && !(sharded && c->no_graph_sharded) && iters < 4096;
    if (want_graph) {
        // graphs bake the grid geometry by value: drop them when it changed
        if (c->graph_grid_owner != &mp.grid || std::memcmp(&c->graph_grid_copy, &g, sizeof(GridDesc)) != 0) {
            for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
            c->graphs.clear();
            c->graph_grid_owner = &mp.grid;
            c->graph_grid_copy = g;
        }
    }
    auto run_range = [&](int it0, int it1, bool fit) -> int {
        GraphKey key{o->residual, k, blocks, it0 + 4096 * it1,
                     threads | (sharded ? 1 << 16 : 0) | (gicp ? 1 << 17 : 0) | (fused_shard ? 1 << 19 : 0) | (fit ? 1 << 21 : 0) | (lb << 22)};
        cudaGraphExec_t exec = nullptr;
        if (want_graph) {
            auto it = c->graphs.find(key);
            if (it != c->graphs.end()) exec = it->second;
        }
        if (want_graph && !exec) {
            cudaGraph_t graph = nullptr;
            // capture on the handle's own stream (a caller-supplied stream may be the legacy default stream, which
            // cannot be captured); the instantiated graph is then launched on c->stream
            cudaStream_t run_stream = c->stream;
            c->stream = c->own_stream;
            cudaError_t ce = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
            if (ce == cudaSuccess) {
                const int64_t before = c->launches;
                enqueue_loop(it0, it1, fit);
                c->graph_launches = c->launches - before;
                c->launches = before;  // counted at replay time below
                ce = cudaStreamEndCapture(c->stream, &graph);
            }
            c->stream = run_stream;  // restored on every path: later calls must stay ordered on the caller's stream
            if (ce == cudaSuccess && enqueue_status == ICP4R_OK) ce = cudaGraphInstantiate(&exec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
            if (ce != cudaSuccess || enqueue_status != ICP4R_OK) {
                // capture is an optimisation: whatever went wrong (an NCCL build that cannot be captured, a stream in a
                // state that refuses capture), the loop still runs as direct launches below
                cudaGetLastError();
                exec = nullptr;
                enqueue_status = ICP4R_OK;
                if (sharded) c->no_graph_sharded = true;
                else c->use_graph = false;
            } else {
                c->graphs[key] = exec;
                c->graph_launch_counts[key] = c->graph_launches;
            }
        }
        if (exec) {
            CK(cudaGraphLaunch(exec, c->stream));
            c->launches += c->graph_launch_counts[key];
        } else {
            enqueue_loop(it0, it1, fit);
            if (enqueue_status != ICP4R_OK) return enqueue_status;
        }
        return ICP4R_OK;
    };
    // With early exit the iterations after convergence are no-op launches (~2 us each): a 64-iteration budget that
    // converges after 8 would spend more time skipping than iterating. Run such loops in chunks and look at the
    // device's `done` flag in between (one 4-byte copy + sync per chunk).
    // The first chunk is short: a tracker that starts from a good prior (the odometry node: the previous pose) converges in
    // two or three outer iterations, and a chunk of 8 then spends more launches skipping than working.
    constexpr int CHUNK = 8, CHUNK0 = 4;
    const bool chunked = o->early_exit && iters > CHUNK + CHUNK / 2 && !sharded && !prof;
    if (!chunked) {
        CKS(run_range(0, iters, true));
    } else {
        for (int it0 = 0, step = CHUNK0; it0 < iters; it0 += step, step = CHUNK) {
            CKS(run_range(it0, std::min(it0 + step, iters), false));
            CK(cudaMemcpyAsync(&hs->done_flag, reinterpret_cast<const char*>(d_st) + offsetof(RegState, done), sizeof(int),
                               cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            if (hs->done_flag) break;
        }
        CKS(run_range(0, 0, true));
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&hs->out, d_out, sizeof(ResultBlock), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->prof_ms.clear();
#ifdef ICP4R_PHASE_TIMING
    if (c->profiling && !sharded) {
        std::vector<unsigned long long> dbg(4 * 160 + 4);
        cudaMemcpy(dbg.data(), d_part + 160 * ICP4R_ACC_LEN, dbg.size() * 8, cudaMemcpyDeviceToHost);
        unsigned long long mx0 = 0, mx1 = 0, mn1 = ~0ull; double avg1 = 0;
        for (int b = 0; b < blocks; ++b) { mx0 = std::max(mx0, dbg[4*b]); mx1 = std::max(mx1, dbg[4*b+1]); mn1 = std::min(mn1, dbg[4*b+1]); avg1 += dbg[4*b+1]; }
        fprintf(stderr, "[phase] blocks %d  warp-loop max %llu  block(min/avg/max) %llu/%.0f/%llu  last: to-final %llu final-reduce %llu solve %llu cycles\n",
                blocks, mx0, mn1, avg1 / blocks, mx1, dbg[4*160], dbg[4*160+1], dbg[4*160+2]);
        cudaMemset(d_part + 160 * ICP4R_ACC_LEN, 0, dbg.size() * 8);
    }
#endif
    if (prof) {
        for (int i = 0; i < iters + 1; ++i) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, c->prof_events[i], c->prof_events[i + 1]));
            c->prof_ms.push_back(ms);
        }
    }
    if (hs->out.xch_timeout) return fail(c, ICP4R_ERR_NCCL, "sharded registration: a peer rank never published its partial sums (timeout)");
    if (iters == 0) {  // no iteration ran: PCL reports convergence by max_iterations
        hs->out.res.converged = 1;
        hs->out.res.iterations = 0;
    }
    if (T_out_host) std::memcpy(T_out_host, hs->out.T, sizeof(hs->out.T));
    if (res_host) *res_host = hs->out.res;
    return ICP4R_OK;
}

// ------------------------------------------------------------------------------------------------ one slab pass
// One linearisation of slab-sharded registration at a given pose: the accumulators of the source points this slab owns,
// without the cross-rank sum and without the solve (MODE_ITER_NOSOLVE of the same kernel the sharded loop runs).
int accumulate_slab(Ctx* c, Map& mp, const float4* d_src, int n, const icp4r_opts* o, const double* T_host, int shard_axis, float slab_lo,
                    float slab_hi, double* acc_out_host) {
    if (!mp.built) return fail(c, ICP4R_ERR_STATE, "registration target has no built map");
    if (o->residual < 0 || o->residual > ICP4R_P2PLANE_3PT) return fail(c, ICP4R_ERR_INVALID, "bad residual kind %d", o->residual);
    if (o->residual == ICP4R_GICP) return fail(c, ICP4R_ERR_UNSUPPORTED, "icp4r_accumulate_slab does not support ICP4R_GICP");
    const int k = knn_k_for(o);
    if (k > ICP4R_MAX_K) return fail(c, ICP4R_ERR_INVALID, "k=%d exceeds ICP4R_MAX_K", k);
    if (o->residual == ICP4R_P2PLANE_KNN && k < 3) return fail(c, ICP4R_ERR_INVALID, "P2PLANE_KNN needs k >= 3");
    const bool sharded = shard_axis >= 0;
    if (sharded && !(o->max_corr_dist > 0.0 && std::isfinite(o->max_corr_dist)))
        return fail(c, ICP4R_ERR_INVALID, "a slab pass needs a finite max_corr_dist (halo guarantee)");
    CKS(reserve(c, c->d_params, sizeof(RegParams)));
    CKS(reserve(c, c->d_state, sizeof(RegState)));
    CKS(reserve(c, c->d_T, 16 * sizeof(double)));
    CKS(reserve(c, c->d_res, sizeof(ResultBlock)));
    CKS(reserve(c, c->d_partials, (size_t)c->sm_count * 4 * ICP4R_ACC_LEN * sizeof(double) + 1024));
    const int max_blocks = c->sm_count * RM_BLOCKS_PER_SM;
    int wpb = (n + max_blocks - 1) / max_blocks;
    wpb = std::min(std::max(wpb, 4), RM_WARPS);
    const int threads = wpb * 32;
    const int blocks = std::min(std::max(1, (n + wpb - 1) / wpb), max_blocks);
    struct Stage {
        RegParams prm;
        double T0[16];
        double acc[ICP4R_ACC_LEN];
    };
    Stage* hs = static_cast<Stage*>(c->h_pinned);
    RegParams& P = hs->prm;
    std::memset(&P, 0, sizeof(P));
    P.src = d_src;
    P.n = n;
    P.residual = o->residual;
    P.k = k;
    P.max_iterations = 1;
    gate_params(o->max_corr_dist, &P.gate_f, &P.gate_r);
    P.plane_thresh = o->plane_thresh;
    P.slack_a = c->slack_a;
    P.slack_b = c->slack_b;
    P.interp_s = (o->interp_s > 0.0 && (o->residual == ICP4R_P2LINE || o->residual == ICP4R_P2PLANE_3PT)) ? o->interp_s : 1.0;
    P.map_sorted = mp.grid.sorted;
    P.map_cell_start = mp.grid.cell_start;
    P.map_coarse = mp.grid.coarse;
    P.map_pts = mp.pts.as<float4>();
    P.map_m = mp.grid.m;
    P.shard_axis = sharded ? shard_axis : -1;
    P.slab_lo = slab_lo;
    P.slab_hi = slab_hi;
    P.stats = c->stats ? c->d_stats.as<unsigned long long>() : nullptr;
    std::memcpy(hs->T0, T_host, sizeof(hs->T0));
    RegParams* d_prm = c->d_params.as<RegParams>();
    RegState* d_st = c->d_state.as<RegState>();
    CK(cudaMemcpyAsync(d_prm, &hs->prm, sizeof(RegParams), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_T.p, hs->T0, sizeof(hs->T0), cudaMemcpyHostToDevice, c->stream));
    init_state_kernel<<<1, 32, 0, c->stream>>>(d_st, c->d_T.as<double>());
    c->launches += 1;
    GridDesc g = mp.grid;
    g.sorted = nullptr;
    g.cell_start = nullptr;
    g.coarse = nullptr;
    g.m = 0;
    if (n > 0) dispatch_iter(c, o->residual, k, MODE_ITER_NOSOLVE, blocks, threads, 1, g, d_prm, d_st, c->d_partials.as<double>(), c->d_res.as<ResultBlock>(), 0, sharded ? 2 : 0);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hs->acc, reinterpret_cast<const char*>(d_st) + offsetof(RegState, acc), sizeof(hs->acc), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::memcpy(acc_out_host, hs->acc, sizeof(hs->acc));
    return ICP4R_OK;
}

// ------------------------------------------------------------------------------------------------ batched scans
// nscan independent scans against the same map in ONE sequence of launches (gridDim.y = scan): a single scan of a
// few thousand points cannot fill the GPU (the iteration kernel is latency-bound, profiles/), several can.
int register_scans_against_map(Ctx* c, Map& mp, const float4* d_src, const int32_t* off_host, int nscan, const icp4r_opts* o,
                               const double* T0s_host, double* T_out_host, icp4r_result* res_host) {
    if (!mp.built) return fail(c, ICP4R_ERR_STATE, "registration target has no built map");
    if (o->residual < 0 || o->residual > ICP4R_P2PLANE_3PT) return fail(c, ICP4R_ERR_INVALID, "bad residual kind %d", o->residual);
    const bool gicp = o->residual == ICP4R_GICP;
    const int kc = std::min(std::max(o->k > 0 ? o->k : 20, 3), ICP4R_MAX_K);  // GICP: neighbours per covariance
    if (gicp) {
        // fast_gicp per scan: the map's covariances are shared (cached per k), every scan gets its own (exhaustive k-NN
        // inside the scan), its own correspondence table and its own Levenberg-Marquardt state; one linearisation launch
        // (gridDim.y = scan) and one LM launch (one block per scan) per outer iteration serve the whole batch.
        int nmax_all = 0;
        for (int b = 0; b < nscan; ++b) nmax_all = std::max(nmax_all, off_host[b + 1] - off_host[b]);
        if (nmax_all > 8192) {  // scans too large for the exhaustive covariance search: one after the other
            for (int b = 0; b < nscan; ++b) {
                icp4r_opts ob = *o;
                if (T0s_host) std::memcpy(ob.T0, T0s_host + 16 * (size_t)b, sizeof(ob.T0));
                CKS(register_against_map(c, mp, d_src + off_host[b], off_host[b + 1] - off_host[b], &ob, -1, 0.f, 0.f,
                                         T_out_host ? T_out_host + 16 * (size_t)b : nullptr, res_host ? res_host + b : nullptr, nullptr));
            }
            return ICP4R_OK;
        }
        CKS(gicp_normals(c, mp, kc));
    }
    const int k = knn_k_for(o);
    if (k > ICP4R_MAX_K) return fail(c, ICP4R_ERR_INVALID, "k=%d exceeds ICP4R_MAX_K", k);
    if (o->residual == ICP4R_P2PLANE_KNN && k < 3) return fail(c, ICP4R_ERR_INVALID, "P2PLANE_KNN needs k >= 3");
    if (o->max_iterations < 0) return fail(c, ICP4R_ERR_INVALID, "negative iteration count");
    struct Stage {
        RegParams prm;
        double T0[16];
        ResultBlock out;
    };
    const int CH = (int)std::min<size_t>(64, c->h_pinned_cap / sizeof(Stage));  // scans per launch sequence
    GridDesc g = mp.grid;  // geometry only (see register_against_map)
    g.sorted = nullptr;
    g.cell_start = nullptr;
    g.coarse = nullptr;
    g.m = 0;
    const int iters = o->max_iterations;
    for (int s0 = 0; s0 < nscan; s0 += CH) {
        const int B = std::min(CH, nscan - s0);
        int nmax = 0;
        for (int b = 0; b < B; ++b) nmax = std::max(nmax, off_host[s0 + b + 1] - off_host[s0 + b]);
        const int bps = std::max(1, c->sm_count * RM_BLOCKS_PER_SM / B);  // blocks per scan: the whole batch is one wave
        int wpb = (nmax + bps - 1) / std::max(bps, 1);
        wpb = std::min(std::max(wpb, 4), RM_WARPS);
        const int threads = wpb * 32;
        const int blocks = std::min(std::max(1, (nmax + wpb - 1) / wpb), bps);
        const void* old_ptrs[4] = {c->bm_params.p, c->bm_state.p, c->bm_res.p, c->bm_partials.p};
        CKS(reserve_grow(c, c->bm_params, (size_t)B * sizeof(RegParams)));
        CKS(reserve_grow(c, c->bm_state, (size_t)B * sizeof(RegState)));
        CKS(reserve_grow(c, c->bm_T0, (size_t)B * 16 * sizeof(double)));
        CKS(reserve_grow(c, c->bm_res, (size_t)B * sizeof(ResultBlock)));
        CKS(reserve_grow(c, c->bm_partials, (size_t)B * blocks * ICP4R_ACC_LEN * sizeof(double)));
        CKS(reserve_grow(c, c->d_nbprev, (size_t)off_host[s0 + B] * ICP4R_MAX_K * sizeof(int32_t)));
        const bool interp = o->interp_s > 0.0 && o->interp_s != 1.0 && (o->residual == ICP4R_P2LINE || o->residual == ICP4R_P2PLANE_3PT);
        const bool use_lb = c->use_hints && c->use_lb && !interp && !gicp && o->max_iterations < 4096 && (nmax + blocks - 1) / blocks >= 8 * wpb;
        if (gicp) {
            const int base = off_host[s0], tot = off_host[s0 + B] - base;
            CKS(reserve_grow(c, c->srcmap.normals, (size_t)std::max(tot, 1) * 3 * sizeof(double)));
            CKS(reserve_grow(c, c->d_gicp_corr, (size_t)std::max(tot, 1) * sizeof(GicpCorr)));
            for (int b = 0; b < B; ++b)
                CKS(gicp_normals_small_to(c, d_src + off_host[s0 + b], off_host[s0 + b + 1] - off_host[s0 + b], kc,
                                          c->srcmap.normals.as<double>() + 3 * (size_t)(off_host[s0 + b] - base)));
        }
        if (use_lb) {
            const void* before = c->d_nbstate.p;
            CKS(reserve_grow(c, c->d_nbstate, (size_t)std::max(off_host[s0 + B], 1) * sizeof(NbState)));
            if (c->d_nbstate.p != before) CK(cudaMemsetAsync(c->d_nbstate.p, 0xFF, c->d_nbstate.cap, c->stream));
            ++c->reg_epoch;
        }
        const bool moved = old_ptrs[0] != c->bm_params.p || old_ptrs[1] != c->bm_state.p || old_ptrs[2] != c->bm_res.p ||
                           old_ptrs[3] != c->bm_partials.p;
        Stage* hs = static_cast<Stage*>(c->h_pinned);
        for (int b = 0; b < B; ++b) {
            RegParams& P = hs[b].prm;
            std::memset(&P, 0, sizeof(P));
            P.src = d_src + off_host[s0 + b];
            P.n = off_host[s0 + b + 1] - off_host[s0 + b];
            P.residual = o->residual;
            P.k = k;
            P.max_iterations = o->max_iterations;
            P.early_exit = o->early_exit;
            gate_params(o->max_corr_dist, &P.gate_f, &P.gate_r);
            P.rot_eps = o->rot_eps;
            P.trans_eps = o->trans_eps;
            P.mse_abs_eps = o->mse_abs_eps;
            P.plane_thresh = o->plane_thresh;
            P.slack_a = c->slack_a;
            P.slack_b = c->slack_b;
    P.slack_a = c->slack_a;
    P.slack_b = c->slack_b;
            P.interp_s = (o->interp_s > 0.0 && (o->residual == ICP4R_P2LINE || o->residual == ICP4R_P2PLANE_3PT)) ? o->interp_s : 1.0;
            P.map_sorted = mp.grid.sorted;
            P.map_cell_start = mp.grid.cell_start;
            P.map_coarse = mp.grid.coarse;
            P.map_pts = mp.pts.as<float4>();
            P.map_m = mp.grid.m;
            P.shard_axis = -1;
            P.stats = c->stats ? c->d_stats.as<unsigned long long>() : nullptr;
            P.nb_prev = c->use_hints ? c->d_nbprev.as<int32_t>() + (size_t)off_host[s0 + b] * ICP4R_MAX_K : nullptr;
            P.nb_state = use_lb ? c->d_nbstate.as<NbState>() + off_host[s0 + b] : nullptr;
            P.epoch = c->reg_epoch & 0x3FFFF;
            if (gicp) {
                const size_t rel = (size_t)(off_host[s0 + b] - off_host[s0]);
                P.src_normals = c->srcmap.normals.as<double>() + 3 * rel;
                P.tgt_normals = mp.normals.as<double>();
                P.tgt_pts = mp.pts.as<float4>();
                P.corr = c->d_gicp_corr.as<GicpCorr>() + rel;
            }
            std::memcpy(hs[b].T0, T0s_host ? T0s_host + 16 * (size_t)(s0 + b) : o->T0, sizeof(hs[b].T0));
        }
        RegParams* d_prm = c->bm_params.as<RegParams>();
        RegState* d_st = c->bm_state.as<RegState>();
        ResultBlock* d_out = c->bm_res.as<ResultBlock>();
        double* d_part = c->bm_partials.as<double>();
        // the staging block interleaves params / T0 / out per scan: two strided copies
        CK(cudaMemcpy2DAsync(d_prm, sizeof(RegParams), &hs[0].prm, sizeof(Stage), sizeof(RegParams), B, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpy2DAsync(c->bm_T0.p, 16 * sizeof(double), hs[0].T0, sizeof(Stage), 16 * sizeof(double), B, cudaMemcpyHostToDevice, c->stream));
        init_state_kernel<<<B, 32, 0, c->stream>>>(d_st, c->bm_T0.as<double>());
        c->launches += 1;
        // iterations [it0, it1) of every scan and, if `fit`, the fitness pass
        auto enqueue = [&](int it0, int it1, bool fit) {
            for (int it = it0; it < it1; ++it) {
                if (gicp) {  // linearise every scan (accumulators left in its st->acc), then one LM step per scan
                    dispatch_iter(c, o->residual, k, MODE_ITER_NOSOLVE, blocks, threads, B, g, d_prm, d_st, d_part, d_out, it, 0);
                    gicp_lm_step(c, d_prm, d_st, it, B);
                } else {
                    dispatch_iter(c, o->residual, k, MODE_ITER, blocks, threads, B, g, d_prm, d_st, d_part, d_out, it, use_lb ? 1 : 0);
                }
            }
            if (fit) dispatch_iter(c, o->residual, k, MODE_FITNESS, blocks, threads, B, g, d_prm, d_st, d_part, d_out, 0, use_lb ? 1 : 0);
        };
        const bool want_graph = c->use_graph && !c->profiling && iters < 4096;
        if (want_graph && (moved || c->graph_grid_owner != &mp.grid || std::memcmp(&c->graph_grid_copy, &g, sizeof(GridDesc)) != 0)) {
            for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
            c->graphs.clear();
            c->graph_grid_owner = &mp.grid;
            c->graph_grid_copy = g;
        }
        auto run_range = [&](int it0, int it1, bool fit) -> int {
            GraphKey key{o->residual, k, blocks, it0 + 4096 * it1,
                         threads | (use_lb ? 1 << 16 : 0) | (gicp ? 1 << 17 : 0) | (1 << 18) | (fit ? 1 << 19 : 0) | (B << 20)};
            cudaGraphExec_t exec = nullptr;
            if (want_graph && c->use_graph) {
                auto itg = c->graphs.find(key);
                if (itg != c->graphs.end()) exec = itg->second;
                if (!exec) {
                    cudaGraph_t graph = nullptr;
                    cudaStream_t run_stream = c->stream;
                    c->stream = c->own_stream;
                    cudaError_t ce = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
                    if (ce == cudaSuccess) {
                        const int64_t before = c->launches;
                        enqueue(it0, it1, fit);
                        c->graph_launches = c->launches - before;
                        c->launches = before;
                        ce = cudaStreamEndCapture(c->stream, &graph);
                    }
                    c->stream = run_stream;  // restored on every path
                    if (ce == cudaSuccess) ce = cudaGraphInstantiate(&exec, graph, 0);
                    if (graph) cudaGraphDestroy(graph);
                    if (ce != cudaSuccess) {  // capture is an optimisation: run the loop as direct launches instead
                        cudaGetLastError();
                        exec = nullptr;
                        c->use_graph = false;
                    } else {
                        c->graphs[key] = exec;
                        c->graph_launch_counts[key] = c->graph_launches;
                    }
                }
            }
            if (exec) {
                CK(cudaGraphLaunch(exec, c->stream));
                c->launches += c->graph_launch_counts[key];
            } else {
                enqueue(it0, it1, fit);
            }
            return ICP4R_OK;
        };
        // loops with early exit and a long budget (fast_gicp: 64) run in chunks with a look at the scans' `done` flags in
        // between, like the single-scan path: the launches after convergence are no-ops but not free
        constexpr int CHUNK = 8;
        if (!(o->early_exit && iters > CHUNK + CHUNK / 2)) {
            CKS(run_range(0, iters, true));
        } else {
            std::vector<int> done(B);
            for (int it0 = 0; it0 < iters; it0 += CHUNK) {
                CKS(run_range(it0, std::min(it0 + CHUNK, iters), false));
                CK(cudaMemcpy2DAsync(done.data(), sizeof(int), reinterpret_cast<const char*>(d_st) + offsetof(RegState, done), sizeof(RegState),
                                     sizeof(int), B, cudaMemcpyDeviceToHost, c->stream));
                CK(cudaStreamSynchronize(c->stream));
                bool all = true;
                for (int b = 0; b < B; ++b) all = all && done[b] != 0;
                if (all) break;
            }
            CKS(run_range(0, 0, true));
        }
        CK(cudaGetLastError());
        CK(cudaMemcpy2DAsync(&hs[0].out, sizeof(Stage), d_out, sizeof(ResultBlock), sizeof(ResultBlock), B, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int b = 0; b < B; ++b) {
            if (iters == 0) {
                hs[b].out.res.converged = 1;
                hs[b].out.res.iterations = 0;
            }
            if (T_out_host) std::memcpy(T_out_host + 16 * (size_t)(s0 + b), hs[b].out.T, sizeof(hs[b].out.T));
            if (res_host) res_host[s0 + b] = hs[b].out.res;
        }
    }
    return ICP4R_OK;
}

}  // namespace icp4r
