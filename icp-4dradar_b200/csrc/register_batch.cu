// register_batch.cu — batched frame-pair registration: one CTA per (source, target) pair, both clouds and a
// per-pair voxel grid resident in shared memory for every iteration of the loop, so a pair costs one read of
// its two clouds from HBM and nothing else (16·(N+M)+64 bytes per registration).
//
// Replaces pcl::IterativeClosestPoint::align for scan-to-scan odometry
// (/root/reference/src/iterative_closest_point.cpp:510-521): 1-NN (exact, ties to the lowest index) ->
// Kabsch/Umeyama in fp64 (or the 6x6 Gauss-Newton form of LidarDistanceFactor, radarFactor.hpp:140-171).
//
// Per pair: (1) bounding box + cell geometry of the target, (2) counting sort of the target into cell order
// inside shared memory (and of the source, by the target cell it starts in), (3) max_iterations x { the exact 1-NN of
// every source point, fp64 accumulators in registers, shuffle + shared-memory reduction in a fixed order, one warp
// solves and updates the pose }, (4) fitness pass.
//
// How the exact 1-NN is kept cheap over the 30 iterations (all of it exact: the answer is always the one an exhaustive
// search with the (d2, index) order would give):
//   * every source point remembers its two nearest target points h1, h2 of the last search and a lower bound LB on its
//     distance to every OTHER target point. A pose update moves the point by delta, so the other points are now at
//     least LB - delta away, and if min(|q - h1|, |q - h2|) < LB - delta the nearer of the two is the unique nearest
//     point: no search at all. One uniform pass over the cloud (32 of 32 lanes active, no loops) settles most points
//     this way once the pose increments get small; keeping two candidates instead of one matters because the usual
//     reason for a failed test is the runner-up coming closer, not a third point.
//   * the points that fail the test are compacted, in order, into a per-warp list and searched: all cells that the box
//     around q with the radius of the farther remembered point touches (it contains the ball in which a better
//     neighbour would have to lie). The same pass yields the new h1, h2 and LB = min(distance of the third nearest
//     point seen, distance from q to the nearest face of that cell box that has unvisited cells behind it).
//   * without a remembered point (first iteration, or nothing inside the gate last time) the radius comes from the
//     nearest point of the query's own cell, or, if that cell is empty, from a cube-shell expansion around it.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ctx.h"
#include "device_math.cuh"
#include "solve_warp.cuh"

namespace icp4r {

constexpr int RB_MAXC = 4096;  // cells per pair

struct BatchParams {
    const float4* src;
    const int32_t* soff;
    const float4* tgt;
    const int32_t* toff;
    int n_pairs, max_n, max_m;
    int max_iterations, early_exit;
    float cell_pts;  // target points per cell (by volume) of the per-pair grid
    float slack;     // extra radius of the bounded search, in cells (a wider box costs candidates and buys a larger LB)
    int reproducible;  // place the source in index order (bit-reproducible sums) instead of with atomics (2.4 % faster)
    int use_hints;  // previous-iteration neighbours kept per source point (16-bit slots: needs max_m < 65535)
    float gate_f, gate_r;
    double rot_eps, trans_eps, mse_abs_eps;
    double T0[16];
    unsigned long long* stats;  // work counters (NULL: off), see icp4r_set_stats
    int* next_pair;  // NULL: pair = blockIdx.x + i * gridDim.x; else the CTAs draw pairs from this counter (zeroed before the launch)
};

struct PairGrid {
    float ox, oy, oz, cell, inv_cell, margin;
    int nx, ny, nz, ncells;
};

__device__ __forceinline__ int cell_of_s(float v, float o, float inv, int dim) {
    float f = floorf(__fmul_rn(__fsub_rn(v, o), inv));
    f = fminf(fmaxf(f, 0.0f), (float)(dim - 1));
    return (int)f;
}

struct BoxHit {
    float d1, d2;  // squared distances of the nearest and second nearest point (INFINITY: none)
    int p1, p2;    // their slots (-1: none)
    float lb;      // every other target point is at least this far away (metres)
};

// Bounded exact search: the caller knows that the nearest point lies in the closed ball of squared radius `hd` around
// q. Every point of every cell the enclosing box touches is looked at — no pruning, the loop body is load / distance /
// a 3-deep insertion — and the third smallest distance seen plus the distance to the faces of the visited cell box
// bound every target point other than the two winners from below.
// EXACT = false orders by distance only (the earlier slot wins a tie); the caller re-runs with EXACT = true, which
// orders by (distance, original index), when the two best distances come out equal — the only case where the
// difference can show.
template <bool EXACT>
__device__ __forceinline__ void box_scan(const PairGrid& g, const float4* __restrict__ s_tgt, const uint32_t* __restrict__ cs, float qx,
                                         float qy, float qz, float margin, float hd, float slack, BoxHit& out, unsigned& cand) {
    const float r = sqrtf(hd) * 1.000001f + margin + slack;
    const int xa = cell_of_s(qx - r, g.ox, g.inv_cell, g.nx), xb = cell_of_s(qx + r, g.ox, g.inv_cell, g.nx);
    const int y0 = cell_of_s(qy - r, g.oy, g.inv_cell, g.ny), y1 = cell_of_s(qy + r, g.oy, g.inv_cell, g.ny);
    const int z0 = cell_of_s(qz - r, g.oz, g.inv_cell, g.nz), z1 = cell_of_s(qz + r, g.oz, g.inv_cell, g.nz);
    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
    int p1 = -1, p2 = -1, i1 = 0x7fffffff, i2 = 0x7fffffff;
    for (int z = z0; z <= z1; ++z) {
        int rowbase = (z * g.ny + y0) * g.nx;
        for (int y = y0; y <= y1; ++y, rowbase += g.nx) {
            const uint32_t s = cs[rowbase + xa], e = cs[rowbase + xb + 1];
            cand += e - s;
            for (uint32_t j = s; j < e; ++j) {
                const float4 c = s_tgt[j];
                const float d = dist2_exact(qx, qy, qz, c.x, c.y, c.z);
                // insertion into (b1, b2, b3) written as selects: every comparison is false for a NaN, fminf drops it
                bool lt1, lt2;
                if (EXACT) {
                    const int ci = __float_as_int(c.w);
                    lt1 = d < b1 || (d == b1 && ci < i1);
                    lt2 = d < b2 || (d == b2 && ci < i2);
                    i2 = lt1 ? i1 : (lt2 ? ci : i2);
                    i1 = lt1 ? ci : i1;
                } else {
                    lt1 = d < b1;
                    lt2 = d < b2;
                }
                b3 = fminf(b3, lt2 ? b2 : d);
                p2 = lt1 ? p1 : (lt2 ? (int)j : p2);
                b2 = lt1 ? b1 : (lt2 ? d : b2);
                p1 = lt1 ? (int)j : p1;
                b1 = lt1 ? d : b1;
            }
        }
    }
    // faces of the visited cell box with cells behind them (the grid spans the target's bounding box: nothing lies outside)
    float bd = 3.4e38f;
    if (xa > 0) bd = fminf(bd, qx - (g.ox + (float)xa * g.cell));
    if (xb < g.nx - 1) bd = fminf(bd, (g.ox + (float)(xb + 1) * g.cell) - qx);
    if (y0 > 0) bd = fminf(bd, qy - (g.oy + (float)y0 * g.cell));
    if (y1 < g.ny - 1) bd = fminf(bd, (g.oy + (float)(y1 + 1) * g.cell) - qy);
    if (z0 > 0) bd = fminf(bd, qz - (g.oz + (float)z0 * g.cell));
    if (z1 < g.nz - 1) bd = fminf(bd, (g.oz + (float)(z1 + 1) * g.cell) - qz);
    out.d1 = b1;
    out.d2 = b2;
    out.p1 = p1;
    out.p2 = p2;
    out.lb = fmaxf(fminf(sqrtf(b3) * 0.999999f, bd - 2.0f * margin), 0.0f);
}
static __device__ __noinline__ void box_scan_exact(const PairGrid& g, const float4* __restrict__ s_tgt, const uint32_t* __restrict__ cs,
                                                   float qx, float qy, float qz, float margin, float hd, float slack, BoxHit& out, unsigned& cand) {
    box_scan<true>(g, s_tgt, cs, qx, qy, qz, margin, hd, slack, out, cand);
}
__device__ __forceinline__ void thread_box_nn(const PairGrid& g, const float4* __restrict__ s_tgt, const uint32_t* __restrict__ cs, float qx,
                                              float qy, float qz, float margin, float hd, float slack, BoxHit& out, unsigned& cand) {
    box_scan<false>(g, s_tgt, cs, qx, qy, qz, margin, hd, slack, out, cand);
    if (out.p2 >= 0 && out.d1 == out.d2) box_scan_exact(g, s_tgt, cs, qx, qy, qz, margin, hd, slack, out, cand);  // a tie for first place
}

// squared distance of the nearest point of the query's own cell (INFINITY if it is empty): a cheap radius for the
// bounded search when nothing is remembered
__device__ __forceinline__ float own_cell_seed(const PairGrid& g, const float4* __restrict__ s_tgt, const uint32_t* __restrict__ cs, float qx,
                                               float qy, float qz, unsigned& cand) {
    const int c = (cell_of_s(qz, g.oz, g.inv_cell, g.nz) * g.ny + cell_of_s(qy, g.oy, g.inv_cell, g.ny)) * g.nx +
                  cell_of_s(qx, g.ox, g.inv_cell, g.nx);
    float best = INFINITY;
    cand += cs[c + 1] - cs[c];
    for (uint32_t j = cs[c], e = cs[c + 1]; j < e; ++j) {
        const float4 t = s_tgt[j];
        best = fminf(best, dist2_exact(qx, qy, qz, t.x, t.y, t.z));
    }
    return best;
}

// exact 1-NN of (qx,qy,qz) over the shared-memory grid without prior knowledge; returns the packed key and the slot
__device__ __forceinline__ uint64_t thread_shell_nn(const PairGrid& g, const float4* __restrict__ s_tgt, const uint32_t* __restrict__ cs,
                                                    float qx, float qy, float qz, float gate_f, float gate_r, float margin, int& best_pos,
                                                    unsigned& cand) {
    const int cx = cell_of_s(qx, g.ox, g.inv_cell, g.nx);
    const int cy = cell_of_s(qy, g.oy, g.inv_cell, g.ny);
    const int cz = cell_of_s(qz, g.oz, g.inv_cell, g.nz);
    int lox = 0, hix = g.nx - 1, loy = 0, hiy = g.ny - 1, loz = 0, hiz = g.nz - 1;
    if (gate_r < 3.0e38f) {
        const float gr = gate_r + margin;
        lox = cell_of_s(qx - gr, g.ox, g.inv_cell, g.nx);
        hix = cell_of_s(qx + gr, g.ox, g.inv_cell, g.nx);
        loy = cell_of_s(qy - gr, g.oy, g.inv_cell, g.ny);
        hiy = cell_of_s(qy + gr, g.oy, g.inv_cell, g.ny);
        loz = cell_of_s(qz - gr, g.oz, g.inv_cell, g.nz);
        hiz = cell_of_s(qz + gr, g.oz, g.inv_cell, g.nz);
    }
    const int rneed = max(max(max(cx - lox, hix - cx), max(cy - loy, hiy - cy)), max(cz - loz, hiz - cz));
    float best_d = INFINITY;   // running best (d2, index): compared as floats first, the packed key is built at the end
    int best_i = 0x7fffffff;
    best_pos = -1;
    int prev = -1;
    for (int R = min(1, rneed); R <= rneed; ++R) {
        const int z0 = max(cz - R, loz), z1 = min(cz + R, hiz);
        const int y0 = max(cy - R, loy), y1 = min(cy + R, hiy);
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
                int xa = max(cx - R, lox), xb = min(cx + R, hix);
                if (best_pos >= 0) {
                    // the running best prunes whole rows and the x-extent of the others (strictly farther cells only,
                    // so a tie with a lower index can never be skipped)
                    const int dy = y - cy, dz = z - cz;
                    float ddy = dy > 0 ? (g.oy + (float)y * g.cell) - qy : (dy < 0 ? qy - (g.oy + (float)(y + 1) * g.cell) : 0.0f);
                    float ddz = dz > 0 ? (g.oz + (float)z * g.cell) - qz : (dz < 0 ? qz - (g.oz + (float)(z + 1) * g.cell) : 0.0f);
                    ddy = fmaxf(ddy - margin, 0.0f);
                    ddz = fmaxf(ddz - margin, 0.0f);
                    const float dyz2 = (ddy * ddy + ddz * ddz) * 0.999999f;
                    const float kd = best_d * 1.000001f;
                    if (dyz2 > kd) continue;
                    const float xr = sqrtf(kd - dyz2) * 1.000001f + margin;
                    xa = max(xa, cell_of_s(qx - xr, g.ox, g.inv_cell, g.nx));
                    xb = min(xb, cell_of_s(qx + xr, g.ox, g.inv_cell, g.nx));
                }
                const int rowbase = (z * g.ny + y) * g.nx;
                const bool fresh = max(abs(y - cy), abs(z - cz)) > prev;
                const int nseg = fresh ? 1 : 2;  // a fresh row is one range; an old row contributes its two end caps
                for (int sgi = 0; sgi < nseg; ++sgi) {
                    int a, b;
                    if (fresh) {
                        a = xa;
                        b = xb;
                    } else if (sgi == 0) {
                        a = xa;
                        b = min(cx - prev - 1, xb);
                    } else {
                        a = max(cx + prev + 1, xa);
                        b = xb;
                    }
                    if (a > b) continue;
                    const uint32_t s = cs[rowbase + a], e = cs[rowbase + b + 1];
                    cand += e - s;
                    for (uint32_t j = s; j < e; ++j) {
                        const float4 c = s_tgt[j];
                        const float d = dist2_exact(qx, qy, qz, c.x, c.y, c.z);
                        if (d <= best_d && d <= gate_f) {  // false for NaN
                            const int ci = __float_as_int(c.w);
                            if (d < best_d || ci < best_i) {
                                best_d = d;
                                best_i = ci;
                                best_pos = (int)j;
                            }
                        }
                    }
                }
            }
        prev = R;
        float bound = 3.4e38f;
        if (cx - R > lox) bound = fminf(bound, qx - (g.ox + (float)(cx - R) * g.cell));
        if (cx + R < hix) bound = fminf(bound, (g.ox + (float)(cx + R + 1) * g.cell) - qx);
        if (cy - R > loy) bound = fminf(bound, qy - (g.oy + (float)(cy - R) * g.cell));
        if (cy + R < hiy) bound = fminf(bound, (g.oy + (float)(cy + R + 1) * g.cell) - qy);
        if (cz - R > loz) bound = fminf(bound, qz - (g.oz + (float)(cz - R) * g.cell));
        if (cz + R < hiz) bound = fminf(bound, (g.oz + (float)(cz + R + 1) * g.cell) - qz);
        if (bound > 3.0e38f) break;
        const float b = bound - margin;
        if (b > 0.0f && best_pos >= 0 && best_d < b * b * 0.99999905f) break;
    }
    return best_pos >= 0 ? pack_key(best_d, best_i) : KEY_EMPTY;
}

#ifdef ICP4R_RB_VERIFY
// debug build only: exhaustive check of one answer
__device__ __noinline__ void rb_verify(const float4* s_tgt, int max_m, int mvalid, float qx, float qy, float qz, float gate_f, int pos, int how,
                                       int i, float lb, float delta) {
    float bd = INFINITY;
    int bi = 0x7fffffff, bp = -1;
    for (int j = 0; j < mvalid; ++j) {
        const float4 c = s_tgt[j];
        const float d = dist2_exact(qx, qy, qz, c.x, c.y, c.z);
        const int ci = __float_as_int(c.w);
        if (d <= gate_f && (d < bd || (d == bd && ci < bi))) {
            bd = d;
            bi = ci;
            bp = j;
        }
    }
    if (bp != pos) printf("RB_VERIFY mismatch how=%d block=%d i=%d got=%d want=%d (d2 want %g) lb=%g delta=%g\n", how, blockIdx.x, i, pos, bp, bd, lb, delta);
}
#endif

#ifdef ICP4R_RB_TIMING
// debug build only: cycles summed over warps (0 phase 1, 1 phase 2, 2 reduce incl. waiting, 3 solve (warp 0), 4 setup (warp 0),
// 5 fitness pass, 6 points searched, 7 passes)
__device__ unsigned long long rb_prof[12];
#define RB_T(var) const long long var = clock64()
#define RB_ADD(slot, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&rb_prof[slot], (unsigned long long)(v)); } while (0)
#else
#define RB_T(var)
#define RB_ADD(slot, v)
#endif

template <int NV, int NW>
__device__ __forceinline__ void block_reduce(double (&acc)[NV], double* s_red /*[NW][32]*/, double* s_tot, int tid) {
    const int lane = tid & 31, w = tid >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double x = acc[v];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
        if (lane == 0) s_red[w * 32 + v] = x;
    }
    __syncthreads();
    if (tid < NV) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NW; ++j) s += s_red[j * 32 + tid];
        s_tot[tid] = s;
    }
    __syncthreads();
}

// contribution of one correspondence; FIT: {count, sum d2}
template <int KIND, bool FIT, int NV>
__device__ __forceinline__ void add_corr(double (&acc)[NV], int& cnt, const double pw[3], const float4& c, float d2) {
    ++cnt;
    if (FIT) {
        acc[1] += (double)d2;
    } else if (KIND == ICP4R_P2P_SVD) {
        const double q[3] = {(double)c.x, (double)c.y, (double)c.z};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            acc[1 + i] += pw[i];
            acc[4 + i] += q[i];
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[7 + 3 * i + j] = fma(pw[i], q[j], acc[7 + 3 * i + j]);
        }
        acc[16] += (double)d2;
    } else {
        contrib_p2p_gn(acc, pw, c.x, c.y, c.z);  // counts in acc[28] itself
    }
}

// One pass over the source cloud of the pair (an iteration, or the fitness pass): warp w owns the contiguous slice
// [w * chunk, (w + 1) * chunk) of the cell-sorted source, so there is no block barrier between the two phases.
// Per source point: s_h1 / s_h2 = slots of the two nearest target points of its last search (0xFFFF: none),
// s_src[].w = LB (see the file header).
template <int KIND, bool FIT, int NV, int NW>
__device__ __forceinline__ void pair_pass(const PairGrid& g, const float4* __restrict__ s_tgt, float4* __restrict__ s_src,
                                          const uint32_t* __restrict__ s_cs, unsigned short* __restrict__ s_h1,
                                          unsigned short* __restrict__ s_h2, unsigned short* __restrict__ s_list, int* __restrict__ s_wcnt,
                                          const double* __restrict__ s_T, const float* __restrict__ s_dA, const BatchParams& P, int n,
                                          int chunk, bool have_prev, double (&acc)[NV], int lane, int w) {
    constexpr int NONE = 0xFFFF;
    const int beg = min(n, w * chunk), end = min(n, beg + chunk);
    unsigned short* list = s_list + beg;
    int nl = 0, cnt = 0;
    unsigned cand = 0;  // squared-distance evaluations of this thread (work counter)
    const unsigned lt = (1u << lane) - 1u;
    const float slack = P.slack * g.cell;
    RB_T(t_p1);
    // ---- phase 1: is one of the two remembered points provably still the nearest one? ------------------------------
    for (int i0 = beg; i0 < end; i0 += 32) {
        const int i = i0 + lane;
        bool need = i < end;
        if (need && have_prev) {
            const int hp1 = (int)s_h1[i], hp2 = (int)s_h2[i];
            if (hp1 != NONE) {
                const float4 p = s_src[i];
                double pw[3];
                xform_point(s_T, p.x, p.y, p.z, pw);
                const float qx = (float)pw[0], qy = (float)pw[1], qz = (float)pw[2];
                const float margin = fmaxf(g.margin, 9.5367431640625e-7f * fmaxf(fabsf(qx), fmaxf(fabsf(qy), fabsf(qz))));
                const float4 h1 = s_tgt[hp1];
                const float4 h2 = s_tgt[hp2 != NONE ? hp2 : hp1];
                const float d1 = dist2_exact(qx, qy, qz, h1.x, h1.y, h1.z);
                const float d2 = hp2 != NONE ? dist2_exact(qx, qy, qz, h2.x, h2.y, h2.z) : INFINITY;
                const bool second_wins = d2 < d1;
                const float dm = second_wins ? d2 : d1;
                // how far the last pose increment D moved this point: q - D^-1 q = (I - R^T) q + R^T t
                const float ex = __fmaf_rn(s_dA[0], qx, __fmaf_rn(s_dA[1], qy, __fmaf_rn(s_dA[2], qz, s_dA[3])));
                const float ey = __fmaf_rn(s_dA[4], qx, __fmaf_rn(s_dA[5], qy, __fmaf_rn(s_dA[6], qz, s_dA[7])));
                const float ez = __fmaf_rn(s_dA[8], qx, __fmaf_rn(s_dA[9], qy, __fmaf_rn(s_dA[10], qz, s_dA[11])));
                const float delta = sqrtf(__fmaf_rn(ex, ex, __fmaf_rn(ey, ey, ez * ez))) * 1.0001f + 2.0f * margin;
                const float lb = p.w - delta;  // every other target point is at least this far away now
                s_src[i].w = lb;
                // (a tie between the two goes to the search, which orders by index)
                if (dm <= P.gate_f && d1 != d2 && sqrtf(dm) * 1.000001f + margin < lb) {  // false for NaN
                    need = false;
                    if (second_wins) {
                        s_h1[i] = (unsigned short)hp2;
                        s_h2[i] = (unsigned short)hp1;
                    }
                    add_corr<KIND, FIT, NV>(acc, cnt, pw, second_wins ? h2 : h1, dm);
#ifdef ICP4R_RB_VERIFY
                    rb_verify(s_tgt, P.max_m, s_cs[g.ncells], qx, qy, qz, P.gate_f, second_wins ? hp2 : hp1, 1, i, lb, delta);
#endif
                }
            }
        }
        const unsigned mask = __ballot_sync(FULL, need);
        if (need) list[nl + __popc(mask & lt)] = (unsigned short)i;
        nl += __popc(mask);
    }
    __syncwarp();
    RB_T(t_p2);
    RB_ADD(0, t_p2 - t_p1);
    RB_ADD(6, nl);
    RB_ADD(7, 1);
    // ---- phase 2: search the rest --------------------------------------------------------------------------------
    // The points that need a search cluster in space, i.e. in a few warps' slices (the source is in cell order): the
    // warps' lists are pooled and the BLOCK's threads take one entry each, so a pass costs one search round (< 256 entries
    // per pair and pass on C4) instead of as many rounds as the fullest warp list needs.
    if (lane == 0) s_wcnt[w] = nl;
    __syncthreads();
    int pre[NW + 1];
    pre[0] = 0;
#pragma unroll
    for (int j = 0; j < NW; ++j) pre[j + 1] = pre[j] + s_wcnt[j];
    const int pooled = pre[NW];
    for (int j = (w << 5) + lane; j < pooled; j += NW * 32) {
        {
            int ww = 0;
#pragma unroll
            for (int t = 1; t < NW; ++t) ww += (j >= pre[t]) ? 1 : 0;
            int off_in = j;
#pragma unroll
            for (int t = 0; t < NW; ++t) off_in = (t == ww) ? j - pre[t] : off_in;
            const int i = (int)s_list[min(n, ww * chunk) + off_in];
            const float4 p = s_src[i];
            double pw[3];
            xform_point(s_T, p.x, p.y, p.z, pw);
            const float qx = (float)pw[0], qy = (float)pw[1], qz = (float)pw[2];
            const float margin = fmaxf(g.margin, 9.5367431640625e-7f * fmaxf(fabsf(qx), fmaxf(fabsf(qy), fabsf(qz))));
            // a radius that is known to hold the nearest point
            float hd = INFINITY;
            if (P.use_hints) {
                if (have_prev && (int)s_h1[i] != NONE) {
                    const int hp1 = (int)s_h1[i], hp2 = (int)s_h2[i];
                    const float4 h1 = s_tgt[hp1];
                    const float d1 = dist2_exact(qx, qy, qz, h1.x, h1.y, h1.z);
                    float d2 = d1;
                    if (hp2 != NONE) {
                        const float4 h2 = s_tgt[hp2];
                        d2 = dist2_exact(qx, qy, qz, h2.x, h2.y, h2.z);
                    }
                    // the farther of the two keeps both inside the box; beyond the gate only the nearer one is a bound
                    const float dmax = fmaxf(d1, d2), dmin = fminf(d1, d2);
                    hd = dmax <= P.gate_f ? dmax : dmin;
                } else {
                    hd = own_cell_seed(g, s_tgt, s_cs, qx, qy, qz, cand);
                }
            }
            int pos = -1;
            float dbest = 0.0f;
            const bool bounded = hd <= fminf(P.gate_f, 3.0e38f);  // false for NaN and for "nothing known" (INFINITY)
            if (bounded) {
                BoxHit hit;
                thread_box_nn(g, s_tgt, s_cs, qx, qy, qz, margin, hd, slack, hit, cand);
                if (hit.d1 <= P.gate_f) {
                    pos = hit.p1;
                    dbest = hit.d1;
                }
                s_h1[i] = pos >= 0 ? (unsigned short)pos : (unsigned short)NONE;
                s_h2[i] = (pos >= 0 && hit.p2 >= 0) ? (unsigned short)hit.p2 : (unsigned short)NONE;
                s_src[i].w = hit.lb;
            } else {
                const uint64_t key = thread_shell_nn(g, s_tgt, s_cs, qx, qy, qz, P.gate_f, P.gate_r, margin, pos, cand);
                dbest = key_d2(key);
                if (P.use_hints) {  // the next pass starts from this point (one remembered point, no bound on the others yet)
                    s_h1[i] = pos >= 0 ? (unsigned short)pos : (unsigned short)NONE;
                    s_h2[i] = (unsigned short)NONE;
                    s_src[i].w = 0.0f;
                }
            }
#ifdef ICP4R_RB_VERIFY
            rb_verify(s_tgt, P.max_m, s_cs[g.ncells], qx, qy, qz, P.gate_f, pos, bounded ? 2 : 3, i, s_src[i].w, 0.f);
#endif
            if (pos >= 0) add_corr<KIND, FIT, NV>(acc, cnt, pw, s_tgt[pos], dbest);
        }
    }
    if (FIT || KIND == ICP4R_P2P_SVD) acc[0] = (double)cnt;
    if (P.stats != nullptr) {
        const unsigned wc = __reduce_add_sync(FULL, cand);
        if (lane == 0) {
            atomicAdd(P.stats + 0, (unsigned long long)nl);
            atomicAdd(P.stats + 1, (unsigned long long)wc);
            atomicAdd(P.stats + 2, (unsigned long long)((end - beg) - nl));
        }
    }
    RB_ADD(1, clock64() - t_p2);
}

template <int KIND, int NT>
__global__ void __launch_bounds__(NT, 2) reg_batch_kernel(const __grid_constant__ BatchParams P, double* __restrict__ T_out,
                                                          icp4r_result* __restrict__ res) {
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_tgt = reinterpret_cast<float4*>(smem_raw);
    float4* s_src = s_tgt + P.max_m;                                // .w = LB (metres) once the loop runs
    uint32_t* s_cs = reinterpret_cast<uint32_t*>(s_src + P.max_n);  // [RB_MAXC + 2] target cell table
    uint32_t* s_cq = s_cs + (RB_MAXC + 2);                          // [RB_MAXC + 2] source cell cursors (spatial sort) ...
    unsigned short* s_list = reinterpret_cast<unsigned short*>(s_cq);  // ... later the per-warp lists of points to search
    const size_t cq_bytes = max((size_t)(RB_MAXC + 2) * 4, ((size_t)(P.max_n + NT) * 2 + 15) & ~(size_t)15);
    unsigned short* s_prev = reinterpret_cast<unsigned short*>(reinterpret_cast<unsigned char*>(s_cq) + cq_bytes);  // [max_n] slot of the nearest point of the last search
    unsigned short* s_prev2 = s_prev + ((P.max_n + 7) & ~7);                                                         // [max_n] slot of the runner-up
    __shared__ double s_red[NW * 32];
    __shared__ double s_tot[32];
    __shared__ double s_T[16];
    __shared__ float s_dA[12];  // last pose increment as the point displacement map q -> (I - R^T) q + R^T t
    __shared__ float s_bb[NW][6];
    __shared__ PairGrid s_g;
    __shared__ uint32_t s_wsum[NW];
    __shared__ int s_wcnt[NW];  // entries of every warp's search list (pair_pass)
    __shared__ int s_flags[4];  // done, converged, iterations, n_corr
    __shared__ double s_misc[2];  // mse_prev, last_cost

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NV = (KIND == ICP4R_P2P_SVD) ? 17 : 29;

    // Pairs differ in cost (searches depend on the geometry), so a static pair -> CTA assignment leaves the SMs of the
    // cheap pairs idle at the end of a launch: with more pairs than CTAs every CTA draws its next pair from a counter.
    __shared__ int s_pair;
    for (int pair = blockIdx.x;; pair += gridDim.x) {
        if (P.next_pair != nullptr) {
            __syncthreads();  // everybody has read s_pair of the previous round
            if (tid == 0) s_pair = atomicAdd(P.next_pair, 1);
            __syncthreads();
            pair = s_pair;
        }
        if (pair >= P.n_pairs) break;
        const int so = P.soff[pair], n = P.soff[pair + 1] - so;
        const int to = P.toff[pair], m = P.toff[pair + 1] - to;
        const float4* __restrict__ gsrc = P.src + so;
        const float4* __restrict__ gtgt = P.tgt + to;
        __syncthreads();  // previous pair fully consumed
        RB_T(t_s0);

        // ---- bounding box of the target ------------------------------------------------------------------
        float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int j = tid; j < m; j += NT) {
            const float4 p = __ldg(gtgt + j);
            if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
                mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
                mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
                mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                mn[a] = fminf(mn[a], __shfl_xor_sync(FULL, mn[a], o));
                mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL, mx[a], o));
            }
        if (lane == 0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                s_bb[w][a] = mn[a];
                s_bb[w][3 + a] = mx[a];
            }
        }
        __syncthreads();
        if (tid == 0) {
            float lo[3], hi[3];
            for (int a = 0; a < 3; ++a) {
                lo[a] = s_bb[0][a];
                hi[a] = s_bb[0][3 + a];
                for (int j = 1; j < NW; ++j) {
                    lo[a] = fminf(lo[a], s_bb[j][a]);
                    hi[a] = fmaxf(hi[a], s_bb[j][3 + a]);
                }
                if (!(lo[a] <= hi[a])) lo[a] = hi[a] = 0.f;  // empty / all non-finite
            }
            const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
            const float emax = fmaxf(ex, fmaxf(ey, ez));
            const float fl = fmaxf(emax * 1e-3f, 1e-6f);
            const float vol = fmaxf(ex, fl) * fmaxf(ey, fl) * fmaxf(ez, fl);
            float cell = cbrtf(vol / fmaxf(1.0f, (float)m / P.cell_pts));  // cell_pts points per cell by volume
            cell = fmaxf(cell, fmaxf(emax * 1e-4f, 1e-6f));
            int nx, ny, nz;
            for (;;) {
                const float fx = floorf(ex / cell) + 1.f, fy = floorf(ey / cell) + 1.f, fz = floorf(ez / cell) + 1.f;
                if (fx * fy * fz <= (float)RB_MAXC) {
                    nx = (int)fx;
                    ny = (int)fy;
                    nz = (int)fz;
                    break;
                }
                cell *= 1.1f;
            }
            float L = 1.0f;
            for (int a = 0; a < 3; ++a) L = fmaxf(L, fmaxf(fabsf(lo[a]), fabsf(hi[a])));
            s_g.ox = lo[0];
            s_g.oy = lo[1];
            s_g.oz = lo[2];
            s_g.cell = cell;
            s_g.inv_cell = 1.0f / cell;
            s_g.margin = L * 9.5367431640625e-7f;
            s_g.nx = nx;
            s_g.ny = ny;
            s_g.nz = nz;
            s_g.ncells = nx * ny * nz;
            s_flags[0] = 0;
            s_flags[1] = 0;
            s_flags[2] = 0;
            s_flags[3] = 0;
            s_misc[0] = INFINITY;
            s_misc[1] = 0.0;
        }
        if (tid < 16) s_T[tid] = P.T0[tid];
        if (tid < 12) s_dA[tid] = 0.0f;
        for (int c0 = tid; c0 < 2 * (RB_MAXC + 2); c0 += NT) s_cs[c0] = 0;  // both tables
        __syncthreads();
        const PairGrid g = s_g;

        // ---- counting sorts into cell order (shared memory): the target (this IS the search structure) and the
        //      source, by the cell its initially-transformed position falls in, so that the 32 queries of a warp
        //      are spatial neighbours: same rows, similar trip counts, broadcast shared-memory reads
        auto tgt_cell = [&](const float4& p) {
            return (cell_of_s(p.z, g.oz, g.inv_cell, g.nz) * g.ny + cell_of_s(p.y, g.oy, g.inv_cell, g.ny)) * g.nx +
                   cell_of_s(p.x, g.ox, g.inv_cell, g.nx);
        };
        auto src_cell = [&](const float4& p) {
            double pw[3];
            xform_point(s_T, p.x, p.y, p.z, pw);
            return (cell_of_s((float)pw[2], g.oz, g.inv_cell, g.nz) * g.ny + cell_of_s((float)pw[1], g.oy, g.inv_cell, g.ny)) * g.nx +
                   cell_of_s((float)pw[0], g.ox, g.inv_cell, g.nx);
        };
        for (int j = tid; j < m; j += NT) {
            const float4 p = __ldg(gtgt + j);
            if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) atomicAdd(&s_cs[tgt_cell(p) + 1], 1u);
        }
        for (int i = tid; i < n; i += NT) {  // the cell of every source point is kept for the placement below
            const int cell = src_cell(__ldg(gsrc + i));
            s_prev[i] = (unsigned short)cell;
            atomicAdd(&s_cq[cell + 1], 1u);
        }
        __syncthreads();
        for (int which = 0; which < 2; ++which) {  // table[c+1] <- exclusive prefix of the counts (the running cursor of cell c)
            uint32_t* tab = which ? s_cq : s_cs;
            constexpr int PER = (RB_MAXC + NT - 1) / NT;
            uint32_t v[PER], sum = 0;
#pragma unroll
            for (int t = 0; t < PER; ++t) {
                const int c0 = tid * PER + t;
                v[t] = c0 < RB_MAXC ? tab[c0 + 1] : 0u;
                sum += v[t];
            }
            uint32_t x = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(FULL, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) s_wsum[w] = x;
            __syncthreads();
            uint32_t base = x - sum;
            for (int j = 0; j < w; ++j) base += s_wsum[j];
#pragma unroll
            for (int t = 0; t < PER; ++t) {
                const int c0 = tid * PER + t;
                if (c0 < RB_MAXC) tab[c0 + 1] = base;
                base += v[t];
            }
            __syncthreads();
        }
        for (int j = tid; j < m; j += NT) {
            const float4 p = __ldg(gtgt + j);
            if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
                const uint32_t pos = atomicAdd(&s_cs[tgt_cell(p) + 1], 1u);
                s_tgt[pos] = make_float4(p.x, p.y, p.z, __int_as_float(j));
            }
        }
        // Which thread sums which source points decides the rounding of the fp64 sums. Default: atomic placement (order
        // inside a cell varies run to run, poses agree to ~1e-14). ICP4R_BATCH_REPRODUCIBLE=1: ascending original index
        // inside a cell — every warp owns a contiguous slice of the cloud, the warps take turns (one barrier per warp and
        // pair) to hand out positions, lanes that share a cell rank themselves with match_any — bit-reproducible.
        if (!P.reproducible) {
            for (int i = tid; i < n; i += NT) s_prev[i] = (unsigned short)atomicAdd(&s_cq[(int)s_prev[i] + 1], 1u);
        } else {
            const int slice = (n + NW - 1) / NW;
            const int beg = w * slice, end = min(n, beg + slice);
            for (int turn = 0; turn < NW; ++turn) {
                if (w == turn) {
                    for (int i0 = beg; i0 < end; i0 += 32) {
                        const int i = i0 + lane;
                        const int cell = i < end ? (int)s_prev[i] : -1 - lane;  // lanes without a point: singleton groups
                        const unsigned grp = __match_any_sync(FULL, cell);
                        if (i < end) {
                            const uint32_t base = s_cq[cell + 1];
                            __syncwarp(grp);
                            if (lane == __ffs(grp) - 1) s_cq[cell + 1] = base + (uint32_t)__popc(grp);
                            s_prev[i] = (unsigned short)(base + (uint32_t)__popc(grp & ((1u << lane) - 1u)));
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
            }
        }
        __syncthreads();
        for (int i = tid; i < n; i += NT) s_src[s_prev[i]] = __ldg(gsrc + i);
        __syncthreads();
        // now s_cs[c] = start of target cell c, s_cs[c+1] = its end; s_cq is free: it holds the search lists from here on
        const int chunk = ((n + NT - 1) / NT) * 32;  // source points per warp
        if (w == 0) RB_ADD(4, clock64() - t_s0);

        // ---- iterations ---------------------------------------------------------------------------------
        for (int it = 0; it < P.max_iterations; ++it) {
            double acc[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[v] = 0.0;
            pair_pass<KIND, false, NV, NW>(g, s_tgt, s_src, s_cs, s_prev, s_prev2, s_list, s_wcnt, s_T, s_dA, P, n, chunk, P.use_hints && it > 0, acc, lane, w);
            RB_T(t_r0);
            block_reduce<NV, NW>(acc, s_red, s_tot, tid);
            RB_T(t_r1);
            RB_ADD(2, t_r1 - t_r0);
            if (KIND == ICP4R_P2P_SVD && w == 0) {
                // Kabsch step: the other warps wait at the barrier below for this, every iteration
                const bool last = (it == P.max_iterations - 1);
                const double cnt = s_tot[0];
                double* Ds = s_red + 16;   // [12] increment, 3x4 row-major
                if (cnt < 3.0) {
                    if (lane == 0) {
                        s_flags[3] = (int)cnt;
                        s_flags[0] = 1;
                        s_flags[1] = 0;
                        s_flags[2] = it;
                    }
                } else {
                    // every lane solves the 3x3 problem redundantly in its own registers (solve_warp.cuh: thread_kabsch);
                    // this is the serial tail of the iteration: latency is all that counts
                    const double inv = 1.0 / cnt;
                    const double pm[3] = {s_tot[1] * inv, s_tot[2] * inv, s_tot[3] * inv};
                    const double qm[3] = {s_tot[4] * inv, s_tot[5] * inv, s_tot[6] * inv};
                    double Hc[9], Rk[9];
#pragma unroll
                    for (int e = 0; e < 9; ++e) Hc[e] = s_tot[7 + e] * inv - pm[e / 3] * qm[e % 3];
                    RB_T(t_k0);
                    thread_kabsch(Hc, Rk);
                    RB_ADD(5, clock64() - t_k0);
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
                    const double tn = warp_compose_entry(Ds, s_T, lane);
                    __syncwarp();
                    if (lane < 12) {
                        s_T[lane] = tn;
                        const int r = lane >> 2, cc = lane & 3;  // displacement map of this increment (see pair_pass)
                        s_dA[lane] = cc < 3 ? (float)((r == cc ? 1.0 : 0.0) - Ds[4 * cc + r])
                                            : (float)((Ds[r] * Ds[3] + Ds[4 + r] * Ds[7]) + Ds[8 + r] * Ds[11]);
                    }
                    if (lane == 0) {
                        s_flags[3] = (int)cnt;
                        const double mse = s_tot[16] / cnt;
                        s_misc[1] = mse;
                        if (P.early_exit) {
                            if (fabs(mse - s_misc[0]) < P.mse_abs_eps) {
                                s_flags[0] = 1;
                                s_flags[1] = 1;
                                s_flags[2] = it + 1;
                            }
                            s_misc[0] = mse;
                        }
                        if (!s_flags[0] && last) {
                            s_flags[0] = 1;
                            s_flags[1] = 1;
                            s_flags[2] = P.max_iterations;
                        }
                    }
                }
            }
            if (KIND != ICP4R_P2P_SVD && tid == 0) {
                double Tc[16], D[16];
                const bool last = (it == P.max_iterations - 1);
                for (int i = 0; i < 16; ++i) Tc[i] = s_T[i];
                {  // Gauss-Newton form: 6x6 Cholesky + SE(3) exponential by one thread (the headline kind is P2P_SVD, above)
                    const double cnt = s_tot[28];
                    s_flags[3] = (int)cnt;
                    double xi[6];
                    if (cnt < 6.0 || chol6_solve(s_tot, s_tot + 21, xi)) {
                        s_flags[0] = 1;
                        s_flags[1] = 0;
                        s_flags[2] = it;
                    } else {
                        se3_exp(xi, D);
                        s_misc[1] = s_tot[27];
                        mat4_mul(D, Tc, Tc);
                        for (int i = 0; i < 16; ++i) s_T[i] = Tc[i];
                        for (int r = 0; r < 3; ++r) {
                            for (int cc = 0; cc < 3; ++cc) s_dA[4 * r + cc] = (float)((r == cc ? 1.0 : 0.0) - D[4 * cc + r]);
                            s_dA[4 * r + 3] = (float)((D[r] * D[3] + D[4 + r] * D[7]) + D[8 + r] * D[11]);
                        }
                        if (P.early_exit) {
                            const double wn = sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]);
                            const double vn = sqrt(xi[3] * xi[3] + xi[4] * xi[4] + xi[5] * xi[5]);
                            if (wn < P.rot_eps && vn < P.trans_eps) {
                                s_flags[0] = 1;
                                s_flags[1] = 1;
                                s_flags[2] = it + 1;
                            }
                        }
                    }
                }
                if (!s_flags[0] && last) {
                    s_flags[0] = 1;
                    s_flags[1] = 1;
                    s_flags[2] = P.max_iterations;
                }
            }
            if (w == 0) RB_ADD(3, clock64() - t_r1);
            __syncthreads();
            if (s_flags[0]) break;
        }

        // ---- fitness pass: mean squared 1-NN distance under the final pose --------------------------
        {
            double fa[2] = {0.0, 0.0};
            pair_pass<KIND, true, 2, NW>(g, s_tgt, s_src, s_cs, s_prev, s_prev2, s_list, s_wcnt, s_T, s_dA, P, n, chunk, P.use_hints && P.max_iterations > 0, fa, lane, w);
            block_reduce<2, NW>(fa, s_red, s_tot, tid);
        }
        if (tid < 16) T_out[(size_t)pair * 16 + tid] = s_T[tid];
        if (tid == 0) {
            icp4r_result r;
            r.converged = (P.max_iterations == 0) ? 1 : s_flags[1];
            r.iterations = s_flags[2];
            r.n_corr = s_flags[3];
            r.n_fitness = (int)s_tot[0];
            r.fitness = s_tot[0] > 0.0 ? s_tot[1] / s_tot[0] : INFINITY;
            r.last_cost = s_misc[1];
            res[pair] = r;
        }
    }
}

constexpr size_t RB_SMEM_LIMIT = 200 * 1024;
static size_t batch_smem_bytes(int max_n, int max_m, int nt) {
    const size_t cq_bytes = std::max((size_t)(RB_MAXC + 2) * 4, ((size_t)(std::max(max_n, 1) + nt) * 2 + 15) & ~(size_t)15);
    return (size_t)(std::max(max_m, 1) + std::max(max_n, 1)) * sizeof(float4) + (RB_MAXC + 2) * sizeof(uint32_t) + cq_bytes +
           2 * (size_t)((std::max(max_n, 1) + 7) & ~7) * sizeof(unsigned short);
}
bool register_batch_fits(int max_n, int max_m) { return max_n < 65535 && max_m < 65535 && batch_smem_bytes(max_n, max_m, 512) <= RB_SMEM_LIMIT; }

template <int KIND, int NT>
static int launch_batch(Ctx* c, const BatchParams& P, size_t smem, double* d_T, icp4r_result* d_res) {
    auto kern = reg_batch_kernel<KIND, NT>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    per_sm = std::max(per_sm, 1);
    const int blocks = std::min(P.n_pairs, c->sm_count * per_sm);
    BatchParams Q = P;
    if (P.n_pairs > blocks) {  // dynamic pair scheduling; each launch gets a counter of its own (launches on two streams overlap)
        CKS(reserve(c, c->d_pairctr, 64 * sizeof(int)));
        Q.next_pair = c->d_pairctr.as<int>() + (c->pairctr_slot++ & 63);
        CK(cudaMemsetAsync(Q.next_pair, 0, sizeof(int), c->stream));
    }
    kern<<<blocks, NT, smem, c->stream>>>(Q, d_T, d_res);
    c->launches += 1;
    CK(cudaGetLastError());
#ifdef ICP4R_RB_TIMING
    {
        unsigned long long h[12], z[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        cudaStreamSynchronize(c->stream);
        cudaMemcpyFromSymbol(h, rb_prof, sizeof(h));
        cudaMemcpyToSymbol(rb_prof, z, sizeof(z));
        const double np = (double)P.n_pairs, nw = NT / 32;
        fprintf(stderr, "[rb timing] per pair, cycles averaged over warps: phase1 %.0f  phase2 %.0f  reduce+wait %.0f | warp 0: solve %.0f setup %.0f | searched points per pair %.0f, passes %.1f | kabsch %.0f cycles per pair\n",
                h[0] / np / nw, h[1] / np / nw, h[2] / np / nw, h[3] / np, h[4] / np, h[6] / np, h[7] / np / nw, h[5] / np);
    }
#endif
    return ICP4R_OK;
}

int register_batch(Ctx* c, const float4* d_src, const int32_t* d_soff, const float4* d_tgt, const int32_t* d_toff, int n_pairs,
                   int max_n, int max_m, const icp4r_opts* o, double* d_T, icp4r_result* d_res) {
    if (n_pairs <= 0) return ICP4R_OK;
    if (o->residual != ICP4R_P2P_SVD && o->residual != ICP4R_P2P_GN)
        return fail(c, ICP4R_ERR_UNSUPPORTED, "batched registration supports P2P_SVD and P2P_GN (got %d)", o->residual);
    // 256 threads per pair also for a single pair (icp4r_register): measured on C1 0.276 ms against 0.340 (384 threads) and
    // 0.439 (512): the wider blocks spill their fp64 accumulators
    int nt = 256;
    if (const char* e = std::getenv("ICP4R_RB_THREADS")) nt = std::atoi(e);
    if (nt != 256 && nt != 384 && nt != 512) nt = 256;
    const size_t smem = batch_smem_bytes(max_n, max_m, nt);
    if (smem > RB_SMEM_LIMIT)
        return fail(c, ICP4R_ERR_UNSUPPORTED, "pair too large for the shared-memory resident kernel (%zu B); use icp4r_register", smem);
    BatchParams P;
    std::memset(&P, 0, sizeof(P));
    P.src = d_src;
    P.soff = d_soff;
    P.tgt = d_tgt;
    P.toff = d_toff;
    P.n_pairs = n_pairs;
    P.max_n = std::max(max_n, 1);
    P.max_m = std::max(max_m, 1);
    P.max_iterations = o->max_iterations;
    P.early_exit = o->early_exit;
    P.use_hints = (c->use_hints && max_m < 65535 && max_n < 65535) ? 1 : 0;
    P.reproducible = c->batch_reproducible ? 1 : 0;
    P.cell_pts = 4.0f;  // measured on C4: 2 -> 520 k, 4-5 -> 565 k, 12 -> 547 k registrations/s (fewer rows per search beat fewer candidates)
    if (const char* e = std::getenv("ICP4R_RB_CELL_PTS")) P.cell_pts = std::max(0.05f, (float)std::atof(e));
    P.slack = 0.0f;
    if (const char* e = std::getenv("ICP4R_RB_SLACK")) P.slack = std::min(std::max(0.0f, (float)std::atof(e)), 4.0f);
    gate_params(o->max_corr_dist, &P.gate_f, &P.gate_r);
    P.rot_eps = o->rot_eps;
    P.trans_eps = o->trans_eps;
    P.mse_abs_eps = o->mse_abs_eps;
    std::memcpy(P.T0, o->T0, sizeof(P.T0));
    P.stats = c->stats ? c->d_stats.as<unsigned long long>() : nullptr;

    const bool svd = o->residual == ICP4R_P2P_SVD;
    if (nt == 512) return svd ? launch_batch<ICP4R_P2P_SVD, 512>(c, P, smem, d_T, d_res) : launch_batch<ICP4R_P2P_GN, 512>(c, P, smem, d_T, d_res);
    if (nt == 384) return svd ? launch_batch<ICP4R_P2P_SVD, 384>(c, P, smem, d_T, d_res) : launch_batch<ICP4R_P2P_GN, 384>(c, P, smem, d_T, d_res);
    return svd ? launch_batch<ICP4R_P2P_SVD, 256>(c, P, smem, d_T, d_res) : launch_batch<ICP4R_P2P_GN, 256>(c, P, smem, d_T, d_res);
}

}  // namespace icp4r
