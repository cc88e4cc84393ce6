// gicp.cu — the pieces of the fast_gicp cost (call site /root/reference/src/radar_odometry.cpp:399-405)
// that are not the fused iteration kernel:
//   * per-point surface normals from each cloud's own k nearest neighbours (fast_gicp's calculate_covariances with
//     the PLANE regularisation keeps only the eigenvector n of the smallest eigenvalue: C = I - (1 - 1e-3) n n^T);
//   * the Levenberg-Marquardt step: damped 6x6 solve, candidate pose, error pass over the stored correspondences,
//     gain ratio, accept / re-damp — one single-block kernel per outer iteration.
// [UPSTREAM, unpinned]: fast_gicp is not vendored by the reference; restated from its published algorithm.
#include <cmath>
#include <cstdlib>

#include "ctx.h"
#include "device_math.cuh"
#include "grid_knn.cuh"
#include "solve_warp.cuh"

namespace icp4r {

// eigenvector of the smallest eigenvalue of a symmetric 3x3 (cyclic Jacobi), same operation order as the oracle
__device__ __forceinline__ void smallest_eigvec3(const double C[9], double n[3]) {
    double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
    for (int i = 0; i < 9; ++i) A[i] = C[i];
    for (int sweep = 0; sweep < 50; ++sweep) {
        const double off = fabs(A[1]) + fabs(A[2]) + fabs(A[5]);
        // converged: rotations by angles below 1e-20 change nothing in double precision (was: 50 sweeps, ~40 of them no-ops
        // that still cost their divisions and square roots)
        if (off <= 1e-20 * (fabs(A[0]) + fabs(A[4]) + fabs(A[8])) || off < 1e-300) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            const double apq = A[3 * p + q];
            if (fabs(apq) < 1e-300) continue;
            const double theta = (A[3 * q + q] - A[3 * p + p]) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double akp = A[3 * k + p], akq = A[3 * k + q];
                A[3 * k + p] = c * akp - s * akq;
                A[3 * k + q] = s * akp + c * akq;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double apk = A[3 * p + k], aqk = A[3 * q + k];
                A[3 * p + k] = c * apk - s * aqk;
                A[3 * q + k] = s * apk + c * aqk;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double vkp = V[3 * k + p], vkq = V[3 * k + q];
                V[3 * k + p] = c * vkp - s * vkq;
                V[3 * k + q] = s * vkp + c * vkq;
            }
        }
    }
    int m = 0;
    if (A[4] < A[0]) m = 1;
    if (A[8] < (m == 0 ? A[0] : A[4])) m = 2;
    const double v0 = m == 0 ? V[0] : (m == 1 ? V[1] : V[2]);
    const double v1 = m == 0 ? V[3] : (m == 1 ? V[4] : V[5]);
    const double v2 = m == 0 ? V[6] : (m == 1 ? V[7] : V[8]);
    const double len = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
    n[0] = v0 / len;
    n[1] = v1 / len;
    n[2] = v2 / len;
}

// covariance of `found` neighbour positions about their mean, divided by k (fast_gicp divides by k_correspondences_), and its
// smallest eigenvector: the one arithmetic, in one order, behind every normals kernel of this file
template <int K, typename NB>
__device__ __forceinline__ void normal_of_neighbours(const NB& nb /* [K][3] floats */, int found, int k, double n[3]) {
    double mean[3] = {0, 0, 0}, C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (j < found) {
            mean[0] += (double)nb[j][0];
            mean[1] += (double)nb[j][1];
            mean[2] += (double)nb[j][2];
        }
    }
    const double fdiv = (double)(found > 0 ? found : 1);
    mean[0] /= fdiv;
    mean[1] /= fdiv;
    mean[2] /= fdiv;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (j < found) {
            const double d[3] = {(double)nb[j][0] - mean[0], (double)nb[j][1] - mean[1], (double)nb[j][2] - mean[2]};
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) C[3 * a + b] += d[a] * d[b];
        }
    }
#pragma unroll
    for (int a = 0; a < 9; ++a) C[a] /= (double)k;
    smallest_eigvec3(C, n);
}

// Self k-NN of a whole cloud, ONE THREAD PER POINT in cell order (k <= 5, the reference's setCorrespondenceRandomness(5)).
// The queries ARE the grid's points, so neighbouring threads sit in the same or adjacent cells and read the same few
// rows of the sorted array (L1 hits); each thread ranks every point of the 3 x 3 x 3 cells around its own cell by
// (d2, index) in registers. The answer is exact when the k-th distance found is smaller than the distance to the nearest
// face of that cell block that has cells behind it; the points where it is not (sparse surroundings: a few per cent) go
// to a list that the warp-per-query kernel below finishes. Same neighbours in the same order -> same normals, bit for bit,
// as the warp-per-query kernel alone (4 x fewer executed instructions per point, no idle lanes).
template <int K>
__global__ void __launch_bounds__(128) normals_brick_kernel(GridDesc g, int k, int* __restrict__ nbslot, int* __restrict__ todo,
                                                            int* __restrict__ todo_n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.m) return;
    const float4 p = __ldg(g.sorted + i);
    const int cx = cell_of(p.x, g.ox, g.inv_cell, g.nx), cy = cell_of(p.y, g.oy, g.inv_cell, g.ny), cz = cell_of(p.z, g.oz, g.inv_cell, g.nz);
    // the k best so far, ascending by (d2, original index): distances, indices and sorted slots in registers. Every
    // candidate runs the same K-step chain of selects — no branch, so the 32 lanes of a warp stay together (a guarded
    // insertion made the warp execute the chain whenever ANY lane inserted: 47 % of the executed instructions).
    float kd[K];
    int ki[K], slot[K];
#pragma unroll
    for (int t = 0; t < K; ++t) {
        kd[t] = INFINITY;
        ki[t] = 0x7fffffff;
        slot[t] = -1;
    }
    const int kk = min(k, K);
    const float margin = fmaxf(g.margin, 9.5367431640625e-7f * fmaxf(fabsf(p.x), fmaxf(fabsf(p.y), fabsf(p.z))));
    // Own cell first — where the cloud is dense it already holds k points and their k-th distance prunes most other
    // cells (a cell is skipped only when even its nearest corner is strictly farther than the k-th distance so far: nothing
    // in it could enter the list, ties included) — then the shell of cells at Chebyshev distance 1, and, if the k-th
    // distance is still not proven smaller than the distance to the block's faces, the shell at distance 2.
    constexpr int RMAX = 2;
    bool exact = false;
    int found = 0;
    for (int R = 0; R <= RMAX && !exact; ++R) {
        const int xa = max(cx - R, 0), xb = min(cx + R, g.nx - 1);
        const int y0 = max(cy - R, 0), y1 = min(cy + R, g.ny - 1);
        const int z0 = max(cz - R, 0), z1 = min(cz + R, g.nz - 1);
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
                float ddy = y > cy ? (g.oy + (float)y * g.cell) - p.y : (y < cy ? p.y - (g.oy + (float)(y + 1) * g.cell) : 0.0f);
                float ddz = z > cz ? (g.oz + (float)z * g.cell) - p.z : (z < cz ? p.z - (g.oz + (float)(z + 1) * g.cell) : 0.0f);
                ddy = fmaxf(ddy - margin, 0.0f);
                ddz = fmaxf(ddz - margin, 0.0f);
                const float dyz2 = (ddy * ddy + ddz * ddz) * 0.999999f;
                float kth = INFINITY;
#pragma unroll
                for (int t = 0; t < K; ++t)
                    if (t == kk - 1) kth = kd[t];
                if (dyz2 > kth * 1.000001f) continue;  // the whole row is too far (never while the list is not full: inf)
                const bool inner_row = max(abs(y - cy), abs(z - cz)) < R;  // only its two end cells belong to this shell
                const uint32_t rowbase = (uint32_t)(z * g.ny + y) * (uint32_t)g.nx;
                for (int x = xa; x <= xb; ++x) {
                    if (inner_row && abs(x - cx) < R) continue;  // visited by an earlier shell
#pragma unroll
                    for (int t = 0; t < K; ++t)
                        if (t == kk - 1) kth = kd[t];
                    float ddx = x > cx ? (g.ox + (float)x * g.cell) - p.x : (x < cx ? p.x - (g.ox + (float)(x + 1) * g.cell) : 0.0f);
                    ddx = fmaxf(ddx - margin, 0.0f);
                    if (dyz2 + ddx * ddx * 0.999999f > kth * 1.000001f) continue;
                    const uint32_t s = __ldg(g.cell_start + rowbase + x), e = __ldg(g.cell_start + rowbase + x + 1);
                    for (uint32_t j = s; j < e; ++j) {
                        const float4 c = __ldg(g.sorted + j);
                        float cd = dist2_exact(p.x, p.y, p.z, c.x, c.y, c.z);
                        int ci = __float_as_int(c.w), cs_ = (int)j;
#pragma unroll
                        for (int t = 0; t < K; ++t) {
                            const bool lt = cd < kd[t] || (cd == kd[t] && ci < ki[t]);
                            const float td = lt ? kd[t] : cd;
                            const int ti = lt ? ki[t] : ci;
                            const int ts = lt ? slot[t] : cs_;
                            kd[t] = lt ? cd : kd[t];
                            ki[t] = lt ? ci : ki[t];
                            slot[t] = lt ? cs_ : slot[t];
                            cd = td;
                            ci = ti;
                            cs_ = ts;
                        }
                    }
                }
            }
        if (R == 0) continue;
        // faces of the visited block with cells behind them (the grid spans the cloud's bounding box: nothing lies outside it)
        float bd = 3.4e38f;
        if (xa > 0) bd = fminf(bd, p.x - (g.ox + (float)xa * g.cell));
        if (xb < g.nx - 1) bd = fminf(bd, (g.ox + (float)(xb + 1) * g.cell) - p.x);
        if (y0 > 0) bd = fminf(bd, p.y - (g.oy + (float)y0 * g.cell));
        if (y1 < g.ny - 1) bd = fminf(bd, (g.oy + (float)(y1 + 1) * g.cell) - p.y);
        if (z0 > 0) bd = fminf(bd, p.z - (g.oz + (float)z0 * g.cell));
        if (z1 < g.nz - 1) bd = fminf(bd, (g.oz + (float)(z1 + 1) * g.cell) - p.z);
        found = 0;
        float dk = 0.0f;
#pragma unroll
        for (int t = 0; t < K; ++t) {
            found += (t < kk && slot[t] >= 0) ? 1 : 0;
            if (t == kk - 1) dk = kd[t];
        }
        exact = bd > 3.0e38f;  // the block is the whole grid
        if (!exact && found == kk) {
            const float b = bd - 2.0f * margin;
            exact = b > 0.0f && dk < b * b * 0.999999f;  // every unvisited point is strictly farther than the k-th found
        }
    }
    // the neighbours' sorted slots (or "not proven": the warp-per-query kernel does that point) go to global memory; the
    // covariances and eigenvectors are a second, register-hungry kernel of their own (normals_from_slots_kernel)
#pragma unroll
    for (int t = 0; t < K; ++t) nbslot[(size_t)i * K + t] = (exact && t < found) ? slot[t] : -1;
    if (!exact) todo[atomicAdd(todo_n, 1)] = i;
}

// second half of the brick path: neighbour slots -> covariance -> smallest eigenvector, one thread per point; points the
// brick pass could not prove (first slot -1) are left to the warp-per-query kernel
template <int K>
__global__ void __launch_bounds__(128) normals_from_slots_kernel(GridDesc g, int k, const int* __restrict__ nbslot, double* __restrict__ normals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.m) return;
    int sl[K];
#pragma unroll
    for (int t = 0; t < K; ++t) sl[t] = __ldg(nbslot + (size_t)i * K + t);
    if (sl[0] < 0) return;
    float nb[K][3];
    int found = 0;
#pragma unroll
    for (int t = 0; t < K; ++t) {
        if (sl[t] >= 0) {
            const float4 c = __ldg(g.sorted + sl[t]);
            nb[t][0] = c.x;
            nb[t][1] = c.y;
            nb[t][2] = c.z;
            ++found;
        } else {
            nb[t][0] = nb[t][1] = nb[t][2] = 0.f;
        }
    }
    double n[3];
    normal_of_neighbours<K>(nb, found, k, n);
    const size_t o = 3 * (size_t)__float_as_int(__ldg(g.sorted + i).w);
    normals[o] = n[0];
    normals[o + 1] = n[1];
    normals[o + 2] = n[2];
}

// k-NN among the grid's own cloud (the point itself included), covariance about the neighbours' mean divided by k,
// smallest eigenvector -> normals[3 * original_index]. A warp searches the neighbours of PARK consecutive points one
// after the other (all lanes cooperate on one query) and parks their coordinates in shared memory; the covariance
// and the Jacobi eigen-solver (a few thousand fp64 instructions) then run with ONE POINT PER LANE instead of
// redundantly on all 32 lanes for every single point.
template <int K>
__global__ void __launch_bounds__(256) normals_kernel(GridDesc g, const float4* __restrict__ pts, int k, int chunk, double* __restrict__ normals,
                                                      const int* __restrict__ list, const int* __restrict__ list_n) {
    constexpr int PARK = K <= 5 ? 32 : (K <= 8 ? 16 : 8);  // upper bound of `chunk` (points a warp parks per round)
    __shared__ WarpSegs segs[8];
    __shared__ float nbuf[8][PARK][K][3];
    __shared__ int nfound[8][PARK], nidx[8][PARK];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int total = list ? *list_n : g.m;  // `list`: only the sorted positions named there (what normals_brick_kernel left over)
    for (int base = warp * chunk; base < total; base += nwarps * chunk) {
        const int cnt = min(chunk, total - base);
        for (int j = 0; j < cnt; ++j) {
            const float4 p = __ldg(g.sorted + (list ? list[base + j] : base + j));
            const uint64_t mine = warp_grid_knn<K>(g, seg_addr(&segs[w]), p.x, p.y, p.z, INFINITY, INFINITY, lane);
            const bool have = (lane < k) && (mine != KEY_EMPTY);
            const int found = __popc(__ballot_sync(FULL, have));
            if (have) {
                const float4 nb = __ldg(pts + key_idx(mine));
                nbuf[w][j][lane][0] = nb.x;
                nbuf[w][j][lane][1] = nb.y;
                nbuf[w][j][lane][2] = nb.z;
            }
            if (lane == 0) {
                nfound[w][j] = found;
                nidx[w][j] = __float_as_int(p.w);
            }
        }
        __syncwarp();
        if (lane < cnt) {
            const int found = nfound[w][lane];
            double n[3];
            normal_of_neighbours<K>(nbuf[w][lane], found, k, n);
            const size_t o = 3 * (size_t)nidx[w][lane];
            normals[o] = n[0];
            normals[o + 1] = n[1];
            normals[o + 2] = n[2];
        }
        __syncwarp();
    }
}

// ---- small clouds (a scan): exhaustive k-NN instead of building a grid that is searched once ------------------------------
__global__ void __launch_bounds__(256) stamp_index_kernel(const float4* __restrict__ pts, int n, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    out[i] = make_float4(p.x, p.y, p.z, __int_as_float(i));
}

// same arithmetic, in the same order, as the parked part of normals_kernel: one thread per point
template <int K>
__global__ void __launch_bounds__(128) normals_from_idx_kernel(const float4* __restrict__ pts, const int32_t* __restrict__ idx,
                                                               const int32_t* __restrict__ found_n, int n, int k, double* __restrict__ normals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int found = found_n[i];
    float nb[K][3];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (j < found) {
            const float4 q = __ldg(pts + idx[(size_t)i * k + j]);
            nb[j][0] = q.x;
            nb[j][1] = q.y;
            nb[j][2] = q.z;
        } else {
            nb[j][0] = nb[j][1] = nb[j][2] = 0.f;
        }
    }
    double mean[3] = {0, 0, 0}, C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (j < found) {
            mean[0] += (double)nb[j][0];
            mean[1] += (double)nb[j][1];
            mean[2] += (double)nb[j][2];
        }
    }
    const double fdiv = (double)(found > 0 ? found : 1);
    mean[0] /= fdiv;
    mean[1] /= fdiv;
    mean[2] /= fdiv;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (j < found) {
            const double d[3] = {(double)nb[j][0] - mean[0], (double)nb[j][1] - mean[1], (double)nb[j][2] - mean[2]};
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) C[3 * a + b] += d[a] * d[b];
        }
    }
#pragma unroll
    for (int a = 0; a < 9; ++a) C[a] /= (double)k;
    double nrm[3];
    smallest_eigvec3(C, nrm);
    normals[3 * (size_t)i] = nrm[0];
    normals[3 * (size_t)i + 1] = nrm[1];
    normals[3 * (size_t)i + 2] = nrm[2];
}

// normals of the n points at d_pts from their own k nearest neighbours, written to `out` (3 doubles per point, device)
int gicp_normals_small_to(Ctx* c, const float4* d_pts, int n, int k, double* out) {
    if (n <= 0) return ICP4R_OK;
    CKS(reserve_grow(c, c->gs_pts, (size_t)n * sizeof(float4)));
    CKS(reserve_grow(c, c->gs_idx, (size_t)n * k * sizeof(int32_t)));
    CKS(reserve_grow(c, c->gs_d2, (size_t)n * k * sizeof(float)));
    CKS(reserve_grow(c, c->gs_found, (size_t)n * sizeof(int32_t)));
    stamp_index_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_pts, n, c->gs_pts.as<float4>());
    c->launches += 1;
    CKS(brute_knn_cloud(c, c->gs_pts.as<float4>(), n, d_pts, n, k, 0.0, c->gs_idx.as<int32_t>(), c->gs_d2.as<float>(), c->gs_found.as<int32_t>()));
    const int blocks = (n + 127) / 128;
    if (k <= 5) normals_from_idx_kernel<5><<<blocks, 128, 0, c->stream>>>(d_pts, c->gs_idx.as<int32_t>(), c->gs_found.as<int32_t>(), n, k, out);
    else if (k <= 8) normals_from_idx_kernel<8><<<blocks, 128, 0, c->stream>>>(d_pts, c->gs_idx.as<int32_t>(), c->gs_found.as<int32_t>(), n, k, out);
    else normals_from_idx_kernel<16><<<blocks, 128, 0, c->stream>>>(d_pts, c->gs_idx.as<int32_t>(), c->gs_found.as<int32_t>(), n, k, out);
    c->launches += 1;
    CK(cudaGetLastError());
    return ICP4R_OK;
}

int gicp_normals_small(Ctx* c, const float4* d_pts, int n, int k, DevBuf& normals) {
    CKS(reserve_grow(c, normals, (size_t)std::max(n, 1) * 3 * sizeof(double)));
    return gicp_normals_small_to(c, d_pts, n, k, normals.as<double>());
}

int gicp_normals(Ctx* c, Map& mp, int k) {
    if (mp.normals_k == k) return ICP4R_OK;
    CKS(reserve_grow(c, mp.normals, (size_t)std::max(mp.m, 1) * 3 * sizeof(double)));
    if (mp.grid.m > 0) {
        // points per warp and round: as many as fit the parking buffers, but a small cloud is spread over all warps
        const int park = k <= 5 ? 32 : (k <= 8 ? 16 : 8);
        const int warps_avail = c->sm_count * 8 * 8;
        const int chunk = std::min(park, std::max(1, (mp.grid.m + warps_avail - 1) / warps_avail));
        const int blocks = std::min((mp.grid.m + 8 * chunk - 1) / (8 * chunk), c->sm_count * 8);
        double* out = mp.normals.as<double>();
        const char* nb_env = std::getenv("ICP4R_NO_BRICK_NORMALS");
        if (k <= 5 && mp.grid.m >= 4096 && !(nb_env && nb_env[0] == '1')) {
            // thread-per-point pass over the 3 x 3 x 3 cell blocks, then the warp-per-query kernel on what it could not prove
            CKS(reserve_grow(c, c->gs_idx, ((size_t)mp.grid.m * 6 + 4) * sizeof(int32_t)));
            int* todo_n = c->gs_idx.as<int>();
            int* todo = todo_n + 4;
            int* nbslot = todo + mp.grid.m;
            CK(cudaMemsetAsync(todo_n, 0, sizeof(int), c->stream));
            normals_brick_kernel<5><<<(mp.grid.m + 127) / 128, 128, 0, c->stream>>>(mp.grid, k, nbslot, todo, todo_n);
            normals_from_slots_kernel<5><<<(mp.grid.m + 127) / 128, 128, 0, c->stream>>>(mp.grid, k, nbslot, out);
            c->launches += 1;
            const int fb_chunk = 4;
            const int fb_blocks = std::max(1, std::min((mp.grid.m / 8 + 8 * fb_chunk - 1) / (8 * fb_chunk), c->sm_count * 8));
            normals_kernel<5><<<fb_blocks, 256, 0, c->stream>>>(mp.grid, mp.pts.as<float4>(), k, fb_chunk, out, todo, todo_n);
            c->launches += 2;
        } else {
            if (k <= 5) normals_kernel<5><<<blocks, 256, 0, c->stream>>>(mp.grid, mp.pts.as<float4>(), k, chunk, out, nullptr, nullptr);
            else if (k <= 8) normals_kernel<8><<<blocks, 256, 0, c->stream>>>(mp.grid, mp.pts.as<float4>(), k, chunk, out, nullptr, nullptr);
            else normals_kernel<16><<<blocks, 256, 0, c->stream>>>(mp.grid, mp.pts.as<float4>(), k, chunk, out, nullptr, nullptr);
            c->launches += 1;
        }
        CK(cudaGetLastError());
    }
    mp.normals_k = k;
    return ICP4R_OK;
}

// ------------------------------------------------------------------------------------------------ LM step
constexpr int LM_THREADS = 1024;

__device__ __forceinline__ bool delta_converged(const double* D /*3x4*/, double rot_eps, double trans_eps) {
    double m = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) m = fmax(m, fabs(D[4 * i + j] - (i == j ? 1.0 : 0.0)) / rot_eps);
        m = fmax(m, fabs(D[4 * i + 3]) / trans_eps);
    }
    return m < 1.0;
}

// fast_gicp LsqRegistration::step_lm for one outer iteration: st->acc holds H (21), g (6), y0, count of the
// linearisation at st->T; P.corr holds each source point's correspondence and L^-1.
__global__ void __launch_bounds__(LM_THREADS) gicp_lm_kernel(const RegParams* __restrict__ prm, RegState* __restrict__ st, int iter) {
    prm += blockIdx.x;  // one block per scan of a batched call
    st += blockIdx.x;
    if (st->done) return;
    __shared__ double Hs[ICP4R_ACC_LEN], Hl[ICP4R_ACC_LEN], Ts[16], Xi[16], Ds[16], xi6[8], red[32];
    __shared__ double s_lambda, s_nu;
    __shared__ int s_state;  // 0: keep trying, 1: step taken (x0 updated or converged without update), 2: failed
    __shared__ RegParams P;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid < ICP4R_ACC_LEN) Hs[tid] = st->acc[tid];
    if (tid < 16) Ts[tid] = st->T[tid];
    if (tid >= 32 && tid < 32 + (int)(sizeof(RegParams) / 4)) reinterpret_cast<uint32_t*>(&P)[tid - 32] = reinterpret_cast<const uint32_t*>(prm)[tid - 32];
    __syncthreads();
    const double cnt = Hs[28], y0 = Hs[27];
    const bool last = (iter == P.max_iterations - 1);
    if (tid == 0) {
        st->n_corr = (int)cnt;
        st->last_cost = y0;
        double lam = st->lm_lambda;
        if (lam < 0.0) {
            double mx = 0.0;
            for (int i = 0; i < 6; ++i) mx = fmax(mx, fabs(Hs[i * 6 - i * (i - 1) / 2]));
            lam = 1e-9 * mx;  // lm_init_lambda_factor_
        }
        s_lambda = lam;
        s_nu = 2.0;
        s_state = cnt < 6.0 ? 2 : 0;
    }
    __syncthreads();
    for (int t = 0; t < 10 && s_state == 0; ++t) {  // lm_max_iterations_
        if (w == 0) {
            if (lane < 21) Hl[lane] = Hs[lane];
            __syncwarp();
            if (lane < 6) Hl[lane * 6 - lane * (lane - 1) / 2] += s_lambda;
            __syncwarp();
            double x = 0.0;
            const bool ok = warp_chol6_solve(Hl, Hs + 21, lane, x);
            if (lane < 6) xi6[lane] = x;
            __syncwarp();
            const double de = warp_se3_exp_entry(xi6, lane);
            if (lane < 12) Ds[lane] = de;
            __syncwarp();
            const double tn = warp_compose_entry(Ds, Ts, lane);
            if (lane < 12) Xi[lane] = tn;
            if (lane == 0 && !ok) s_state = 2;
        }
        __syncthreads();
        if (s_state != 0) break;
        // y_i = sum |L^-1 (b - xi a)|^2 over the stored correspondences
        double part = 0.0;
        for (int i = tid; i < P.n; i += LM_THREADS) {
            const GicpCorr cr = P.corr[i];
            if (cr.idx < 0) continue;
            const float4 a = __ldg(P.src + i);
            double pa[3];
            xform_point(Xi, a.x, a.y, a.z, pa);
            const float4 b = __ldg(P.tgt_pts + cr.idx);
            const double e0 = (double)b.x - pa[0], e1 = (double)b.y - pa[1], e2 = (double)b.z - pa[2];
            const double f0 = cr.li[0] * e0, f1 = cr.li[1] * e0 + cr.li[2] * e1, f2 = (cr.li[3] * e0 + cr.li[4] * e1) + cr.li[5] * e2;
            part += (f0 * f0 + f1 * f1) + f2 * f2;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
        if (lane == 0) red[w] = part;
        __syncthreads();
        if (tid == 0) {
            double yi = 0.0;
            for (int k = 0; k < LM_THREADS / 32; ++k) yi += red[k];
            double den = 0.0;
            for (int i = 0; i < 6; ++i) den += xi6[i] * (s_lambda * xi6[i] - Hs[21 + i]);
            const double rho = (y0 - yi) / den;
            if (rho < 0) {
                if (delta_converged(Ds, P.rot_eps, P.trans_eps)) {
                    s_state = 1;  // fast_gicp returns true here without moving x0
                } else {
                    s_lambda = s_nu * s_lambda;
                    s_nu = 2.0 * s_nu;
                }
            } else {
                for (int i = 0; i < 12; ++i) st->T[i] = Xi[i];
                const double f = 2.0 * rho - 1.0;
                s_lambda = s_lambda * fmax(1.0 / 3.0, 1.0 - f * f * f);
                s_state = 1;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        st->lm_lambda = s_lambda;
        if (s_state != 1) {  // "lm not converged!!" or too few correspondences
            st->done = 1;
            st->converged = 0;
            st->iterations = iter;
        } else {
            const bool conv = delta_converged(Ds, P.rot_eps, P.trans_eps);
            st->lm_last_conv = conv ? 1 : 0;
            if (P.early_exit && conv) {
                st->done = 1;
                st->converged = 1;
                st->iterations = iter + 1;
            } else if (last) {
                st->done = 1;
                st->converged = P.early_exit ? 0 : (conv ? 1 : 0);
                st->iterations = P.max_iterations;
            }
        }
    }
}

int gicp_lm_step(Ctx* c, const RegParams* d_prm, RegState* d_st, int iter, int nscan) {
    gicp_lm_kernel<<<nscan, LM_THREADS, 0, c->stream>>>(d_prm, d_st, iter);
    c->launches += 1;
    return ICP4R_OK;
}

}  // namespace icp4r
