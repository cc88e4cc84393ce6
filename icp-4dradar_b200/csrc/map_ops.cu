// map_ops.cu — map maintenance next to the kNN path:
//   * Add_Points with voxel down-sampling (/root/reference/third_party/ikd-Tree/ikd_Tree.cpp:422-457), exact
//     sequential semantics evaluated voxel-parallel, with an exact block-sequential fallback;
//   * Sector_Search (/root/reference/third_party/ikd-Tree/ikd_Tree.cpp:415-419,1098-1140,1434-1448) as one
//     filter + compaction pass over the map (the reference walks the whole tree, pruning commented out).
#include <cmath>
#include <cstring>

#include "ctx.h"
#include "device_math.cuh"
#include "grid_knn.cuh"

namespace icp4r {

// ---- voxel box of a point, float arithmetic exactly as ikd_Tree.cpp:432-440 ---------------------------------
struct VoxBox {
    float mn[3], mx[3], mid[3];
    int t[3];  // floor(p / ds) per axis (the voxel's identity)
};

__device__ __forceinline__ int vox_floor(float v, float ds) {
    float f = floorf(__fdiv_rn(v, ds));
    f = fminf(fmaxf(f, -2.0e9f), 2.0e9f);
    return (int)f;
}

__device__ __forceinline__ VoxBox vox_box(float x, float y, float z, float ds) {
    VoxBox b;
    const float p[3] = {x, y, z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float fl = floorf(__fdiv_rn(p[a], ds));
        b.t[a] = (int)fminf(fmaxf(fl, -2.0e9f), 2.0e9f);
        b.mn[a] = __fmul_rn(fl, ds);
        b.mx[a] = __fadd_rn(b.mn[a], ds);
        // mid = min + (max - min) / 2.0  evaluated in double, stored as float (ikd_Tree.cpp:438-440)
        b.mid[a] = (float)__dadd_rn((double)b.mn[a], __ddiv_rn((double)__fsub_rn(b.mx[a], b.mn[a]), 2.0));
    }
    return b;
}

__device__ __forceinline__ bool in_box(const VoxBox& b, float x, float y, float z) {  // ikd_Tree.cpp:1034
    return b.mn[0] <= x && b.mx[0] > x && b.mn[1] <= y && b.mx[1] > y && b.mn[2] <= z && b.mx[2] > z;
}

__device__ __forceinline__ bool same_point(float ax, float ay, float az, float bx, float by, float bz) {  // ikd_Tree.cpp:1422
    return fabs((double)__fsub_rn(ax, bx)) < 1e-6 && fabs((double)__fsub_rn(ay, by)) < 1e-6 && fabs((double)__fsub_rn(az, bz)) < 1e-6;
}

struct DsResult {   // per voxel leader
    int winner;     // global index of the survivor (-1: none / nothing changed)
    int collapsed;  // 1: every other point of the box dies
};

// One thread per new point; the first point of each voxel (the leader) replays the reference's sequential rule for
// all new points of that voxel in order. Existing points come from the map grid cells that overlap the box.
// hazard[0] is raised when a point's voxel identity and its box membership disagree (float rounding at a voxel
// face): voxels then interact and the caller switches to the exact block-sequential kernel.
__global__ void __launch_bounds__(128) ds_voxel_kernel(GridDesc g, const float4* __restrict__ pts /*all, incl. the n new at [m, m+n)*/,
                                                       int m, int n, float ds, DsResult* __restrict__ res, int* __restrict__ counter,
                                                       int* __restrict__ hazard) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[m + i];
    const VoxBox b = vox_box(p.x, p.y, p.z, ds);
    res[i].winner = -1;
    res[i].collapsed = 0;
    if (!in_box(b, p.x, p.y, p.z)) atomicOr(hazard, 1);
    {   // would the point also fall inside a neighbouring voxel's box (faces that do not meet exactly in float)?
        const float pc[3] = {p.x, p.y, p.z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float fl = floorf(__fdiv_rn(pc[a], ds));
            const float prev_mx = __fadd_rn(__fmul_rn(fl - 1.0f, ds), ds), next_mn = __fmul_rn(fl + 1.0f, ds);
            if (pc[a] < prev_mx || pc[a] >= next_mn) atomicOr(hazard, 1);
        }
    }
    // leader = no earlier new point with the same voxel
    for (int j = 0; j < i; ++j) {
        const float4 q = pts[m + j];
        if (vox_floor(q.x, ds) == b.t[0] && vox_floor(q.y, ds) == b.t[1] && vox_floor(q.z, ds) == b.t[2]) return;
    }
    // existing valid points in the box
    int cnt = 0;
    float best_d = INFINITY;
    int best = -1;
    if (g.m > 0) {
        const int x0 = cell_of(b.mn[0], g.ox, g.inv_cell, g.nx), x1 = cell_of(b.mx[0], g.ox, g.inv_cell, g.nx);
        const int y0 = cell_of(b.mn[1], g.oy, g.inv_cell, g.ny), y1 = cell_of(b.mx[1], g.oy, g.inv_cell, g.ny);
        const int z0 = cell_of(b.mn[2], g.oz, g.inv_cell, g.nz), z1 = cell_of(b.mx[2], g.oz, g.inv_cell, g.nz);
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
                const uint32_t rb = (uint32_t)(z * g.ny + y) * (uint32_t)g.nx;
                const uint32_t s = g.cell_start[rb + x0], e = g.cell_start[rb + x1 + 1];
                for (uint32_t t = s; t < e; ++t) {
                    const float4 c = g.sorted[t];
                    const bool inside = in_box(b, c.x, c.y, c.z);
                    const bool same_vox = vox_floor(c.x, ds) == b.t[0] && vox_floor(c.y, ds) == b.t[1] && vox_floor(c.z, ds) == b.t[2];
                    if (inside != same_vox) atomicOr(hazard, 1);
                    if (!inside) continue;
                    ++cnt;
                    const float d = dist2_exact(c.x, c.y, c.z, b.mid[0], b.mid[1], b.mid[2]);
                    const int id = __float_as_int(c.w);
                    if (d < best_d || (d == best_d && id < best)) {
                        best_d = d;
                        best = id;
                    }
                }
            }
    }
    // replay the new points of this voxel in order (ikd_Tree.cpp:443-457)
    int events = 0, collapsed = 0;
    for (int j = i; j < n; ++j) {
        const float4 q = pts[m + j];
        if (j > i && !(vox_floor(q.x, ds) == b.t[0] && vox_floor(q.y, ds) == b.t[1] && vox_floor(q.z, ds) == b.t[2])) continue;
        const float dn = dist2_exact(q.x, q.y, q.z, b.mid[0], b.mid[1], b.mid[2]);
        const bool existing_wins = best >= 0 && best_d < dn;  // strict: the new point wins ties (:447)
        bool same = !existing_wins;
        if (existing_wins) {
            const float4 w = pts[best];
            same = same_point(q.x, q.y, q.z, w.x, w.y, w.z);
        }
        if (cnt > 1 || same) {
            if (!existing_wins) {
                best = m + j;
                best_d = dn;
            }
            cnt = 1;
            collapsed = 1;
            ++events;
        }
    }
    res[i].winner = collapsed ? best : -1;
    res[i].collapsed = collapsed;
    if (events) atomicAdd(counter, events);
}

// apply: each leader whose voxel collapsed kills every other existing point of its box; new points are valid only
// when they are their voxel's final survivor
__global__ void __launch_bounds__(128) ds_apply_kernel(GridDesc g, const float4* __restrict__ pts, uint8_t* __restrict__ valid, int m, int n,
                                                       float ds, const DsResult* __restrict__ res) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    valid[m + i] = 0;
    const DsResult r = res[i];
    if (!r.collapsed) return;
    const float4 p = pts[m + i];
    const VoxBox b = vox_box(p.x, p.y, p.z, ds);
    if (g.m > 0) {
        const int x0 = cell_of(b.mn[0], g.ox, g.inv_cell, g.nx), x1 = cell_of(b.mx[0], g.ox, g.inv_cell, g.nx);
        const int y0 = cell_of(b.mn[1], g.oy, g.inv_cell, g.ny), y1 = cell_of(b.mx[1], g.oy, g.inv_cell, g.ny);
        const int z0 = cell_of(b.mn[2], g.oz, g.inv_cell, g.nz), z1 = cell_of(b.mx[2], g.oz, g.inv_cell, g.nz);
        for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
                const uint32_t rb = (uint32_t)(z * g.ny + y) * (uint32_t)g.nx;
                const uint32_t s = g.cell_start[rb + x0], e = g.cell_start[rb + x1 + 1];
                for (uint32_t t = s; t < e; ++t) {
                    const float4 c = g.sorted[t];
                    if (in_box(b, c.x, c.y, c.z)) valid[__float_as_int(c.w)] = (__float_as_int(c.w) == r.winner) ? 1 : 0;
                }
            }
    }
}
__global__ void ds_apply_new_kernel(uint8_t* __restrict__ valid, int n, const DsResult* __restrict__ res) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && res[i].collapsed) valid[res[i].winner] = 1;
}

// Exact fallback: ONE block replays the batch point by point over the whole (unindexed) point array, exactly like
// the oracle. O((m + n) * n / 1024) — only used when a rounding hazard was detected or on request.
__global__ void __launch_bounds__(1024) ds_sequential_kernel(const float4* __restrict__ pts, uint8_t* __restrict__ valid, int m, int n, float ds,
                                                             int* __restrict__ counter) {
    __shared__ int s_cnt[32];
    __shared__ unsigned long long s_best[32];
    __shared__ int s_decision[3];  // collapse?, winner, total count
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int events = 0;
    for (int i = 0; i < n; ++i) {
        const float4 p = pts[m + i];
        const VoxBox b = vox_box(p.x, p.y, p.z, ds);
        int cnt = 0;
        unsigned long long best = ~0ull;
        for (int j = tid; j < m + i; j += 1024) {
            if (!valid[j]) continue;
            const float4 c = pts[j];
            if (!in_box(b, c.x, c.y, c.z)) continue;
            ++cnt;
            const unsigned long long key = pack_key(dist2_exact(c.x, c.y, c.z, b.mid[0], b.mid[1], b.mid[2]), j);
            best = key < best ? key : best;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cnt += __shfl_xor_sync(FULL, cnt, o);
            const unsigned long long ob = __shfl_xor_sync(FULL, best, o);
            best = ob < best ? ob : best;
        }
        if (lane == 0) {
            s_cnt[w] = cnt;
            s_best[w] = best;
        }
        __syncthreads();
        if (tid == 0) {
            int tc = 0;
            unsigned long long tb = ~0ull;
            for (int k = 0; k < 32; ++k) {
                tc += s_cnt[k];
                tb = s_best[k] < tb ? s_best[k] : tb;
            }
            const float dn = dist2_exact(p.x, p.y, p.z, b.mid[0], b.mid[1], b.mid[2]);
            const bool existing_wins = tb != ~0ull && key_d2(tb) < dn;
            bool same = !existing_wins;
            if (existing_wins) {
                const float4 wp = pts[key_idx(tb)];
                same = same_point(p.x, p.y, p.z, wp.x, wp.y, wp.z);
            }
            s_decision[0] = (tc > 1 || same) ? 1 : 0;
            s_decision[1] = existing_wins ? key_idx(tb) : (m + i);
            s_decision[2] = tc;
        }
        __syncthreads();
        const int collapse = s_decision[0], winner = s_decision[1];
        if (collapse) {
            if (s_decision[2] > 0)
                for (int j = tid; j < m + i; j += 1024) {
                    if (!valid[j]) continue;
                    const float4 c = pts[j];
                    if (in_box(b, c.x, c.y, c.z)) valid[j] = 0;
                }
            __syncthreads();
            if (tid == 0) {
                valid[m + i] = 0;
                valid[winner] = 1;
            }
            ++events;
        } else if (tid == 0) {
            valid[m + i] = 0;
        }
        __syncthreads();
    }
    if (tid == 0) *counter = events;
}

// mp.m existing points (grid built over them), the n new points already copied to pts[m, m+n). Updates valid[].
int map_downsample_add(Ctx* c, Map& mp, int n, int* n_replaced_host, bool force_sequential) {
    const int m = mp.m;
    CKS(reserve(c, c->d_scratch, 64));
    int* d_ctr = c->d_scratch.as<int>();
    CK(cudaMemsetAsync(d_ctr, 0, 2 * sizeof(int), c->stream));
    CKS(reserve(c, c->d_partials, std::max((size_t)n * sizeof(DsResult), (size_t)c->sm_count * 4 * ICP4R_ACC_LEN * sizeof(double) + 1024)));
    DsResult* d_res = c->d_partials.as<DsResult>();
    int h[2] = {0, 0};
    bool sequential = force_sequential;
    if (!sequential) {
        ds_voxel_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(mp.grid, mp.pts.as<float4>(), m, n, mp.ds_voxel, d_res, d_ctr, d_ctr + 1);
        c->launches += 1;
        CK(cudaMemcpyAsync(h, d_ctr, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        sequential = h[1] != 0;
    }
    if (sequential) {
        CK(cudaMemsetAsync(d_ctr, 0, 2 * sizeof(int), c->stream));
        ds_sequential_kernel<<<1, 1024, 0, c->stream>>>(mp.pts.as<float4>(), mp.valid.as<uint8_t>(), m, n, mp.ds_voxel, d_ctr);
        c->launches += 1;
        CK(cudaMemcpyAsync(h, d_ctr, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    } else {
        // ds_apply_kernel zeroes valid[m + i] and kills the losers; then the surviving new points are switched on
        ds_apply_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(mp.grid, mp.pts.as<float4>(), mp.valid.as<uint8_t>(), m, n, mp.ds_voxel, d_res);
        ds_apply_new_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(mp.valid.as<uint8_t>(), n, d_res);
        c->launches += 2;
    }
    CK(cudaGetLastError());
    if (n_replaced_host) *n_replaced_host = h[0];
    return ICP4R_OK;
}

// ---- Sector_Search --------------------------------------------------------------------------------------------
// calc_heading (ikd_Tree.cpp:1434-1448) in the reference's float/double mix; asin is evaluated in double and
// rounded to float (the reference calls the float overload), which differs only within an ulp of the thresholds.
__device__ __forceinline__ float heading_of(float ax, float ay, float az, float bx, float by, float bz) {
    const float s = sqrtf(dist2_exact(ax, ay, az, bx, by, bz));
    const float as = (float)asin((double)__fdiv_rn(__fsub_rn(ax, bx), s));
    float h;
    if (__fsub_rn(ay, by) < 0.f) h = (float)(180.0 + (double)__fmul_rn(as, 180.f) / M_PI);
    else h = (float)((double)__fmul_rn(-as, 180.f) / M_PI);
    if (h > 180.f && h < 360.f) h = h - 360.f;
    return h;
}

__global__ void __launch_bounds__(256) sector_kernel(const float4* __restrict__ sorted, int m, float cx, float cy, float cz, float radius,
                                                     float heading, int32_t* __restrict__ out, int cap, int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    int id = -1;
    if (i < m) {
        const float4 p = sorted[i];
        id = __float_as_int(p.w);
        const float dh = fabsf(__fsub_rn(heading_of(p.x, p.y, p.z, cx, cy, cz), heading));
        // ikd_Tree.cpp:1114-1116: (alive && in radius && dh < 60) || dh > 300 ; deleted points are never returned here
        hit = (dist2_exact(p.x, p.y, p.z, cx, cy, cz) <= __fmul_rn(radius, radius) && dh < 60.f) || dh > 300.f;
    }
    const unsigned b = __ballot_sync(FULL, hit);
    if (b == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(b) - 1) base = atomicAdd(count, __popc(b));
    base = __shfl_sync(FULL, base, __ffs(b) - 1);
    if (hit) {
        const int slot = base + __popc(b & ((1u << lane) - 1u));
        if (slot < cap) out[slot] = id;
    }
}

// ---- Box_Search / Radius_Search: the same one-pass filter + warp-aggregated compaction over the sorted map -----------
// (ikd_Tree.cpp:401-412,1024-1095; half-open box test `min <= p < max`, sphere test d2 <= radius * radius in float)
struct RegionQuery {
    int kind;  // 0 box, 1 sphere
    float a[3], b[3];  // box min / max, or centre (a) and radius^2 (b[0])
};
__device__ __forceinline__ bool in_region(const RegionQuery& q, const float4& p) {
    if (q.kind == 0) return q.a[0] <= p.x && q.b[0] > p.x && q.a[1] <= p.y && q.b[1] > p.y && q.a[2] <= p.z && q.b[2] > p.z;
    return dist2_exact(p.x, p.y, p.z, q.a[0], q.a[1], q.a[2]) <= q.b[0];
}
__global__ void __launch_bounds__(256) region_kernel(const float4* __restrict__ sorted, int m, RegionQuery q, int32_t* __restrict__ out, int cap,
                                                     int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool hit = false;
    int id = -1;
    if (i < m) {
        const float4 p = sorted[i];
        id = __float_as_int(p.w);
        hit = in_region(q, p);
    }
    const unsigned b = __ballot_sync(FULL, hit);
    if (b == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(b) - 1) base = atomicAdd(count, __popc(b));
    base = __shfl_sync(FULL, base, __ffs(b) - 1);
    if (hit) {
        const int slot = base + __popc(b & ((1u << lane) - 1u));
        if (slot < cap) out[slot] = id;
    }
}

int map_region_search(Ctx* c, const Map& mp, int kind, const float a[3], const float b[3], int32_t* d_out, int cap, int* n_out_host) {
    CKS(reserve(c, c->d_scratch, 64));
    int* d_cnt = c->d_scratch.as<int>();
    CK(cudaMemsetAsync(d_cnt, 0, sizeof(int), c->stream));
    RegionQuery q;
    q.kind = kind;
    for (int i = 0; i < 3; ++i) {
        q.a[i] = a[i];
        q.b[i] = b[i];
    }
    const int m = mp.grid.m;
    if (m > 0) {
        region_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(mp.grid.sorted, m, q, d_out, cap, d_cnt);
        c->launches += 1;
    }
    CK(cudaGetLastError());
    int h = 0;
    CK(cudaMemcpyAsync(&h, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *n_out_host = h;
    return ICP4R_OK;
}

// ---- Delete_Point_Boxes / Add_Point_Boxes / Delete_Points (ikd_Tree.cpp:500-565,656-824) ------------------------------
// valid[] is the alive flag; userdel[] remembers that a point was removed by a delete call (and may come back with
// Add_Point_Boxes) as opposed to down-sampling (point_downsample_deleted: never comes back).
__global__ void __launch_bounds__(256) box_flag_kernel(const float4* __restrict__ pts, uint8_t* __restrict__ valid, uint8_t* __restrict__ userdel,
                                                       int m, const float* __restrict__ boxes6, int nb, int revive, int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool change = false;
    if (i < m) {
        const bool candidate = revive ? (!valid[i] && userdel[i]) : (valid[i] != 0);
        if (candidate) {
            const float4 p = pts[i];
            for (int b = 0; b < nb && !change; ++b) {
                const float* q = boxes6 + 6 * b;
                change = q[0] <= p.x && q[3] > p.x && q[1] <= p.y && q[4] > p.y && q[2] <= p.z && q[5] > p.z;
            }
            if (change) {
                valid[i] = revive ? 1 : 0;
                userdel[i] = revive ? 0 : 1;
            }
        }
    }
    const unsigned bal = __ballot_sync(FULL, change);
    if (bal && (threadIdx.x & 31) == 0) atomicAdd(count, __popc(bal));
}

// one warp walks the requests in order (each sees the deletions of the previous ones, like the reference's loop);
// for a request the lanes scan the cells within EPSS of the point and the lowest insertion index that matches wins
__global__ void __launch_bounds__(32) delete_points_kernel(GridDesc g, const float4* __restrict__ req, int n, uint8_t* __restrict__ valid,
                                                          uint8_t* __restrict__ userdel, int* __restrict__ count) {
    const int lane = threadIdx.x;
    int deleted = 0;
    for (int r = 0; r < n; ++r) {
        const float4 t = req[r];
        int best = 0x7fffffff;
        if (isfinite(t.x) && isfinite(t.y) && isfinite(t.z)) {
            const float e = 2e-6f + g.margin;
            const int x0 = cell_of(t.x - e, g.ox, g.inv_cell, g.nx), x1 = cell_of(t.x + e, g.ox, g.inv_cell, g.nx);
            const int y0 = cell_of(t.y - e, g.oy, g.inv_cell, g.ny), y1 = cell_of(t.y + e, g.oy, g.inv_cell, g.ny);
            const int z0 = cell_of(t.z - e, g.oz, g.inv_cell, g.nz), z1 = cell_of(t.z + e, g.oz, g.inv_cell, g.nz);
            for (int z = z0; z <= z1; ++z)
                for (int y = y0; y <= y1; ++y) {
                    const uint32_t rowbase = (uint32_t)(z * g.ny + y) * (uint32_t)g.nx;
                    const uint32_t s = g.cell_start[rowbase + x0], eidx = g.cell_start[rowbase + x1 + 1];
                    for (uint32_t j = s + lane; j < eidx; j += 32) {
                        const float4 p = g.sorted[j];
                        const int id = __float_as_int(p.w);
                        // same_point (ikd_Tree.cpp:1422-1424): float differences, compared with EPSS in double
                        if (valid[id] && (double)fabsf(__fsub_rn(p.x, t.x)) < 1e-6 && (double)fabsf(__fsub_rn(p.y, t.y)) < 1e-6 &&
                            (double)fabsf(__fsub_rn(p.z, t.z)) < 1e-6)
                            best = min(best, id);
                    }
                }
        }
        best = __reduce_min_sync(FULL, best);
        if (best != 0x7fffffff) {
            if (lane == 0) {
                valid[best] = 0;
                userdel[best] = 1;
            }
            ++deleted;
        }
        __syncwarp();
    }
    if (lane == 0) *count = deleted;
}

int map_box_flags(Ctx* c, Map& mp, const float* d_boxes6, int nb, bool revive, int* n_changed_host) {
    *n_changed_host = 0;
    if (nb <= 0 || mp.m <= 0) return ICP4R_OK;
    CKS(reserve(c, c->d_scratch, 64));
    int* d_cnt = c->d_scratch.as<int>();
    CK(cudaMemsetAsync(d_cnt, 0, sizeof(int), c->stream));
    box_flag_kernel<<<(mp.m + 255) / 256, 256, 0, c->stream>>>(mp.pts.as<float4>(), mp.valid.as<uint8_t>(), mp.userdel.as<uint8_t>(), mp.m, d_boxes6, nb,
                                                               revive ? 1 : 0, d_cnt);
    c->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(n_changed_host, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return ICP4R_OK;
}

int map_delete_points(Ctx* c, Map& mp, const float4* d_req, int n, int* n_deleted_host) {
    *n_deleted_host = 0;
    if (n <= 0 || mp.grid.m <= 0) return ICP4R_OK;
    CKS(reserve(c, c->d_scratch, 64));
    int* d_cnt = c->d_scratch.as<int>();
    CK(cudaMemsetAsync(d_cnt, 0, sizeof(int), c->stream));
    delete_points_kernel<<<1, 32, 0, c->stream>>>(mp.grid, d_req, n, mp.valid.as<uint8_t>(), mp.userdel.as<uint8_t>(), d_cnt);
    c->launches += 1;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(n_deleted_host, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return ICP4R_OK;
}

int map_sector(Ctx* c, const Map& mp, const float centre[3], float radius, float heading, int32_t* d_out, int cap, int* n_out_host) {
    CKS(reserve(c, c->d_scratch, 64));
    int* d_cnt = c->d_scratch.as<int>();
    CK(cudaMemsetAsync(d_cnt, 0, sizeof(int), c->stream));
    const int m = mp.grid.m;
    if (m > 0) {
        sector_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(mp.grid.sorted, m, centre[0], centre[1], centre[2], radius, heading, d_out, cap, d_cnt);
        c->launches += 1;
    }
    CK(cudaGetLastError());
    int h = 0;
    CK(cudaMemcpyAsync(&h, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *n_out_host = h;
    return ICP4R_OK;
}

}  // namespace icp4r
