// voxel_filter.cu — centroid-per-leaf down-sampling of a cloud.
//
// Replaces the pcl::VoxelGrid<PointXYZI> pass the scan-to-map node runs over the whole accumulated map every frame
// (/root/reference/src/radar_odometry.cpp:426-429, leaf 0.5 m). PCL is a dependency of the reference, not part of
// its tree; the algorithm restated here is VoxelGrid::applyFilter of PCL 1.8 with its defaults (downsample_all_data,
// min_points_per_voxel 0, no field filter):
//   min/max over the finite points -> min_b = floor(min * inv_leaf), div_b = max_b - min_b + 1
//   leaf index of a point = (floor(x*inv) - min_b.x) + (floor(y*inv) - min_b.y) * div_b.x + (floor(z*inv) - min_b.z) * div_b.x*div_b.y
//   points sorted by leaf index; one output point per occupied leaf, ascending leaf index: float sums of x, y, z,
//   intensity divided by the count (pcl::CentroidPoint accumulators).
// PCL sorts with std::sort, so the order in which a leaf's points are added (and with it the last float bit of the
// centroid) is unspecified there; here it is ascending input index (stable radix sort), which the oracle restates.
//
// Device plan (HBM-bound integer/byte work, no tensor cores): min/max reduction (16 B/pt read), key pass (16 B read,
// 8 B written), stable LSB radix sort of (leaf, index) pairs (radix_sort.cu, 16 B/pt per 8-bit digit), head flags +
// block counts, one-block scan of the block counts, and a leaf pass in which the thread that finds a leaf head
// sums that leaf's points in order (random 16 B gathers) and writes the centroid to its compacted slot.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "ctx.h"
#include "device_math.cuh"

namespace icp4r {

namespace {
constexpr int VG_THREADS = 256;
constexpr int VG_TILE = 1024;  // points per block in the head/count/leaf passes (4 per thread)

__device__ __forceinline__ int vg_f2ord(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
inline float vg_ord2f(int i) {
    const int j = i >= 0 ? i : i ^ 0x7fffffff;
    float f;
    std::memcpy(&f, &j, 4);
    return f;
}

__global__ void vg_minmax_init(int* bb) {
    if (threadIdx.x < 3) bb[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) bb[threadIdx.x] = (int)0x80000000;
    else if (threadIdx.x == 6) bb[6] = 0;
}

__global__ void __launch_bounds__(VG_THREADS) vg_minmax_kernel(const float4* __restrict__ pts, const uint8_t* __restrict__ valid, int n,
                                                              int* __restrict__ bb) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (valid && !valid[i]) continue;
        const float4 p = __ldg(pts + i);
        if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) continue;
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
        ++cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(FULL, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL, mx[a], o));
        }
        cnt += __shfl_xor_sync(FULL, cnt, o);
    }
    // one set of atomics per block, not per warp (they serialise on seven words)
    __shared__ float s_mn[VG_THREADS / 32][3], s_mx[VG_THREADS / 32][3];
    __shared__ int s_cnt[VG_THREADS / 32];
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            s_mn[w][a] = mn[a];
            s_mx[w][a] = mx[a];
        }
        s_cnt[w] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int j = 1; j < VG_THREADS / 32; ++j) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                mn[a] = fminf(mn[a], s_mn[j][a]);
                mx[a] = fmaxf(mx[a], s_mx[j][a]);
            }
            cnt += s_cnt[j];
        }
        if (cnt > 0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                atomicMin(bb + a, vg_f2ord(mn[a]));
                atomicMax(bb + 3 + a, vg_f2ord(mx[a]));
            }
            atomicAdd(bb + 6, cnt);
        }
    }
}

struct VgDesc {
    float inv[3];
    int min_b[3];
    int mul1, mul2;      // div_b.x, div_b.x * div_b.y
    uint32_t invalid;    // key of skipped points (== number of leaves of the index space: sorts last)
};

__global__ void __launch_bounds__(VG_THREADS) vg_key_kernel(const float4* __restrict__ pts, const uint8_t* __restrict__ valid, int n, VgDesc d,
                                                           uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = __ldg(pts + i);
    uint32_t key = d.invalid;
    if ((!valid || valid[i]) && isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        // PCL: static_cast<int>(floor(p.x * inverse_leaf_size[0]) - static_cast<float>(min_b[0])) — float arithmetic
        const int i0 = (int)__fsub_rn(floorf(__fmul_rn(p.x, d.inv[0])), (float)d.min_b[0]);
        const int i1 = (int)__fsub_rn(floorf(__fmul_rn(p.y, d.inv[1])), (float)d.min_b[1]);
        const int i2 = (int)__fsub_rn(floorf(__fmul_rn(p.z, d.inv[2])), (float)d.min_b[2]);
        key = (uint32_t)(i0 + i1 * d.mul1 + i2 * d.mul2);
    }
    keys[i] = key;
    vals[i] = (uint32_t)i;
}

// heads of the leaves in the sorted key array, counted per tile
__global__ void __launch_bounds__(VG_THREADS) vg_count_kernel(const uint32_t* __restrict__ keys, int nv, int* __restrict__ tile_cnt) {
    __shared__ int wsum[VG_THREADS / 32];
    const int base = blockIdx.x * VG_TILE;
    int c = 0;
#pragma unroll
    for (int u = 0; u < VG_TILE / VG_THREADS; ++u) {
        const int i = base + u * VG_THREADS + threadIdx.x;
        if (i < nv && (i == 0 || keys[i] != keys[i - 1])) ++c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int j = 0; j < VG_THREADS / 32; ++j) t += wsum[j];
        tile_cnt[blockIdx.x] = t;
    }
}

// exclusive scan of the tile counts by one block; total -> tile_cnt[tiles]
__global__ void __launch_bounds__(1024) vg_scan_kernel(int* __restrict__ tile_cnt, int tiles) {
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b0 = 0; b0 < tiles; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const int v = i < tiles ? tile_cnt[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            int s = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(FULL, s, o);
                if (lane >= o) s += y;
            }
            wsum[lane] = s;  // inclusive over warps
        }
        __syncthreads();
        const int before = carry + (w > 0 ? wsum[w - 1] : 0) + inc - v;
        if (i < tiles) tile_cnt[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_cnt[tiles] = carry;
}

// every leaf head sums its leaf in input order and writes the centroid to slot (tile offset + rank inside the tile)
__global__ void __launch_bounds__(VG_THREADS) vg_leaf_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ keys,
                                                            const uint32_t* __restrict__ vals, int nv, const int* __restrict__ tile_off,
                                                            float4* __restrict__ out, int cap) {
    __shared__ int wsum[VG_THREADS / 32];
    const int base = blockIdx.x * VG_TILE;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // thread t owns the 4 consecutive entries base + 4t .. base + 4t + 3, so ranks follow the sorted order
    const int i0 = base + threadIdx.x * (VG_TILE / VG_THREADS);
    bool head[VG_TILE / VG_THREADS];
    int c = 0;
#pragma unroll
    for (int u = 0; u < VG_TILE / VG_THREADS; ++u) {
        const int i = i0 + u;
        head[u] = i < nv && (i == 0 || keys[i] != keys[i - 1]);
        c += head[u];
    }
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    int before = tile_off[blockIdx.x] + inc - c;
    for (int j = 0; j < w; ++j) before += wsum[j];
#pragma unroll
    for (int u = 0; u < VG_TILE / VG_THREADS; ++u) {
        if (!head[u]) continue;
        const int slot = before++;
        if (slot >= cap) continue;
        const uint32_t key = keys[i0 + u];
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        int cnt = 0;
        for (int i = i0 + u; i < nv && keys[i] == key; ++i) {
            const float4 p = __ldg(pts + vals[i]);
            sx = __fadd_rn(sx, p.x);
            sy = __fadd_rn(sy, p.y);
            sz = __fadd_rn(sz, p.z);
            si = __fadd_rn(si, p.w);
            ++cnt;
        }
        const float fn = (float)cnt;
        out[slot] = make_float4(__fdiv_rn(sx, fn), __fdiv_rn(sy, fn), __fdiv_rn(sz, fn), __fdiv_rn(si, fn));
    }
}
}  // namespace

int voxel_grid(Ctx* c, const float4* d_pts, const uint8_t* d_valid, int n, float leaf, float4* d_out, int cap, int* n_out_host) {
    *n_out_host = 0;
    if (!(leaf > 0.0f) || !std::isfinite(leaf)) return fail(c, ICP4R_ERR_INVALID, "voxel grid: leaf size must be positive");
    if (n <= 0) return ICP4R_OK;
    CKS(reserve(c, c->d_scratch, 64));
    int* d_bb = c->d_scratch.as<int>();
    vg_minmax_init<<<1, 32, 0, c->stream>>>(d_bb);
    vg_minmax_kernel<<<std::max(1, std::min((n + 8 * VG_THREADS - 1) / (8 * VG_THREADS), c->sm_count * 8)), VG_THREADS, 0, c->stream>>>(d_pts, d_valid, n, d_bb);
    c->launches += 2;
    int bb[7];
    CK(cudaMemcpyAsync(bb, d_bb, sizeof(bb), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const int nfinite = bb[6];
    if (nfinite == 0) return ICP4R_OK;
    const float inv = 1.0f / leaf;
    VgDesc d;
    long long div[3];
    for (int a = 0; a < 3; ++a) {
        const float mn = vg_ord2f(bb[a]), mx = vg_ord2f(bb[3 + a]);
        d.inv[a] = inv;
        d.min_b[a] = (int)std::floor(mn * inv);
        const int max_b = (int)std::floor(mx * inv);
        div[a] = (long long)max_b - d.min_b[a] + 1;
        // PCL's own guard (voxel_grid.hpp: "Leaf size is too small for the input dataset. Integer indices would overflow.")
        const long long dx = (long long)((mx - mn) * inv) + 1;
        if (dx > 0x7fffffffLL) return fail(c, ICP4R_ERR_INVALID, "voxel grid: leaf size too small for the extent of the cloud");
    }
    const double cells = (double)div[0] * (double)div[1] * (double)div[2];
    if (cells > 2147483647.0) return fail(c, ICP4R_ERR_INVALID, "voxel grid: leaf size too small for the extent of the cloud (leaf index would overflow)");
    d.mul1 = (int)div[0];
    d.mul2 = (int)(div[0] * div[1]);
    d.invalid = (uint32_t)(div[0] * div[1] * div[2]);
    int bits = 1;
    while (bits < 32 && (1ull << bits) <= (unsigned long long)d.invalid) ++bits;

    CKS(reserve_grow(c, c->vg_keys, (size_t)n * 2 * sizeof(uint32_t)));  // the map this runs over grows every frame: geometric growth, not a cudaMalloc per call
    CKS(reserve_grow(c, c->vg_vals, (size_t)n * 2 * sizeof(uint32_t)));
    uint32_t *ka = c->vg_keys.as<uint32_t>(), *kb = ka + n, *va = c->vg_vals.as<uint32_t>(), *vb = va + n;
    vg_key_kernel<<<(n + VG_THREADS - 1) / VG_THREADS, VG_THREADS, 0, c->stream>>>(d_pts, d_valid, n, d, ka, va);
    c->launches += 1;
    uint32_t *ks = nullptr, *vs = nullptr;
    CKS(radix_sort_pairs(c, ka, kb, va, vb, n, bits, c->vg_sort, &ks, &vs));
    const int nv = nfinite;  // skipped points carry the largest key and sit behind the finite ones
    const int tiles = (nv + VG_TILE - 1) / VG_TILE;
    CKS(reserve_grow(c, c->vg_tiles, (size_t)(tiles + 1) * sizeof(int)));
    int* d_tiles = c->vg_tiles.as<int>();
    vg_count_kernel<<<tiles, VG_THREADS, 0, c->stream>>>(ks, nv, d_tiles);
    vg_scan_kernel<<<1, 1024, 0, c->stream>>>(d_tiles, tiles);
    vg_leaf_kernel<<<tiles, VG_THREADS, 0, c->stream>>>(d_pts, ks, vs, nv, d_tiles, d_out, cap);
    c->launches += 3;
    CK(cudaGetLastError());
    int total = 0;
    CK(cudaMemcpyAsync(&total, d_tiles + tiles, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *n_out_host = total;
    return ICP4R_OK;
}

}  // namespace icp4r
