// grid.cu — the device map: voxel-grid build (replaces KD_TREE::Build / BuildTree,
// /root/reference/third_party/ikd-Tree/ikd_Tree.cpp:354-365,582-630) and the stand-alone kNN kernels behind
// icp4r_map_knn / icp4r_map_knn_brute (replace Nearest_Search, ikd_Tree.cpp:368-398).
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "ctx.h"
#include "device_math.cuh"
#include "grid_knn.cuh"

namespace icp4r {

// ------------------------------------------------------------------------------------------------ bbox
__device__ __forceinline__ int f2ord(float f) {  // order-preserving float -> int
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
static inline float ord2f(int i) {
    const int j = i >= 0 ? i : i ^ 0x7fffffff;
    float f;
    std::memcpy(&f, &j, 4);
    return f;
}

__global__ void bbox_init(int* bb) {
    if (threadIdx.x < 3) bb[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) bb[threadIdx.x] = (int)0x80000000;
    else if (threadIdx.x == 6) bb[6] = 0;
}

__global__ void __launch_bounds__(256) bbox_kernel(const float4* __restrict__ pts, const uint8_t* __restrict__ valid, int m,
                                                   int* __restrict__ bb) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    int cnt = 0;
    const int stride = gridDim.x * blockDim.x;
    // four independent loads in flight per thread (one dependent load per trip left the kernel at a fifth of the HBM rate)
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < m; i0 += 4 * stride) {
        float4 p[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * stride;
            ok[u] = i < m;
            if (ok[u]) {
                p[u] = pts[i];
                ok[u] = valid[i] != 0;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (ok[u] && isfinite(p[u].x) && isfinite(p[u].y) && isfinite(p[u].z)) {
                mn[0] = fminf(mn[0], p[u].x); mx[0] = fmaxf(mx[0], p[u].x);
                mn[1] = fminf(mn[1], p[u].y); mx[1] = fmaxf(mx[1], p[u].y);
                mn[2] = fminf(mn[2], p[u].z); mx[2] = fmaxf(mx[2], p[u].z);
                ++cnt;
            }
        }
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
    // one set of atomics per BLOCK (per warp they serialised on the seven words: 18 us for a 110 k-point sub-map)
    __shared__ float s_mn[8][3], s_mx[8][3];
    __shared__ int s_cnt[8];
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
        for (int j = 1; j < 8; ++j) {
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
                atomicMin(bb + a, f2ord(mn[a]));
                atomicMax(bb + 3 + a, f2ord(mx[a]));
            }
            atomicAdd(bb + 6, cnt);
        }
    }
}

// ------------------------------------------------------------------------------------------------ keys
__global__ void __launch_bounds__(256) key_kernel(const float4* __restrict__ pts, const uint8_t* __restrict__ valid, int m,
                                                  GridDesc g, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float4 p = pts[i];
    uint32_t key = (uint32_t)g.ncells;  // invalid / non-finite points sort behind every cell
    if (valid[i] && isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        const int cx = cell_of(p.x, g.ox, g.inv_cell, g.nx);
        const int cy = cell_of(p.y, g.oy, g.inv_cell, g.ny);
        const int cz = cell_of(p.z, g.oz, g.inv_cell, g.nz);
        key = (uint32_t)(cz * g.ny + cy) * (uint32_t)g.nx + (uint32_t)cx;
    }
    keys[i] = key;
    vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) gather_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ vals, int n,
                                                     float4* __restrict__ sorted) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = vals[i];
    const float4 p = pts[j];
    sorted[i] = make_float4(p.x, p.y, p.z, __uint_as_float(j));
}

// cell_start[c] = first sorted position whose key >= c  (c in [0, ncells])
__global__ void __launch_bounds__(256) cell_start_kernel(const uint32_t* __restrict__ keys, int n, int ncells,
                                                         uint32_t* __restrict__ cell_start) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > ncells) return;
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < (uint32_t)c) lo = mid + 1;
        else hi = mid;
    }
    cell_start[c] = (uint32_t)lo;
}

__global__ void __launch_bounds__(256) occupied_kernel(const uint32_t* __restrict__ cell_start, int ncells, int* __restrict__ out) {
    int cnt = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += gridDim.x * blockDim.x)
        cnt += cell_start[c + 1] > cell_start[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(out, cnt);
}

// coarse occupancy: one WARP per block of 8 x 8 x 8 cells; its lanes take two of the block's 64 x-row segments each
// (one thread per block walked 128 dependent-latency loads: 22 us for a sub-map whose coarse table fits one thread block)
__global__ void __launch_bounds__(256) coarse_count_kernel(const uint32_t* __restrict__ cs, GridDesc g, uint32_t* __restrict__ coarse) {
    const int cnx = (g.nx + 7) >> 3, cny = (g.ny + 7) >> 3, cnz = (g.nz + 7) >> 3;
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= cnx * cny * cnz) return;
    const int Z = t / (cnx * cny), rem = t - Z * (cnx * cny), Y = rem / cnx, X = rem - Y * cnx;
    const int xa = X << 3, xb = min(xa + 8, g.nx);
    uint32_t cnt = 0;
#pragma unroll
    for (int r = lane; r < 64; r += 32) {
        const int z = (Z << 3) + (r >> 3), y = (Y << 3) + (r & 7);
        if (z < g.nz && y < g.ny) {
            const uint32_t rowbase = (uint32_t)(z * g.ny + y) * (uint32_t)g.nx;
            cnt += __ldg(cs + rowbase + xb) - __ldg(cs + rowbase + xa);
        }
    }
    cnt = __reduce_add_sync(FULL, cnt);
    if (lane == 0) coarse[t] = cnt;
}

static int build_coarse(Ctx* c, Map& mp, GridDesc& g) {
    const size_t nc = (size_t)((g.nx + 7) >> 3) * (size_t)((g.ny + 7) >> 3) * (size_t)((g.nz + 7) >> 3);
    CKS(reserve_grow(c, mp.coarse, nc * sizeof(uint32_t)));
    coarse_count_kernel<<<(unsigned)((nc + 7) / 8), 256, 0, c->stream>>>(mp.cell_start.as<uint32_t>(), g, mp.coarse.as<uint32_t>());
    c->launches += 1;
    g.coarse = mp.coarse.as<uint32_t>();
    return ICP4R_OK;
}

// ------------------------------------------------------------------------------------------------ bucket build
// Large maps: instead of an LSB radix sort of (key, index) pairs (3-4 passes of 20 B per point and pass, then a random
// 16-byte gather per point) the points themselves are moved twice:
//   1. scatter (x, y, z, index) into the fixed-capacity slot range of the point's BUCKET = 2^S consecutive cell keys
//      (one returning atomic per point on a few thousand cursors; arrival order), then scan the bucket counts;
//   2. one block per bucket loads its ~0.8 k points into shared memory, counting-sorts them by cell, orders every
//      cell's points by index (the order the stable radix sort would give: the result is bit-identical), writes them
//      out coalesced and fills the bucket's slice of the cell table — no binary search, no gather.
// 80 B of traffic per point instead of ~125 B, none of it a per-point random read. Falls back to the radix path when a
// bucket would not fit shared memory (strongly non-uniform maps).
constexpr int BK_THREADS = 256;
constexpr int BK_CAP = 2048;      // slots per bucket (scratch array and the sorting block's shared memory)
constexpr int BK_MAX_S = 12;      // at most 4096 cell keys per bucket
// The scatter is bound by the L2's atomic throughput, not by bandwidth and not by the SMs (ncu: issue slots 7 % busy,
// long-scoreboard stall 129 per issued instruction, DRAM at 19 %): 20 M returning atomics in ~0.4 ms = ~25 per clock over
// the whole L2. Spreading the cursors over separate 128-byte lines (stride 64 words) changed nothing (418 vs 410 us), so
// it is the read-modify-write rate of the slices, not same-line serialisation: packed cursors stay.
constexpr int BK_CSTRIDE = 1;     // uint32 words between two cursors

__device__ __forceinline__ uint32_t key_of_point(const GridDesc& g, const float4& p) {
    const int cx = cell_of(p.x, g.ox, g.inv_cell, g.nx);
    const int cy = cell_of(p.y, g.oy, g.inv_cell, g.ny);
    const int cz = cell_of(p.z, g.oz, g.inv_cell, g.nz);
    return (uint32_t)(cz * g.ny + cy) * (uint32_t)g.nx + (uint32_t)cx;
}

// one block: exclusive scan of the bucket counts (base[nb] = total), info = {largest bucket}. Every thread owns a contiguous
// run of buckets (one pass, two barriers; chunks of 1024 with three barriers each took 26 us for 20 k buckets)
__global__ void __launch_bounds__(1024) bk_scan_kernel(const uint32_t* __restrict__ cnt, int nb, uint32_t* __restrict__ base,
                                                       uint32_t* __restrict__ info) {
    __shared__ uint32_t wsum[32], wmax[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int per = (nb + 1023) / 1024;
    const int i0 = min(nb, tid * per), i1 = min(nb, i0 + per);
    uint32_t sum = 0, vmax = 0;
    for (int i = i0; i < i1; ++i) {
        const uint32_t v = cnt[(size_t)i * BK_CSTRIDE];
        sum += v;
        vmax = max(vmax, v);
    }
    uint32_t x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, x, o);
        if (lane >= o) x += y;
    }
    vmax = __reduce_max_sync(FULL, vmax);
    if (lane == 31) wsum[w] = x;
    if (lane == 0) wmax[w] = vmax;
    __syncthreads();
    uint32_t run = x - sum;
    for (int j = 0; j < w; ++j) run += wsum[j];
    for (int i = i0; i < i1; ++i) {
        base[i] = run;
        run += cnt[(size_t)i * BK_CSTRIDE];
    }
    if (tid == 1023) {
        base[nb] = run;  // the last thread's run ends at the total (threads past the end own empty runs)
        uint32_t m = 0;
        for (int j = 0; j < 32; ++j) m = max(m, wmax[j]);
        info[0] = m;
    }
}

__global__ void __launch_bounds__(256) bk_scatter_kernel(const float4* __restrict__ pts, const uint8_t* __restrict__ valid, int m, GridDesc g,
                                                         int S, uint32_t* __restrict__ cursor, float4* __restrict__ tmp) {
    const int stride = gridDim.x * blockDim.x;
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < m; i0 += 4 * stride) {
        float4 p[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * stride;
            ok[u] = i < m;
            if (ok[u]) {
                p[u] = pts[i];
                ok[u] = valid[i] != 0;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (ok[u] && isfinite(p[u].x) && isfinite(p[u].y) && isfinite(p[u].z)) {
                const uint32_t b = key_of_point(g, p[u]) >> S;
                const uint32_t slot = atomicAdd(cursor + (size_t)b * BK_CSTRIDE, 1u);  // a bucket that overflows is detected by the scan: radix path
                if (slot < (uint32_t)BK_CAP)
                    tmp[(size_t)b * BK_CAP + slot] = make_float4(p[u].x, p[u].y, p[u].z, __uint_as_float((uint32_t)(i0 + u * stride)));
            }
    }
}

// block b sorts bucket b: keys [b << S, (b + 1) << S)
__global__ void __launch_bounds__(BK_THREADS) bk_sort_kernel(const float4* __restrict__ tmp, const uint32_t* __restrict__ base, GridDesc g, int S,
                                                            float4* __restrict__ sorted, uint32_t* __restrict__ cell_start) {
    extern __shared__ __align__(16) unsigned char bk_smem[];
    float4* s_pts = reinterpret_cast<float4*>(bk_smem);                        // [BK_CAP]
    unsigned short* s_fk = reinterpret_cast<unsigned short*>(s_pts + BK_CAP);  // [BK_CAP] cell of a point, relative to the bucket
    unsigned short* s_perm = s_fk + BK_CAP;                                    // [BK_CAP] point at a sorted slot
    uint32_t* s_start = reinterpret_cast<uint32_t*>(s_perm + BK_CAP);          // [2^S + 1]
    uint32_t* s_cur = s_start + ((1 << S) + 1);                                // [2^S]
    __shared__ uint32_t wsum[BK_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t b = blockIdx.x;
    const uint32_t off = base[b], nbk = base[b + 1] - off;
    const int ncell = 1 << S;
    for (int c = tid; c <= ncell; c += BK_THREADS) s_start[c] = 0;
    for (int c = tid; c < ncell; c += BK_THREADS) s_cur[c] = 0;
    __syncthreads();
    for (uint32_t j = tid; j < nbk; j += BK_THREADS) {
        const float4 p = tmp[(size_t)b * BK_CAP + j];
        s_pts[j] = p;
        const uint32_t fk = key_of_point(g, p) - (b << S);
        s_fk[j] = (unsigned short)fk;
        atomicAdd(&s_start[fk + 1], 1u);
    }
    __syncthreads();
    {   // s_start[c + 1] <- inclusive prefix of the counts, i.e. s_start[c] = first slot of cell c
        const int per = (ncell + BK_THREADS - 1) / BK_THREADS;
        uint32_t sum = 0;
        for (int t = 0; t < per; ++t) {
            const int c = tid * per + t;
            if (c < ncell) sum += s_start[c + 1];
        }
        uint32_t x = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        uint32_t run = x - sum;
        for (int j = 0; j < w; ++j) run += wsum[j];
        for (int t = 0; t < per; ++t) {
            const int c = tid * per + t;
            if (c < ncell) {
                run += s_start[c + 1];
                s_start[c + 1] = run;
            }
        }
    }
    __syncthreads();
    for (uint32_t j = tid; j < nbk; j += BK_THREADS) {
        const uint32_t fk = s_fk[j];
        s_perm[s_start[fk] + atomicAdd(&s_cur[fk], 1u)] = (unsigned short)j;
    }
    __syncthreads();
    // ascending index inside every cell (what the stable sort of the radix path produces): every point counts the points
    // of its cell with a smaller index — its rank — and goes straight to its final position. (One thread insertion-sorting
    // a whole cell left the block waiting at the barrier for the thread with the fullest cell: barrier stall 10.9 per issue.)
    for (uint32_t a = tid; a < nbk; a += BK_THREADS) {
        const unsigned short pj = s_perm[a];
        const uint32_t fk = s_fk[pj];
        const uint32_t s0 = s_start[fk], e0 = s_start[fk + 1];
        const float4 me = s_pts[pj];
        const uint32_t idx = __float_as_uint(me.w);
        uint32_t rank = 0;
        for (uint32_t q = s0; q < e0; ++q) rank += __float_as_uint(s_pts[s_perm[q]].w) < idx ? 1u : 0u;
        sorted[off + s0 + rank] = me;
    }
    for (int c = tid; c < ncell; c += BK_THREADS) {
        const uint32_t key = (b << S) + (uint32_t)c;
        if (key <= (uint32_t)g.ncells) cell_start[key] = off + s_start[c];
    }
}

// *built = true if the grid (sorted points + cell table) was built, false if the caller must take the radix path
static int bucket_build(Ctx* c, Map& mp, const GridDesc& g, int m, int nvalid, bool* built) {
    *built = false;
    const char* e = std::getenv("ICP4R_BUCKET_MIN");
    // worth it for large maps only: below a few million points the radix path is faster (a 290 k-point sector sub-map spent
    // 136 us in the scatter alone — few buckets, contended cursors — against ~75 us for the whole radix sort), and a map
    // whose buckets overflowed once (surfaces: strongly non-uniform occupancy) is not tried again until it is rebuilt from scratch
    const int min_pts = e ? std::atoi(e) : 4000000;
    if (nvalid < min_pts || min_pts < 0 || mp.no_bucket) return ICP4R_OK;
    // about 0.8 k points per bucket by volume (slots: BK_CAP)
    const double ppc = (double)nvalid / std::max(g.ncells, 1);
    int S = (int)std::lround(std::log2(std::max(800.0 / std::max(ppc, 1e-9), 1.0)));
    S = std::min(std::max(S, 2), BK_MAX_S);
    const long long nb_ll = ((long long)g.ncells + 1 + (1ll << S) - 1) >> S;
    if (nb_ll * BK_CAP > 4ll * std::max(m, 1) + (1 << 20)) return ICP4R_OK;  // mostly empty buckets (surfaces in a big volume): not worth the scratch
    const int nb = (int)nb_ll;
    CKS(reserve_grow(c, c->d_scratch, ((size_t)nb * BK_CSTRIDE + nb + 16) * sizeof(uint32_t)));
    uint32_t* cursor = c->d_scratch.as<uint32_t>();   // [nb * BK_CSTRIDE] points per bucket, one cursor per 256 bytes
    uint32_t* base = cursor + (size_t)nb * BK_CSTRIDE;  // [nb + 1]
    uint32_t* info = base + nb + 1;                   // [1]
    CKS(reserve_grow(c, mp.sorted_alt, (size_t)nb * BK_CAP * sizeof(float4)));
    CK(cudaMemsetAsync(cursor, 0, (size_t)nb * BK_CSTRIDE * sizeof(uint32_t), c->stream));
    const int blocks = std::min((m + 1023) / 1024, c->sm_count * 8);
    bk_scatter_kernel<<<blocks, 256, 0, c->stream>>>(mp.pts.as<float4>(), mp.valid.as<uint8_t>(), m, g, S, cursor, mp.sorted_alt.as<float4>());
    bk_scan_kernel<<<1, 1024, 0, c->stream>>>(cursor, nb, base, info);
    c->launches += 2;
    uint32_t h_max = 0;
    CK(cudaMemcpyAsync(&h_max, info, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (h_max > (uint32_t)BK_CAP) {
        mp.no_bucket = true;
        return ICP4R_OK;
    }
    const size_t smem = (size_t)BK_CAP * (sizeof(float4) + 2 * sizeof(unsigned short)) + ((size_t)2 * (1 << S) + 1) * sizeof(uint32_t);
    static bool attr_set = false;
    if (!attr_set) {
        CK(cudaFuncSetAttribute(bk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)((size_t)BK_CAP * (sizeof(float4) + 2 * sizeof(unsigned short)) + ((size_t)2 * (1 << BK_MAX_S) + 1) * sizeof(uint32_t))));
        attr_set = true;
    }
    bk_sort_kernel<<<nb, BK_THREADS, smem, c->stream>>>(mp.sorted_alt.as<float4>(), base, g, S, mp.sorted.as<float4>(), mp.cell_start.as<uint32_t>());
    c->launches += 1;
    CK(cudaGetLastError());
    *built = true;
    return ICP4R_OK;
}

int map_reserve(Ctx* c, Map& mp, int cap) {
    if ((size_t)cap * sizeof(float4) <= mp.pts.cap) return ICP4R_OK;
    // grow geometrically, keep contents
    size_t want = (size_t)cap;
    size_t have = mp.pts.cap / sizeof(float4);
    if (want < have * 2) want = have * 2;
    DevBuf np, nv, nu;
    CKS(reserve(c, np, want * sizeof(float4)));
    CKS(reserve(c, nv, want));
    CKS(reserve(c, nu, want));
    CK(cudaMemsetAsync(nu.p, 0, want, c->stream));
    if (mp.m > 0) {
        CK(cudaMemcpyAsync(np.p, mp.pts.p, (size_t)mp.m * sizeof(float4), cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaMemcpyAsync(nv.p, mp.valid.p, (size_t)mp.m, cudaMemcpyDeviceToDevice, c->stream));
        if (mp.userdel.p) CK(cudaMemcpyAsync(nu.p, mp.userdel.p, (size_t)mp.m, cudaMemcpyDeviceToDevice, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    release(mp.pts);
    release(mp.valid);
    release(mp.userdel);
    mp.pts = np;
    mp.valid = nv;
    mp.userdel = nu;
    return ICP4R_OK;
}

namespace {
struct Trace {  // ICP4R_TRACE=1: wall time of the build stages (each stage is followed by a stream sync)
    bool on;
    cudaStream_t st;
    std::chrono::steady_clock::time_point t0;
    explicit Trace(cudaStream_t s) : st(s) {
        const char* e = std::getenv("ICP4R_TRACE");
        on = e && e[0] == '1';
        t0 = std::chrono::steady_clock::now();
    }
    void mark(const char* what) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[icp4r trace] %-22s %8.1f us\n", what, std::chrono::duration<double, std::micro>(t1 - t0).count());
        t0 = t1;
    }
};
}  // namespace

int map_rebuild_grid(Ctx* c, Map& mp) {
    const int m = mp.m;
    Trace tr(c->stream);
    mp.built = false;
    mp.normals_k = 0;  // any cached GICP normals belong to the previous point set
    mp.grid = GridDesc{};
    if (m <= 0) {
        mp.m_valid = 0;
        CKS(reserve(c, mp.cell_start, 2 * sizeof(uint32_t)));
        CK(cudaMemsetAsync(mp.cell_start.p, 0, 2 * sizeof(uint32_t), c->stream));
        GridDesc g{};
        g.nx = g.ny = g.nz = 1;
        g.ncells = 1;
        g.cell = g.inv_cell = 1.f;
        g.m = 0;
        g.sorted = mp.sorted.as<float4>();
        g.cell_start = mp.cell_start.as<uint32_t>();
        CKS(build_coarse(c, mp, g));
        mp.grid = g;
        mp.built = true;
        return ICP4R_OK;
    }
    // 1. bounding box of the valid points
    CKS(reserve(c, c->d_scratch, 4096));
    int* d_bb = c->d_scratch.as<int>();
    bbox_init<<<1, 32, 0, c->stream>>>(d_bb);
    const int bblocks = std::max(1, std::min((m + 4095) / 4096, c->sm_count * 8));  // >= 16 points per thread
    bbox_kernel<<<bblocks, 256, 0, c->stream>>>(mp.pts.as<float4>(), mp.valid.as<uint8_t>(), m, d_bb);
    c->launches += 2;
    int h_bb[8];
    CK(cudaMemcpyAsync(h_bb, d_bb, 7 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const int nvalid = h_bb[6];
    mp.m_valid = nvalid;
    tr.mark("bbox");
    float mn[3], mx[3];
    for (int a = 0; a < 3; ++a) {
        mn[a] = nvalid ? ord2f(h_bb[a]) : 0.f;
        mx[a] = nvalid ? ord2f(h_bb[3 + a]) : 0.f;
        mp.bb_min[a] = mn[a];
        mp.bb_max[a] = mx[a];
    }
    // 2. grid geometry: ~8 points per cell by volume unless the caller fixed the cell size
    double ext[3];
    double L = 1.0;
    for (int a = 0; a < 3; ++a) {
        ext[a] = (double)mx[a] - (double)mn[a];
        L = std::max(L, std::max(std::fabs((double)mn[a]), std::fabs((double)mx[a])));
    }
    const double ext_tight[3] = {ext[0], ext[1], ext[2]};
    if (mp.padded) {
        // a growing map (odometry): leave room around the bounding box so that the next batches merge into this
        // grid instead of forcing a new sort; exactness never depends on the geometry
        const double emax_t = std::max(ext[0], std::max(ext[1], ext[2]));
        const double cell_guess = mp.user_cell > 0.f ? (double)mp.user_cell : (mp.hint_cell > 0.f ? (double)mp.hint_cell : 0.005 * emax_t);
        for (int a = 0; a < 3; ++a) {
            // proportional to the axis' own extent (a ground vehicle's map grows in x and y, hardly in z), a few cells at least
            const double pad = std::max(0.15 * ext[a], 4.0 * cell_guess);
            mn[a] = (float)((double)mn[a] - pad);
            mx[a] = (float)((double)mx[a] + pad);
            ext[a] = (double)mx[a] - (double)mn[a];
            L = std::max(L, std::max(std::fabs((double)mn[a]), std::fabs((double)mx[a])));
        }
    }
    const double emax = std::max(ext[0], std::max(ext[1], ext[2]));
    double cell = mp.user_cell > 0.f ? (double)mp.user_cell : (double)mp.hint_cell;
    if (!(cell > 0.0)) {
        const double floor_e = std::max(emax * 1e-3, 1e-6);
        const double vol = std::max(ext_tight[0], floor_e) * std::max(ext_tight[1], floor_e) * std::max(ext_tight[2], floor_e);
        cell = std::cbrt(vol / std::max(1.0, nvalid / 8.0));
        cell = std::max(cell, std::max(emax * 1e-4, 1e-6));
    }
    const double max_cells = 64.0 * 1024 * 1024;
    GridDesc g{};
    // The volume estimate under-counts density when the points lie on surfaces (ground, walls). After the first
    // build the number of occupied cells gives the real points-per-occupied-cell figure; if it is far from the
    // target (~4: the 3x3x3 block around a query then holds a few dozen candidates and the 5th neighbour is
    // still nearer than the block's faces) the cell is rescaled once or twice and the grid rebuilt.
    for (int pass = 0; pass < 3; ++pass) {
        int nx, ny, nz;
        for (;;) {
            const double fx = std::floor(ext[0] / cell) + 1, fy = std::floor(ext[1] / cell) + 1, fz = std::floor(ext[2] / cell) + 1;
            if (fx * fy * fz <= max_cells && fx < 2e6 && fy < 2e6 && fz < 2e6) {
                nx = (int)fx;
                ny = (int)fy;
                nz = (int)fz;
                break;
            }
            cell *= 1.26;  // coarsen: exactness does not depend on the cell size
        }
        g = GridDesc{};
        g.ox = mn[0];
        g.oy = mn[1];
        g.oz = mn[2];
        g.cell = (float)cell;
        g.inv_cell = 1.0f / g.cell;
        g.nx = nx;
        g.ny = ny;
        g.nz = nz;
        g.ncells = nx * ny * nz;
        g.m = nvalid;
        g.margin = (float)(L * 9.5367431640625e-7);
        // 3. keys + stable radix sort by key
        CKS(reserve_grow(c, mp.keys_a, (size_t)m * 4));
        CKS(reserve_grow(c, mp.keys_b, (size_t)m * 4));
        CKS(reserve_grow(c, mp.vals_a, (size_t)m * 4));
        CKS(reserve_grow(c, mp.vals_b, (size_t)m * 4));
        CKS(reserve_grow(c, mp.sorted, (size_t)std::max(m, 1) * sizeof(float4)));
        CKS(reserve_grow(c, mp.cell_start, ((size_t)g.ncells + 2) * sizeof(uint32_t)));
        tr.mark("reserve");
        uint32_t *ks = nullptr, *vs = nullptr;
        bool bucketed = false;
        CKS(bucket_build(c, mp, g, m, nvalid, &bucketed));  // large maps: two moves of the points, no radix sort
        if (bucketed) {
            tr.mark("bucket build");
        } else {
            key_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(mp.pts.as<float4>(), mp.valid.as<uint8_t>(), m, g,
                                                                 mp.keys_a.as<uint32_t>(), mp.vals_a.as<uint32_t>());
            c->launches += 1;
            tr.mark("keys");
            int bits = 1;
            while ((1ll << bits) <= (long long)g.ncells) ++bits;  // key == ncells must be representable
            CKS(radix_sort_pairs(c, mp.keys_a.as<uint32_t>(), mp.keys_b.as<uint32_t>(), mp.vals_a.as<uint32_t>(),
                                 mp.vals_b.as<uint32_t>(), m, bits, c->d_scratch, &ks, &vs));
            tr.mark("radix sort");
            // 4. cell table by binary search over the sorted keys
            cell_start_kernel<<<(g.ncells + 1 + 255) / 256, 256, 0, c->stream>>>(ks, m, g.ncells, mp.cell_start.as<uint32_t>());
            c->launches += 1;
            tr.mark("cell table");
        }
        bool again = false;
        if (!(mp.user_cell > 0.f) && !mp.quick_build && pass < 2 && nvalid >= 64) {
            int* d_occ = c->d_scratch.as<int>();
            CK(cudaMemsetAsync(d_occ, 0, sizeof(int), c->stream));
            occupied_kernel<<<std::min((g.ncells + 255) / 256, c->sm_count * 8), 256, 0, c->stream>>>(mp.cell_start.as<uint32_t>(),
                                                                                                    g.ncells, d_occ);
            c->launches += 1;
            int occ = 0;
            CK(cudaMemcpyAsync(&occ, d_occ, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            tr.mark("occupancy");
            if (tr.on) fprintf(stderr, "[icp4r trace] pass %d cell %.3f cells %d occupied %d\n", pass, cell, g.ncells, occ);
            const double per = (double)nvalid / std::max(occ, 1);
            if (per > 8.0 || per < 2.0) {
                // occupied cells scale between c^2 (surfaces) and c^3 (volumes): take the gentler exponent
                const double f = std::sqrt(4.0 / per);
                const double nc = std::max(cell * std::min(std::max(f, 0.25), 4.0), std::max(emax * 1e-4, 1e-6));
                if (std::fabs(nc - cell) > 0.1 * cell) {
                    cell = nc;
                    again = true;
                }
            }
        }
        if (again) continue;
        // 5. points into sorted order (valid ones come first)
        if (nvalid > 0 && !bucketed) {
            gather_kernel<<<(nvalid + 255) / 256, 256, 0, c->stream>>>(mp.pts.as<float4>(), vs, nvalid, mp.sorted.as<float4>());
            c->launches += 1;
        }
        tr.mark("gather");
        break;
    }
    g.sorted = mp.sorted.as<float4>();
    g.cell_start = mp.cell_start.as<uint32_t>();
    CKS(build_coarse(c, mp, g));
    tr.mark("coarse table");
    CK(cudaGetLastError());
    mp.grid = g;
    mp.built = true;
    mp.hint_cell = g.cell;
    mp.valid_at_build = nvalid;
    return ICP4R_OK;
}

// ------------------------------------------------------------------------------------------------ incremental append
// Add_Points(.., false) of a batch that lies inside the current grid: instead of sorting all M + n points again
// (about ten passes over the map), sort the n new keys and MERGE: every old point moves up by the number of new
// points in lower cells, every new point lands behind the old points of its cell (stable: insertion order inside a
// cell is kept, which is what the full sort produces), and the cell table is shifted by the same counts. One read
// and one write of the sorted array (32 B per map point) — the HBM floor for keeping the array contiguous.
namespace {
__device__ __forceinline__ int lower_bound_u32(const uint32_t* __restrict__ a, int lo, int hi, uint32_t v) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// keys of the new points; info = {outside flag, finite count, ordered min xyz, ordered max xyz}
__global__ void __launch_bounds__(256) inc_key_kernel(const float4* __restrict__ pts, int base, int n, GridDesc g, uint32_t* __restrict__ keys,
                                                      uint32_t* __restrict__ vals, int* __restrict__ info) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[base + i];
    uint32_t key = (uint32_t)g.ncells;
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
        const int cx = __float2int_rd(__fmul_rn(__fsub_rn(p.x, g.ox), g.inv_cell));
        const int cy = __float2int_rd(__fmul_rn(__fsub_rn(p.y, g.oy), g.inv_cell));
        const int cz = __float2int_rd(__fmul_rn(__fsub_rn(p.z, g.oz), g.inv_cell));
        if (cx < 0 || cx >= g.nx || cy < 0 || cy >= g.ny || cz < 0 || cz >= g.nz) {
            atomicOr(info, 1);
        } else {
            key = (uint32_t)(cz * g.ny + cy) * (uint32_t)g.nx + (uint32_t)cx;  // == cell_of() for an inside point
        }
        atomicAdd(info + 1, 1);
        atomicMin(info + 2, f2ord(p.x)); atomicMax(info + 5, f2ord(p.x));
        atomicMin(info + 3, f2ord(p.y)); atomicMax(info + 6, f2ord(p.y));
        atomicMin(info + 4, f2ord(p.z)); atomicMax(info + 7, f2ord(p.z));
    }
    keys[i] = key;
    vals[i] = (uint32_t)(base + i);
}

__global__ void inc_info_init(int* info) {
    const int t = threadIdx.x;
    if (t < 2) info[t] = 0;
    else if (t < 5) info[t] = 0x7fffffff;
    else if (t < 8) info[t] = (int)0x80000000;
}

// small batches (a radar / lidar frame): one block sorts the (key, index) pairs in shared memory. The pairs are
// unique, so sorting the packed 64-bit words gives the stable order (ascending index inside a cell).
constexpr int SB_MAX = 4096;
// Bitonic network over 4096 packed words, 4 consecutive words per thread: strides 1 and 2 stay inside a thread, strides
// 4..64 are lane exchanges (shuffles), only strides >= 128 go through shared memory (15 of the 78 steps).
__device__ __forceinline__ unsigned long long cmpx(unsigned long long a, unsigned long long b, bool keep_min) {
    return (a < b) == keep_min ? a : b;
}
// Frame-sized batches (n <= SB_MAX) sorted by RANK: every block stages all n packed (key, index) words in shared memory and
// each of its threads counts the words smaller than its own — n broadcast reads and compares, no barrier after the load —
// then writes its element to that position. n / 128 blocks on as many SMs: ~8 us for 3,000 keys, where the one-block bitonic
// network below (78 compare-exchange steps, 15 of them through shared memory with a block barrier each) took 30 us.
constexpr int RK_THREADS = 128;
__global__ void __launch_bounds__(RK_THREADS) rank_sort_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, int n,
                                                              uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ unsigned long long s[SB_MAX];
    for (int i = threadIdx.x; i < n; i += RK_THREADS) s[i] = ((unsigned long long)keys[i] << 32) | vals[i];
    __syncthreads();
    const int i = blockIdx.x * RK_THREADS + threadIdx.x;
    if (i >= n) return;
    const unsigned long long mine = s[i];  // words are unique: the index sits in the low half
    int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    int j = 0;
    for (; j + 4 <= n; j += 4) {
        r0 += s[j] < mine;
        r1 += s[j + 1] < mine;
        r2 += s[j + 2] < mine;
        r3 += s[j + 3] < mine;
    }
    for (; j < n; ++j) r0 += s[j] < mine;
    const int rank = (r0 + r1) + (r2 + r3);
    keys_out[rank] = (uint32_t)(mine >> 32);
    vals_out[rank] = (uint32_t)mine;
}

__global__ void __launch_bounds__(1024) small_sort_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, int n,
                                                         uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ unsigned long long s[SB_MAX];
    const int t = threadIdx.x;
    unsigned long long e[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = 4 * t + r;
        e[r] = i < n ? ((unsigned long long)keys[i] << 32) | vals[i] : ~0ull;
    }
    for (int k = 2; k <= SB_MAX; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 128) {  // partner lives in another warp
#pragma unroll
                for (int r = 0; r < 4; ++r) s[4 * t + r] = e[r];
                __syncthreads();
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int i = 4 * t + r;
                    const unsigned long long o = s[i ^ j];
                    e[r] = cmpx(e[r], o, ((i & j) == 0) == ((i & k) == 0));
                }
                __syncthreads();
            } else if (j >= 4) {  // partner lane, same slot
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int i = 4 * t + r;
                    const unsigned long long o = __shfl_xor_sync(FULL, e[r], j >> 2);
                    e[r] = cmpx(e[r], o, ((i & j) == 0) == ((i & k) == 0));
                }
            } else {  // inside the thread
                const bool asc = ((4 * t) & k) == 0;  // k >= 4 here or the 4-group shares the direction bit... see below
                if (j == 2) {
                    const bool a0 = k == 2 ? true : asc;  // k == 2 never has j == 2
                    const unsigned long long x0 = e[0], x1 = e[1], x2 = e[2], x3 = e[3];
                    e[0] = cmpx(x0, x2, a0);
                    e[2] = cmpx(x2, x0, !a0);
                    e[1] = cmpx(x1, x3, a0);
                    e[3] = cmpx(x3, x1, !a0);
                } else {  // j == 1: direction bit k may split the thread's four words when k == 2
                    const bool a01 = ((4 * t + 0) & k) == 0, a23 = ((4 * t + 2) & k) == 0;
                    const unsigned long long x0 = e[0], x1 = e[1], x2 = e[2], x3 = e[3];
                    e[0] = cmpx(x0, x1, a01);
                    e[1] = cmpx(x1, x0, !a01);
                    e[2] = cmpx(x2, x3, a23);
                    e[3] = cmpx(x3, x2, !a23);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = 4 * t + r;
        if (i < n) {
            keys_out[i] = (uint32_t)(e[r] >> 32);
            vals_out[i] = (uint32_t)e[r];
        }
    }
}

constexpr int INC_THREADS = 256;
// search bounds of every block of the two merge passes below, so that the passes themselves start with a
// (usually empty) range instead of a chain of dependent loads: bnd[b] = lower_bound(first key of block b); the entry
// behind the last block is nf
// both tables in ONE launch (they are independent: blocks [0, gb_old) fill the old-point bounds, the rest the cell bounds)
__global__ void __launch_bounds__(256) inc_bounds_kernel(const float4* __restrict__ sorted_old, int m_old, GridDesc g, const uint32_t* __restrict__ nkeys,
                                                        int nf, int ob, int* __restrict__ bnd_old, int gb_old, int cb, int* __restrict__ bnd_cell) {
    if ((int)blockIdx.x < gb_old) {
        const int b = blockIdx.x * blockDim.x + threadIdx.x;
        if (b > ob) return;
        if (b == ob) {
            bnd_old[b] = nf;
            return;
        }
        const float4 p = sorted_old[(size_t)b * INC_THREADS];
        const uint32_t key = (uint32_t)(cell_of(p.z, g.oz, g.inv_cell, g.nz) * g.ny + cell_of(p.y, g.oy, g.inv_cell, g.ny)) * (uint32_t)g.nx +
                             (uint32_t)cell_of(p.x, g.ox, g.inv_cell, g.nx);
        bnd_old[b] = lower_bound_u32(nkeys, 0, nf, key);
    } else {
        const int b = ((int)blockIdx.x - gb_old) * blockDim.x + threadIdx.x;
        if (b > cb) return;
        bnd_cell[b] = b == cb ? nf : lower_bound_u32(nkeys, 0, nf, (uint32_t)b * INC_THREADS);
    }
}
// old sorted point j moves to j + (new points in lower cells)
__global__ void __launch_bounds__(INC_THREADS) inc_merge_old_kernel(const float4* __restrict__ sorted_old, int m_old, GridDesc g,
                                                                   const uint32_t* __restrict__ nkeys, const int* __restrict__ bnd,
                                                                   float4* __restrict__ sorted_new) {
    const int j = blockIdx.x * INC_THREADS + threadIdx.x;
    if (j >= m_old) return;
    const float4 p = sorted_old[j];
    // the block's points are consecutive in cell order: its keys lie between the first key of this block and the
    // first key of the next one, so the answer lies in [bnd[b], bnd[b+1]] — an empty range for most blocks
    const int lo = __ldg(bnd + blockIdx.x), hi = __ldg(bnd + blockIdx.x + 1);
    int sh = lo;
    if (lo < hi) {
        const uint32_t key = (uint32_t)(cell_of(p.z, g.oz, g.inv_cell, g.nz) * g.ny + cell_of(p.y, g.oy, g.inv_cell, g.ny)) * (uint32_t)g.nx +
                             (uint32_t)cell_of(p.x, g.ox, g.inv_cell, g.nx);
        sh = lower_bound_u32(nkeys, lo, hi, key);
    }
    sorted_new[j + sh] = p;
}

// new point i (i-th in key order) lands behind the old points of its cell and the new ones before it
__global__ void __launch_bounds__(256) inc_merge_new_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ nkeys,
                                                           const uint32_t* __restrict__ nvals, int nf, const uint32_t* __restrict__ cs_old,
                                                           float4* __restrict__ sorted_new, GridDesc g, uint32_t* __restrict__ coarse) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    const uint32_t k = nkeys[i], idx = nvals[i];
    const float4 p = pts[idx];
    sorted_new[(uint32_t)i + cs_old[k + 1]] = make_float4(p.x, p.y, p.z, __uint_as_float(idx));
    // the coarse occupancy table counts this point too (cheaper than recounting the whole table)
    const uint32_t row = k / (uint32_t)g.nx, cx = k - row * (uint32_t)g.nx, cz = row / (uint32_t)g.ny, cy = row - cz * (uint32_t)g.ny;
    const uint32_t cnx = (uint32_t)(g.nx + 7) >> 3, cny = (uint32_t)(g.ny + 7) >> 3;
    atomicAdd(coarse + ((size_t)(cz >> 3) * cny + (cy >> 3)) * cnx + (cx >> 3), 1u);
}

// cell_start[c] += new points in cells < c (in place; runs after inc_merge_new_kernel has read the old table)
__global__ void __launch_bounds__(INC_THREADS) inc_cell_kernel(uint32_t* __restrict__ cs, int ncells, const uint32_t* __restrict__ nkeys,
                                                              const int* __restrict__ bnd) {
    const int c = blockIdx.x * INC_THREADS + threadIdx.x;
    if (c > ncells) return;
    const int lo = __ldg(bnd + blockIdx.x), hi = __ldg(bnd + blockIdx.x + 1);
    const int add = lo < hi ? lower_bound_u32(nkeys, lo, hi, (uint32_t)c) : lo;
    if (add) cs[c] += (uint32_t)add;
}
}  // namespace

int map_append_incremental(Ctx* c, Map& mp, int n_new, bool* merged) {
    *merged = false;
    const char* off = std::getenv("ICP4R_NO_INCREMENTAL");
    if (off && off[0] == '1') return ICP4R_OK;
    if (!mp.built || n_new <= 0 || mp.grid.m <= 0 || mp.grid.m != mp.m_valid) return ICP4R_OK;
    // density drift: the cell size was chosen for the point count of the last full build
    if ((long long)mp.m_valid + n_new > (long long)mp.valid_at_build * 3 / 2 + 4096) return ICP4R_OK;
    Trace tr(c->stream);
    const GridDesc g = mp.grid;
    CKS(reserve_grow(c, mp.ik_a, (size_t)n_new * 4));
    CKS(reserve_grow(c, mp.ik_b, (size_t)n_new * 4));
    CKS(reserve_grow(c, mp.iv_a, (size_t)n_new * 4));
    CKS(reserve_grow(c, mp.iv_b, (size_t)n_new * 4));
    CKS(reserve(c, c->d_scratch, 4096));
    int* d_info = c->d_scratch.as<int>();
    inc_info_init<<<1, 32, 0, c->stream>>>(d_info);
    inc_key_kernel<<<(n_new + 255) / 256, 256, 0, c->stream>>>(mp.pts.as<float4>(), mp.m, n_new, g, mp.ik_a.as<uint32_t>(),
                                                               mp.iv_a.as<uint32_t>(), d_info);
    c->launches += 2;
    int info[8];
    CK(cudaMemcpyAsync(info, d_info, sizeof(info), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (info[0] != 0) {  // a point outside the grid: rebuild with slack so that the following batches fit
        mp.padded = true;
        return ICP4R_OK;
    }
    const int nf = info[1];
    tr.mark("inc keys");
    if (nf > 0) {
        int bits = 1;
        while ((1ll << bits) <= (long long)g.ncells) ++bits;
        uint32_t *ks, *vs;
        if (n_new <= SB_MAX) {
            ks = mp.ik_b.as<uint32_t>();
            vs = mp.iv_b.as<uint32_t>();
            static const bool bitonic = []() { const char* e = std::getenv("ICP4R_SMALL_SORT_BITONIC"); return e && e[0] == '1'; }();
            if (bitonic) small_sort_kernel<<<1, 1024, 0, c->stream>>>(mp.ik_a.as<uint32_t>(), mp.iv_a.as<uint32_t>(), n_new, ks, vs);
            else rank_sort_kernel<<<(n_new + RK_THREADS - 1) / RK_THREADS, RK_THREADS, 0, c->stream>>>(mp.ik_a.as<uint32_t>(), mp.iv_a.as<uint32_t>(), n_new, ks, vs);
            c->launches += 1;
        } else {
            CKS(radix_sort_pairs(c, mp.ik_a.as<uint32_t>(), mp.ik_b.as<uint32_t>(), mp.iv_a.as<uint32_t>(), mp.iv_b.as<uint32_t>(), n_new, bits,
                                 c->d_scratch, &ks, &vs));
        }
        tr.mark("inc sort");
        const int m_old = g.m, m_new = m_old + nf;
        if (mp.sorted_alt.cap < (size_t)m_new * sizeof(float4)) CKS(reserve(c, mp.sorted_alt, ((size_t)m_new + (size_t)m_new / 4) * sizeof(float4)));
        float4* s_new = mp.sorted_alt.as<float4>();
        const int ob = (m_old + INC_THREADS - 1) / INC_THREADS, cb = (g.ncells + 1 + INC_THREADS - 1) / INC_THREADS;
        CKS(reserve_grow(c, mp.inc_bnd, ((size_t)(ob + 1) + (size_t)(cb + 1)) * sizeof(int)));
        int* bnd_old = mp.inc_bnd.as<int>();
        int* bnd_cell = bnd_old + (ob + 1);
        const int gb_old = (ob + 1 + 255) / 256, gb_cell = (cb + 1 + 255) / 256;
        inc_bounds_kernel<<<gb_old + gb_cell, 256, 0, c->stream>>>(g.sorted, m_old, g, ks, nf, ob, bnd_old, gb_old, cb, bnd_cell);
        inc_merge_old_kernel<<<ob, INC_THREADS, 0, c->stream>>>(g.sorted, m_old, g, ks, bnd_old, s_new);
        inc_merge_new_kernel<<<(nf + 255) / 256, 256, 0, c->stream>>>(mp.pts.as<float4>(), ks, vs, nf, g.cell_start, s_new, g, mp.coarse.as<uint32_t>());
        inc_cell_kernel<<<cb, INC_THREADS, 0, c->stream>>>(mp.cell_start.as<uint32_t>(), g.ncells, ks, bnd_cell);
        c->launches += 4;
        CK(cudaGetLastError());
        tr.mark("inc merge");
        std::swap(mp.sorted, mp.sorted_alt);
        mp.grid.sorted = mp.sorted.as<float4>();
        mp.grid.m = m_new;
        mp.m_valid = m_new;
        for (int a = 0; a < 3; ++a) {
            mp.bb_min[a] = std::min(mp.bb_min[a], ord2f(info[2 + a]));
            mp.bb_max[a] = std::max(mp.bb_max[a], ord2f(info[5 + a]));
        }
        const float Lnew = std::max(std::max(std::fabs(ord2f(info[2])), std::fabs(ord2f(info[5]))),
                                    std::max(std::max(std::fabs(ord2f(info[3])), std::fabs(ord2f(info[6]))),
                                             std::max(std::fabs(ord2f(info[4])), std::fabs(ord2f(info[7])))));
        mp.grid.margin = std::max(mp.grid.margin, Lnew * 9.5367431640625e-7f);
    }
    mp.normals_k = 0;
    *merged = true;
    return ICP4R_OK;
}

// ------------------------------------------------------------------------------------------------ gate
void gate_params(double max_dist, float* gate_f, float* gate_r) {
    if (!(max_dist > 0.0) || std::isinf(max_dist)) {
        *gate_f = INFINITY;
        *gate_r = INFINITY;
        return;
    }
    const double g2 = max_dist * max_dist;  // ikd_Tree.cpp:880
    float f = (float)g2;
    if ((double)f > g2) f = std::nextafterf(f, 0.0f);  // largest float with (double)f <= g2
    *gate_f = f;
    float r = (float)(max_dist * (1.0 + 1e-6));
    if ((double)r < max_dist * (1.0 + 1e-6)) r = std::nextafterf(r, INFINITY);
    *gate_r = r;
}

// ------------------------------------------------------------------------------------------------ grid kNN kernel
template <int K>
__global__ void __launch_bounds__(256, (K <= 5 ? 4 : 2)) grid_knn_kernel(GridDesc g, const float4* __restrict__ q, int nq, int k, float gate_f,
                                                       float gate_r, int32_t* __restrict__ idx, float* __restrict__ d2,
                                                       int32_t* __restrict__ found) {
    __shared__ WarpSegs segs[8];
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = warp; i < nq; i += nwarps) {
#ifdef ICP4R_KNN_TIMING
        const long long t_begin = clock64();
#endif
        const float4 p = __ldg(q + i);
        const uint64_t mine = warp_grid_knn<K>(g, seg_addr(&segs[threadIdx.x >> 5]), p.x, p.y, p.z, gate_f, gate_r, lane);
        const bool have = (lane < k) && (mine != KEY_EMPTY);
        if (lane < k) {
            idx[(size_t)i * k + lane] = have ? key_idx(mine) : -1;
            d2[(size_t)i * k + lane] = have ? key_d2(mine) : INFINITY;
        }
        const unsigned b = __ballot_sync(FULL, have);
        if (found && lane == 0) found[i] = __popc(b);
#ifdef ICP4R_KNN_TIMING
        if (found && lane == 0) found[i] = (int)(clock64() - t_begin);  // dev probe: cycles spent on this query
#endif
    }
}

template <int K>
static int launch_grid_knn(Ctx* c, const GridDesc& g, const float4* q, int nq, int k, float gate_f, float gate_r, int32_t* idx,
                           float* d2, int32_t* found) {
    const int wpb = 8;
    int blocks = (nq + wpb - 1) / wpb;
    blocks = std::min(blocks, c->sm_count * 8);
    grid_knn_kernel<K><<<blocks, wpb * 32, 0, c->stream>>>(g, q, nq, k, gate_f, gate_r, idx, d2, found);
    c->launches += 1;
    CK(cudaGetLastError());
    return ICP4R_OK;
}

int grid_knn(Ctx* c, const Map& mp, const float4* q, int nq, int k, double max_dist, int32_t* idx, float* d2, int32_t* found) {
    if (nq <= 0) return ICP4R_OK;
    float gf, gr;
    gate_params(max_dist, &gf, &gr);
    const GridDesc& g = mp.grid;
    if (k == 1) return launch_grid_knn<1>(c, g, q, nq, k, gf, gr, idx, d2, found);
    if (k == 2) return launch_grid_knn<2>(c, g, q, nq, k, gf, gr, idx, d2, found);
    if (k <= 5) return launch_grid_knn<5>(c, g, q, nq, k, gf, gr, idx, d2, found);
    if (k <= 8) return launch_grid_knn<8>(c, g, q, nq, k, gf, gr, idx, d2, found);
    return launch_grid_knn<16>(c, g, q, nq, k, gf, gr, idx, d2, found);
}

// ------------------------------------------------------------------------------------------------ exhaustive kNN
// Thread per query, target tiles staged in shared memory (every thread reads the same candidate: broadcast),
// the target range split across blockIdx.y so small query counts still fill the GPU; a second kernel merges.
constexpr int BF_THREADS = 128;
constexpr int BF_TILE = 1024;

template <int K>
__global__ void __launch_bounds__(BF_THREADS) brute_knn_kernel(const float4* __restrict__ sorted, int m, const float4* __restrict__ q,
                                                               int nq, float gate_f, int splits, uint64_t* __restrict__ part) {
    __shared__ float4 tile[BF_TILE];
    const int qi = blockIdx.x * BF_THREADS + threadIdx.x;
    const int sp = blockIdx.y;
    const int chunk = (m + splits - 1) / splits;
    const int lo = sp * chunk, hi = min(m, lo + chunk);
    float4 p = make_float4(0, 0, 0, 0);
    if (qi < nq) p = q[qi];
    TopK<K> list;
    list.clear();
    for (int base = lo; base < hi; base += BF_TILE) {
        const int cnt = min(BF_TILE, hi - base);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += BF_THREADS) tile[t] = sorted[base + t];
        __syncthreads();
        for (int t = 0; t < cnt; ++t) {
            const float4 cpt = tile[t];
            const float d = dist2_exact(p.x, p.y, p.z, cpt.x, cpt.y, cpt.z);
            if (d <= gate_f) list.insert(pack_key(d, __float_as_int(cpt.w)));
        }
    }
    if (qi < nq) {
#pragma unroll
        for (int j = 0; j < K; ++j) part[((size_t)qi * splits + sp) * K + j] = list.key[j];
    }
}

template <int K>
__global__ void __launch_bounds__(128) brute_merge_kernel(const uint64_t* __restrict__ part, int nq, int splits, int k,
                                                          int32_t* __restrict__ idx, float* __restrict__ d2,
                                                          int32_t* __restrict__ found) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    TopK<K> list;
    list.clear();
    for (int s = 0; s < splits; ++s)
#pragma unroll
        for (int j = 0; j < K; ++j) list.insert(part[((size_t)qi * splits + s) * K + j]);
    int f = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (j < k) {
            const bool have = list.key[j] != KEY_EMPTY;
            idx[(size_t)qi * k + j] = have ? key_idx(list.key[j]) : -1;
            d2[(size_t)qi * k + j] = have ? key_d2(list.key[j]) : INFINITY;
            f += have;
        }
    }
    if (found) found[qi] = f;
}

template <int K>
static int launch_brute(Ctx* c, const float4* sorted, int m, const float4* q, int nq, int k, float gate_f, int32_t* idx, float* d2,
                        int32_t* found) {
    const int qblocks = (nq + BF_THREADS - 1) / BF_THREADS;
    int splits = std::max(1, std::min((c->sm_count * 4 + qblocks - 1) / qblocks, (m + BF_TILE - 1) / BF_TILE));
    splits = std::min(splits, 65535);
    // own scratch: d_partials is baked into captured registration loops and must never move
    CKS(reserve_grow(c, c->bf_part, (size_t)nq * splits * K * sizeof(uint64_t)));
    uint64_t* part = c->bf_part.as<uint64_t>();
    brute_knn_kernel<K><<<dim3(qblocks, splits), BF_THREADS, 0, c->stream>>>(sorted, m, q, nq, gate_f, splits, part);
    brute_merge_kernel<K><<<(nq + 127) / 128, 128, 0, c->stream>>>(part, nq, splits, k, idx, d2, found);
    c->launches += 2;
    CK(cudaGetLastError());
    return ICP4R_OK;
}

// exhaustive kNN of q over an arbitrary cloud whose w lanes carry the indices to report
int brute_knn_cloud(Ctx* c, const float4* cloud_xyzi, int m, const float4* q, int nq, int k, double max_dist, int32_t* idx, float* d2,
                    int32_t* found) {
    if (nq <= 0) return ICP4R_OK;
    float gf, gr;
    gate_params(max_dist, &gf, &gr);
    if (k == 1) return launch_brute<1>(c, cloud_xyzi, m, q, nq, k, gf, idx, d2, found);
    if (k == 2) return launch_brute<2>(c, cloud_xyzi, m, q, nq, k, gf, idx, d2, found);
    if (k <= 5) return launch_brute<5>(c, cloud_xyzi, m, q, nq, k, gf, idx, d2, found);
    if (k <= 8) return launch_brute<8>(c, cloud_xyzi, m, q, nq, k, gf, idx, d2, found);
    return launch_brute<16>(c, cloud_xyzi, m, q, nq, k, gf, idx, d2, found);
}

int brute_knn(Ctx* c, const Map& mp, const float4* q, int nq, int k, double max_dist, int32_t* idx, float* d2, int32_t* found) {
    return brute_knn_cloud(c, mp.grid.sorted, mp.grid.m, q, nq, k, max_dist, idx, d2, found);
}

// ------------------------------------------------------------------------------------------------ transform
struct Pose12 {
    double T[12];
};
// pointAssociateToMap (/root/reference/src/radar_odometry.cpp:137-145): p' = R p + t in double, stored as float
__global__ void __launch_bounds__(256) transform_kernel(Pose12 P, const float4* __restrict__ in, int n, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    double pw[3];
    xform_point(P.T, p.x, p.y, p.z, pw);
    out[i] = make_float4((float)pw[0], (float)pw[1], (float)pw[2], p.w);
}

int transform_points(Ctx* c, const double* T_host, const float4* d_in, int n, float4* d_out) {
    Pose12 P;
    for (int i = 0; i < 12; ++i) P.T[i] = T_host[i];
    transform_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(P, d_in, n, d_out);
    c->launches += 1;
    CK(cudaGetLastError());
    return ICP4R_OK;
}

}  // namespace icp4r
