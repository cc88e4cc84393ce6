// radix_sort.cu — hand-written stable LSB radix sort of (uint32 key, uint32 value) pairs, 8-bit digits.
//
// Used to order map points by voxel key (grid.cu).  Three kernels per digit pass:
//   rs_hist    per-tile digit histogram (shared-memory atomics)        reads 4 B / element
//   rs_scan    one block per digit: exclusive scan of that digit's counts across tiles
//   rs_scatter stable ranks inside the tile by warp match-any + per-warp running counts, then scatter
//                                                                       reads 8 B, writes 8 B / element
// A tile is 2048 consecutive elements (256 threads x 8); warp w of a block owns the 256 consecutive
// elements [w*256, w*256+256) of the tile and walks them in 8 rounds of 32 so that order (warp, round,
// lane) == index order, which is what makes the scatter stable.
#include "ctx.h"
#include "device_math.cuh"

namespace icp4r {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS) rs_hist(const uint32_t* __restrict__ keys, int n, int shift, int tiles,
                                                      uint32_t* __restrict__ counts) {
    __shared__ uint32_t h[256];
    const int tid = threadIdx.x, tile = blockIdx.x;
    h[tid] = 0;
    __syncthreads();
    const int base = tile * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const int g = base + i * RS_THREADS + tid;
        if (g < n) atomicAdd(&h[(keys[g] >> shift) & 255u], 1u);
    }
    __syncthreads();
    counts[(size_t)tid * tiles + tile] = h[tid];
}

// block d: exclusive scan over tiles of counts[d][*]; totals[d] = sum
__global__ void __launch_bounds__(RS_THREADS) rs_scan(uint32_t* __restrict__ counts, int tiles,
                                                      uint32_t* __restrict__ totals) {
    __shared__ uint32_t wsum[RS_THREADS / 32];
    __shared__ uint32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    uint32_t* row = counts + (size_t)blockIdx.x * tiles;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < tiles; base += RS_THREADS) {
        const int i = base + tid;
        const uint32_t v = i < tiles ? row[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        uint32_t woff = 0;
        for (int j = 0; j < w; ++j) woff += wsum[j];
        const uint32_t carry = carry_s;
        if (i < tiles) row[i] = carry + woff + x - v;
        __syncthreads();
        if (tid == RS_THREADS - 1) carry_s = carry + woff + x;
        __syncthreads();
    }
    if (tid == 0) totals[blockIdx.x] = carry_s;
}

__global__ void __launch_bounds__(RS_THREADS)
    rs_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
               uint32_t* __restrict__ vals_out, int n, int shift, int tiles, const uint32_t* __restrict__ counts,
               const uint32_t* __restrict__ totals) {
    __shared__ uint32_t whist[RS_THREADS / 32][256];
    __shared__ uint32_t dbase[256];
    __shared__ uint32_t wsum[RS_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, tile = blockIdx.x;

    // exclusive scan of the 256 digit totals (every block recomputes it: 256 values)
    {
        const uint32_t v = totals[tid];
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[w] = x;
#pragma unroll
        for (int j = 0; j < RS_THREADS / 32; ++j) whist[j][tid] = 0;
        __syncthreads();
        uint32_t woff = 0;
        for (int j = 0; j < w; ++j) woff += wsum[j];
        dbase[tid] = woff + x - v + counts[(size_t)tid * tiles + tile];
    }
    __syncthreads();

    uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
    const int wbase = tile * RS_TILE + w * (32 * RS_ITEMS);
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int g = wbase + r * 32 + lane;
        const bool ok = g < n;
        key[r] = ok ? keys_in[g] : 0xffffffffu;
        val[r] = ok ? vals_in[g] : 0u;
        const uint32_t dg = ok ? ((key[r] >> shift) & 255u) : 256u;
        const uint32_t mask = __match_any_sync(FULL, dg);
        uint32_t prior = 0;
        if (ok) prior = whist[w][dg];
        __syncwarp();
        if (ok && (mask & lt) == 0) whist[w][dg] = prior + __popc(mask);  // lowest lane of the group
        __syncwarp();
        rank[r] = prior + __popc(mask & lt);
    }
    __syncthreads();
    // Stage the tile in digit order through shared memory, then write it out: consecutive threads then hold
    // consecutive elements of the same digit and their global writes are contiguous runs instead of a 32-way scatter.
    __shared__ uint32_t toff[256];               // first tile-local slot of every digit
    __shared__ uint32_t stage_k[RS_TILE], stage_v[RS_TILE];
    {   // thread d: per-warp counts -> per-warp offsets inside digit d's run; the digit's tile count -> toff by a block scan
        uint32_t run = 0;
#pragma unroll
        for (int j = 0; j < RS_THREADS / 32; ++j) {
            const uint32_t t = whist[j][tid];
            whist[j][tid] = run;
            run += t;
        }
        uint32_t x = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        uint32_t woff = 0;
        for (int j = 0; j < w; ++j) woff += wsum[j];
        toff[tid] = woff + x - run;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int g = wbase + r * 32 + lane;
        if (g < n) {
            const uint32_t dg = (key[r] >> shift) & 255u;
            const uint32_t slot = toff[dg] + whist[w][dg] + rank[r];
            stage_k[slot] = key[r];
            stage_v[slot] = val[r];
        }
    }
    __syncthreads();
    const int tile_n = min(RS_TILE, n - tile * RS_TILE);
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const int i = r * RS_THREADS + tid;
        if (i < tile_n) {
            const uint32_t k = stage_k[i];
            const uint32_t dg = (k >> shift) & 255u;
            const uint32_t pos = dbase[dg] + ((uint32_t)i - toff[dg]);
            keys_out[pos] = k;
            vals_out[pos] = stage_v[i];
        }
    }
}

int radix_sort_pairs(Ctx* c, uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, int n, int bits,
                     DevBuf& scratch, uint32_t** keys_out, uint32_t** vals_out) {
    *keys_out = keys_a;
    *vals_out = vals_a;
    if (n <= 0) return ICP4R_OK;
    const int tiles = (n + RS_TILE - 1) / RS_TILE;
    CKS(reserve_grow(c, scratch, ((size_t)256 * tiles + 256) * sizeof(uint32_t)));  // callers sort growing maps frame after frame
    uint32_t* counts = scratch.as<uint32_t>();
    uint32_t* totals = counts + (size_t)256 * tiles;
    if (bits < 1) bits = 1;
    const int passes = (bits + 7) / 8;
    uint32_t *ki = keys_a, *ko = keys_b, *vi = vals_a, *vo = vals_b;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        rs_hist<<<tiles, RS_THREADS, 0, c->stream>>>(ki, n, shift, tiles, counts);
        rs_scan<<<256, RS_THREADS, 0, c->stream>>>(counts, tiles, totals);
        rs_scatter<<<tiles, RS_THREADS, 0, c->stream>>>(ki, vi, ko, vo, n, shift, tiles, counts, totals);
        c->launches += 3;
        uint32_t* t = ki;
        ki = ko;
        ko = t;
        t = vi;
        vi = vo;
        vo = t;
    }
    CK(cudaGetLastError());
    *keys_out = ki;
    *vals_out = vi;
    return ICP4R_OK;
}

}  // namespace icp4r
