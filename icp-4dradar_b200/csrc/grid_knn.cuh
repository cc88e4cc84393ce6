// grid_knn.cuh — exact k-nearest-neighbour search of one query by one warp over the voxel grid.
//
// Replaces KD_TREE::Search (/root/reference/third_party/ikd-Tree/ikd_Tree.cpp:877-1021) with the same
// result set: the k valid points of smallest float d2 = (dx*dx+dy*dy)+dz*dz with (double)d2 <= max_dist^2,
// ascending; the product's tie rule (lowest index first) replaces the reference's traversal order.
//
// Search = cube-shell expansion around the query's cell.  After the cube of radius R has been scanned,
// every unscanned point lies beyond one of the cube faces that still has cells of interest behind it, so
// its distance is at least `bound` = the smallest query-to-face distance (minus a rounding margin).  The
// search stops as soon as the current k-th best d2 is strictly below bound^2 (nothing unscanned can enter,
// not even on a tie) or the cube covers the whole gate box.
//
// Memory access: the cells of one x-row are contiguous in the sorted array, so a shell is a few dozen
// contiguous ranges.  Each lane fetches the bounds of one row (all rows' cell-table loads in flight at once),
// a warp scan turns the range lengths into one virtual candidate index space kept in shared memory, and the
// lanes then stride over that space four candidates at a time — coalesced float4 loads with 4 independent
// requests in flight per lane instead of one dependent round trip per row.  Every lane keeps a private sorted
// top-K in registers; the 32 lists are merged with warp reductions.
#pragma once
#include "ctx.h"
#include "device_math.cuh"

namespace icp4r {

constexpr int GRID_RING_CAP = 8;
constexpr int KNN_SEGS = 64;  // two ranges per lane

struct WarpSegs {  // per-warp shared-memory scratch
    uint32_t start[KNN_SEGS];
    uint32_t pre[KNN_SEGS + 1];
};
// The scratch is addressed through its 32-bit shared-space address with explicit ld.shared / st.shared: handing a C++
// reference to shared memory through the (inlined) search functions made the compiler rebuild the shared-window base
// (S2R SR_CgaCtaId + LEA) and the warp index at every access — 7 % of the executed instructions of the fused kernel.
using SegAddr = uint32_t;
__device__ __forceinline__ SegAddr seg_addr(WarpSegs* sg) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(sg);
    uint32_t pinned;  // an opaque move: the compiler keeps the value in a register instead of re-deriving it at every use
    asm volatile("mov.u32 %0, %1;" : "=r"(pinned) : "r"(a));
    return pinned;
}
__device__ __forceinline__ uint32_t lds_u32(SegAddr a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_u32(SegAddr a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
constexpr uint32_t SEG_PRE = KNN_SEGS * 4;  // byte offset of pre[] inside WarpSegs

__device__ __forceinline__ int cell_of(float v, float o, float inv, int dim) {
    const int c = __float2int_rd(__fmul_rn(__fsub_rn(v, o), inv));  // floor; NaN -> 0, +-inf saturate
    return min(max(c, 0), dim - 1);
}

template <int K>
__device__ __forceinline__ void consider(float qx, float qy, float qz, const float4& c, float gate_f, uint64_t kth, TopK<K>& list) {
    const float d = dist2_exact(qx, qy, qz, c.x, c.y, c.z);
    if (d <= gate_f) {  // false for NaN
        const uint64_t key = pack_key(d, __float_as_int(c.w));
        if (key < kth) list.insert(key);
    }
}

// scan the ranges described by this lane's (s0,e0) and (s1,e1) together with the other 31 lanes' ranges
template <int K>
__device__ __forceinline__ void scan_segments(const float4* __restrict__ sorted, SegAddr sg, uint32_t s0, uint32_t e0, uint32_t s1,
                                              uint32_t e1, int lane, float qx, float qy, float qz, float gate_f, uint64_t kth,
                                              TopK<K>& list, unsigned* cand = nullptr) {
    const uint32_t l0 = e0 - s0, l1 = e1 - s1;
    uint32_t inc = l0 + l1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += y;
    }
    const uint32_t total = __shfl_sync(FULL, inc, 31);
    if (total == 0) return;
    if (cand) *cand += total;  // work counter (icp4r_set_stats); the pointer is a compile-time null elsewhere
    const uint32_t exc = inc - (l0 + l1);
    // compact the non-empty ranges (in lane order) so the per-lane walk below only steps over real ones
    const unsigned b0 = __ballot_sync(FULL, l0 > 0), b1 = __ballot_sync(FULL, l1 > 0);
    const unsigned lt = (1u << lane) - 1u;
    const int slot = __popc(b0 & lt) + __popc(b1 & lt);
    const int nseg = __popc(b0) + __popc(b1);
    __syncwarp();
    if (l0 > 0) {
        sts_u32(sg + 4u * slot, s0);
        sts_u32(sg + SEG_PRE + 4u * slot, exc);
    }
    if (l1 > 0) {
        sts_u32(sg + 4u * (slot + (l0 > 0)), s1);
        sts_u32(sg + SEG_PRE + 4u * (slot + (l0 > 0)), exc + l0);
    }
    if (lane == 0) sts_u32(sg + SEG_PRE + 4u * nseg, total);
    __syncwarp();
    int r = 0;
    for (uint32_t t0 = lane; t0 < total; t0 += 128) {
        float4 c[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t t = t0 + 32u * u;
            ok[u] = t < total;
            if (ok[u]) {
                while (t >= lds_u32(sg + SEG_PRE + 4u * (r + 1))) ++r;  // pre[nseg] == total > t terminates the walk
                c[u] = __ldg(sorted + (lds_u32(sg + 4u * r) + (t - lds_u32(sg + SEG_PRE + 4u * r))));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (ok[u]) consider<K>(qx, qy, qz, c[u], gate_f, kth, list);
    }
}

// Coarse occupancy: entry (X, Y, Z) counts the points of the 8 x 8 x 8 cells [8X, 8X+8) x ... ; dims ceil(n / 8).
constexpr int COARSE_SHIFT = 3;
constexpr int COARSE_MAX_R = 24;
// Squared distance within which at least K points are guaranteed to lie (-1: not found within COARSE_MAX_R coarse cells).
template <int K>
__device__ __forceinline__ float coarse_bound(const GridDesc& g, const uint32_t* __restrict__ coarse, int cx, int cy, int cz, float qx,
                                              float qy, float qz, float margin, int lane) {
    const int cnx = (g.nx + 7) >> COARSE_SHIFT, cny = (g.ny + 7) >> COARSE_SHIFT, cnz = (g.nz + 7) >> COARSE_SHIFT;
    const int X = cx >> COARSE_SHIFT, Y = cy >> COARSE_SHIFT, Z = cz >> COARSE_SHIFT;
    int Rc = 0;
    for (;;) {
        const int x0 = max(X - Rc, 0), x1 = min(X + Rc, cnx - 1), y0 = max(Y - Rc, 0), y1 = min(Y + Rc, cny - 1);
        const int z0 = max(Z - Rc, 0), z1 = min(Z + Rc, cnz - 1);
        const int sx = x1 - x0 + 1, sy = y1 - y0 + 1, sz = z1 - z0 + 1;
        const int ncube = sx * sy * sz;
        uint32_t part = 0;
        for (int t = lane; t < ncube; t += 32) {
            const int tz = t / (sx * sy), rem = t - tz * (sx * sy), ty = rem / sx, tx = rem - ty * sx;
            part += __ldg(coarse + ((size_t)(z0 + tz) * cny + (y0 + ty)) * cnx + (x0 + tx));
        }
        part = __reduce_add_sync(FULL, part);
        if (part >= (uint32_t)K) {
            const float cw = g.cell * (float)(1 << COARSE_SHIFT);
            const float lo[3] = {g.ox + (float)x0 * cw, g.oy + (float)y0 * cw, g.oz + (float)z0 * cw};
            const float hi[3] = {fminf(g.ox + (float)(x1 + 1) * cw, g.ox + (float)g.nx * g.cell),
                                 fminf(g.oy + (float)(y1 + 1) * cw, g.oy + (float)g.ny * g.cell),
                                 fminf(g.oz + (float)(z1 + 1) * cw, g.oz + (float)g.nz * g.cell)};
            const float q[3] = {qx, qy, qz};
            float d2 = 0.0f;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float b = fmaxf(fabsf(q[a] - lo[a]), fabsf(hi[a] - q[a])) + 2.0f * margin;
                d2 += b * b;
            }
            return d2 * 1.00001f;
        }
        if (sx == cnx && sy == cny && sz == cnz) return -1.0f;  // the whole map holds fewer than K points
        if (Rc >= COARSE_MAX_R) return -1.0f;
        Rc = Rc < 4 ? Rc + 1 : Rc + (Rc >> 1);  // 0 1 2 3 4 6 9 13 19 28
    }
}

// One pruned pass over the cells that the ball of squared radius `bound2` (clipped to the gate box) touches: the
// fallback for queries whose shell search hit the ring cap. Kept out of line — it is rare and must not cost the
// common path registers or instructions. `bound2` must bound the K-th distance (>= K valid points within it).
template <int K>
static __device__ __noinline__ uint64_t wide_ball_knn(const GridDesc& g, const float4* __restrict__ sorted, const uint32_t* __restrict__ cs,
                                                      SegAddr sg, float qx, float qy, float qz, float gate_f, float br, float bound2,
                                                      float margin, int lane) {
    const int cx = cell_of(qx, g.ox, g.inv_cell, g.nx), cy = cell_of(qy, g.oy, g.inv_cell, g.ny), cz = cell_of(qz, g.oz, g.inv_cell, g.nz);
    const int lox = cell_of(qx - br, g.ox, g.inv_cell, g.nx), hix = cell_of(qx + br, g.ox, g.inv_cell, g.nx);
    const int loy = cell_of(qy - br, g.oy, g.inv_cell, g.ny), hiy = cell_of(qy + br, g.oy, g.inv_cell, g.ny);
    const int loz = cell_of(qz - br, g.oz, g.inv_cell, g.nz), hiz = cell_of(qz + br, g.oz, g.inv_cell, g.nz);
    const int sy = hiy - loy + 1, sz = hiz - loz + 1, rows = sy * sz;
    const float kd = bound2 * 1.000001f;
    const uint64_t bound_key = ((uint64_t)__float_as_uint(bound2) << 32) | 0xFFFFFFFFull;
    TopK<K> list;
    list.clear();
    for (int rb = 0; rb < rows; rb += 32) {
        const int r = rb + lane;
        uint32_t s0 = 0, e0 = 0;
        if (r < rows) {
            const int rz = r / sy, y = loy + (r - rz * sy), z = loz + rz;
            const int dy = y - cy, dz = z - cz;
            float ddy = dy > 0 ? (g.oy + (float)y * g.cell) - qy : (dy < 0 ? qy - (g.oy + (float)(y + 1) * g.cell) : 0.0f);
            float ddz = dz > 0 ? (g.oz + (float)z * g.cell) - qz : (dz < 0 ? qz - (g.oz + (float)(z + 1) * g.cell) : 0.0f);
            ddy = fmaxf(ddy - margin, 0.0f);
            ddz = fmaxf(ddz - margin, 0.0f);
            const float dyz2 = (ddy * ddy + ddz * ddz) * 0.999999f;
            if (!(dyz2 > kd)) {
                const float xr = sqrtf(kd - dyz2) * 1.000001f + margin;
                const int xa = max(lox, cell_of(qx - xr, g.ox, g.inv_cell, g.nx)), xb = min(hix, cell_of(qx + xr, g.ox, g.inv_cell, g.nx));
                if (xa <= xb) {
                    const uint32_t rowbase = (uint32_t)(z * g.ny + y) * (uint32_t)g.nx;
                    s0 = __ldg(cs + rowbase + xa);
                    e0 = __ldg(cs + rowbase + xb + 1);
                }
            }
        }
        if (__any_sync(FULL, e0 > s0)) scan_segments<K>(sorted, sg, s0, e0, 0u, 0u, lane, qx, qy, qz, gate_f, bound_key, list);
    }
    return warp_merge_topk<K>(list, lane);
}

// Returns, in lane r < K, the r-th nearest neighbour's packed key (KEY_EMPTY if fewer exist).
// WIDE = false leaves the coarse-table fallback out (the fused iteration kernel: its register and instruction budget
// is tuned for the common path, and its queries are gated scan points near the map).
// `g` carries the geometry only; the sorted points, the cell table and their length are passed next to it so that a
// captured launch (whose GridDesc is baked in by value) can pick them up from device memory at run time.
//
// `hint` >= 0: the caller knows K distinct valid points with d2 <= hint (the neighbours found for this source point at
// the previous pose). The k-th distance can then not exceed `hint`, so ONE pass over the cells that intersect the ball
// of that radius (clipped to the gate box) sees every point that can be part of the answer: no own-cell probe, no
// shell-by-shell proof, one merge. The result is the same set, the hint only removes work.
template <int K, bool WIDE = true>
__device__ __forceinline__ uint64_t warp_grid_knn(const GridDesc& g, const float4* __restrict__ sorted, const uint32_t* __restrict__ cs,
                                                  const uint32_t* __restrict__ coarse, int m_sorted, SegAddr sg, float qx, float qy,
                                                  float qz, float gate_f, float gate_r, int lane, float hint = -1.0f,
                                                  unsigned* cand = nullptr, float* cover2 = nullptr) {
    // *cover2 (optional): every valid point with d2 below this value (and inside the gate) has been ranked by the search,
    // i.e. a point that is not among the returned ones is at least that far away (squared) or beyond the gate
    const float qmax = fmaxf(fabsf(qx), fmaxf(fabsf(qy), fabsf(qz)));
    const float margin = fmaxf(g.margin, 9.5367431640625e-7f * qmax);  // 2^-20 * magnitude

    const int cx = cell_of(qx, g.ox, g.inv_cell, g.nx);
    const int cy = cell_of(qy, g.oy, g.inv_cell, g.ny);
    const int cz = cell_of(qz, g.oz, g.inv_cell, g.nz);
    int lox = 0, hix = g.nx - 1, loy = 0, hiy = g.ny - 1, loz = 0, hiz = g.nz - 1;
    const float gr = gate_r < 3.0e38f ? gate_r + margin : 3.4e38f;
    bool hinted = (hint >= 0.0f) && (hint < 3.0e38f);
    // cells that can hold part of the answer: the gate box, or the (smaller) box around the hint ball
    float br = hinted ? fminf(gr, sqrtf(hint) * 1.000001f + margin) : gr;
    int rneed;
    for (;;) {
        if (br < 3.0e38f) {
            lox = cell_of(qx - br, g.ox, g.inv_cell, g.nx);
            hix = cell_of(qx + br, g.ox, g.inv_cell, g.nx);
            loy = cell_of(qy - br, g.oy, g.inv_cell, g.ny);
            hiy = cell_of(qy + br, g.oy, g.inv_cell, g.ny);
            loz = cell_of(qz - br, g.oz, g.inv_cell, g.nz);
            hiz = cell_of(qz + br, g.oz, g.inv_cell, g.nz);
        } else {
            lox = 0, hix = g.nx - 1, loy = 0, hiy = g.ny - 1, loz = 0, hiz = g.nz - 1;
        }
        rneed = max(max(max(cx - lox, hix - cx), max(cy - loy, hiy - cy)), max(cz - loz, hiz - cz));
        if (!hinted || rneed <= GRID_RING_CAP) break;
        hinted = false;  // a hint ball wider than the ring cap: search without it
        br = gr;
    }
    // candidates farther than the hint can not be part of the answer (an equal distance still can)
    const uint64_t hint_key = ((uint64_t)__float_as_uint(hint) << 32) | 0xFFFFFFFFull;

    TopK<K> list;
    list.clear();
    uint64_t mine = KEY_EMPTY, kth = KEY_EMPTY;
    if (cover2) *cover2 = 0.0f;  // unknown until one of the proven exits below sets it
    int prev = -1;  // radius already scanned
    bool done = false;
    // Dense neighbourhoods (walls): look at the query's own cell first. Its k-th distance then prunes the rows and
    // the x-extent of the next shells, so a 3x3x3 block that would hold hundreds of candidates shrinks to the few
    // cells that can still contain a closer point.
    int R0 = hinted ? rneed : min(1, rneed);
    if (rneed > 0 && !hinted) {
        const uint32_t ci = (uint32_t)(cz * g.ny + cy) * (uint32_t)g.nx + (uint32_t)cx;
        if (__ldg(cs + ci + 1) - __ldg(cs + ci) >= 2u * K) R0 = 0;
    }
    for (int R = R0; R <= rneed && R <= GRID_RING_CAP; ++R) {
        // rows (dy, dz) of the shell, 32 at a time: each lane fetches the cell ranges of one row
        const int side = 2 * R + 1, rows = side * side;
        const float inv_side = 1.0f / (float)side;  // r / side without an integer division (exact for these small r)
        const bool prune = hinted || kth != KEY_EMPTY;
        const float kd = (hinted ? hint : key_d2(kth)) * 1.000001f;  // inflated: a pruned cell can not even hold a tie
        for (int rb = 0; rb < rows; rb += 32) {
            const int r = rb + lane;
            uint32_t s0 = 0, e0 = 0, s1 = 0, e1 = 0;
            if (r < rows) {
                const int rz = (int)(((float)r + 0.5f) * inv_side);
                const int dy = r - rz * side - R, dz = rz - R;
                const int y = cy + dy, z = cz + dz;
                if (y >= loy && y <= hiy && z >= loz && z <= hiz) {
                    int xa = max(cx - R, lox), xb = min(cx + R, hix);
                    bool skip = false;
                    if (prune) {
                        // lower bound of the distance from the query to this row's y/z slab (0 inside the slab)
                        float ddy = dy > 0 ? (g.oy + (float)y * g.cell) - qy : (dy < 0 ? qy - (g.oy + (float)(y + 1) * g.cell) : 0.0f);
                        float ddz = dz > 0 ? (g.oz + (float)z * g.cell) - qz : (dz < 0 ? qz - (g.oz + (float)(z + 1) * g.cell) : 0.0f);
                        ddy = fmaxf(ddy - margin, 0.0f);
                        ddz = fmaxf(ddz - margin, 0.0f);
                        const float dyz2 = (ddy * ddy + ddz * ddz) * 0.999999f;
                        if (dyz2 > kd) {
                            skip = true;
                        } else {
                            const float xr = sqrtf(kd - dyz2) * 1.000001f + margin;
                            xa = max(xa, cell_of(qx - xr, g.ox, g.inv_cell, g.nx));
                            xb = min(xb, cell_of(qx + xr, g.ox, g.inv_cell, g.nx));
                        }
                    }
                    if (!skip) {
                        const uint32_t rowbase = (uint32_t)(z * g.ny + y) * (uint32_t)g.nx;
                        if (max(abs(dy), abs(dz)) > prev) {  // new row: whole x range
                            if (xa <= xb) {
                                s0 = __ldg(cs + rowbase + xa);
                                e0 = __ldg(cs + rowbase + xb + 1);
                            }
                        } else {  // row already scanned up to radius prev: only the two end caps
                            const int xl = min(cx - prev - 1, xb), xr2 = max(cx + prev + 1, xa);
                            if (xa <= xl) {
                                s0 = __ldg(cs + rowbase + xa);
                                e0 = __ldg(cs + rowbase + xl + 1);
                            }
                            if (xr2 <= xb) {
                                s1 = __ldg(cs + rowbase + xr2);
                                e1 = __ldg(cs + rowbase + xb + 1);
                            }
                        }
                    }
                }
            }
            scan_segments<K>(sorted, sg, s0, e0, s1, e1, lane, qx, qy, qz, gate_f, hinted ? hint_key : kth, list, cand);
        }
        prev = R;
        if (mine != KEY_EMPTY) list.insert(mine);  // carry the previous shells' winners (lanes < K)
        mine = warp_merge_topk<K>(list, lane);
        list.clear();
        kth = __shfl_sync(FULL, mine, K - 1);
        if (hinted) {  // the pass covered the whole hint box
            done = true;
            if (cover2) *cover2 = hint * 0.999999f;
            break;
        }

        float bound = 3.4e38f;
        if (cx - R > lox) bound = fminf(bound, qx - (g.ox + (float)(cx - R) * g.cell));
        if (cx + R < hix) bound = fminf(bound, (g.ox + (float)(cx + R + 1) * g.cell) - qx);
        if (cy - R > loy) bound = fminf(bound, qy - (g.oy + (float)(cy - R) * g.cell));
        if (cy + R < hiy) bound = fminf(bound, (g.oy + (float)(cy + R + 1) * g.cell) - qy);
        if (cz - R > loz) bound = fminf(bound, qz - (g.oz + (float)(cz - R) * g.cell));
        if (cz + R < hiz) bound = fminf(bound, (g.oz + (float)(cz + R + 1) * g.cell) - qz);
        if (bound > 3.0e38f) {
            done = true;  // cube covers the gate box
            if (cover2) *cover2 = INFINITY;
            break;
        }
        const float b = bound - margin;
        if (b > 0.0f && kth != KEY_EMPTY && key_d2(kth) < b * b * 0.99999905f) {
            done = true;
            if (cover2) *cover2 = b * b * 0.99999905f;
            break;
        }
    }
    if (WIDE && !done && rneed > 0 && prev < rneed && coarse != nullptr) {
        // Ring cap reached (a query far from the map's points, or a gate much wider than the cells). The coarse
        // occupancy table tells how far one has to go to be SURE of K points: that distance bounds the K-th distance
        // like the previous-iteration hint does, and one pruned pass over the ball settles the query. The pass may be
        // wide, but an empty row costs two loads — far less than scanning the whole map.
        const float cb = coarse_bound<K>(g, coarse, cx, cy, cz, qx, qy, qz, margin, lane);
        if (cb >= 0.0f) {
            const float brw = fminf(gr, sqrtf(cb) * 1.000001f + margin);
            // rows the pass would touch (clipped to the grid) against the points an exhaustive scan would read
            const float wy = (float)(cell_of(qy + brw, g.oy, g.inv_cell, g.ny) - cell_of(qy - brw, g.oy, g.inv_cell, g.ny) + 1);
            const float wz = (float)(cell_of(qz + brw, g.oz, g.inv_cell, g.nz) - cell_of(qz - brw, g.oz, g.inv_cell, g.nz) + 1);
            if (wy * wz * 3.0f < (float)m_sorted)
                return wide_ball_knn<K>(g, sorted, cs, sg, qx, qy, qz, gate_f, brw, fminf(cb, gate_f), margin, lane);
        }
    }
    if (!done && rneed > 0 && prev < rneed) {
        // no usable bound (tiny map, fewer than K points, or a ball that would cover most of the map): exhaustive scan
        list.clear();
        scan_segments<K>(sorted, sg, 0u, lane == 0 ? (uint32_t)m_sorted : 0u, 0u, 0u, lane, qx, qy, qz, gate_f, KEY_EMPTY, list, cand);
        mine = warp_merge_topk<K>(list, lane);
        if (cover2) *cover2 = INFINITY;
    }
    return mine;
}

}  // namespace icp4r

namespace icp4r {
template <int K>
__device__ __forceinline__ uint64_t warp_grid_knn(const GridDesc& g, SegAddr sg, float qx, float qy, float qz, float gate_f,
                                                  float gate_r, int lane) {
    return warp_grid_knn<K>(g, g.sorted, g.cell_start, g.coarse, g.m, sg, qx, qy, qz, gate_f, gate_r, lane);
}
}  // namespace icp4r
