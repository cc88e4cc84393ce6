// ctx.h — host-side state behind an icp4r_handle and the internal entry points of each translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/icp4r.h"

namespace icp4r {

// ---- error plumbing -------------------------------------------------------------------------------
struct Ctx;
int fail(Ctx* c, int code, const char* fmt, ...);

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return icp4r::fail(c, ICP4R_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,       \
                               cudaGetErrorString(e__));                                              \
    } while (0)
#define CKS(call)                       \
    do {                                \
        int s__ = (call);               \
        if (s__ != ICP4R_OK) return s__; \
    } while (0)

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <typename T>
    T* as() const { return static_cast<T*>(p); }
};
int reserve(Ctx* c, DevBuf& b, size_t bytes);
int reserve_grow(Ctx* c, DevBuf& b, size_t bytes);
void release(DevBuf& b);

// ---- uniform voxel grid over a point set (replaces the k-d tree) ----------------------------------
// Cells are keyed row-major with x fastest: key = (cz * ny + cy) * nx + cx, so the cells of one x-row are
// contiguous in the sorted array and a (2R+1)^3 neighbourhood is (2R+1)^2 contiguous ranges.
struct GridDesc {
    float ox, oy, oz;     // origin (lower corner)
    float cell, inv_cell;
    int nx, ny, nz;
    int ncells;
    int m;                // points in the sorted array
    float margin;         // absolute slack used by the termination bound (covers float rounding)
    const float4* sorted; // x, y, z, bit-cast original index
    const uint32_t* cell_start;  // [ncells + 1]
    const uint32_t* coarse;      // points per block of 8 x 8 x 8 cells, dims ceil(n / 8) (bound for far queries)
};

struct Map {
    DevBuf pts;         // float4 [cap] insertion order (x, y, z, intensity)
    DevBuf valid;       // uint8  [cap]
    DevBuf userdel;     // uint8  [cap]: removed by Delete_Points / Delete_Point_Boxes (Add_Point_Boxes can bring it back)
    DevBuf sorted;      // float4 [m_sorted]
    DevBuf cell_start;  // uint32 [ncells + 1]
    DevBuf coarse;      // uint32 [ceil(nx/8) * ceil(ny/8) * ceil(nz/8)]
    DevBuf keys_a, keys_b, vals_a, vals_b;  // radix sort ping-pong
    int m = 0;          // points ever offered (index space)
    int m_valid = 0;
    bool built = false;
    float user_cell = 0.f;
    float hint_cell = 0.f;  // refined cell size of the previous build of this map (Add_Points starts from it)
    float ds_voxel = 0.2f;  // KD_TREE default downsample_size, ikd_Tree.h:196
    GridDesc grid{};
    float bb_min[3] = {0, 0, 0}, bb_max[3] = {0, 0, 0};
    DevBuf normals;     // double[3 * m]: GICP surface normals by insertion index (valid while normals_k != 0)
    int normals_k = 0;  // neighbours they were estimated from; reset to 0 whenever the grid is rebuilt
    // incremental Add_Points (no down-sampling): new points are merged into the sorted array instead of re-sorting
    DevBuf sorted_alt;               // the other half of the sorted ping-pong
    DevBuf ik_a, ik_b, iv_a, iv_b;   // keys / indices of the batch being merged
    DevBuf inc_bnd;                  // per-block search bounds of the merge passes
    int valid_at_build = 0;          // valid points at the last full build (density drift -> full rebuild)
    bool quick_build = false;        // skip the occupancy-based cell refinement (short-lived clouds: GICP source)
    bool no_bucket = false;          // the bucket build overflowed on this map once: radix path from then on (reset by Build)
    bool padded = false;             // grid built with slack around the bounding box (set once an append fell outside)
};

// ---- peer-memory exchange for slab-sharded registration (one process per GPU, NVLink P2P) --------------------
constexpr int XCH_MAXW = 8;
struct Xch {  // lives on every rank; peers write their partial sums straight into it
    // [parity][writer rank][2 words per accumulator]: every 8-byte word carries 32 bits of a double and the 32-bit epoch
    // it belongs to, so a word is valid the moment it is seen with the expected epoch — no flag, no fence
    unsigned long long ll[2][XCH_MAXW][2 * ICP4R_ACC_LEN];
    unsigned long long seq;  // exchanges this rank has completed (identical on all ranks)
};
struct XchTable {  // device-resident view of the communicator
    Xch* peer[XCH_MAXW];  // peer[r] = rank r's Xch mapped into this process (own rank: the local buffer)
    int rank, world;
};

struct GicpCorr {  // per source point, written by the linearisation, read by the LM error passes
    int idx;       // target index or -1
    int pad;
    double li[6];  // L^-1 (lower triangle, row-major) with (C_B + R C_A R^T) = L L^T
};

// device-resident registration state (one per handle)
struct RegState {
    double T[16];
    double acc[ICP4R_ACC_LEN];
    double mse_prev;
    double last_cost;
    double lm_lambda;    // GICP Levenberg-Marquardt damping (< 0: not initialised)
    int lm_last_conv;
    int pad2;
    double fit_sum;
    int fit_cnt;
    int xch_timeout;     // set when a peer never published (fused sharded flavour)
    int done;
    int converged;
    int iterations;
    int n_corr;
    unsigned ticket;
    unsigned ticket_fit;
    int loop_epoch;      // persistent loop: iterations completed (published by the block that solved, see reg_loop_kernel)
    // the last pose increment D as the displacement map q -> q - D^-1 q = (I - R^T) q + R^T t (row-major 3x4) and the
    // iteration whose solve produced it (-1: none yet): a source point moved by |dA q| between that pass and the next
    float dA[12];
    int last_pass;
    int pad3;
};

struct NbState {   // per source point, next to nb_prev
    float lb;      // every map point other than the k remembered neighbours was at least this far away (metres) ...
    int tag;       // ... at the pass with this tag (call epoch * 4096 + iteration)
};

struct RegParams {  // kernel parameters that change per call; lives in device memory so graphs stay valid
    const float4* src;
    int n;
    int residual;
    int k;
    int max_iterations;
    int early_exit;
    float gate_f;        // largest float whose double is <= max_dist^2 (+inf when ungated)
    float gate_r;        // search radius in metres incl. rounding slack (+inf when ungated)
    double rot_eps, trans_eps, mse_abs_eps, plane_thresh;
    double interp_s;     // P2LINE / P2PLANE_3PT: interpolation ratio of the functors (1 = full pose)
    double* dump_pose;
    double* dump_acc;
    int32_t* dump_idx;
    // sharded registration
    int shard_axis;      // -1: not sharded
    float slab_lo, slab_hi;
    // GICP
    const double* src_normals;
    const double* tgt_normals;
    const float4* tgt_pts;  // target points by insertion index
    GicpCorr* corr;
    // fused cross-rank sum over peer memory (NULL: single rank or the NCCL flavour)
    const XchTable* xt;
    // the map's buffers: read from here (not from the by-value GridDesc) so that captured loops survive Add_Points
    const float4* map_sorted;
    const uint32_t* map_cell_start;
    const uint32_t* map_coarse;
    const float4* map_pts;
    int map_m;
    // neighbours found at the previous iteration, K per source point (NULL: hints switched off)
    int32_t* nb_prev;
    // work counters (NULL: off), see icp4r_set_stats
    unsigned long long* stats;
    // per-point bound that lets a point keep its neighbours without a search (NULL: off), and this call's epoch
    NbState* nb_state;
    int epoch;
    int pad_epoch;
    float slack_a, slack_b;  // proof flavour: the search ball is widened by slack_a x (last displacement) unless that exceeds slack_b x r
};

struct GraphKey {
    int kind, k, blocks, iters, flags;
    bool operator<(const GraphKey& o) const {
        if (kind != o.kind) return kind < o.kind;
        if (k != o.k) return k < o.k;
        if (blocks != o.blocks) return blocks < o.blocks;
        if (iters != o.iters) return iters < o.iters;
        return flags < o.flags;
    }
};

struct Ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // host->device copies of large batched calls run ahead of the compute stream
    DevBuf d_pairctr;                    // pair counters of the batched kernel's dynamic scheduling (one per launch in flight)
    unsigned pairctr_slot = 0;
    cudaStream_t aux_stream = nullptr;   // second compute stream of those calls: consecutive chunk kernels overlap their tails
    cudaEvent_t copy_events[36] = {};   // [chunk] copy done, [32] start, [33] aux stream done
    std::string err;
    int64_t launches = 0;

    Map map;      // the handle's persistent map
    Map tmp;      // transient target of icp4r_register
    Map srcmap;   // transient index over the SOURCE cloud (GICP source covariances)
    DevBuf d_src_normals, d_gicp_corr;

    DevBuf d_src, d_q, d_idx, d_d2, d_found, d_scratch, d_partials, d_state, d_params, d_T, d_res;
    // layout of the caller's point-cloud INPUT rows (icp4r_set_point_layout): byte stride and byte offset of w (< 0: none)
    int pt_stride = 16, pt_woff = 12;
    DevBuf d_raw;  // raw strided rows of a host cloud on their way to the device repack
    DevBuf d_dump_pose, d_dump_acc, d_dump_idx;
    DevBuf b_src, b_tgt, b_soff, b_toff, b_T, b_res;  // batched registration
    DevBuf bm_params, bm_state, bm_T0, bm_res, bm_partials;  // batched scans against the map
    void* h_pinned = nullptr;  // small pinned staging block
    size_t h_pinned_cap = 0;

    std::map<GraphKey, cudaGraphExec_t> graphs;
    std::map<GraphKey, int64_t> graph_launch_counts;
    int64_t graph_launches = 0;
    bool no_graph_sharded = false;
    const GridDesc* graph_grid_owner = nullptr;  // graphs bake the grid by value: identity of what they baked
    GridDesc graph_grid_copy{};
    bool use_graph = true;
    bool profiling = false;
    DevBuf d_stats;  // 8 x uint64 work counters, allocated by icp4r_set_stats
    bool stats = false;
    std::vector<cudaEvent_t> prof_events;
    std::vector<float> prof_ms;

    // sharding (NCCL loaded lazily with dlopen)
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
    DevBuf d_xch, d_xt, d_nbprev, d_nbstate;
    int reg_epoch = 0;  // bumped per registration call: NbState entries of earlier calls never match
    DevBuf bf_part;                           // exhaustive k-NN: per-split partial lists
    DevBuf gs_pts, gs_idx, gs_d2, gs_found;  // GICP: exhaustive k-NN scratch for the scan's own normals
    DevBuf vg_keys, vg_vals, vg_sort, vg_tiles, vg_out;  // voxel-grid centroid filter
    bool batch_reproducible = false;  // ICP4R_BATCH_REPRODUCIBLE=1 (see register_batch.cu)
    float submap_cell = 0.f;  // icp4r_register_submap: occupancy-refined cell size of the last refined sub-map (0: none yet)
    int submap_m = 0, submap_uses = 0;
    float slack_a = 4.0f, slack_b = 0.5f;  // ICP4R_LB_SLACK_A / ICP4R_LB_SLACK_B (A/B measurements)
    bool coop_ok = true;      // cleared when a cooperative launch was refused
    bool use_persist = false;  // ICP4R_PERSIST=1: single-scan loops as ONE cooperative launch (reg_loop_kernel) instead of a graph of per-iteration launches
    bool use_lb = true;     // ICP4R_NO_LB=1 turns the keep-the-neighbours-without-a-search proof off (A/B measurements)
    bool use_hints = true;  // ICP4R_NO_HINTS=1 turns the previous-iteration search bound off (A/B measurements)            // local exchange buffer and the peer table
    void* xch_peers[XCH_MAXW] = {nullptr};  // peer mappings opened with cudaIpcOpenMemHandle
    bool xch_ready = false;
};

inline int64_t& launches(Ctx* c) { return c->launches; }

// ---- internal entry points --------------------------------------------------------------------------
// radix_sort.cu: stable LSB radix sort of (key, value) pairs, `bits` significant key bits. Result in
// keys_out/vals_out (which alias one of the two ping-pong pairs).
int radix_sort_pairs(Ctx* c, uint32_t* keys_a, uint32_t* keys_b, uint32_t* vals_a, uint32_t* vals_b, int n, int bits,
                     DevBuf& scratch, uint32_t** keys_out, uint32_t** vals_out);

// grid.cu
int map_reserve(Ctx* c, Map& mp, int cap);
int map_rebuild_grid(Ctx* c, Map& mp);  // (re)sort all valid points into the grid
int map_append_incremental(Ctx* c, Map& mp, int n_new, bool* merged);  // merge pts[m, m+n_new) into the grid if it fits
int grid_knn(Ctx* c, const Map& mp, const float4* q, int nq, int k, double max_dist, int32_t* idx, float* d2,
             int32_t* found);
int brute_knn_cloud(Ctx* c, const float4* cloud_xyzi, int m, const float4* q, int nq, int k, double max_dist, int32_t* idx, float* d2,
                    int32_t* found);
int gicp_normals_small(Ctx* c, const float4* d_pts, int n, int k, DevBuf& normals);
int gicp_normals_small_to(Ctx* c, const float4* d_pts, int n, int k, double* out);  // same, into a caller-sized device array  // own-cloud k-NN normals without a grid
int brute_knn(Ctx* c, const Map& mp, const float4* q, int nq, int k, double max_dist, int32_t* idx, float* d2,
              int32_t* found);
void gate_params(double max_dist, float* gate_f, float* gate_r);

// map_ops.cu
int map_downsample_add(Ctx* c, Map& mp, int n, int* n_replaced_host, bool force_sequential);
int voxel_grid(Ctx* c, const float4* d_pts, const uint8_t* d_valid, int n, float leaf, float4* d_out, int cap, int* n_out_host);
int map_region_search(Ctx* c, const Map& mp, int kind, const float a[3], const float b[3], int32_t* d_out, int cap, int* n_out_host);
int map_box_flags(Ctx* c, Map& mp, const float* d_boxes6, int nb, bool revive, int* n_changed_host);
int map_delete_points(Ctx* c, Map& mp, const float4* d_req, int n, int* n_deleted_host);
int map_sector(Ctx* c, const Map& mp, const float centre[3], float radius, float heading, int32_t* d_out, int cap, int* n_out_host);

// register_map.cu
int register_against_map(Ctx* c, Map& mp, const float4* d_src, int n, const icp4r_opts* o, int shard_axis,
                         float slab_lo, float slab_hi, double* T_out_host, icp4r_result* res_host,
                         const icp4r_dump* dump_dev);

int accumulate_slab(Ctx* c, Map& mp, const float4* d_src, int n, const icp4r_opts* o, const double* T_host, int shard_axis, float slab_lo,
                    float slab_hi, double* acc_out_host);

int register_scans_against_map(Ctx* c, Map& mp, const float4* d_src, const int32_t* off_host, int nscan, const icp4r_opts* o,
                               const double* T0s_host, double* T_out_host, icp4r_result* res_host);

// register_batch.cu
bool register_batch_fits(int max_n, int max_m);  // does a pair of these sizes fit the shared-memory resident kernel?
int register_batch(Ctx* c, const float4* d_src, const int32_t* d_soff, const float4* d_tgt, const int32_t* d_toff,
                   int n_pairs, int max_n, int max_m, const icp4r_opts* o, double* d_T, icp4r_result* d_res);

// gicp.cu
int gicp_normals(Ctx* c, Map& mp, int k);  // fills mp.normals for every valid point of mp (cached per k)
int gicp_lm_step(Ctx* c, const RegParams* d_prm, RegState* d_st, int iter, int nscan = 1);  // one block per scan

// doppler.cu
int doppler_filter(Ctx* c, const float* d_rec, int n, int iterations, uint64_t seed, double sigma, double split, uint8_t* d_mask,
                   void* out_host, bool sync_now = true);

// shard.cu
int shard_allreduce(Ctx* c, double* d_buf, int count);

}  // namespace icp4r

struct icp4r_ctx : icp4r::Ctx {};
