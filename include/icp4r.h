/* icp4r.h — C ABI of libicp4r_cuda: the B200 (sm_100a) drop-in for the registration hot path of
 * invokermyself/ICP-4DRadar.
 *
 * The reference has no FFI of its own; its hot path sits behind two C++ call shapes:
 *   - the ikd-Tree shape   KD_TREE<PointType>::{Build, Add_Points, Nearest_Search, Sector_Search, ...}
 *                          (/root/reference/third_party/ikd-Tree/ikd_Tree.h:227-251, used at
 *                          /root/reference/src/radar_odometry.cpp:92,347-348,390,396), and
 *   - the PCL Registration shape  {setInputSource, setInputTarget, align, hasConverged, getFitnessScore,
 *                          getFinalTransformation} (/root/reference/src/iterative_closest_point.cpp:510-521,
 *                          /root/reference/src/radar_odometry.cpp:399-411).
 * The header-only C++ adapters in icp-4dradar_b200/adapters/icp4r/ present those two shapes on top of
 * the entry points below; INTEGRATION.md shows the two-line change at each reference call site.
 *
 * Conventions
 *   - every function returns an icp4r_status (0 = ok); nothing throws or calls back across the boundary;
 *     icp4r_last_error(h) describes the last failure on that handle.
 *   - one handle = one CUDA device + one stream + one caller thread at a time: a handle is stateful (staging buffers,
 *     the transient target of icp4r_register, the error text, captured launch graphs) and is NOT thread-safe; calls on
 *     one handle from several threads must be serialised by the caller. Distinct handles are independent and may be used
 *     concurrently (that is how multi-GPU batches run; the C++ adapters keep one handle per thread).
 *   - points are packed float rows  x, y, z, w  (w = intensity, carried not used); `mem` says whether the
 *     caller's pointers (inputs AND outputs of that call) are host (ICP4R_HOST) or device (ICP4R_DEVICE).
 *   - poses are row-major double[16]; the update convention is the left perturbation T <- exp(xi^) T with
 *     xi = (omega, v), p' = R p + t.
 *   - indices returned by the map are positions in insertion order (Build order, then every point ever
 *     offered to Add_Points, kept or not); missing neighbour slots hold idx = -1, d2 = +inf.
 *   - exact kNN: float d2 = (dx*dx + dy*dy) + dz*dz evaluated without FMA contraction
 *     (calc_dist, ikd_Tree.cpp:1427-1431), kept iff (double)d2 <= max_dist^2 (ikd_Tree.cpp:880,895),
 *     ordered by (d2, index) — ties go to the lowest index.
 *   - there is no CPU fallback: without a usable CUDA device icp4r_create fails.
 */
#ifndef ICP4R_H
#define ICP4R_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICP4R_VERSION_MAJOR 0
#define ICP4R_VERSION_MINOR 1
#define ICP4R_MAX_K 16
#define ICP4R_ACC_LEN 32

typedef struct icp4r_ctx* icp4r_handle;

typedef enum icp4r_status {
    ICP4R_OK = 0,
    ICP4R_ERR_INVALID = 1,     /* bad argument */
    ICP4R_ERR_CUDA = 2,        /* CUDA runtime error (text in icp4r_last_error) */
    ICP4R_ERR_NOMEM = 3,
    ICP4R_ERR_STATE = 4,       /* e.g. map query before Build */
    ICP4R_ERR_NCCL = 5,
    ICP4R_ERR_UNSUPPORTED = 6
} icp4r_status;

typedef enum icp4r_mem { ICP4R_HOST = 0, ICP4R_DEVICE = 1 } icp4r_mem;

/* residual kinds; definitions follow /root/reference/include/radarFactor.hpp */
typedef enum icp4r_residual {
    ICP4R_P2P_SVD = 0,     /* LidarDistanceFactor objective solved in closed form (Kabsch/Umeyama), what
                              pcl::IterativeClosestPoint does at iterative_closest_point.cpp:510-514     */
    ICP4R_P2P_GN = 1,      /* LidarDistanceFactor (radarFactor.hpp:140-171), Gauss-Newton 6x6           */
    ICP4R_P2PLANE_KNN = 2, /* LidarPlaneNormFactor (radarFactor.hpp:105-137), plane from the k neighbours */
    ICP4R_P2LINE = 3,      /* RadarEdgeFactor (radarFactor.hpp:11-54), line through the 2 nearest, s = opts.interp_s */
    ICP4R_P2PLANE_3PT = 5, /* LidarPlaneFactor (radarFactor.hpp:56-103), plane through the 3 nearest points, s = opts.interp_s */
    ICP4R_GICP = 4         /* fast_gicp cost (radar_odometry.cpp:399-405): k-NN plane-regularised covariances of both
                              clouds, 1-NN Mahalanobis residual, Levenberg-Marquardt; opts.k = CorrespondenceRandomness */
} icp4r_residual;

typedef struct icp4r_opts {
    int32_t residual;        /* icp4r_residual */
    int32_t k;               /* neighbours for P2PLANE_KNN (3..ICP4R_MAX_K, default 5); ignored otherwise */
    int32_t max_iterations;  /* PCL default 10 */
    int32_t early_exit;      /* 0: run exactly max_iterations */
    double max_corr_dist;    /* <= 0 or +inf: ungated (PCL / fast_gicp defaults) */
    double rot_eps;          /* early exit (GN kinds): |omega| < rot_eps && |v| < trans_eps */
    double trans_eps;
    double mse_abs_eps;      /* early exit (P2P_SVD): |mse_i - mse_{i-1}| < mse_abs_eps (PCL 1e-12) */
    double plane_thresh;     /* P2PLANE_KNN: all k neighbours within this distance of the plane (0.2) */
    double T0[16];           /* initial guess, row-major */
    double interp_s;         /* P2LINE / P2PLANE_3PT: the interpolation ratio s of RadarEdgeFactor / LidarPlaneFactor
                                (radarFactor.hpp:26-32,78-84): the residual is evaluated at slerp(I, q, s) p + s t instead of
                                q p + t (A-LOAM's motion-distortion model; one s per call). 1 (default) = the full pose;
                                values <= 0 are treated as 1. Ignored by the other kinds. */
} icp4r_opts;

typedef struct icp4r_result {
    int32_t converged;       /* PCL semantics: max_iterations reached or a criterion met */
    int32_t iterations;
    int32_t n_corr;          /* correspondences used by the last iteration */
    int32_t n_fitness;
    double fitness;          /* mean squared 1-NN distance after the final transform (getFitnessScore) */
    double last_cost;
} icp4r_result;

/* Optional per-iteration debug dumps (caller-owned, same memory space as the call's `mem`; any may be NULL).
 *   pose: [max_iterations][16]  pose BEFORE iteration i
 *   acc : [max_iterations][ICP4R_ACC_LEN]
 *         P2P_SVD -> n, sum p'(3), sum q(3), sum p' q^T (9, row-major), sum d2
 *         GN kinds -> H upper triangle row-major (21), g = J^T r (6), cost, n
 *   idx : [max_iterations][n][k] neighbour indices */
typedef struct icp4r_dump {
    double* pose;
    double* acc;
    int32_t* idx;
} icp4r_dump;

/* ---- lifecycle ------------------------------------------------------------------------------------ */
int icp4r_create(int device, icp4r_handle* out);
int icp4r_destroy(icp4r_handle h);
const char* icp4r_last_error(icp4r_handle h);
const char* icp4r_version(void);
int icp4r_default_opts(icp4r_opts* o);
/* run on an existing cudaStream_t (e.g. the harness' current stream) instead of the handle's own */
int icp4r_set_stream(icp4r_handle h, void* cuda_stream);
/* Layout of the point-cloud INPUT rows (every `*_xyzw` input parameter) of the calls that follow, host or device:
 * row i starts at byte i * stride_bytes and holds x, y, z as floats at bytes 0, 4, 8 and w at byte w_offset_bytes
 * (< 0: no w, 0 is stored). Default (16, 12) = packed x, y, z, w. pcl::PointXYZI, the reference's PointType
 * (/root/reference/include/radar_odometry.h typedef, 32-byte rows, intensity at byte 16), is (32, 16): the adapters pass
 * the reference's clouds as they lie in memory — the rows cross the bus unmodified and one device kernel repacks them,
 * no host-side pack loop. stride_bytes: multiple of 4 in [12, 4096]. Point OUTPUTS are always packed x, y, z, w;
 * the 5-float Doppler records (icp4r_doppler_*) have their own fixed layout. */
int icp4r_set_point_layout(icp4r_handle h, int32_t stride_bytes, int32_t w_offset_bytes);
/* Completion: every call that returns something in HOST memory (poses, results, counts, host output arrays) has
 * finished when it returns. Outputs the caller asked for in DEVICE memory (mem = ICP4R_DEVICE: neighbour tables,
 * transformed points, batched poses/results, dumps, filtered clouds) are ordered on the handle's stream: use them on
 * that stream, or call icp4r_synchronize first. */
int icp4r_synchronize(icp4r_handle h);
/* number of kernels this handle has launched so far (graph-replayed kernels included) */
int icp4r_launch_count(icp4r_handle h, int64_t* out);
/* measurement aid: with profiling on, icp4r_register / icp4r_register_map launch their iteration kernels
 * one by one (no graph) with a CUDA event between launches; icp4r_last_profile then returns the device
 * time of each launch in milliseconds: max_iterations iteration kernels followed by the fitness pass. */
int icp4r_set_profiling(icp4r_handle h, int on);
int icp4r_last_profile(icp4r_handle h, float* ms_out, int32_t cap, int32_t* n_out);
/* measurement aid: work counters of the registration kernels (icp4r_register*, all flavours), summed over the calls
 * since the last icp4r_get_stats: out[0] = neighbour searches run, out[1] = squared-distance evaluations (candidates
 * looked at), out[2] = source points whose previous neighbours were proven still nearest without a search (batched
 * pairs), out[3..7] reserved. Counting costs a few instructions per search; it is off by default. */
int icp4r_set_stats(icp4r_handle h, int on);
int icp4r_get_stats(icp4r_handle h, int64_t out[8]);

/* ---- map: replaces KD_TREE<PointType> (ikd_Tree.h:227-251) ---------------------------------------- */
/* Build (ikd_Tree.cpp:354-365): discards any previous map. cell_size <= 0 picks one from the density. */
int icp4r_map_build(icp4r_handle h, const float* xyzw, int32_t n, int mem, float cell_size);
/* set_downsample_param (ikd_Tree.h:232) */
int icp4r_map_set_downsample(icp4r_handle h, float voxel);
/* Add_Points (ikd_Tree.cpp:422-497). downsample_on = 0 appends; 1 keeps one point per voxel, the one
 * nearest the voxel centre (new point wins ties). *n_replaced = the reference's return value. */
int icp4r_map_add_points(icp4r_handle h, const float* xyzw, int32_t n, int mem, int downsample_on,
                         int32_t* n_replaced);
/* size() / validnum() (ikd_Tree.h:234-235) */
int icp4r_map_size(icp4r_handle h, int32_t* size, int32_t* valid);
/* tree_range() (ikd_Tree.h:248): min xyz, max xyz over valid points */
int icp4r_map_range(icp4r_handle h, float out6[6]);
/* Nearest_Search (ikd_Tree.cpp:368-398) for nq queries at once. idx/d2: [nq,k]; found: [nq] (may be NULL) */
int icp4r_map_knn(icp4r_handle h, const float* q_xyzw, int32_t nq, int mem, int32_t k, double max_dist,
                  int32_t* idx, float* d2, int32_t* found);
/* same contract, forced through the exhaustive tiled kernel (no grid) — the on-device cross-check */
int icp4r_map_knn_brute(icp4r_handle h, const float* q_xyzw, int32_t nq, int mem, int32_t k, double max_dist,
                        int32_t* idx, float* d2, int32_t* found);
/* Sector_Search (ikd_Tree.cpp:415-419,1098-1140): indices of matching points, unordered; *n_out may exceed cap */
int icp4r_map_sector(icp4r_handle h, const float centre_xyz[3], float radius, float heading_deg, int mem,
                     int32_t* idx_out, int32_t cap, int32_t* n_out);
/* copy of the stored points in insertion order (valid flags optional) */
int icp4r_map_points(icp4r_handle h, int mem, float* xyzw_out, uint8_t* valid_out, int32_t cap);

/* ---- registration: replaces the PCL / fast_gicp align() calls -------------------------------------- */
/* source vs an explicit target cloud (iterative_closest_point.cpp:510-521): builds a transient index */
int icp4r_register(icp4r_handle h, const float* src_xyzw, int32_t n, const float* tgt_xyzw, int32_t m, int mem,
                   const icp4r_opts* opts, double T_out[16], icp4r_result* res, const icp4r_dump* dump);
/* source vs the handle's map (radar_odometry.cpp:390-411 with the ikd-Tree map as the target) */
int icp4r_register_map(icp4r_handle h, const float* src_xyzw, int32_t n, int mem, const icp4r_opts* opts,
                       double T_out[16], icp4r_result* res, const icp4r_dump* dump);
/* source vs a SUBSET of the handle's map given by point indices (icp4r_map_sector's output; same memory space as the
 * call): what radar_odometry.cpp:396-405 does with SubMap. The subset is gathered on the device into the handle's
 * transient target, indexed and registered like icp4r_register — the sub-map never crosses the bus. */
int icp4r_register_submap(icp4r_handle h, const float* src_xyzw, int32_t n, const int32_t* idx, int32_t n_idx, int mem,
                          const icp4r_opts* opts, double T_out[16], icp4r_result* res);
/* n_scans independent scans against the handle's map in ONE sequence of launches (throughput mode: one scan of a
 * few thousand points leaves most of the GPU idle). src_xyzw holds the scans back to back (memory space `mem`);
 * off [n_scans+1] (points), T0s ([n_scans][16] initial guesses, or NULL = opts->T0 for all), T_out [n_scans][16]
 * and res [n_scans] are HOST memory. All residual kinds except ICP4R_GICP. */
int icp4r_register_map_batch(icp4r_handle h, const float* src_xyzw, const int32_t* off, int32_t n_scans, int mem,
                             const icp4r_opts* opts, const double* T0s, double* T_out, icp4r_result* res);
/* n_pairs independent (source, target) pairs, CSR-style offsets ([n_pairs+1], in points), one CTA per pair
 * with both clouds resident in shared memory. P2P_SVD and P2P_GN. T_out: [n_pairs][16]; res: [n_pairs].
 * opts->T0 is the initial guess of every pair. Offsets and results follow `mem` too. */
int icp4r_register_batch(icp4r_handle h, const float* src_xyzw, const int32_t* src_off, const float* tgt_xyzw,
                         const int32_t* tgt_off, int32_t n_pairs, int mem, const icp4r_opts* opts, double* T_out,
                         icp4r_result* res);

/* ---- slab-sharded large maps (one process per GPU) -------------------------------------------------- */
/* rank 0 makes an id, the host program broadcasts the 128 bytes (any transport), every rank joins */
int icp4r_shard_unique_id(char id_out[128]);
int icp4r_shard_init(icp4r_handle h, const char id[128], int rank, int world);
/* Optional faster cross-rank sum: instead of an NCCL call per iteration, each rank's iteration kernel writes its
 * partial sums straight into every peer's exchange buffer over NVLink (CUDA IPC mapping) and adds the peers'
 * contributions itself, so a sharded iteration stays ONE kernel. Every rank exports a 64-byte handle, the host
 * program gathers them (any transport), every rank imports the full list (handles = world x 64 bytes, rank order).
 * One process per GPU, at most 8 ranks; once imported, icp4r_register_sharded uses this path. */
int icp4r_shard_ipc_export(icp4r_handle h, unsigned char handle_out[64]);
int icp4r_shard_ipc_import(icp4r_handle h, const unsigned char* handles, int rank, int world);
/* registration against the union of all ranks' maps: every rank passes the same source; each source
 * point is owned by the rank whose slab [slab_lo, slab_hi) along `axis` contains its transformed
 * position; the 29 partial accumulators are summed across ranks every iteration; all ranks return the
 * same pose. Requires a gated search and a halo >= max_corr_dist around each rank's map. */
int icp4r_register_sharded(icp4r_handle h, const float* src_xyzw, int32_t n, int mem, const icp4r_opts* opts,
                           int axis, float slab_lo, float slab_hi, double T_out[16], icp4r_result* res);

/* One linearisation of a slab-sharded registration at the pose T, WITHOUT the cross-rank sum and the solve: acc_out
 * receives this slab's partial accumulators (layout of icp4r_dump.acc) over the source points whose transformed
 * position falls in [slab_lo, slab_hi) along `axis` (axis = -1: every point, the whole map). For host programs that
 * bring their own collective (MPI, gloo, ...) and for checking the ownership / halo rule on a single GPU. acc_out is
 * HOST memory. All residual kinds except ICP4R_GICP. */
int icp4r_accumulate_slab(icp4r_handle h, const float* src_xyzw, int32_t n, int mem, const icp4r_opts* opts, const double T[16],
                          int axis, float slab_lo, float slab_hi, double acc_out[ICP4R_ACC_LEN]);

/* ---- Doppler static-point filter (next to the path: the step before registration in icp4radar) -------------------- */
/* records: n x 5 floats x, y, z, intensity, v_r (the reference's .bin layout, iterative_closest_point.cpp:373-377).
 * Two-point sine-model RANSAC (fitSineRansac, :85-128), static/dynamic split by the signed residual (:392-403) and
 * least-squares ego velocity over the static points (:412-427). iterations <= 0 -> 0.2 * n (:389). */
typedef struct icp4r_doppler_opts {
    int32_t iterations;
    int32_t reserved;
    uint64_t seed;     /* sample indices come from a counter-based generator (the reference: unseeded random_device) */
    double sigma;      /* inlier band of the RANSAC score, reference default 0.5 */
    double split;      /* dynamic if residual > split, reference 0.2 */
} icp4r_doppler_opts;
typedef struct icp4r_doppler_result {
    double A, b, score;   /* best model v_r cos(beta) = A cos(alpha + b) and its inlier count */
    double velocity[3];
    int32_t n_static;
    int32_t best_iteration;
} icp4r_doppler_result;
int icp4r_doppler_filter(icp4r_handle h, const float* xyziv, int32_t n, int mem, const icp4r_doppler_opts* opts,
                         uint8_t* static_mask, icp4r_doppler_result* res);

/* The static points of a radar frame: icp4r_doppler_filter followed by an order-preserving compaction of the records the
 * filter keeps, as packed x, y, z, intensity rows — the cloud the reference hands to its registration
 * (iterative_closest_point.cpp:392-407). At most `cap` rows are written; *n_out = number of static points. */
int icp4r_doppler_static_points(icp4r_handle h, const float* xyziv, int32_t n, int mem, const icp4r_doppler_opts* opts,
                                float* xyzw_out, int32_t cap, int32_t* n_out, icp4r_doppler_result* res);

/* ---- helpers on the path ----------------------------------------------------------------------------- */
/* p' = R p + t in double, written back as float (pointAssociateToMap, radar_odometry.cpp:137-145) */
int icp4r_transform_points(icp4r_handle h, const double T[16], const float* xyzw, int32_t n, int mem,
                           float* xyzw_out);

/* ---- the rest of the KD_TREE surface (ikd_Tree.h:243-249) ----------------------------------------------- */
/* KD_TREE::Box_Search (ikd_Tree.cpp:401-405,1024-1051): indices of the valid points with min <= p < max on every
 * axis. Output order is unspecified (the reference: tree order); *n_out = number of hits, at most cap written. */
int icp4r_map_box_search(icp4r_handle h, const float box_min[3], const float box_max[3], int mem, int32_t* idx_out,
                         int32_t cap, int32_t* n_out);
/* KD_TREE::Radius_Search (ikd_Tree.cpp:408-412,1054-1095): valid points with float d2 <= radius * radius. */
int icp4r_map_radius_search(icp4r_handle h, const float centre_xyz[3], float radius, int mem, int32_t* idx_out,
                            int32_t cap, int32_t* n_out);
/* KD_TREE::Delete_Point_Boxes (ikd_Tree.cpp:544-565): deletes every valid point inside any of the boxes
 * (boxes6 = n_boxes x {min xyz, max xyz}, host memory, half-open like Box_Search); *n_deleted = how many. */
int icp4r_map_delete_boxes(icp4r_handle h, const float* boxes6, int32_t n_boxes, int32_t* n_deleted);
/* KD_TREE::Add_Point_Boxes (ikd_Tree.cpp:500-519): points inside the boxes that Delete_Points / Delete_Point_Boxes
 * removed come back (points removed by down-sampling do not). The reference can only revive points that no
 * re-balancing rebuild has purged yet — which ones depends on its tree shape; here every such point is restorable. */
int icp4r_map_add_boxes(icp4r_handle h, const float* boxes6, int32_t n_boxes, int32_t* n_restored);
/* KD_TREE::Delete_Points (ikd_Tree.cpp:522-541): for each requested point, in order, ONE valid point with
 * |dx|, |dy|, |dz| < 1e-6 (same_point, :1422) is deleted — the one with the lowest index (the reference: the first on
 * its descent path, which can miss a copy when coordinates tie on a split axis). */
int icp4r_map_delete_points(icp4r_handle h, const float* xyzw, int32_t n, int mem, int32_t* n_deleted);

/* One frame of scan-to-map odometry (the loop body of radar_odometry.cpp:380-421 with registration before insertion):
 * register the scan against the handle's map starting from T_io, write the estimated pose back to T_io, transform the
 * scan with it (pointAssociateToMap) and append it to the map (Add_Points(.., false)) — one host-to-device copy of
 * the scan, nothing else crosses the bus. With an empty map the scan is inserted at T_io (first frame).
 * downsample_on must be 0 (use icp4r_map_add_points for down-sampled insertion). */
int icp4r_odometry_step(icp4r_handle h, const float* scan_xyzw, int32_t n, int mem, const icp4r_opts* opts,
                        int downsample_on, double T_io[16], icp4r_result* res);

/* Centroid-per-leaf down-sampling: pcl::VoxelGrid<PointXYZI>::filter with setLeafSize(leaf, leaf, leaf) as the
 * scan-to-map node runs it over the accumulated map every frame (radar_odometry.cpp:426-429). One output point per
 * occupied leaf (float mean of x, y, z, intensity), ascending leaf index (x fastest). Non-finite points are
 * skipped. xyzw == NULL: filter the handle's map (deleted points skipped, n ignored). At most `cap` points are
 * written; *n_out receives the number of occupied leaves (call again with a larger buffer if it exceeds cap).
 * A leaf so small that the leaf index would overflow 31 bits is an error (PCL passes the cloud through instead). */
int icp4r_voxel_grid(icp4r_handle h, const float* xyzw, int32_t n, int mem, float leaf, float* xyzw_out,
                     int32_t cap, int32_t* n_out);

#ifdef __cplusplus
}
#endif
#endif /* ICP4R_H */
