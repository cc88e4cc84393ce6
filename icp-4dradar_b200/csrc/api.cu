// api.cu — the extern "C" boundary of libicp4r_cuda (include/icp4r.h). Argument checking, host<->device
// staging on the handle's stream, and dispatch into the kernels. There is no CPU fallback anywhere.
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "ctx.h"

namespace icp4r {

static thread_local std::string g_create_err;

int fail(Ctx* c, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_err = buf;
    return code;
}

// like reserve, but over-allocates by 50 % so a buffer that grows a little on every call (the map's sort and
// index arrays in an odometry loop) is not freed and re-allocated each time
int reserve_grow(Ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return ICP4R_OK;
    return reserve(c, b, bytes + bytes / 2);
}

int reserve(Ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return ICP4R_OK;
    if (b.p) {
        cudaStreamSynchronize(c->stream);  // nobody may still be reading the old block
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes < 256 ? 256 : bytes;
    const cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(c, ICP4R_ERR_NOMEM, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return ICP4R_OK;
}

void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

// order-preserving compaction of the radar records the Doppler filter kept, as x, y, z, intensity rows (one block)
__global__ void __launch_bounds__(1024) compact_static_kernel(const float* __restrict__ rec, const uint8_t* __restrict__ mask, int n,
                                                             float4* __restrict__ out, int cap, int* __restrict__ n_out) {
    __shared__ int wcnt[32];
    __shared__ int carry_s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n; b0 += 1024) {
        const int i = b0 + tid;
        const bool keep = i < n && mask[i] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wcnt[w] = __popc(bal);
        __syncthreads();
        int before = carry_s;
        for (int j = 0; j < w; ++j) before += wcnt[j];
        if (keep) {
            const int pos = before + __popc(bal & ((1u << lane) - 1u));
            if (pos < cap) out[pos] = make_float4(rec[5 * (size_t)i], rec[5 * (size_t)i + 1], rec[5 * (size_t)i + 2], rec[5 * (size_t)i + 3]);
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int j = 0; j < 32; ++j) t += wcnt[j];
            carry_s += t;
        }
        __syncthreads();
    }
    if (tid == 0) *n_out = carry_s;
}

// target of a sub-map registration: the map's points at the given indices (out-of-range indices give a NaN row, which
// every search ignores)
__global__ void __launch_bounds__(256) gather_subset_kernel(const float4* __restrict__ pts, int m, const int32_t* __restrict__ idx, int n_idx,
                                                           float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_idx) return;
    const int j = idx[i];
    out[i] = (j >= 0 && j < m) ? pts[j] : make_float4(NAN, NAN, NAN, 0.f);
}

int shard_unique_id(char id_out[128]);
int shard_init(Ctx* c, const char id_in[128], int rank, int world);
void shard_destroy(Ctx* c);
int shard_ipc_export(Ctx* c, unsigned char out[64]);
int shard_ipc_import(Ctx* c, const unsigned char* handles, int rank, int world);
void shard_ipc_close(Ctx* c);
int transform_points(Ctx* c, const double* T_host, const float4* d_in, int n, float4* d_out);

static void free_map(Map& m) {
    release(m.pts);
    release(m.valid);
    release(m.userdel);
    release(m.sorted);
    release(m.cell_start);
    release(m.keys_a);
    release(m.keys_b);
    release(m.vals_a);
    release(m.vals_b);
    release(m.normals);
    release(m.coarse);
    release(m.sorted_alt);
    release(m.ik_a);
    release(m.ik_b);
    release(m.iv_a);
    release(m.iv_b);
    release(m.inc_bnd);
    m.m = m.m_valid = 0;
    m.built = false;
}

static void drop_graphs(Ctx* c) {
    for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
    c->graphs.clear();
    c->graph_grid_owner = nullptr;
}

// copy a caller buffer onto the device (or alias it when it already lives there)
static int stage_in(Ctx* c, DevBuf& b, const void* user, size_t bytes, int mem, const void** dev) {
    if (bytes == 0) {
        CKS(reserve(c, b, 256));
        *dev = b.p;
        return ICP4R_OK;
    }
    if (mem == ICP4R_DEVICE) {
        *dev = user;
        return ICP4R_OK;
    }
    CKS(reserve(c, b, bytes));
    CK(cudaMemcpyAsync(b.p, user, bytes, cudaMemcpyHostToDevice, c->stream));
    *dev = b.p;
    return ICP4R_OK;
}

// rows of `stride` bytes (x, y, z at byte 0, 4, 8; w at byte `woff`, or none) -> packed x, y, z, w
__global__ void __launch_bounds__(256) repack_rows_kernel(const unsigned char* __restrict__ raw, size_t n, int stride, int woff, int vec,
                                                          float4* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* row = reinterpret_cast<const float*>(raw + i * (size_t)stride);
    float4 p;
    if (vec) {  // 16-byte aligned rows (pcl::PointXYZI: 32 bytes, aligned base): one vector load for x, y, z
        const float4 v = *reinterpret_cast<const float4*>(row);
        p = make_float4(v.x, v.y, v.z, woff == 12 ? v.w : 0.f);
    } else {
        p = make_float4(row[0], row[1], row[2], 0.f);
    }
    if (woff >= 0 && !(vec && woff == 12)) p.w = row[woff >> 2];
    out[i] = p;
}

static bool packed_layout(const Ctx* c) { return c->pt_stride == 16 && c->pt_woff == 12; }
static const float* row_at(const Ctx* c, const float* base, size_t i) {
    return reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(base) + i * (size_t)c->pt_stride);
}

// the caller's rows [0, n) as packed float4 at `dst` (device memory), whatever their layout and memory space
static int unpack_points_to(Ctx* c, const float* user, size_t n, int mem, float4* dst) {
    if (n == 0) return ICP4R_OK;
    if (packed_layout(c)) {
        CK(cudaMemcpyAsync(dst, user, n * sizeof(float4), mem == ICP4R_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
        return ICP4R_OK;
    }
    const unsigned char* raw = reinterpret_cast<const unsigned char*>(user);
    if (mem == ICP4R_HOST) {  // the raw rows cross the bus as they are (no host-side pack loop), the device repacks them
        CKS(reserve_grow(c, c->d_raw, n * (size_t)c->pt_stride));
        CK(cudaMemcpyAsync(c->d_raw.p, user, n * (size_t)c->pt_stride, cudaMemcpyHostToDevice, c->stream));
        raw = static_cast<const unsigned char*>(c->d_raw.p);
    }
    // vector loads need 16-byte aligned rows: the stride AND the caller's base pointer (a device pointer may be anything)
    const int vec = ((c->pt_stride & 15) == 0 && (reinterpret_cast<uintptr_t>(raw) & 15) == 0) ? 1 : 0;
    if ((reinterpret_cast<uintptr_t>(raw) & 3) != 0) return fail(c, ICP4R_ERR_INVALID, "point rows must be 4-byte aligned");
    repack_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(raw, n, c->pt_stride, c->pt_woff, vec, dst);
    c->launches += 1;
    CK(cudaGetLastError());
    return ICP4R_OK;
}

// a point-cloud input of a call: packed float4 on the device (aliased when it already is exactly that)
static int stage_points(Ctx* c, DevBuf& b, const float* user, size_t n, int mem, const void** dev) {
    if (packed_layout(c)) return stage_in(c, b, user, n * sizeof(float4), mem, dev);
    CKS(reserve(c, b, std::max<size_t>(n, 16) * sizeof(float4)));
    CKS(unpack_points_to(c, user, n, mem, static_cast<float4*>(b.p)));
    *dev = b.p;
    return ICP4R_OK;
}

static int set_points(Ctx* c, Map& mp, const float* xyzw, int n, int mem, int offset) {
    CKS(map_reserve(c, mp, offset + n));
    if (n > 0) {
        CKS(unpack_points_to(c, xyzw, (size_t)n, mem, mp.pts.as<float4>() + offset));
        CK(cudaMemsetAsync(mp.valid.as<uint8_t>() + offset, 1, (size_t)n, c->stream));
        CK(cudaMemsetAsync(mp.userdel.as<uint8_t>() + offset, 0, (size_t)n, c->stream));
    }
    return ICP4R_OK;
}

static bool bad_mem(int mem) { return mem != ICP4R_HOST && mem != ICP4R_DEVICE; }

}  // namespace icp4r

using namespace icp4r;

#define HCHECK(h)                       \
    if (!(h)) return ICP4R_ERR_INVALID; \
    Ctx* c = (h);                       \
    c->err.clear();                     \
    if (cudaSetDevice(c->device) != cudaSuccess) return fail(c, ICP4R_ERR_CUDA, "cudaSetDevice(%d) failed", c->device)


extern "C" {

const char* icp4r_version(void) { return "libicp4r_cuda 0.1 (sm_100a)"; }

int icp4r_default_opts(icp4r_opts* o) {
    if (!o) return ICP4R_ERR_INVALID;
    std::memset(o, 0, sizeof(*o));
    o->residual = ICP4R_P2P_SVD;
    o->k = 5;
    o->max_iterations = 10;  // PCL default; iterative_closest_point.cpp:513 leaves it untouched
    o->early_exit = 0;
    o->max_corr_dist = 0.0;
    o->rot_eps = 2e-3;  // fast_gicp defaults
    o->trans_eps = 5e-4;
    o->mse_abs_eps = 1e-12;
    o->plane_thresh = 0.2;
    for (int i = 0; i < 16; ++i) o->T0[i] = (i % 5 == 0) ? 1.0 : 0.0;
    o->interp_s = 1.0;
    return ICP4R_OK;
}

int icp4r_create(int device, icp4r_handle* out) {
    if (!out) return ICP4R_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(nullptr, ICP4R_ERR_CUDA, "no CUDA device (%s); libicp4r_cuda has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, ICP4R_ERR_INVALID, "device %d out of range (0..%d)", device, ndev - 1);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, ICP4R_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, ICP4R_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(nullptr, ICP4R_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                    prop.minor);
    icp4r_ctx* c = new (std::nothrow) icp4r_ctx();
    if (!c) return ICP4R_ERR_NOMEM;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete c;
        return fail(nullptr, ICP4R_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    c->stream = c->own_stream;
    c->h_pinned_cap = 64 * 1024;
    if ((e = cudaMallocHost(&c->h_pinned, c->h_pinned_cap)) != cudaSuccess) {
        cudaStreamDestroy(c->own_stream);
        delete c;
        return fail(nullptr, ICP4R_ERR_NOMEM, "cudaMallocHost: %s", cudaGetErrorString(e));
    }
    const char* ng = std::getenv("ICP4R_NO_GRAPH");
    c->use_graph = !(ng && ng[0] == '1');
    const char* nh = std::getenv("ICP4R_NO_HINTS");
    c->use_hints = !(nh && nh[0] == '1');
    {
        const char* nl = std::getenv("ICP4R_NO_LB");
        c->use_lb = !(nl && nl[0] == '1');
        // measured on B200 (scripts/probe_map_kernels.py): the persistent loop is bit-identical but SLOWER than the captured
        // graph of per-iteration launches (C2 single 0.485 vs 0.436 ms, C5 0.762 vs 0.695 ms) — a kernel boundary inside a
        // graph costs ~1 us here, the in-kernel hand-over a release + an acquire round trip + the pose reload — so it is opt-in
        if (const char* e = std::getenv("ICP4R_LB_SLACK_A")) c->slack_a = (float)std::atof(e);
        if (const char* e = std::getenv("ICP4R_LB_SLACK_B")) c->slack_b = (float)std::atof(e);
        const char* np_ = std::getenv("ICP4R_PERSIST");
        c->use_persist = np_ && np_[0] == '1';
    }
    const char* br = std::getenv("ICP4R_BATCH_REPRODUCIBLE");
    c->batch_reproducible = br && br[0] == '1';
    *out = c;
    return ICP4R_OK;
}

int icp4r_destroy(icp4r_handle h) {
    if (!h) return ICP4R_ERR_INVALID;
    Ctx* c = h;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->own_stream != c->stream) cudaStreamSynchronize(c->own_stream);
    drop_graphs(c);
    for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
    if (c->copy_stream) {
        for (auto& e : c->copy_events)
            if (e) cudaEventDestroy(e);
        cudaStreamDestroy(c->copy_stream);
        if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    }
    shard_destroy(c);
    shard_ipc_close(c);
    release(c->d_xch);
    release(c->d_xt);
    release(c->d_nbprev);
    release(c->d_raw);
    release(c->d_pairctr);
    release(c->d_nbstate);
    release(c->bf_part);
    release(c->gs_pts);
    release(c->gs_idx);
    release(c->gs_d2);
    release(c->gs_found);
    release(c->vg_keys);
    release(c->vg_vals);
    release(c->vg_sort);
    release(c->vg_tiles);
    release(c->vg_out);
    free_map(c->map);
    free_map(c->tmp);
    free_map(c->srcmap);
    release(c->d_src_normals);
    release(c->d_gicp_corr);
    DevBuf* bufs[] = {&c->d_src, &c->d_q, &c->d_idx, &c->d_d2, &c->d_found, &c->d_scratch, &c->d_partials, &c->d_state, &c->d_params,
                      &c->d_T, &c->d_res, &c->d_dump_pose, &c->d_dump_acc, &c->d_dump_idx, &c->b_src, &c->b_tgt, &c->b_soff,
                      &c->b_toff, &c->b_T, &c->b_res, &c->bm_params, &c->bm_state, &c->bm_T0, &c->bm_res, &c->bm_partials};
    for (DevBuf* b : bufs) release(*b);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    cudaStreamDestroy(c->own_stream);
    delete h;
    return ICP4R_OK;
}

const char* icp4r_last_error(icp4r_handle h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int icp4r_set_stream(icp4r_handle h, void* s) {
    HCHECK(h);
    cudaStreamSynchronize(c->stream);
    c->stream = s ? static_cast<cudaStream_t>(s) : c->own_stream;
    return ICP4R_OK;
}

int icp4r_set_point_layout(icp4r_handle h, int32_t stride_bytes, int32_t w_offset_bytes) {
    HCHECK(h);
    if (stride_bytes < 12 || (stride_bytes & 3) != 0 || stride_bytes > 4096 || (w_offset_bytes >= 0 && ((w_offset_bytes & 3) != 0 || w_offset_bytes + 4 > stride_bytes)))
        return fail(c, ICP4R_ERR_INVALID, "icp4r_set_point_layout: stride %d / w offset %d (rows are 4-byte aligned, x y z first)", stride_bytes, w_offset_bytes);
    c->pt_stride = stride_bytes;
    c->pt_woff = w_offset_bytes < 0 ? -1 : w_offset_bytes;
    return ICP4R_OK;
}

int icp4r_synchronize(icp4r_handle h) {
    HCHECK(h);
    CK(cudaStreamSynchronize(c->stream));
    return ICP4R_OK;
}

int icp4r_launch_count(icp4r_handle h, int64_t* out) {
    if (!h || !out) return ICP4R_ERR_INVALID;
    *out = h->launches;
    return ICP4R_OK;
}

int icp4r_set_profiling(icp4r_handle h, int on) {
    HCHECK(h);
    c->profiling = on != 0;
    return ICP4R_OK;
}

int icp4r_last_profile(icp4r_handle h, float* ms_out, int32_t cap, int32_t* n_out) {
    HCHECK(h);
    if (!n_out || cap < 0 || (cap > 0 && !ms_out)) return fail(c, ICP4R_ERR_INVALID, "icp4r_last_profile: bad arguments");
    *n_out = (int32_t)c->prof_ms.size();
    for (int i = 0; i < cap && i < (int)c->prof_ms.size(); ++i) ms_out[i] = c->prof_ms[i];
    return ICP4R_OK;
}

int icp4r_set_stats(icp4r_handle h, int on) {
    HCHECK(h);
    if (on) {
        CKS(reserve(c, c->d_stats, 8 * sizeof(unsigned long long)));
        CK(cudaMemsetAsync(c->d_stats.p, 0, 8 * sizeof(unsigned long long), c->stream));
    }
    c->stats = on != 0;
    // captured loops bake the per-call parameter block's address, not its content: nothing to invalidate
    return ICP4R_OK;
}

int icp4r_get_stats(icp4r_handle h, int64_t out[8]) {
    HCHECK(h);
    if (!out) return fail(c, ICP4R_ERR_INVALID, "icp4r_get_stats: null output");
    for (int i = 0; i < 8; ++i) out[i] = 0;
    if (!c->d_stats.p) return ICP4R_OK;
    CK(cudaMemcpyAsync(out, c->d_stats.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemsetAsync(c->d_stats.p, 0, 8 * sizeof(unsigned long long), c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return ICP4R_OK;
}

// ---- map ------------------------------------------------------------------------------------------------

int icp4r_map_build(icp4r_handle h, const float* xyzw, int32_t n, int mem, float cell_size) {
    HCHECK(h);
    if (n < 0 || (n > 0 && !xyzw) || bad_mem(mem)) return fail(c, ICP4R_ERR_INVALID, "icp4r_map_build: bad arguments");
    Map& mp = c->map;
    mp.m = 0;
    mp.m_valid = 0;
    mp.built = false;
    mp.user_cell = cell_size;
    mp.hint_cell = 0.f;
    mp.no_bucket = false;
    mp.padded = false;  // Build discards the previous map, its growth history included (a padded grid has ~2x the cells)
    CKS(set_points(c, mp, xyzw, n, mem, 0));
    mp.m = n;
    CKS(map_rebuild_grid(c, mp));
    return ICP4R_OK;
}

int icp4r_map_set_downsample(icp4r_handle h, float voxel) {
    HCHECK(h);
    if (!(voxel > 0.f)) return fail(c, ICP4R_ERR_INVALID, "voxel must be positive");
    c->map.ds_voxel = voxel;
    return ICP4R_OK;
}

int icp4r_map_add_points(icp4r_handle h, const float* xyzw, int32_t n, int mem, int downsample_on, int32_t* n_replaced) {
    HCHECK(h);
    if (n < 0 || (n > 0 && !xyzw) || bad_mem(mem)) return fail(c, ICP4R_ERR_INVALID, "icp4r_map_add_points: bad arguments");
    if (n_replaced) *n_replaced = 0;
    Map& mp = c->map;
    if (n == 0) return ICP4R_OK;
    if (!downsample_on) {
        CKS(set_points(c, mp, xyzw, n, mem, mp.m));
        bool merged = false;
        CKS(map_append_incremental(c, mp, n, &merged));  // merge into the existing grid when the batch fits
        mp.m += n;
        if (!merged) CKS(map_rebuild_grid(c, mp));
        return ICP4R_OK;
    }
    // down-sampling: the reference inserts point by point, each seeing the effect of the previous ones. The device
    // evaluates that rule voxel-parallel over chunks (the chunk bound keeps the leader search cheap); between
    // chunks the grid is rebuilt so the next chunk sees the survivors.
    if (!mp.built) CKS(map_rebuild_grid(c, mp));
    const char* fs = std::getenv("ICP4R_DS_SEQUENTIAL");
    const bool force_seq = fs && fs[0] == '1';
    const int CH = 8192;
    int total = 0;
    for (int off = 0; off < n; off += CH) {
        const int cn = std::min(CH, n - off);
        CKS(set_points(c, mp, row_at(c, xyzw, (size_t)off), cn, mem, mp.m));
        CK(cudaMemsetAsync(mp.valid.as<uint8_t>() + mp.m, 0, (size_t)cn, c->stream));  // not inserted yet
        int rep = 0;
        CKS(map_downsample_add(c, mp, cn, &rep, force_seq));
        total += rep;
        mp.m += cn;
        CKS(map_rebuild_grid(c, mp));
    }
    if (n_replaced) *n_replaced = total;
    return ICP4R_OK;
}

int icp4r_map_size(icp4r_handle h, int32_t* size, int32_t* valid) {
    HCHECK(h);
    if (size) *size = c->map.m;
    if (valid) *valid = c->map.m_valid;
    return ICP4R_OK;
}

int icp4r_map_range(icp4r_handle h, float out6[6]) {
    HCHECK(h);
    if (!out6) return fail(c, ICP4R_ERR_INVALID, "null output");
    if (!c->map.built) return fail(c, ICP4R_ERR_STATE, "map not built");
    for (int a = 0; a < 3; ++a) {
        out6[a] = c->map.bb_min[a];
        out6[3 + a] = c->map.bb_max[a];
    }
    return ICP4R_OK;
}

static int map_knn_common(Ctx* c, bool brute, const float* q, int32_t nq, int mem, int32_t k, double max_dist, int32_t* idx, float* d2,
                          int32_t* found) {
    if (nq < 0 || k < 1 || k > ICP4R_MAX_K || bad_mem(mem) || (nq > 0 && (!q || !idx || !d2)))
        return fail(c, ICP4R_ERR_INVALID, "icp4r_map_knn: bad arguments (nq=%d k=%d)", nq, k);
    if (!c->map.built) return fail(c, ICP4R_ERR_STATE, "icp4r_map_knn before icp4r_map_build");
    if (nq == 0) return ICP4R_OK;
    const void* dq = nullptr;
    CKS(stage_points(c, c->d_q, q, (size_t)nq, mem, &dq));
    int32_t* di = idx;
    float* dd = d2;
    int32_t* df = found;
    if (mem == ICP4R_HOST) {
        CKS(reserve(c, c->d_idx, (size_t)nq * k * 4));
        CKS(reserve(c, c->d_d2, (size_t)nq * k * 4));
        CKS(reserve(c, c->d_found, (size_t)nq * 4));
        di = c->d_idx.as<int32_t>();
        dd = c->d_d2.as<float>();
        df = c->d_found.as<int32_t>();
    }
    if (brute) CKS(brute_knn(c, c->map, static_cast<const float4*>(dq), nq, k, max_dist, di, dd, df));
    else CKS(grid_knn(c, c->map, static_cast<const float4*>(dq), nq, k, max_dist, di, dd, df));
    if (mem == ICP4R_HOST) {
        CK(cudaMemcpyAsync(idx, di, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(d2, dd, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, c->stream));
        if (found) CK(cudaMemcpyAsync(found, df, (size_t)nq * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return ICP4R_OK;
}

int icp4r_map_knn(icp4r_handle h, const float* q, int32_t nq, int mem, int32_t k, double max_dist, int32_t* idx, float* d2,
                  int32_t* found) {
    HCHECK(h);
    return map_knn_common(c, false, q, nq, mem, k, max_dist, idx, d2, found);
}

int icp4r_map_knn_brute(icp4r_handle h, const float* q, int32_t nq, int mem, int32_t k, double max_dist, int32_t* idx, float* d2,
                        int32_t* found) {
    HCHECK(h);
    return map_knn_common(c, true, q, nq, mem, k, max_dist, idx, d2, found);
}

int icp4r_map_sector(icp4r_handle h, const float centre_xyz[3], float radius, float heading_deg, int mem, int32_t* idx_out,
                     int32_t cap, int32_t* n_out) {
    HCHECK(h);
    if (!centre_xyz || !n_out || cap < 0 || (cap > 0 && !idx_out) || bad_mem(mem)) return fail(c, ICP4R_ERR_INVALID, "icp4r_map_sector: bad arguments");
    if (!c->map.built) return fail(c, ICP4R_ERR_STATE, "icp4r_map_sector before icp4r_map_build");
    int32_t* d_out = idx_out;
    if (mem == ICP4R_HOST) {
        CKS(reserve(c, c->d_idx, (size_t)std::max(cap, 1) * 4));
        d_out = c->d_idx.as<int32_t>();
    }
    int cnt = 0;
    CKS(map_sector(c, c->map, centre_xyz, radius, heading_deg, d_out, cap, &cnt));
    if (mem == ICP4R_HOST && cap > 0 && cnt > 0) {
        CK(cudaMemcpyAsync(idx_out, d_out, (size_t)std::min(cnt, cap) * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    *n_out = cnt;
    return ICP4R_OK;
}

static int region_search(Ctx* c, int kind, const float a[3], const float b[3], int mem, int32_t* idx_out, int32_t cap, int32_t* n_out) {
    if (!c->map.built) return fail(c, ICP4R_ERR_STATE, "search before icp4r_map_build");
    int32_t* d_out = idx_out;
    if (mem == ICP4R_HOST) {
        CKS(reserve(c, c->d_idx, (size_t)std::max(cap, 1) * 4));
        d_out = c->d_idx.as<int32_t>();
    }
    int cnt = 0;
    CKS(map_region_search(c, c->map, kind, a, b, d_out, cap, &cnt));
    if (mem == ICP4R_HOST && cap > 0 && cnt > 0) {
        CK(cudaMemcpyAsync(idx_out, d_out, (size_t)std::min(cnt, cap) * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    *n_out = cnt;
    return ICP4R_OK;
}

int icp4r_map_box_search(icp4r_handle h, const float box_min[3], const float box_max[3], int mem, int32_t* idx_out, int32_t cap, int32_t* n_out) {
    HCHECK(h);
    if (!box_min || !box_max || !n_out || cap < 0 || (cap > 0 && !idx_out) || bad_mem(mem))
        return fail(c, ICP4R_ERR_INVALID, "icp4r_map_box_search: bad arguments");
    return region_search(c, 0, box_min, box_max, mem, idx_out, cap, n_out);
}

int icp4r_map_radius_search(icp4r_handle h, const float centre_xyz[3], float radius, int mem, int32_t* idx_out, int32_t cap, int32_t* n_out) {
    HCHECK(h);
    if (!centre_xyz || !n_out || cap < 0 || (cap > 0 && !idx_out) || bad_mem(mem))
        return fail(c, ICP4R_ERR_INVALID, "icp4r_map_radius_search: bad arguments");
    const float b[3] = {radius * radius, 0.f, 0.f};  // ikd_Tree.cpp:1070: calc_dist(...) <= radius * radius
    return region_search(c, 1, centre_xyz, b, mem, idx_out, cap, n_out);
}

static int box_flags(Ctx* c, const float* boxes6, int32_t n_boxes, bool revive, int32_t* n_changed) {
    if (n_changed) *n_changed = 0;
    if (n_boxes < 0 || (n_boxes > 0 && !boxes6)) return fail(c, ICP4R_ERR_INVALID, "bad box list");
    Map& mp = c->map;
    if (!mp.built) return fail(c, ICP4R_ERR_STATE, "box operation before icp4r_map_build");
    if (n_boxes == 0 || mp.m == 0) return ICP4R_OK;
    const void* dboxes = nullptr;
    CKS(stage_in(c, c->d_q, boxes6, (size_t)n_boxes * 6 * sizeof(float), ICP4R_HOST, &dboxes));
    int changed = 0;
    CKS(map_box_flags(c, mp, static_cast<const float*>(dboxes), n_boxes, revive, &changed));
    if (changed > 0) CKS(map_rebuild_grid(c, mp));
    if (n_changed) *n_changed = changed;
    return ICP4R_OK;
}

int icp4r_map_delete_boxes(icp4r_handle h, const float* boxes6, int32_t n_boxes, int32_t* n_deleted) {
    HCHECK(h);
    return box_flags(c, boxes6, n_boxes, false, n_deleted);
}

int icp4r_map_add_boxes(icp4r_handle h, const float* boxes6, int32_t n_boxes, int32_t* n_restored) {
    HCHECK(h);
    return box_flags(c, boxes6, n_boxes, true, n_restored);
}

int icp4r_map_delete_points(icp4r_handle h, const float* xyzw, int32_t n, int mem, int32_t* n_deleted) {
    HCHECK(h);
    if (n_deleted) *n_deleted = 0;
    if (n < 0 || (n > 0 && !xyzw) || bad_mem(mem)) return fail(c, ICP4R_ERR_INVALID, "icp4r_map_delete_points: bad arguments");
    Map& mp = c->map;
    if (!mp.built) return fail(c, ICP4R_ERR_STATE, "icp4r_map_delete_points before icp4r_map_build");
    if (n == 0) return ICP4R_OK;
    const void* dreq = nullptr;
    CKS(stage_points(c, c->d_q, xyzw, (size_t)n, mem, &dreq));
    int deleted = 0;
    CKS(map_delete_points(c, mp, static_cast<const float4*>(dreq), n, &deleted));
    if (deleted > 0) CKS(map_rebuild_grid(c, mp));
    if (n_deleted) *n_deleted = deleted;
    return ICP4R_OK;
}

int icp4r_map_points(icp4r_handle h, int mem, float* xyzw_out, uint8_t* valid_out, int32_t cap) {
    HCHECK(h);
    if (bad_mem(mem) || cap < 0) return fail(c, ICP4R_ERR_INVALID, "icp4r_map_points: bad arguments");
    const int n = std::min(cap, c->map.m);
    const cudaMemcpyKind kind = mem == ICP4R_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (n > 0 && xyzw_out) CK(cudaMemcpyAsync(xyzw_out, c->map.pts.p, (size_t)n * sizeof(float4), kind, c->stream));
    if (n > 0 && valid_out) CK(cudaMemcpyAsync(valid_out, c->map.valid.p, (size_t)n, kind, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return ICP4R_OK;
}

// ---- registration -----------------------------------------------------------------------------------------

static int k_of(const icp4r_opts* o) {
    return (o->residual == ICP4R_P2P_SVD || o->residual == ICP4R_P2P_GN || o->residual == ICP4R_GICP) ? 1 : (o->residual == ICP4R_P2LINE ? 2 : (o->residual == ICP4R_P2PLANE_3PT ? 3 : (o->k > 0 ? o->k : 5)));
}

// dumps requested with host pointers are produced on the device and copied back afterwards
struct DumpStage {
    icp4r_dump dev{nullptr, nullptr, nullptr};
    const icp4r_dump* user = nullptr;
    size_t pose_b = 0, acc_b = 0, idx_b = 0;
    bool host = false;
};

static int dump_prepare(Ctx* c, const icp4r_dump* dump, int mem, int n, const icp4r_opts* o, DumpStage& ds) {
    ds.user = dump;
    if (!dump) return ICP4R_OK;
    const size_t it = (size_t)std::max(o->max_iterations, 0);
    ds.pose_b = it * 16 * sizeof(double);
    ds.acc_b = it * ICP4R_ACC_LEN * sizeof(double);
    ds.idx_b = it * (size_t)n * k_of(o) * sizeof(int32_t);
    if (mem == ICP4R_DEVICE) {
        ds.dev = *dump;
        return ICP4R_OK;
    }
    ds.host = true;
    if (dump->pose) {
        CKS(reserve(c, c->d_dump_pose, ds.pose_b));
        ds.dev.pose = c->d_dump_pose.as<double>();
    }
    if (dump->acc) {
        CKS(reserve(c, c->d_dump_acc, ds.acc_b));
        ds.dev.acc = c->d_dump_acc.as<double>();
    }
    if (dump->idx) {
        CKS(reserve(c, c->d_dump_idx, ds.idx_b));
        ds.dev.idx = c->d_dump_idx.as<int32_t>();
    }
    if (ds.dev.pose && ds.pose_b) CK(cudaMemsetAsync(ds.dev.pose, 0, ds.pose_b, c->stream));
    if (ds.dev.acc && ds.acc_b) CK(cudaMemsetAsync(ds.dev.acc, 0, ds.acc_b, c->stream));
    if (ds.dev.idx && ds.idx_b) CK(cudaMemsetAsync(ds.dev.idx, 0xff, ds.idx_b, c->stream));
    return ICP4R_OK;
}

static int dump_finish(Ctx* c, DumpStage& ds) {
    if (!ds.user || !ds.host) return ICP4R_OK;
    if (ds.dev.pose && ds.pose_b) CK(cudaMemcpyAsync(ds.user->pose, ds.dev.pose, ds.pose_b, cudaMemcpyDeviceToHost, c->stream));
    if (ds.dev.acc && ds.acc_b) CK(cudaMemcpyAsync(ds.user->acc, ds.dev.acc, ds.acc_b, cudaMemcpyDeviceToHost, c->stream));
    if (ds.dev.idx && ds.idx_b) CK(cudaMemcpyAsync(ds.user->idx, ds.dev.idx, ds.idx_b, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return ICP4R_OK;
}

int icp4r_register_map(icp4r_handle h, const float* src, int32_t n, int mem, const icp4r_opts* opts, double T_out[16],
                       icp4r_result* res, const icp4r_dump* dump) {
    HCHECK(h);
    if (!opts || n < 0 || (n > 0 && !src) || bad_mem(mem)) return fail(c, ICP4R_ERR_INVALID, "icp4r_register_map: bad arguments");
    const void* dsrc = nullptr;
    CKS(stage_points(c, c->d_src, src, (size_t)n, mem, &dsrc));
    DumpStage ds;
    CKS(dump_prepare(c, dump, mem, n, opts, ds));
    CKS(register_against_map(c, c->map, static_cast<const float4*>(dsrc), n, opts, -1, 0.f, 0.f, T_out, res, dump ? &ds.dev : nullptr));
    return dump_finish(c, ds);
}

// One frame of scan-to-map odometry in one call: register against the map, move the scan with the estimated pose,
// append it. The scan crosses the bus once and the transformed points never leave the device.
int icp4r_odometry_step(icp4r_handle h, const float* scan, int32_t n, int mem, const icp4r_opts* opts, int downsample_on, double T_io[16],
                        icp4r_result* res) {
    HCHECK(h);
    if (!opts || !T_io || n < 0 || (n > 0 && !scan) || bad_mem(mem)) return fail(c, ICP4R_ERR_INVALID, "icp4r_odometry_step: bad arguments");
    if (downsample_on) return fail(c, ICP4R_ERR_UNSUPPORTED, "icp4r_odometry_step: use icp4r_map_add_points for down-sampled insertion");
    Map& mp = c->map;
    const void* dsrc = nullptr;
    CKS(stage_points(c, c->d_src, scan, (size_t)n, mem, &dsrc));
    icp4r_result r;
    std::memset(&r, 0, sizeof(r));
    if (mp.built && mp.m > 0) {
        icp4r_opts o = *opts;
        std::memcpy(o.T0, T_io, sizeof(o.T0));
        CKS(register_against_map(c, mp, static_cast<const float4*>(dsrc), n, &o, -1, 0.f, 0.f, T_io, &r, nullptr));
    } else {
        r.converged = 1;  // first frame: nothing to register against, the prior pose stands
    }
    if (res) *res = r;
    if (n == 0) return ICP4R_OK;
    // p_w = R p + t straight into the map's point array (pointAssociateToMap + Add_Points(.., false))
    CKS(map_reserve(c, mp, mp.m + n));
    CKS(transform_points(c, T_io, static_cast<const float4*>(dsrc), n, mp.pts.as<float4>() + mp.m));
    CK(cudaMemsetAsync(mp.valid.as<uint8_t>() + mp.m, 1, (size_t)n, c->stream));
    CK(cudaMemsetAsync(mp.userdel.as<uint8_t>() + mp.m, 0, (size_t)n, c->stream));
    bool merged = false;
    CKS(map_append_incremental(c, mp, n, &merged));
    mp.m += n;
    if (!merged) CKS(map_rebuild_grid(c, mp));
    return ICP4R_OK;
}

int icp4r_register_map_batch(icp4r_handle h, const float* src, const int32_t* off, int32_t n_scans, int mem, const icp4r_opts* opts,
                             const double* T0s, double* T_out, icp4r_result* res) {
    HCHECK(h);
    if (!opts || n_scans < 0 || bad_mem(mem) || (n_scans > 0 && (!off || !T_out || !res)))
        return fail(c, ICP4R_ERR_INVALID, "icp4r_register_map_batch: bad arguments");
    if (n_scans == 0) return ICP4R_OK;
    // offsets, initial guesses, poses and results are small and always host memory; only the points follow `mem`
    for (int i = 0; i < n_scans; ++i)
        if (off[i + 1] < off[i]) return fail(c, ICP4R_ERR_INVALID, "offsets must be non-decreasing (scan %d)", i);
    const size_t total = (size_t)(off[n_scans] - off[0]);
    if (total > 0 && !src) return fail(c, ICP4R_ERR_INVALID, "null cloud pointer");
    const void* dsrc = nullptr;
    CKS(stage_points(c, c->d_src, src, (size_t)off[n_scans], mem, &dsrc));
    return register_scans_against_map(c, c->map, static_cast<const float4*>(dsrc), off, n_scans, opts, T0s, T_out, res);
}

// registration against a transient target (the handle's c->tmp map): an explicit cloud (tgt, m) or, with idx != nullptr,
// the subset idx[0..n_idx) of the handle's persistent map gathered on the device
static int register_transient(Ctx* c, const float* src, int32_t n, const float* tgt, int32_t m, int mem, const int32_t* idx, int32_t n_idx,
                              const icp4r_opts* opts, double T_out[16], icp4r_result* res, const icp4r_dump* dump) {
    // transient index over the target, rebuilt per call like PCL's setInputTarget kd-tree
    Map& mp = c->tmp;
    mp.m = 0;
    mp.built = false;
    mp.user_cell = 0.f;
    mp.hint_cell = 0.f;
    if (idx != nullptr) m = n_idx;
    {   // a small target that serves one registration: the volume-estimated cell is good enough, the occupancy-based
        // refinement (a device round trip and up to two more sorts) costs more than it saves
        const char* e = std::getenv("ICP4R_TMP_REFINE");
        mp.quick_build = m <= 16384 && !(e && e[0] == '1');
    }
    if (idx == nullptr) {
        CKS(set_points(c, mp, tgt, m, mem, 0));
    } else {
        // Sub-maps of consecutive frames look alike (the sector moves a little, the map grows a little): the cell size
        // the occupancy refinement found for one serves the next ones as it is — one sort instead of up to three and no
        // look at the occupancy from the host. Refined again when the sub-map's size drifted by a quarter or after 64 uses.
        // (The MAP's own cell size is no substitute: a growing map keeps the geometry of its first frames while its
        // density rises a hundredfold, and the covariance pass over a sub-map with such coarse cells scans hundreds of
        // candidates per point.)
        const bool reuse = c->submap_cell > 0.f && c->submap_uses < 64 && (double)m > 0.75 * c->submap_m && (double)m < 1.33 * c->submap_m;
        if (reuse) {
            mp.hint_cell = c->submap_cell;
            mp.quick_build = true;
            ++c->submap_uses;
        } else {
            mp.quick_build = false;
            c->submap_cell = -1.f;  // recorded after the build below
        }
        CKS(map_reserve(c, mp, m));
        const void* didx = nullptr;
        CKS(stage_in(c, c->d_idx, idx, (size_t)m * sizeof(int32_t), mem, &didx));
        if (m > 0) {
            gather_subset_kernel<<<(m + 255) / 256, 256, 0, c->stream>>>(c->map.pts.as<float4>(), c->map.m, static_cast<const int32_t*>(didx), m,
                                                                        mp.pts.as<float4>());
            c->launches += 1;
            CK(cudaMemsetAsync(mp.valid.p, 1, (size_t)m, c->stream));
            CK(cudaMemsetAsync(mp.userdel.p, 0, (size_t)m, c->stream));
        }
    }
    mp.m = m;
    CKS(map_rebuild_grid(c, mp));
    if (idx != nullptr && c->submap_cell < 0.f) {
        c->submap_cell = mp.grid.cell;  // (finer cells were measured slower: x0.5 -> +16 % per frame — 8x the cell table for every pass that walks it)
        c->submap_m = m;
        c->submap_uses = 0;
    }
    const void* dsrc = nullptr;
    CKS(stage_points(c, c->d_src, src, (size_t)n, mem, &dsrc));
    DumpStage ds;
    CKS(dump_prepare(c, dump, mem, n, opts, ds));
    CKS(register_against_map(c, mp, static_cast<const float4*>(dsrc), n, opts, -1, 0.f, 0.f, T_out, res, dump ? &ds.dev : nullptr));
    return dump_finish(c, ds);
}

int icp4r_register(icp4r_handle h, const float* src, int32_t n, const float* tgt, int32_t m, int mem, const icp4r_opts* opts,
                   double T_out[16], icp4r_result* res, const icp4r_dump* dump) {
    HCHECK(h);
    if (!opts || n < 0 || m < 0 || (n > 0 && !src) || (m > 0 && !tgt) || bad_mem(mem))
        return fail(c, ICP4R_ERR_INVALID, "icp4r_register: bad arguments");
    // Point-to-point kinds on clouds that fit one SM's shared memory: the whole registration — target grid, every
    // iteration, the solve, the fitness pass — runs as ONE launch of the resident kernel of icp4r_register_batch (a batch
    // of one pair): no per-iteration launch, no grid in global memory, no cross-block reduction.
    {
        const char* e = std::getenv("ICP4R_REGISTER_VIA_MAP");
        const bool kind_ok = opts->residual == ICP4R_P2P_SVD || opts->residual == ICP4R_P2P_GN;
        if (kind_ok && !dump && !c->profiling && !(e && e[0] == '1') && register_batch_fits(n, m) && opts->max_iterations >= 0) {
            struct Stage {
                int32_t soff[2], toff[2];
                double T[16];
                icp4r_result res;
            };
            Stage* hs = static_cast<Stage*>(c->h_pinned);
            hs->soff[0] = 0, hs->soff[1] = n, hs->toff[0] = 0, hs->toff[1] = m;
            const void *dsrc = nullptr, *dtgt = nullptr;
            CKS(stage_points(c, c->b_src, src, (size_t)n, mem, &dsrc));
            CKS(stage_points(c, c->b_tgt, tgt, (size_t)m, mem, &dtgt));
            CKS(reserve(c, c->b_soff, 4 * sizeof(int32_t)));
            CKS(reserve(c, c->b_T, 16 * sizeof(double)));
            CKS(reserve(c, c->b_res, sizeof(icp4r_result)));
            CK(cudaMemcpyAsync(c->b_soff.p, hs->soff, 4 * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            CKS(register_batch(c, static_cast<const float4*>(dsrc), c->b_soff.as<int32_t>(), static_cast<const float4*>(dtgt),
                               c->b_soff.as<int32_t>() + 2, 1, n, m, opts, c->b_T.as<double>(), c->b_res.as<icp4r_result>()));
            CK(cudaMemcpyAsync(hs->T, c->b_T.p, 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(&hs->res, c->b_res.p, sizeof(icp4r_result), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            if (T_out) std::memcpy(T_out, hs->T, sizeof(hs->T));
            if (res) *res = hs->res;
            return ICP4R_OK;
        }
    }
    return register_transient(c, src, n, tgt, m, mem, nullptr, 0, opts, T_out, res, dump);
}

int icp4r_register_submap(icp4r_handle h, const float* src, int32_t n, const int32_t* idx, int32_t n_idx, int mem, const icp4r_opts* opts,
                          double T_out[16], icp4r_result* res) {
    HCHECK(h);
    if (!opts || n < 0 || n_idx < 0 || (n > 0 && !src) || (n_idx > 0 && !idx) || bad_mem(mem))
        return fail(c, ICP4R_ERR_INVALID, "icp4r_register_submap: bad arguments");
    if (!c->map.built) return fail(c, ICP4R_ERR_STATE, "icp4r_register_submap: the handle has no map");
    static const int32_t none = 0;
    return register_transient(c, src, n, nullptr, 0, mem, n_idx > 0 ? idx : &none, n_idx, opts, T_out, res, nullptr);
}

int icp4r_register_batch(icp4r_handle h, const float* src, const int32_t* src_off, const float* tgt, const int32_t* tgt_off,
                         int32_t n_pairs, int mem, const icp4r_opts* opts, double* T_out, icp4r_result* res) {
    HCHECK(h);
    if (!opts || n_pairs < 0 || bad_mem(mem) || (n_pairs > 0 && (!src_off || !tgt_off || !T_out || !res)))
        return fail(c, ICP4R_ERR_INVALID, "icp4r_register_batch: bad arguments");
    if (n_pairs == 0) return ICP4R_OK;
    // offsets are needed on the host to size shared memory
    std::vector<int32_t> so(n_pairs + 1), to(n_pairs + 1);
    if (mem == ICP4R_DEVICE) {
        CK(cudaMemcpyAsync(so.data(), src_off, (size_t)(n_pairs + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(to.data(), tgt_off, (size_t)(n_pairs + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    } else {
        std::memcpy(so.data(), src_off, (size_t)(n_pairs + 1) * 4);
        std::memcpy(to.data(), tgt_off, (size_t)(n_pairs + 1) * 4);
    }
    int max_n = 0, max_m = 0;
    for (int i = 0; i < n_pairs; ++i) {
        const int a = so[i + 1] - so[i], b = to[i + 1] - to[i];
        if (a < 0 || b < 0) return fail(c, ICP4R_ERR_INVALID, "offsets must be non-decreasing (pair %d)", i);
        max_n = std::max(max_n, a);
        max_m = std::max(max_m, b);
    }
    const size_t ns = (size_t)so[n_pairs], nt = (size_t)to[n_pairs];
    if ((ns > 0 && !src) || (nt > 0 && !tgt)) return fail(c, ICP4R_ERR_INVALID, "null cloud pointer");
    // a strided input layout (icp4r_set_point_layout): both clouds are repacked on the device first; from here on the
    // POINTS are packed device rows (pmem), offsets and outputs keep following `mem`
    int pmem = mem;
    if (!packed_layout(c)) {
        const void *ps = nullptr, *pt = nullptr;
        CKS(stage_points(c, c->b_src, src, ns, mem, &ps));
        CKS(stage_points(c, c->b_tgt, tgt, nt, mem, &pt));
        src = static_cast<const float*>(ps);
        tgt = static_cast<const float*>(pt);
        pmem = ICP4R_DEVICE;
    }
    if (!register_batch_fits(max_n, max_m)) {
        // A pair that does not fit one SM's shared memory (n + m above ~12 k points): every pair of the batch goes through
        // the map path instead — target grid in global memory, one launch per iteration — one after the other. Same
        // results (both loops are checked against the same oracle), lower throughput; the resident kernel is the fast
        // path for radar-sized clouds.
        if (opts->residual != ICP4R_P2P_SVD && opts->residual != ICP4R_P2P_GN)
            return fail(c, ICP4R_ERR_UNSUPPORTED, "batched registration supports P2P_SVD and P2P_GN (got %d)", opts->residual);
        std::vector<double> Th((size_t)n_pairs * 16);
        std::vector<icp4r_result> Rh(n_pairs);
        const int keep_stride = c->pt_stride, keep_woff = c->pt_woff;
        if (pmem != mem) c->pt_stride = 16, c->pt_woff = 12;  // the rows are packed by now
        int rc_pairs = ICP4R_OK;
        for (int i = 0; i < n_pairs && rc_pairs == ICP4R_OK; ++i)
            rc_pairs = register_transient(c, src + 4 * (size_t)so[i], so[i + 1] - so[i], tgt + 4 * (size_t)to[i], to[i + 1] - to[i], pmem, nullptr, 0, opts,
                                          Th.data() + 16 * (size_t)i, &Rh[i], nullptr);
        c->pt_stride = keep_stride, c->pt_woff = keep_woff;
        CKS(rc_pairs);
        if (mem == ICP4R_DEVICE) {
            CK(cudaMemcpyAsync(T_out, Th.data(), Th.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(res, Rh.data(), Rh.size() * sizeof(icp4r_result), cudaMemcpyHostToDevice, c->stream));
            CK(cudaStreamSynchronize(c->stream));
        } else {
            std::memcpy(T_out, Th.data(), Th.size() * sizeof(double));
            std::memcpy(res, Rh.data(), Rh.size() * sizeof(icp4r_result));
        }
        return ICP4R_OK;
    }
    // Large host-resident batches: the clouds are copied in chunks on a copy stream that runs ahead of the kernels, so the
    // bus transfer of chunk k+1 overlaps the registration of chunk k (the clouds of a pair range are contiguous). Many
    // small chunks keep the part of the copy nobody can hide — the first chunk — short; their kernels alternate between
    // two compute streams so that the half-empty last wave of one chunk's persistent CTAs is filled by the next chunk's
    // instead of idling the SMs (measured with 8 chunks on one stream: 14.6 ms of the 127.6 ms C4 step were exposed).
    constexpr int NCHUNK = 32, EV_START = 32, EV_AUX = 33;
    if (mem == ICP4R_HOST && pmem == ICP4R_HOST && n_pairs >= 512 && (ns + nt) * sizeof(float4) >= ((size_t)32 << 20)) {
        const int nchunk = std::min(NCHUNK, n_pairs / 64);  // at least 64 pairs per chunk
        if (!c->copy_stream) {
            CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
            CK(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
            for (auto& e : c->copy_events) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        CKS(reserve(c, c->b_src, std::max<size_t>(ns, 1) * sizeof(float4)));
        CKS(reserve(c, c->b_tgt, std::max<size_t>(nt, 1) * sizeof(float4)));
        CKS(reserve(c, c->b_soff, (size_t)(n_pairs + 1) * 4));
        CKS(reserve(c, c->b_toff, (size_t)(n_pairs + 1) * 4));
        CKS(reserve(c, c->b_T, (size_t)n_pairs * 16 * sizeof(double)));
        CKS(reserve(c, c->b_res, (size_t)n_pairs * sizeof(icp4r_result)));
        CK(cudaMemcpyAsync(c->b_soff.p, src_off, (size_t)(n_pairs + 1) * 4, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->b_toff.p, tgt_off, (size_t)(n_pairs + 1) * 4, cudaMemcpyHostToDevice, c->stream));
        // the staging buffers may still be read by earlier work of this handle: copies and the second compute stream start after it
        CK(cudaEventRecord(c->copy_events[EV_START], c->stream));
        CK(cudaStreamWaitEvent(c->copy_stream, c->copy_events[EV_START], 0));
        CK(cudaStreamWaitEvent(c->aux_stream, c->copy_events[EV_START], 0));
        const int per = (n_pairs + nchunk - 1) / nchunk;
        for (int k = 0; k < nchunk; ++k) {
            const int p0 = std::min(k * per, n_pairs), p1 = std::min(p0 + per, n_pairs);
            if (p1 > p0) {
                const size_t s0 = (size_t)so[p0], s1 = (size_t)so[p1], t0 = (size_t)to[p0], t1 = (size_t)to[p1];
                if (s1 > s0) CK(cudaMemcpyAsync(c->b_src.as<float4>() + s0, src + 4 * s0, (s1 - s0) * sizeof(float4), cudaMemcpyHostToDevice, c->copy_stream));
                if (t1 > t0) CK(cudaMemcpyAsync(c->b_tgt.as<float4>() + t0, tgt + 4 * t0, (t1 - t0) * sizeof(float4), cudaMemcpyHostToDevice, c->copy_stream));
            }
            CK(cudaEventRecord(c->copy_events[k], c->copy_stream));
        }
        cudaStream_t main_stream = c->stream;
        int rc_chunks = ICP4R_OK;
        for (int k = 0; k < nchunk && rc_chunks == ICP4R_OK; ++k) {
            const int p0 = std::min(k * per, n_pairs), p1 = std::min(p0 + per, n_pairs);
            cudaStream_t cs = (k & 1) ? c->aux_stream : main_stream;
            if (cudaStreamWaitEvent(cs, c->copy_events[k], 0) != cudaSuccess) rc_chunks = ICP4R_ERR_CUDA;
            if (p1 > p0 && rc_chunks == ICP4R_OK) {
                c->stream = cs;  // register_batch launches on the handle's stream
                rc_chunks = register_batch(c, c->b_src.as<float4>(), c->b_soff.as<int32_t>() + p0, c->b_tgt.as<float4>(), c->b_toff.as<int32_t>() + p0,
                                           p1 - p0, max_n, max_m, opts, c->b_T.as<double>() + (size_t)p0 * 16, c->b_res.as<icp4r_result>() + p0);
                c->stream = main_stream;
            }
        }
        CK(cudaEventRecord(c->copy_events[EV_AUX], c->aux_stream));
        CK(cudaStreamWaitEvent(c->stream, c->copy_events[EV_AUX], 0));
        if (rc_chunks != ICP4R_OK) {
            cudaStreamSynchronize(c->stream);
            return rc_chunks == ICP4R_ERR_CUDA ? fail(c, ICP4R_ERR_CUDA, "icp4r_register_batch: stream wait failed") : rc_chunks;
        }
        CK(cudaMemcpyAsync(T_out, c->b_T.p, (size_t)n_pairs * 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(res, c->b_res.p, (size_t)n_pairs * sizeof(icp4r_result), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        return ICP4R_OK;
    }
    const void *dsrc = nullptr, *dtgt = nullptr, *dso = nullptr, *dto = nullptr;
    CKS(stage_in(c, c->b_src, src, ns * sizeof(float4), pmem, &dsrc));
    CKS(stage_in(c, c->b_tgt, tgt, nt * sizeof(float4), pmem, &dtgt));
    CKS(stage_in(c, c->b_soff, src_off, (size_t)(n_pairs + 1) * 4, mem, &dso));
    CKS(stage_in(c, c->b_toff, tgt_off, (size_t)(n_pairs + 1) * 4, mem, &dto));
    double* dT = T_out;
    icp4r_result* dR = res;
    if (mem == ICP4R_HOST) {
        CKS(reserve(c, c->b_T, (size_t)n_pairs * 16 * sizeof(double)));
        CKS(reserve(c, c->b_res, (size_t)n_pairs * sizeof(icp4r_result)));
        dT = c->b_T.as<double>();
        dR = c->b_res.as<icp4r_result>();
    }
    CKS(register_batch(c, static_cast<const float4*>(dsrc), static_cast<const int32_t*>(dso), static_cast<const float4*>(dtgt),
                       static_cast<const int32_t*>(dto), n_pairs, max_n, max_m, opts, dT, dR));
    if (mem == ICP4R_HOST) {
        CK(cudaMemcpyAsync(T_out, dT, (size_t)n_pairs * 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(res, dR, (size_t)n_pairs * sizeof(icp4r_result), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return ICP4R_OK;
}

// ---- sharding ---------------------------------------------------------------------------------------------

int icp4r_shard_unique_id(char id_out[128]) {
    if (!id_out) return ICP4R_ERR_INVALID;
    return shard_unique_id(id_out);
}

int icp4r_shard_init(icp4r_handle h, const char id[128], int rank, int world) {
    HCHECK(h);
    if (!id) return fail(c, ICP4R_ERR_INVALID, "null id");
    return shard_init(c, id, rank, world);
}

int icp4r_shard_ipc_export(icp4r_handle h, unsigned char handle_out[64]) {
    HCHECK(h);
    if (!handle_out) return fail(c, ICP4R_ERR_INVALID, "null output");
    return shard_ipc_export(c, handle_out);
}

int icp4r_shard_ipc_import(icp4r_handle h, const unsigned char* handles, int rank, int world) {
    HCHECK(h);
    if (!handles) return fail(c, ICP4R_ERR_INVALID, "null handles");
    return shard_ipc_import(c, handles, rank, world);
}

int icp4r_register_sharded(icp4r_handle h, const float* src, int32_t n, int mem, const icp4r_opts* opts, int axis, float slab_lo,
                           float slab_hi, double T_out[16], icp4r_result* res) {
    HCHECK(h);
    if (!opts || n < 0 || (n > 0 && !src) || bad_mem(mem) || axis < 0 || axis > 2)
        return fail(c, ICP4R_ERR_INVALID, "icp4r_register_sharded: bad arguments");
    const void* dsrc = nullptr;
    CKS(stage_points(c, c->d_src, src, (size_t)n, mem, &dsrc));
    return register_against_map(c, c->map, static_cast<const float4*>(dsrc), n, opts, axis, slab_lo, slab_hi, T_out, res, nullptr);
}

int icp4r_accumulate_slab(icp4r_handle h, const float* src, int32_t n, int mem, const icp4r_opts* opts, const double T[16], int axis,
                          float slab_lo, float slab_hi, double acc_out[ICP4R_ACC_LEN]) {
    HCHECK(h);
    if (!opts || !T || !acc_out || n < 0 || (n > 0 && !src) || bad_mem(mem) || axis < -1 || axis > 2)
        return fail(c, ICP4R_ERR_INVALID, "icp4r_accumulate_slab: bad arguments");
    const void* dsrc = nullptr;
    CKS(stage_points(c, c->d_src, src, (size_t)n, mem, &dsrc));
    return accumulate_slab(c, c->map, static_cast<const float4*>(dsrc), n, opts, T, axis, slab_lo, slab_hi, acc_out);
}

// ---- Doppler filter ---------------------------------------------------------------------------------------------

int icp4r_doppler_filter(icp4r_handle h, const float* xyziv, int32_t n, int mem, const icp4r_doppler_opts* opts, uint8_t* static_mask,
                         icp4r_doppler_result* res) {
    HCHECK(h);
    if (!opts || !res || n < 0 || (n > 0 && !xyziv) || bad_mem(mem)) return fail(c, ICP4R_ERR_INVALID, "icp4r_doppler_filter: bad arguments");
    const void* drec = nullptr;
    CKS(stage_in(c, c->d_src, xyziv, (size_t)n * 5 * sizeof(float), mem, &drec));
    uint8_t* dmask = static_mask;
    if (mem == ICP4R_HOST && static_mask) {
        CKS(reserve(c, c->d_found, (size_t)std::max(n, 1)));
        dmask = c->d_found.as<uint8_t>();
    }
    static_assert(sizeof(icp4r_doppler_result) == 56, "icp4r_doppler_result layout");
    CKS(doppler_filter(c, static_cast<const float*>(drec), n, opts->iterations, opts->seed, opts->sigma, opts->split, dmask, c->h_pinned));
    std::memcpy(res, c->h_pinned, sizeof(icp4r_doppler_result));
    if (mem == ICP4R_HOST && static_mask && n > 0) {
        CK(cudaMemcpyAsync(static_mask, dmask, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return ICP4R_OK;
}

int icp4r_doppler_static_points(icp4r_handle h, const float* xyziv, int32_t n, int mem, const icp4r_doppler_opts* opts, float* xyzw_out,
                                int32_t cap, int32_t* n_out, icp4r_doppler_result* res) {
    HCHECK(h);
    if (!opts || !res || !n_out || n < 0 || cap < 0 || (n > 0 && !xyziv) || (cap > 0 && !xyzw_out) || bad_mem(mem))
        return fail(c, ICP4R_ERR_INVALID, "icp4r_doppler_static_points: bad arguments");
    const void* drec = nullptr;
    CKS(stage_in(c, c->d_src, xyziv, (size_t)n * 5 * sizeof(float), mem, &drec));
    CKS(reserve(c, c->d_found, (size_t)std::max(n, 1)));
    uint8_t* dmask = c->d_found.as<uint8_t>();
    // the filter's result, the compaction and the count of static points are enqueued back to back: ONE host round trip
    CKS(doppler_filter(c, static_cast<const float*>(drec), n, opts->iterations, opts->seed, opts->sigma, opts->split, dmask, c->h_pinned, false));
    float4* dout = reinterpret_cast<float4*>(xyzw_out);
    if (mem == ICP4R_HOST) {
        CKS(reserve(c, c->gs_pts, (size_t)std::max(cap, 1) * sizeof(float4)));
        dout = c->gs_pts.as<float4>();
    }
    CKS(reserve(c, c->d_scratch, 4096));
    int* d_n = c->d_scratch.as<int>();
    compact_static_kernel<<<1, 1024, 0, c->stream>>>(static_cast<const float*>(drec), dmask, n, dout, cap, d_n);
    c->launches += 1;
    int* h_n = reinterpret_cast<int*>(static_cast<char*>(c->h_pinned) + 256);
    CK(cudaMemcpyAsync(h_n, d_n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::memcpy(res, c->h_pinned, sizeof(icp4r_doppler_result));
    *n_out = *h_n;
    if (mem == ICP4R_HOST && std::min(*n_out, cap) > 0) {
        CK(cudaMemcpyAsync(xyzw_out, dout, (size_t)std::min(*n_out, cap) * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return ICP4R_OK;
}

// ---- helpers ------------------------------------------------------------------------------------------------

}  // extern "C"

namespace icp4r {
int transform_points(Ctx* c, const double* T_host, const float4* d_in, int n, float4* d_out);
}

extern "C" int icp4r_transform_points(icp4r_handle h, const double T[16], const float* xyzw, int32_t n, int mem, float* xyzw_out) {
    HCHECK(h);
    if (!T || n < 0 || (n > 0 && (!xyzw || !xyzw_out)) || bad_mem(mem)) return fail(c, ICP4R_ERR_INVALID, "icp4r_transform_points: bad arguments");
    if (n == 0) return ICP4R_OK;
    const void* din = nullptr;
    CKS(stage_points(c, c->d_q, xyzw, (size_t)n, mem, &din));
    float4* dout = reinterpret_cast<float4*>(xyzw_out);
    if (mem == ICP4R_HOST) {
        CKS(reserve(c, c->d_src, (size_t)n * sizeof(float4)));
        dout = c->d_src.as<float4>();
    }
    CKS(transform_points(c, T, static_cast<const float4*>(din), n, dout));
    if (mem == ICP4R_HOST) {
        CK(cudaMemcpyAsync(xyzw_out, dout, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return ICP4R_OK;
}

extern "C" int icp4r_voxel_grid(icp4r_handle h, const float* xyzw, int32_t n, int mem, float leaf, float* xyzw_out, int32_t cap, int32_t* n_out) {
    HCHECK(h);
    if (!n_out || n < 0 || cap < 0 || (cap > 0 && !xyzw_out) || bad_mem(mem)) return fail(c, ICP4R_ERR_INVALID, "icp4r_voxel_grid: bad arguments");
    const float4* din = nullptr;
    const uint8_t* dvalid = nullptr;
    if (xyzw) {
        const void* p;
        CKS(stage_points(c, c->d_q, xyzw, (size_t)n, mem, &p));
        din = static_cast<const float4*>(p);
    } else {  // the handle's own map (deleted points skipped)
        if (!c->map.built) return fail(c, ICP4R_ERR_STATE, "icp4r_voxel_grid(NULL) before icp4r_map_build");
        din = c->map.pts.as<float4>();
        dvalid = c->map.valid.as<uint8_t>();
        n = c->map.m;
    }
    float4* dout = reinterpret_cast<float4*>(xyzw_out);
    if (mem == ICP4R_HOST) {
        CKS(reserve(c, c->vg_out, (size_t)std::max(cap, 1) * sizeof(float4)));
        dout = c->vg_out.as<float4>();
    }
    int cnt = 0;
    CKS(voxel_grid(c, din, dvalid, n, leaf, dout, cap, &cnt));
    if (mem == ICP4R_HOST && cap > 0 && cnt > 0) {
        CK(cudaMemcpyAsync(xyzw_out, dout, (size_t)std::min(cnt, cap) * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    *n_out = cnt;
    return ICP4R_OK;
}
