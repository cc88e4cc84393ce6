// shard.cu — cross-rank sum of the per-iteration accumulators for slab-sharded maps.
// One process per GPU; the only collective on the path is an all-reduce of the 29 fp64 accumulators
// (21 J^T J + 6 J^T r + cost + count) per iteration.  NCCL is bound at run time with dlopen so the library has
// no link-time dependency and shares the copy the host program (torch) already loaded.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "ctx.h"

namespace icp4r {

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi& api() {
    static NcclApi a;
    if (a.lib) return a;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) return a;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.lib, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.lib, "ncclCommInitRank"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(a.lib, "ncclAllReduce"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.lib, "ncclCommDestroy"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(a.lib, "ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy && a.GetErrorString;
    return a;
}
}  // namespace

int shard_unique_id(char id_out[128]) {
    NcclApi& a = api();
    if (!a.ok) return ICP4R_ERR_NCCL;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    if (a.GetUniqueId(&id) != ncclSuccess) return ICP4R_ERR_NCCL;
    std::memcpy(id_out, &id, 128);
    return ICP4R_OK;
}

int shard_init(Ctx* c, const char id_in[128], int rank, int world) {
    NcclApi& a = api();
    if (!a.ok) return fail(c, ICP4R_ERR_NCCL, "libnccl.so.2 could not be loaded");
    if (world < 1 || rank < 0 || rank >= world) return fail(c, ICP4R_ERR_INVALID, "bad rank/world %d/%d", rank, world);
    if (c->nccl_comm) {
        a.CommDestroy(static_cast<ncclComm_t>(c->nccl_comm));
        c->nccl_comm = nullptr;
    }
    ncclUniqueId id;
    std::memcpy(&id, id_in, 128);
    ncclComm_t comm;
    const ncclResult_t r = a.CommInitRank(&comm, world, id, rank);
    if (r != ncclSuccess) return fail(c, ICP4R_ERR_NCCL, "ncclCommInitRank: %s", a.GetErrorString(r));
    c->nccl_comm = comm;
    c->rank = rank;
    c->world = world;
    return ICP4R_OK;
}

void shard_destroy(Ctx* c) {
    if (c->nccl_comm && api().ok) api().CommDestroy(static_cast<ncclComm_t>(c->nccl_comm));
    c->nccl_comm = nullptr;
}

// ---- peer-memory flavour: exchange buffers mapped across processes with CUDA IPC ------------------------------------
int shard_ipc_export(Ctx* c, unsigned char out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    // the exchange counts its rounds on every rank (Xch.seq); exporting again would zero this rank's count while the
    // peers keep theirs and every later exchange would time out: one export per handle
    if (c->d_xch.p != nullptr)
        return fail(c, ICP4R_ERR_STATE, "icp4r_shard_ipc_export was already called on this handle (create a new handle to re-join)");
    CKS(reserve(c, c->d_xch, sizeof(Xch)));
    CK(cudaMemsetAsync(c->d_xch.p, 0, sizeof(Xch), c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaIpcMemHandle_t hnd;
    CK(cudaIpcGetMemHandle(&hnd, c->d_xch.p));
    std::memcpy(out, &hnd, 64);
    return ICP4R_OK;
}

int shard_ipc_import(Ctx* c, const unsigned char* handles, int rank, int world) {
    if (world < 1 || world > XCH_MAXW || rank < 0 || rank >= world)
        return fail(c, ICP4R_ERR_INVALID, "peer exchange supports 1..%d ranks (got rank %d of %d)", XCH_MAXW, rank, world);
    if (!c->d_xch.p) return fail(c, ICP4R_ERR_STATE, "icp4r_shard_ipc_export has not been called on this handle");
    XchTable t;
    std::memset(&t, 0, sizeof(t));
    t.rank = rank;
    t.world = world;
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            t.peer[r] = c->d_xch.as<Xch>();
            continue;
        }
        cudaIpcMemHandle_t hnd;
        std::memcpy(&hnd, handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(c, ICP4R_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        c->xch_peers[r] = p;
        t.peer[r] = static_cast<Xch*>(p);
    }
    CKS(reserve(c, c->d_xt, sizeof(XchTable)));
    CK(cudaMemcpyAsync(c->d_xt.p, &t, sizeof(t), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->rank = rank;
    c->world = world;
    c->xch_ready = true;
    return ICP4R_OK;
}

void shard_ipc_close(Ctx* c) {
    for (int r = 0; r < XCH_MAXW; ++r)
        if (c->xch_peers[r]) {
            cudaIpcCloseMemHandle(c->xch_peers[r]);
            c->xch_peers[r] = nullptr;
        }
    c->xch_ready = false;
}

int shard_allreduce(Ctx* c, double* d_buf, int count) {
    if (c->world <= 1 && !c->nccl_comm) return ICP4R_OK;  // single rank: the sum is the local value
    NcclApi& a = api();
    if (!a.ok || !c->nccl_comm) return fail(c, ICP4R_ERR_STATE, "icp4r_shard_init has not been called");
    const ncclResult_t r = a.AllReduce(d_buf, d_buf, (size_t)count, ncclDouble, ncclSum, static_cast<ncclComm_t>(c->nccl_comm), c->stream);
    if (r != ncclSuccess) return fail(c, ICP4R_ERR_NCCL, "ncclAllReduce: %s", a.GetErrorString(r));
    return ICP4R_OK;
}

}  // namespace icp4r
