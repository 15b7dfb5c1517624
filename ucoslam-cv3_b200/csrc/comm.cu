// comm.cu — the one real exchange step of the path: a communicator for the reduced-Hessian all-reduce of sharded bundle
// adjustment (BASELINE config 5) and the top-k merge of a row-sharded descriptor map (config 4).
//
// One process per GPU; the process group itself (rendezvous, rank numbering) belongs to the caller (torch.distributed under
// torchrun, MPI, ...).  The caller has rank 0 produce a 128-byte id with uco_b200_comm_unique_id, distributes it by whatever
// channel it owns, and every rank calls uco_b200_comm_create.  Collectives run on the context's stream over NCCL (NVLink 5 /
// NVSwitch on a B200 node).  NCCL is bound at run time (dlopen of libnccl.so.2: under Python that is the copy torch already
// loaded), so the library has no link-time dependency on it and single-GPU users never touch it.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <mutex>
#include <string>

namespace {
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) {
            api.err = std::string("dlopen(libnccl.so.2) failed: ") + (dlerror() ? dlerror() : "?");
            return;
        }
        auto sym = [&](const char* n) {
            void* p = dlsym(h, n);
            if (!p) api.err = std::string("libnccl: missing symbol ") + n;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.ok = api.err.empty();
    });
    return api;
}
}  // namespace

struct uco_b200_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    uco_b200_ctx* ctx = nullptr;
};

int uco_comm_rank(const uco_b200_comm* c) { return c ? c->rank : 0; }
int uco_comm_world(const uco_b200_comm* c) { return c ? c->world : 1; }

// op: 0 = f64 sum, 1 = f64 max, 2 = u8 sum, 3 = i32 sum; in place when send == recv
int uco_comm_allreduce(uco_b200_comm* c, const void* send, void* recv, size_t count, int op, cudaStream_t s) {
    if (!c || c->world == 1) {
        if (send != recv) {
            const size_t es = op == 2 ? 1 : (op == 3 ? 4 : 8);
            cudaError_t e = cudaMemcpyAsync(recv, send, count * es, cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) return UCO_E_CUDA;
        }
        return UCO_OK;
    }
    NcclApi& api = nccl();
    const ncclDataType_t dt = op == 2 ? ncclUint8 : (op == 3 ? ncclInt32 : ncclFloat64);
    const ncclRedOp_t ro = op == 1 ? ncclMax : ncclSum;
    ncclResult_t r = api.AllReduce(send, recv, count, dt, ro, c->comm, s);
    if (r != ncclSuccess) return uco_fail(c->ctx, UCO_E_CUDA, "ncclAllReduce -> %s", api.GetErrorString(r));
    c->ctx->launches++;
    return UCO_OK;
}

// every rank contributes `bytes` bytes; recv holds world * bytes in rank order
int uco_comm_allgather(uco_b200_comm* c, const void* send, void* recv, size_t bytes, cudaStream_t s) {
    if (!c || c->world == 1) {
        if (send != recv && cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, s) != cudaSuccess) return UCO_E_CUDA;
        return UCO_OK;
    }
    NcclApi& api = nccl();
    ncclResult_t r = api.AllGather(send, recv, bytes, ncclUint8, c->comm, s);
    if (r != ncclSuccess) return uco_fail(c->ctx, UCO_E_CUDA, "ncclAllGather -> %s", api.GetErrorString(r));
    c->ctx->launches++;
    return UCO_OK;
}

extern "C" {

int uco_b200_comm_unique_id(uint8_t* id128) {
    if (!id128) return UCO_E_INVALID;
    NcclApi& api = nccl();
    if (!api.ok) return UCO_E_INVALID;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (api.GetUniqueId(&id) != ncclSuccess) return UCO_E_CUDA;
    memcpy(id128, &id, 128);
    return UCO_OK;
}

int uco_b200_comm_create(uco_b200_ctx* ctx, const uint8_t* id128, int rank, int world, uco_b200_comm** out) {
    if (!ctx || !out) return UCO_E_INVALID;
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return uco_fail(ctx, UCO_E_INVALID, "comm_create: rank %d of %d", rank, world);
    cudaSetDevice(ctx->device);
    uco_b200_comm* c = new uco_b200_comm();
    c->rank = rank;
    c->world = world;
    c->ctx = ctx;
    if (world > 1) {
        if (!id128) { delete c; return uco_fail(ctx, UCO_E_INVALID, "comm_create: null id"); }
        NcclApi& api = nccl();
        if (!api.ok) { delete c; return uco_fail(ctx, UCO_E_INVALID, "comm_create: %s", api.err.c_str()); }
        ncclUniqueId id;
        memcpy(&id, id128, 128);
        ncclResult_t r = api.CommInitRank(&c->comm, world, id, rank);
        if (r != ncclSuccess) { delete c; return uco_fail(ctx, UCO_E_CUDA, "ncclCommInitRank -> %s", api.GetErrorString(r)); }
    }
    *out = c;
    return UCO_OK;
}

void uco_b200_comm_destroy(uco_b200_comm* c) {
    if (!c) return;
    if (c->comm) nccl().CommDestroy(c->comm);
    delete c;
}

}  // extern "C"
