// comm.cu -- multi-GPU plumbing of stage 2: one process per GPU, one NCCL communicator per ctx.
//
// The reference has no multi-device path at all (its only hint is the commented-out "cross-device access is used
// for faster model averaging over pcie", DeepWalk.java:43).  Here the walk stage shards by walk id without any
// collective; the skip-gram stage on the large synthetic configs is data-parallel over corpus shards and exchanges
// the per-GPU embedding deltas over NVLink every few thousand sentences, combined per row by the number of ranks that
// touched the row (sgns.cu calls dge_comm_* below; DESIGN.md 3.4 for why not the plain sum).
//
// NCCL is bound lazily with dlopen: libdge.so has no link-time dependency on it, a single-GPU host never loads
// it, and inside a process that already carries a libnccl.so.2 (e.g. torch.distributed in bench.py) the same
// copy is reused.
#include "dge_internal.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>

struct dge_nccl_api {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};
static dge_nccl_api g_nccl;

static const char *nccl_load() {
    if (g_nccl.handle) return nullptr;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return "cannot dlopen libnccl.so.2";
#define DGE_SYM(field, sym)                                         \
    *(void **)(&g_nccl.field) = dlsym(h, sym);                      \
    if (!g_nccl.field) { dlclose(h); return "libnccl lacks " sym; }
    DGE_SYM(GetUniqueId, "ncclGetUniqueId");
    DGE_SYM(CommInitRank, "ncclCommInitRank");
    DGE_SYM(CommDestroy, "ncclCommDestroy");
    DGE_SYM(AllReduce, "ncclAllReduce");
    DGE_SYM(Broadcast, "ncclBroadcast");
    DGE_SYM(GetErrorString, "ncclGetErrorString");
    DGE_SYM(GetVersion, "ncclGetVersion");
#undef DGE_SYM
    g_nccl.handle = h;
    return nullptr;
}

#define DGE_NCCL(ctx, expr)                                                                                  \
    do {                                                                                                     \
        ncclResult_t _r = (expr);                                                                            \
        if (_r != ncclSuccess)                                                                               \
            return dge_fail((ctx), DGE_E_COMM, std::string(#expr) + ": " + g_nccl.GetErrorString(_r));      \
    } while (0)

// ---- internal entry points used by sgns.cu
int dge_comm_allreduce_sum_f32(dge_ctx *ctx, float *buf, size_t n) {
    if (!ctx->comm || ctx->world <= 1) return DGE_OK;
    // NCCL element counts are size_t; split anyway into <= 2^30-element calls to bound the staging NCCL allocates
    const size_t chunk = (size_t)1 << 30;
    for (size_t off = 0; off < n; off += chunk) {
        size_t len = n - off < chunk ? n - off : chunk;
        DGE_NCCL(ctx, g_nccl.AllReduce(buf + off, buf + off, len, ncclFloat32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    }
    return DGE_OK;
}
int dge_comm_allreduce_sum_u64(dge_ctx *ctx, unsigned long long *buf, size_t n) {
    if (!ctx->comm || ctx->world <= 1) return DGE_OK;
    DGE_NCCL(ctx, g_nccl.AllReduce(buf, buf, n, ncclUint64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    return DGE_OK;
}
int dge_comm_allreduce_max_u64(dge_ctx *ctx, unsigned long long *buf, size_t n) {
    if (!ctx->comm || ctx->world <= 1) return DGE_OK;
    DGE_NCCL(ctx, g_nccl.AllReduce(buf, buf, n, ncclUint64, ncclMax, (ncclComm_t)ctx->comm, ctx->stream));
    return DGE_OK;
}

extern "C" {

int dge_comm_unique_id(void *id, size_t bytes) {
    if (!id || bytes < sizeof(ncclUniqueId))
        return dge_fail(nullptr, DGE_E_INVALID, "dge_comm_unique_id: buffer must hold DGE_COMM_ID_BYTES bytes");
    const char *err = nccl_load();
    if (err) return dge_fail(nullptr, DGE_E_COMM, std::string("dge_comm_unique_id: ") + err);
    ncclUniqueId u;
    DGE_NCCL(nullptr, g_nccl.GetUniqueId(&u));
    memset(id, 0, bytes);
    memcpy(id, &u, sizeof(u));
    return DGE_OK;
}

int dge_comm_init(dge_ctx *ctx, int rank, int world, const void *id, size_t bytes) {
    if (!ctx) return dge_fail(nullptr, DGE_E_INVALID, "dge_comm_init: ctx is NULL");
    if (world < 1 || rank < 0 || rank >= world || !id || bytes < sizeof(ncclUniqueId))
        return dge_fail(ctx, DGE_E_INVALID, "dge_comm_init: bad rank / world / id");
    if (ctx->comm) return dge_fail(ctx, DGE_E_INVALID, "dge_comm_init: ctx already has a communicator");
    const char *err = nccl_load();
    if (err) return dge_fail(ctx, DGE_E_COMM, std::string("dge_comm_init: ") + err);
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t comm = nullptr;
    DGE_NCCL(ctx, g_nccl.CommInitRank(&comm, world, u, rank));
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    return DGE_OK;
}

int dge_comm_shape(const dge_ctx *ctx, int *rank, int *world) {
    if (!ctx) return dge_fail(nullptr, DGE_E_INVALID, "dge_comm_shape: ctx is NULL");
    if (rank) *rank = ctx->rank;
    if (world) *world = ctx->world;
    return DGE_OK;
}

void dge_comm_destroy(dge_ctx *ctx) {
    if (!ctx || !ctx->comm) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
    ctx->rank = 0;
    ctx->world = 1;
}

int dge_comm_nccl_version(void) {
    if (nccl_load()) return 0;
    int v = 0;
    g_nccl.GetVersion(&v);
    return v;
}

} // extern "C"
