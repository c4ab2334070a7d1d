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


// all-gather of n_per_rank u64 values per rank into host memory out[world * n_per_rank], as a sum all-reduce of a
// zero-padded vector (a few hundred bytes: shard sizes, parameter fingerprints, cudaIpc handles).  Collective.
int dge_comm_allgather_u64(dge_ctx *ctx, const unsigned long long *mine, int n_per_rank, unsigned long long *out) {
    const int world = ctx->comm ? ctx->world : 1, rank = ctx->comm ? ctx->rank : 0;
    const size_t total = (size_t)world * (size_t)n_per_rank;
    if (world <= 1) { memcpy(out, mine, sizeof(unsigned long long) * (size_t)n_per_rank); return DGE_OK; }
    unsigned long long *d = nullptr;
    DGE_CUDA(ctx, dge_malloc(ctx, &d, total));
    cudaMemsetAsync(d, 0, total * sizeof(unsigned long long), ctx->stream);
    cudaMemcpyAsync(d + (size_t)rank * n_per_rank, mine, sizeof(unsigned long long) * (size_t)n_per_rank, cudaMemcpyHostToDevice, ctx->stream);
    int rc = dge_comm_allreduce_sum_u64(ctx, d, total);
    cudaError_t e = cudaMemcpyAsync(out, d, total * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dge_free(ctx, d);
    if (rc != DGE_OK) return rc;
    if (e != cudaSuccess) return dge_fail(ctx, DGE_E_CUDA, std::string("dge_comm_allgather_u64: ") + cudaGetErrorString(e));
    return DGE_OK;
}

// Collective agreement on a status: every rank passes its local status (DGE_OK or a negative dge_status) and all
// ranks return the WORST one, so that an error on one rank (a failed cudaMalloc, a parameter mismatch) makes every
// rank leave together instead of leaving the others hung in the next collective.
int dge_comm_agree(dge_ctx *ctx, int local_status, const char *what) {
    if (!ctx->comm || ctx->world <= 1) return local_status;
    unsigned long long mine = (unsigned long long)(-(long long)local_status); // 0 = ok, larger = worse
    std::vector<unsigned long long> all((size_t)ctx->world);
    int rc = dge_comm_allgather_u64(ctx, &mine, 1, all.data());
    if (rc != DGE_OK) return rc;
    int worst = DGE_OK, who = -1;
    for (int r = 0; r < ctx->world; r++)
        if (-(int)all[r] < worst) { worst = -(int)all[r]; who = r; }
    if (worst != DGE_OK && local_status == DGE_OK)
        return dge_fail(ctx, worst, std::string(what) + ": rank " + std::to_string(who) + " failed with status " + std::to_string(worst) +
                                        "; every rank leaves the collective");
    return worst != DGE_OK ? local_status : DGE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Data-parallel skip-gram: the exchange of the embedding deltas.
//
// Every rank trains a slice of its own corpus shard on its replica `cur` of syn0 / syn1neg; `base` is the state all
// replicas started the slice from.  The exchange combines the per-rank deltas d_r = cur_r - base per ROW
//     base += sum_r d_r / div(row),   cur_r = base  on every rank
// with div chosen by the combine rule (dge.h DGE_COMBINE_*).
//
// Two implementations of the same arithmetic:
//  * PEER (default on one NVSwitch box): ONE kernel per exchange over NVLink peer memory.  The replicas are mapped
//    into every process (cudaIpc); rank q owns rows [q V / W, (q + 1) V / W) of both tables: it loads the W replicas
//    of a row (W - 1 of them remote, 128-bit loads over NVLink), forms the deltas against ITS slice of base (the only
//    copy of base that exists: V / W rows per table), applies the rule and stores the new row into every replica
//    (remote 128-bit stores) -- reduce-scatter, combine rule, all-gather and the base update fused, with no staging,
//    no separate delta / apply passes and no per-row statistics exchange.  NVLink traffic per rank and exchange:
//    (W - 1) / W of both tables in, the same out -- the minimum any all-reduce moves.  Two 8-byte NCCL all-reduces
//    are the cross-GPU barriers (all replicas trained / all replicas rewritten) and carry the error status.
//  * NCCL (fallback when peer mapping is unavailable; A/B baseline): delta pass, ncclAllReduce of both tables and of
//    the per-row statistics, apply pass; `base` is a full copy per table.
struct dge_dp {
    dge_ctx *ctx = nullptr;
    float *cur[2] = {nullptr, nullptr};   // this rank's replicas (the model's tables)
    float *base[2] = {nullptr, nullptr};  // PEER: rows [row_lo, row_hi); NCCL: all rows
    float *aux = nullptr;                 // NCCL: per row and table {contributors, sum_r |d_r|^2}
    unsigned long long *flag = nullptr;   // device word for the barrier all-reduce (carries the error status)
    float *peer[2][DGE_DP_MAX_WORLD];     // PEER: replica of every rank, mapped here
    int32_t V = 0, stride = 0, n4 = 0;
    int64_t row_lo = 0, row_hi = 0;
    int combine = DGE_COMBINE_ALIGNED;
    bool use_peer = false;
    float ms = 0.f;                       // device time spent in exchanges
    int exchanges = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

__device__ __forceinline__ float dp_div(int combine, int world, float contributors, float sq_of_sum, float sum_of_sq) {
    const float c = contributors > 1.f ? contributors : 1.f;
    switch (combine) {
        case DGE_COMBINE_MEAN: return (float)world;
        case DGE_COMBINE_CONTRIBUTORS: return c;
        case DGE_COMBINE_SQRT: return sqrtf(c);
        case DGE_COMBINE_ALIGNED: { // deltas that point the same way are averaged, orthogonal ones summed: 1 <= div <= contributors
            if (!(sum_of_sq > 0.f)) return 1.f;
            const float a = __fdiv_rn(sq_of_sum, sum_of_sq);
            return a > 1.f ? a : 1.f;
        }
        default: return 1.f; // DGE_COMBINE_SUM
    }
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float4 ld_sys4(const float4 *p) { // peer (or local) replica row: bypass L1, coherent at the owner's L2
    float4 r;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_sys4(float4 *p, const float4 &v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct dp_peers { float *cur[2][DGE_DP_MAX_WORLD]; };

// PEER exchange: one warp per (table, row) of this rank's row slice; lane l owns float4 slots l, l + 32, ... of the row.
template <int SLOTS> // float4 slots per lane: rows of up to 32 * SLOTS slots
__global__ void __launch_bounds__(256)
k_dp_exchange_peer(const dp_peers P, float *__restrict__ base0, float *__restrict__ base1, int64_t row_lo, int64_t row_hi,
                   int32_t stride, int32_t n4, int world, int combine) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t rows = row_hi - row_lo;
    for (int64_t u = warp; u < 2 * rows; u += n_warps) {
        const int t = u >= rows;
        const int64_t lr = t ? u - rows : u;                  // row within the slice
        const size_t off = (size_t)(row_lo + lr) * (size_t)stride;
        float4 *b = reinterpret_cast<float4 *>((t ? base1 : base0) + (size_t)lr * (size_t)stride);
        float4 bs[SLOTS], sum[SLOTS];
#pragma unroll
        for (int v = 0; v < SLOTS; v++) {
            const int q = lane + 32 * v;
            bs[v] = q < n4 ? b[q] : make_float4(0.f, 0.f, 0.f, 0.f);
            sum[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float sum_of_sq = 0.f, contributors = 0.f;
        for (int r0 = 0; r0 < world; r0 += 4) {               // four replicas in flight per lane
            float4 x[4][SLOTS];
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int v = 0; v < SLOTS; v++) {
                    const int q = lane + 32 * v;
                    x[j][v] = (r0 + j < world && q < n4) ? ld_sys4(reinterpret_cast<const float4 *>(P.cur[t][r0 + j] + off) + q) : bs[v];
                }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float sq = 0.f;
#pragma unroll
                for (int v = 0; v < SLOTS; v++) {
                    const float4 d = make_float4(x[j][v].x - bs[v].x, x[j][v].y - bs[v].y, x[j][v].z - bs[v].z, x[j][v].w - bs[v].w);
                    sum[v].x += d.x; sum[v].y += d.y; sum[v].z += d.z; sum[v].w += d.w;
                    sq += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
                }
                sq = warp_sum(sq);                              // |d_r|^2 of the whole row (0 for the padding replicas)
                sum_of_sq += sq;
                contributors += sq > 0.f ? 1.f : 0.f;
            }
        }
        float sq_of_sum = 0.f;
#pragma unroll
        for (int v = 0; v < SLOTS; v++) sq_of_sum += sum[v].x * sum[v].x + sum[v].y * sum[v].y + sum[v].z * sum[v].z + sum[v].w * sum[v].w;
        sq_of_sum = warp_sum(sq_of_sum);
        if (contributors == 0.f) continue;                     // nobody moved the row: every replica still equals base
        const float div = dp_div(combine, world, contributors, sq_of_sum, sum_of_sq);
#pragma unroll
        for (int v = 0; v < SLOTS; v++) {
            const int q = lane + 32 * v;
            if (q >= n4) continue;
            const float4 nv = make_float4(bs[v].x + __fdiv_rn(sum[v].x, div), bs[v].y + __fdiv_rn(sum[v].y, div),
                                          bs[v].z + __fdiv_rn(sum[v].z, div), bs[v].w + __fdiv_rn(sum[v].w, div));
            b[q] = nv;
            for (int r = 0; r < world; r++) st_sys4(reinterpret_cast<float4 *>(P.cur[t][r] + off) + q, nv);
        }
    }
    __threadfence_system();
}

// NCCL exchange, pass 1: cur -= base in place (the delta that is all-reduced), per-row statistics
__global__ void __launch_bounds__(256)
k_dp_delta(float *__restrict__ c0, const float *__restrict__ b0, float *__restrict__ c1, const float *__restrict__ b1, int32_t V,
           int32_t stride, int32_t n4, float *__restrict__ aux) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < 2 * (int64_t)V; u += n_warps) {
        const int t = u >= V;
        const size_t off = (size_t)(t ? u - V : u) * (size_t)stride;
        float4 *c = reinterpret_cast<float4 *>((t ? c1 : c0) + off);
        const float4 *b = reinterpret_cast<const float4 *>((t ? b1 : b0) + off);
        float sq = 0.f;
        for (int q = lane; q < n4; q += 32) {
            const float4 x = c[q], y = b[q];
            const float4 d = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
            c[q] = d;
            sq += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
        }
        sq = warp_sum(sq);
        if (lane == 0) { aux[2 * u] = sq > 0.f ? 1.f : 0.f; aux[2 * u + 1] = sq; }
    }
}
// pass 2: cur holds sum_r d_r, aux the summed statistics: base += sum / div; cur = base
__global__ void __launch_bounds__(256)
k_dp_apply(float *__restrict__ c0, float *__restrict__ b0, float *__restrict__ c1, float *__restrict__ b1, int32_t V, int32_t stride,
           int32_t n4, const float *__restrict__ aux, int combine, int world) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp; u < 2 * (int64_t)V; u += n_warps) {
        const int t = u >= V;
        const size_t off = (size_t)(t ? u - V : u) * (size_t)stride;
        float4 *c = reinterpret_cast<float4 *>((t ? c1 : c0) + off);
        float4 *b = reinterpret_cast<float4 *>((t ? b1 : b0) + off);
        const float contributors = aux[2 * u], sum_of_sq = aux[2 * u + 1];
        float sq = 0.f;
        for (int q = lane; q < n4; q += 32) { const float4 x = c[q]; sq += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w; }
        sq = warp_sum(sq);
        const float div = dp_div(combine, world, contributors, sq, sum_of_sq);
        for (int q = lane; q < n4; q += 32) {
            const float4 x = c[q], y = b[q];
            const float4 nv = make_float4(y.x + __fdiv_rn(x.x, div), y.y + __fdiv_rn(x.y, div), y.z + __fdiv_rn(x.z, div), y.w + __fdiv_rn(x.w, div));
            b[q] = nv; c[q] = nv;
        }
    }
}

// Cross-GPU barrier on the ctx stream that also agrees on an error flag: returns (on every rank alike) whether any
// rank raised it.  The host reads the flag back only when `check` is set (the stream stays asynchronous otherwise).
static int dp_barrier(dge_dp *dp, int local_error, bool check, int *any_error) {
    dge_ctx *ctx = dp->ctx;
    unsigned long long h = local_error ? 1ULL : 0ULL;
    cudaMemcpyAsync(dp->flag, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream);
    int rc = dge_comm_allreduce_max_u64(ctx, dp->flag, 1);
    if (rc != DGE_OK) return rc;
    if (check) {
        cudaError_t e = cudaMemcpyAsync(&h, dp->flag, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return dge_fail(ctx, DGE_E_CUDA, std::string("data-parallel barrier: ") + cudaGetErrorString(e));
        if (any_error) *any_error = h != 0;
    }
    return DGE_OK;
}

void dge_dp_release_cache(dge_ctx *ctx) {
    for (int r = 0; r < DGE_DP_MAX_WORLD; r++)
        if (ctx->dp_peer_mapped[r]) { cudaIpcCloseMemHandle(ctx->dp_peer_ptr[r]); ctx->dp_peer_mapped[r] = false; ctx->dp_peer_ptr[r] = nullptr; }
    if (ctx->dp_arena) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->dp_arena); ctx->dp_arena = nullptr; ctx->dp_arena_bytes = 0; }
}

// Collective.  The arena only grows; when ANY rank has to reallocate, every rank first drops its mappings of the peers'
// arenas (nobody frees memory another process still has mapped), then the owners reallocate.
int dge_dp_arena(dge_ctx *ctx, size_t bytes, float **out) {
    *out = nullptr;
    unsigned long long grow = bytes > ctx->dp_arena_bytes ? 1ULL : 0ULL;
    std::vector<unsigned long long> all((size_t)ctx->world);
    int rc = dge_comm_allgather_u64(ctx, &grow, 1, all.data());
    if (rc != DGE_OK) return rc;
    bool any = false;
    for (int r = 0; r < ctx->world; r++) any = any || all[r] != 0;
    int local = DGE_OK;
    if (any) {
        cudaStreamSynchronize(ctx->stream);
        for (int r = 0; r < DGE_DP_MAX_WORLD; r++)
            if (ctx->dp_peer_mapped[r]) { cudaIpcCloseMemHandle(ctx->dp_peer_ptr[r]); ctx->dp_peer_mapped[r] = false; ctx->dp_peer_ptr[r] = nullptr; }
        rc = dge_comm_agree(ctx, DGE_OK, "dge_dp_arena (mappings closed)"); // barrier: every mapping of every arena is gone
        if (rc != DGE_OK) return rc;
        if (grow) {
            if (ctx->dp_arena) { cudaFree(ctx->dp_arena); ctx->dp_arena = nullptr; ctx->dp_arena_bytes = 0; }
            void *q = nullptr;
            if (cudaMalloc(&q, bytes) != cudaSuccess)
                local = dge_fail(ctx, DGE_E_CUDA, std::string("dge_dp_arena: cudaMalloc of the replica arena failed (") + cudaGetErrorString(cudaGetLastError()) + ")");
            else { ctx->dp_arena = (float *)q; ctx->dp_arena_bytes = bytes; }
        }
        local = dge_comm_agree(ctx, local, "dge_dp_arena");
        if (local != DGE_OK) return local;
    }
    *out = ctx->dp_arena;
    return DGE_OK;
}

void dge_dp_end(dge_dp *dp) {
    if (!dp) return;
    dge_ctx *ctx = dp->ctx;
    cudaStreamSynchronize(ctx->stream);
    dge_free(ctx, dp->base[0]); dge_free(ctx, dp->base[1]); dge_free(ctx, dp->aux); dge_free(ctx, dp->flag);
    if (dp->e0) cudaEventDestroy(dp->e0);
    if (dp->e1) cudaEventDestroy(dp->e1);
    delete dp;
}

// Collective.  syn0 / syn1neg: this rank's replicas, identical on every rank at this point; they must come from
// cudaMalloc (not the stream-ordered pool) for the peer mapping.  transport: DGE_TRANSPORT_AUTO / _PEER / _NCCL.
int dge_dp_begin(dge_ctx *ctx, float *syn0, float *syn1neg, int32_t V, int32_t stride, int32_t n4, int combine, int transport,
                 dge_dp **out) {
    *out = nullptr;
    const int world = ctx->world, rank = ctx->rank;
    dge_dp *dp = new dge_dp();
    dp->ctx = ctx; dp->cur[0] = syn0; dp->cur[1] = syn1neg; dp->V = V; dp->stride = stride; dp->n4 = n4; dp->combine = combine;
    for (int r = 0; r < DGE_DP_MAX_WORLD; r++) dp->peer[0][r] = dp->peer[1][r] = nullptr;
    int local = DGE_OK;
    if (cudaEventCreate(&dp->e0) != cudaSuccess || cudaEventCreate(&dp->e1) != cudaSuccess || dge_malloc(ctx, &dp->flag, 1) != cudaSuccess)
        local = dge_fail(ctx, DGE_E_CUDA, "dge_dp_begin: event / flag allocation failed");
    // ---- peer mapping of the replica arenas (every rank must succeed, or all fall back to NCCL together).  syn0 / syn1neg
    // live in the rank's arena (dge_dp_arena); a peer's mapping is kept in the ctx and reused while its handle is unchanged.
    bool peer_ok = transport != DGE_TRANSPORT_NCCL && world <= DGE_DP_MAX_WORLD && n4 <= 128 && V > 0 && syn0 == ctx->dp_arena;
    const size_t off1 = (size_t)(syn1neg - syn0);   // the same on every rank: V and the pitch are global
    unsigned long long mine[9];
    memset(mine, 0, sizeof(mine));
    if (peer_ok) {
        cudaIpcMemHandle_t h0;
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        if (cudaIpcGetMemHandle(&h0, ctx->dp_arena) != cudaSuccess) { peer_ok = false; cudaGetLastError(); }
        else memcpy(mine, &h0, 64);
    }
    mine[8] = peer_ok ? 1 : 0;
    std::vector<unsigned long long> all((size_t)world * 9);
    int rc = dge_comm_allgather_u64(ctx, mine, 9, all.data());
    if (rc != DGE_OK) { dge_dp_end(dp); return rc; }
    for (int r = 0; r < world; r++) peer_ok = peer_ok && all[(size_t)r * 9 + 8] == 1;
    if (peer_ok) {
        for (int r = 0; r < world && peer_ok; r++) {
            if (r == rank) { dp->peer[0][r] = syn0; dp->peer[1][r] = syn1neg; continue; }
            const void *hbytes = &all[(size_t)r * 9];
            if (!(ctx->dp_peer_mapped[r] && memcmp(ctx->dp_peer_handle[r], hbytes, 64) == 0)) {
                if (ctx->dp_peer_mapped[r]) { cudaIpcCloseMemHandle(ctx->dp_peer_ptr[r]); ctx->dp_peer_mapped[r] = false; }
                cudaIpcMemHandle_t h;
                memcpy(&h, hbytes, 64);
                void *q = nullptr;
                if (cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { peer_ok = false; cudaGetLastError(); break; }
                ctx->dp_peer_ptr[r] = q; ctx->dp_peer_mapped[r] = true;
                memcpy(ctx->dp_peer_handle[r], hbytes, 64);
            }
            dp->peer[0][r] = (float *)ctx->dp_peer_ptr[r];
            dp->peer[1][r] = (float *)ctx->dp_peer_ptr[r] + off1;
        }
    }
    // all ranks must have mapped all peers
    {
        unsigned long long ok = peer_ok ? 1 : 0;
        std::vector<unsigned long long> oks((size_t)world);
        rc = dge_comm_allgather_u64(ctx, &ok, 1, oks.data());
        if (rc != DGE_OK) { dge_dp_end(dp); return rc; }
        for (int r = 0; r < world; r++) peer_ok = peer_ok && oks[r] == 1;
    }
    if (!peer_ok && transport == DGE_TRANSPORT_PEER)
        local = dge_fail(ctx, DGE_E_COMM, "dge_dp_begin: peer mapping of the replicas (cudaIpc over NVLink) is unavailable and transport = PEER was demanded");
    dp->use_peer = peer_ok;
    // ---- base: the row slice of this rank (PEER) or everything (NCCL)
    if (local == DGE_OK) {
        dp->row_lo = peer_ok ? (int64_t)V * rank / world : 0;
        dp->row_hi = peer_ok ? (int64_t)V * (rank + 1) / world : V;
        const size_t rows = (size_t)(dp->row_hi - dp->row_lo), nb = (rows ? rows : 1) * (size_t)stride;
        if (dge_malloc(ctx, &dp->base[0], nb) != cudaSuccess || dge_malloc(ctx, &dp->base[1], nb) != cudaSuccess ||
            (!peer_ok && dge_malloc(ctx, &dp->aux, 4 * (size_t)(V ? V : 1)) != cudaSuccess))
            local = dge_fail(ctx, DGE_E_CUDA, "dge_dp_begin: cudaMalloc of the delta base failed");
        else {
            cudaMemcpyAsync(dp->base[0], syn0 + (size_t)dp->row_lo * stride, rows * (size_t)stride * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream);
            cudaMemcpyAsync(dp->base[1], syn1neg + (size_t)dp->row_lo * stride, rows * (size_t)stride * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream);
        }
    }
    rc = dge_comm_agree(ctx, local, "dge_dp_begin");
    if (rc != DGE_OK) { dge_dp_end(dp); return rc; }
    ctx->phase_ms["sgns_transport"] = peer_ok ? 1.f : 2.f; // 1 = peer-memory kernel over NVLink, 2 = NCCL all-reduce
    *out = dp;
    return DGE_OK;
}

// One exchange, stream-ordered after the training launches of the slice.  local_error != 0 tells the other ranks
// that this rank is in trouble; with `check` the host waits and every rank learns whether anyone was (collective).
int dge_dp_exchange(dge_dp *dp, int local_error, bool check, int *any_error) {
    dge_ctx *ctx = dp->ctx;
    cudaStream_t st = ctx->stream;
    const int grid = ctx->sm_count * 8;
    cudaEventRecord(dp->e0, st);
    int rc = dp_barrier(dp, local_error, false, nullptr);      // every replica has finished its slice
    if (rc != DGE_OK) return rc;
    if (dp->use_peer) {
        dp_peers P;
        for (int t = 0; t < 2; t++)
            for (int r = 0; r < DGE_DP_MAX_WORLD; r++) P.cur[t][r] = dp->peer[t][r];
        if (dp->row_hi > dp->row_lo) {
            if (dp->n4 <= 32) k_dp_exchange_peer<1><<<grid, 256, 0, st>>>(P, dp->base[0], dp->base[1], dp->row_lo, dp->row_hi, dp->stride, dp->n4, ctx->world, dp->combine);
            else if (dp->n4 <= 64) k_dp_exchange_peer<2><<<grid, 256, 0, st>>>(P, dp->base[0], dp->base[1], dp->row_lo, dp->row_hi, dp->stride, dp->n4, ctx->world, dp->combine);
            else k_dp_exchange_peer<4><<<grid, 256, 0, st>>>(P, dp->base[0], dp->base[1], dp->row_lo, dp->row_hi, dp->stride, dp->n4, ctx->world, dp->combine);
            ctx->launches++;
        }
    } else {
        const size_t nel = (size_t)dp->V * (size_t)dp->stride;
        k_dp_delta<<<grid, 256, 0, st>>>(dp->cur[0], dp->base[0], dp->cur[1], dp->base[1], dp->V, dp->stride, dp->n4, dp->aux);
        rc = dge_comm_allreduce_sum_f32(ctx, dp->cur[0], nel);
        if (rc == DGE_OK) rc = dge_comm_allreduce_sum_f32(ctx, dp->cur[1], nel);
        if (rc == DGE_OK) rc = dge_comm_allreduce_sum_f32(ctx, dp->aux, 4 * (size_t)dp->V);
        if (rc != DGE_OK) return rc;
        k_dp_apply<<<grid, 256, 0, st>>>(dp->cur[0], dp->base[0], dp->cur[1], dp->base[1], dp->V, dp->stride, dp->n4, dp->aux, dp->combine, ctx->world);
        ctx->launches += 2;
    }
    int launch_err = cudaGetLastError() != cudaSuccess;
    rc = dp_barrier(dp, launch_err, check, any_error);         // every replica has been rewritten
    if (rc != DGE_OK) return rc;
    cudaEventRecord(dp->e1, st);
    if (check) {
        float ms = 0.f;
        if (cudaEventSynchronize(dp->e1) == cudaSuccess && cudaEventElapsedTime(&ms, dp->e0, dp->e1) == cudaSuccess) dp->ms += ms;
    }
    dp->exchanges++;
    return DGE_OK;
}
float dge_dp_ms(const dge_dp *dp) { return dp ? dp->ms : 0.f; }
bool dge_dp_uses_peer_memory(const dge_dp *dp) { return dp && dp->use_peer; }

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
    dge_dp_release_cache(ctx);
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
