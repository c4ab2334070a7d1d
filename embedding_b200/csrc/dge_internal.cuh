// dge_internal.cuh -- shared internals of libdge.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <string>
#include <vector>
#include "../../include/dge.h"

#define DGE_WARP 32

struct dge_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // phase timers
    cudaEvent_t tev0 = nullptr, tev1 = nullptr; // dge_timer_start / dge_timer_stop (bench.py's timed region)
    void *comm = nullptr;                       // ncclComm_t when dge_comm_init was called (comm.cu)
    int rank = 0, world = 1;
    std::string err;
    std::map<std::string, float> phase_ms;
    int64_t launches = 0;
    cudaMemPool_t pool = nullptr;               // PRIVATE stream-ordered pool of this ctx (not the device's default pool)
    // data-parallel skip-gram (comm.cu): the replica arena the ranks train in (cudaMalloc: mappable by the other ranks),
    // kept across dge_sgns_train calls together with the peers' mappings of THEIR arenas -- cudaIpcOpenMemHandle costs
    // ~20 ms per peer, so it is paid once per communicator, not once per epoch
    float *dp_arena = nullptr;
    size_t dp_arena_bytes = 0;
    void *dp_peer_ptr[16] = {nullptr};
    unsigned char dp_peer_handle[16][64] = {{0}};
    bool dp_peer_mapped[16] = {false};
    int refs = 1;                               // the ctx itself + every live graph / corpus / model / flows handle
    bool closed = false;                        // dge_destroy was called; torn down when the last handle is freed
};
// Handles keep their ctx alive: a *_free after dge_destroy is safe (the ctx is torn down by the last release).
void dge_ctx_retain(dge_ctx *ctx);
void dge_ctx_release(dge_ctx *ctx);

// One 32-byte sector per edge: a walk step is ONE dependent random load.  The record carries the
// chosen column's and the alias column's destination together with their CSR row (start, degree),
// so the next step needs no row_ptr lookup.
struct __align__(32) dge_edge_rec {
    double prob;      // alias-table acceptance threshold of this column
    int32_t dst;      // col[i]
    int32_t adst;     // col[alias[i]]  (alias == -1  =>  dst)
    uint32_t start0;  // row_ptr[dst]
    uint32_t deg0;    // degree(dst)
    uint32_t start1;  // row_ptr[adst]
    uint32_t deg1;    // degree(adst)
};
static_assert(sizeof(dge_edge_rec) == 32, "edge record must be one 32 B sector");

struct dge_graph {
    dge_ctx *ctx = nullptr;
    int32_t nv = 0, ns = 0;
    int64_t ne = 0;
    int64_t *row_ptr = nullptr;   // [nv+1]
    int32_t *col = nullptr;       // [ne]
    double *w = nullptr;          // [ne]
    double *prob = nullptr;       // [ne]
    int32_t *alias = nullptr;     // [ne]
    double *out_degree = nullptr; // [nv]
    int32_t *sources = nullptr;   // [ns]
    double *src_w = nullptr;      // [ns] out_degree of each source
    double *src_prob = nullptr;   // [ns]
    int32_t *src_alias = nullptr; // [ns]
    double *sws = nullptr;        // [1] device copy of sourceWeightSum
    double source_weight_sum = 0;
    dge_edge_rec *rec = nullptr;  // [ne]
    dge_edge_rec *srec = nullptr; // [ns]
    int32_t *v_layer = nullptr;   // [nv] labels by vertex id, only for graphs built by dge_crosstime_graph_build
    int32_t *v_region = nullptr;  // [nv] region INDEX (position in the host's region-id array)
};

struct dge_flows {
    dge_ctx *ctx = nullptr;
    int32_t n = 0;           // regions
    int32_t *F = nullptr;    // dense trip counts [n][24][n] (src, hour of day, dst)
    int64_t trips = 0;       // records added with dge_flows_add_trips
};

int dge_graph_build_device(dge_ctx *ctx, int32_t nv, int64_t ne, const int32_t *d_src, const int32_t *d_dst,
                           const double *d_w, int32_t ns, const int32_t *sources, const double *out_degree,
                           const double *source_weight_sum, dge_graph **out);
// exclusive scan of int32 flags into int64 positions (graph.cu); pos has n+1 entries
int dge_scan_i32(dge_ctx *ctx, const int32_t *d_in, int32_t n, int64_t *d_pos);

struct dge_corpus {
    dge_ctx *ctx = nullptr;
    int64_t n = 0;          // walks
    int32_t L = 0;          // positions
    int32_t n_ids = 0;      // id space (graph vertices)
    int32_t *tok = nullptr; // POSITION-major [L][n]: the walk kernel's stores are coalesced
};

struct dge_model {
    dge_ctx *ctx = nullptr;
    int32_t V = 0, dim = 0, stride = 0; // stride = dim rounded up to 4 floats (zero padded)
    int64_t pairs = 0;   // (centre, context) updates executed
    int64_t words = 0;   // in-vocabulary tokens of the corpus
    float *syn0 = nullptr, *syn1neg = nullptr; // device [V*stride]
    float *syn0_alloc = nullptr, *syn1neg_alloc = nullptr; // what was allocated (the tables may start at an offset inside: placement)
    int32_t *id_of_word = nullptr;             // device [V]
};

// every handle is created and deleted through these: the handle holds a reference on its ctx
template <typename H> static inline H *dge_new_handle(dge_ctx *ctx) { H *h = new H(); h->ctx = ctx; dge_ctx_retain(ctx); return h; }
template <typename H> static inline void dge_delete_handle(H *h) { if (!h) return; dge_ctx *c = h->ctx; delete h; dge_ctx_release(c); }

// ---- multi-GPU (comm.cu); no-ops without a communicator
int dge_comm_allreduce_sum_f32(dge_ctx *ctx, float *buf, size_t n);
int dge_comm_allreduce_sum_u64(dge_ctx *ctx, unsigned long long *buf, size_t n);
int dge_comm_allreduce_max_u64(dge_ctx *ctx, unsigned long long *buf, size_t n);
int dge_comm_allgather_u64(dge_ctx *ctx, const unsigned long long *mine, int n_per_rank, unsigned long long *out_host);
int dge_comm_agree(dge_ctx *ctx, int local_status, const char *what);
// data-parallel exchange of the skip-gram replicas (comm.cu)
#define DGE_DP_MAX_WORLD 16
struct dge_dp;
int dge_dp_arena(dge_ctx *ctx, size_t bytes, float **out);   // collective: the rank's replica arena of at least `bytes`
void dge_dp_release_cache(dge_ctx *ctx);                      // closes the peer mappings, frees the arena
int dge_dp_begin(dge_ctx *ctx, float *syn0, float *syn1neg, int32_t V, int32_t stride, int32_t n4, int combine, int transport,
                 dge_dp **out);
int dge_dp_exchange(dge_dp *dp, int local_error, bool check, int *any_error);
void dge_dp_end(dge_dp *dp);
float dge_dp_ms(const dge_dp *dp);
bool dge_dp_uses_peer_memory(const dge_dp *dp);

// ---- error plumbing
void dge_set_error(dge_ctx *ctx, const std::string &msg);
int dge_fail(dge_ctx *ctx, int code, const std::string &msg);
#define DGE_CUDA(ctx, expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return dge_fail((ctx), DGE_E_CUDA,                                                    \
                            std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                                ":" + std::to_string(__LINE__) + ")");                            \
    } while (0)
#define DGE_LAUNCH_CHECK(ctx)                 \
    do {                                      \
        (ctx)->launches++;                    \
        DGE_CUDA((ctx), cudaGetLastError());  \
    } while (0)

// phase timing with CUDA events on ctx->stream
struct dge_phase_timer {
    dge_ctx *ctx;
    const char *name;
    dge_phase_timer(dge_ctx *c, const char *n) : ctx(c), name(n) { cudaEventRecord(ctx->ev0, ctx->stream); }
    void stop() {
        cudaEventRecord(ctx->ev1, ctx->stream);
        cudaEventSynchronize(ctx->ev1);
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->phase_ms[name] = ms;
    }
};

// Device memory comes from the ctx's OWN stream-ordered pool on the ctx stream (context.cu creates it with the release
// threshold "never" and destroys it with the ctx, so the process-wide default pool that torch / NCCL may share is left
// alone): the multi-GB token / scratch buffers of one step are reused by the next step without a round trip to the
// driver.  Frees are stream-ordered too, so they never synchronise the device.
template <typename T>
static inline cudaError_t dge_malloc(dge_ctx *ctx, T **p, size_t n) {
    if (ctx->pool) return cudaMallocFromPoolAsync((void **)p, (n ? n : 1) * sizeof(T), ctx->pool, ctx->stream);
    return cudaMallocAsync((void **)p, (n ? n : 1) * sizeof(T), ctx->stream);
}
static inline void dge_free(dge_ctx *ctx, void *p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}

// ---- Philox4x32-10 (Salmon et al., SC'11).  Same stream definition as oracle/dge_oracle.c.
__host__ __device__ static inline void dge_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        if (r) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform on the java.util.Random.nextDouble() grid: (bits >> 11) * 2^-53
__host__ __device__ static inline double dge_u53(uint32_t lo, uint32_t hi) {
    uint64_t bits = ((uint64_t)hi << 32) | lo;
    return (double)(bits >> 11) * 0x1.0p-53;
}
