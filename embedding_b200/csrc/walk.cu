// walk.cu -- stage 1b on device: weighted random walks over the packed alias records.
//
// Reference behaviour reproduced (embedding/src/main/java/embedding/):
//   LayeredGraph.sampleVertexSequence()    :232-252  (alias sampler)
//   LayeredGraph.sampleVertexSequence_OV() :260-279  (CDF sampler, the reference's slow baseline)
//   Vertex.sampleNextVertex()              :104-116  one uniform drives column and coin
//   CrossTimeGraph.sampleSequenceHelper    :134-140  the loop over numSamples walks
// The unseeded shared java.util.Random (:14) is replaced by a counter-based Philox4x32-10 stream
// per walk, so output is reproducible and independent of how walks are split over GPUs.
//
// Layout: one thread per walk; a step is ONE dependent 32-byte (256-bit) load (dge_edge_rec, a full DRAM
// sector) because the record carries the CSR row of both possible destinations.  Tokens are
// stored position-major [L][n] so every store instruction of a warp is one 128 B line.
#include "dge_internal.cuh"

__device__ __forceinline__ dge_edge_rec ld_rec(const dge_edge_rec *p) {
    // ONE 256-bit read-only load of the record's sector (sm_100 LDG.256): the L2-resident configs are bound by L1
    // request throughput (ncu: l1tex 97 % busy with two 128-bit requests per step), so one request per step it is.
    // Records are 32 bytes and 32-byte aligned (dge_edge_rec, arrays from the pool are 256-byte aligned).
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    dge_edge_rec r;
    r.prob = __longlong_as_double((long long)a);
    r.dst = (int32_t)(uint32_t)b; r.adst = (int32_t)(uint32_t)(b >> 32);
    r.start0 = (uint32_t)c; r.deg0 = (uint32_t)(c >> 32); r.start1 = (uint32_t)d; r.deg1 = (uint32_t)(d >> 32);
    return r;
}

// draw `t` of walk `wid`: Philox block t/2 gives two 53-bit uniforms
struct walk_rng {
    uint32_t w0, w1, k0, k1;
    uint32_t r[4];
    __device__ __forceinline__ double next(uint32_t t) {
        if ((t & 1) == 0) dge_philox4x32_10(w0, w1, t >> 1, 0u, k0, k1, r);
        return (t & 1) ? dge_u53(r[2], r[3]) : dge_u53(r[0], r[1]);
    }
};

__global__ void __launch_bounds__(256)
k_walk_alias(const dge_edge_rec *__restrict__ rec, const dge_edge_rec *__restrict__ srec, int32_t ns, int64_t n_walks,
             int64_t first_walk_id, int32_t L, uint64_t seed, int32_t *__restrict__ tok) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_walks) return;
    uint64_t wid = (uint64_t)(first_walk_id + i);
    walk_rng g;
    g.w0 = (uint32_t)wid; g.w1 = (uint32_t)(wid >> 32); g.k0 = (uint32_t)seed; g.k1 = (uint32_t)(seed >> 32);
    const dge_edge_rec *row = srec;
    uint32_t k = (uint32_t)ns;
    int32_t j = 0;
    for (; j < L; j++) {
        if (k == 0) break; // dead end (Java: sampleNextVertex() == null) or empty source list
        double x = g.next((uint32_t)j);
        // LayeredGraph.java:108-115
        double xk = __dmul_rn(x, (double)k);
        int32_t c = __double2int_rz(xk);
        double y = __dsub_rn(xk, (double)c);
        dge_edge_rec r = ld_rec(row + c);
        bool first = y < r.prob;
        int32_t v = first ? r.dst : r.adst;
        row = rec + (first ? r.start0 : r.start1);
        k = first ? r.deg0 : r.deg1;
        tok[(int64_t)j * n_walks + i] = v;
    }
    for (; j < L; j++) tok[(int64_t)j * n_walks + i] = -1;
}

// CDF sampler (LayeredGraph.java:89-98, :260-279): linear scan of the row's weights.
__global__ void __launch_bounds__(256)
k_walk_cdf(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col, const double *__restrict__ wc,
           const double *__restrict__ od, const int32_t *__restrict__ sources, int32_t ns,
           const double *__restrict__ sws, int64_t n_walks, int64_t first_walk_id, int32_t L, uint64_t seed,
           int32_t *__restrict__ tok) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_walks) return;
    uint64_t wid = (uint64_t)(first_walk_id + i);
    walk_rng g;
    g.w0 = (uint32_t)wid; g.w1 = (uint32_t)(wid >> 32); g.k0 = (uint32_t)seed; g.k1 = (uint32_t)(seed >> 32);
    int32_t j = 0;
    int32_t v = -1;
    if (L > 0 && ns > 0) {
        double s = __dmul_rn(g.next(0u), *sws), cnt = 0.0;
        for (int32_t q = 0; q < ns; q++) {
            cnt = __dadd_rn(cnt, od[sources[q]]);
            if (cnt >= s) { v = sources[q]; break; }
        }
    }
    if (v >= 0) {
        tok[i] = v;
        j = 1;
        for (; j < L; j++) {
            int64_t b = row_ptr[v], e = row_ptr[v + 1];
            if (b == e) break;
            double s = __dmul_rn(g.next((uint32_t)j), od[v]), cnt = 0.0;
            int32_t nn = -1;
            for (int64_t q = b; q < e; q++) {
                cnt = __dadd_rn(cnt, wc[q]);
                if (cnt >= s) { nn = col[q]; break; }
            }
            if (nn < 0) break;
            tok[(int64_t)j * n_walks + i] = nn;
            v = nn;
        }
    }
    for (; j < L; j++) tok[(int64_t)j * n_walks + i] = -1;
}

// position-major [L][n] -> walk-major [n][L] through a shared-memory tile (both sides coalesced)
__global__ void k_tokens_to_walk_major(const int32_t *__restrict__ tok, int64_t n, int32_t L, int32_t *__restrict__ out) {
    __shared__ int32_t tile[32][33];
    int64_t i0 = (int64_t)blockIdx.x * 32;
    int32_t j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t i = i0 + threadIdx.x;
        int32_t j = j0 + r;
        if (i < n && j < L) tile[r][threadIdx.x] = tok[(int64_t)j * n + i];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t i = i0 + r;
        int32_t j = j0 + threadIdx.x;
        if (i < n && j < L) out[i * L + j] = tile[threadIdx.x][r];
    }
}
// same transposition, narrowed to 16 bits (0xFFFF = padding); n_ids <= 65535 is checked by the caller
__global__ void k_tokens_to_walk_major_u16(const int32_t *__restrict__ tok, int64_t n, int32_t L, uint16_t *__restrict__ out) {
    __shared__ int32_t tile[32][33];
    int64_t i0 = (int64_t)blockIdx.x * 32;
    int32_t j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t i = i0 + threadIdx.x;
        int32_t j = j0 + r;
        if (i < n && j < L) tile[r][threadIdx.x] = tok[(int64_t)j * n + i];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t i = i0 + r;
        int32_t j = j0 + threadIdx.x;
        if (i < n && j < L) out[i * L + j] = (uint16_t)tile[threadIdx.x][r]; // -1 -> 0xFFFF
    }
}
__global__ void k_tokens_to_pos_major(const int32_t *__restrict__ in, int64_t n, int32_t L, int32_t *__restrict__ tok) {
    __shared__ int32_t tile[32][33];
    int64_t i0 = (int64_t)blockIdx.x * 32;
    int32_t j0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t i = i0 + r;
        int32_t j = j0 + threadIdx.x;
        if (i < n && j < L) tile[r][threadIdx.x] = in[i * L + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t i = i0 + threadIdx.x;
        int32_t j = j0 + r;
        if (i < n && j < L) tok[(int64_t)j * n + i] = tile[threadIdx.x][r];
    }
}

__global__ void k_count_tokens(const int32_t *__restrict__ tok, int64_t total, int32_t n_ids, unsigned long long *cnt,
                               int *bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long c = 0;
    for (; i < total; i += stride) {
        int32_t t = tok[i];
        c += t >= 0;
        if (t < -1 || t >= n_ids) *bad = 1;
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(cnt, c);
}

__global__ void k_relabel(int32_t *__restrict__ tok, int64_t n, int32_t L, const int32_t *__restrict__ map,
                          int32_t old_ids, int32_t new_ids, int32_t pos_stride, int *bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = n * L, stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int32_t t = tok[i];
        if (t < 0) continue;
        int32_t j = (int32_t)(i / n); // position-major layout
        int64_t v = t < old_ids ? (int64_t)map[t] + (int64_t)j * pos_stride : -1;
        if (v < 0 || v >= new_ids) { *bad = 1; continue; }
        tok[i] = (int32_t)v;
    }
}

extern "C" {

int dge_corpus_relabel(dge_corpus *c, const int32_t *id_map, int32_t new_n_ids, int32_t position_stride) {
    if (!c) return dge_fail(nullptr, DGE_E_INVALID, "dge_corpus_relabel: corpus is NULL");
    dge_ctx *ctx = c->ctx;
    if ((!id_map && c->n_ids > 0) || new_n_ids < 0) return dge_fail(ctx, DGE_E_INVALID, "dge_corpus_relabel: bad arguments");
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    int32_t *d_map = nullptr;
    int *d_bad = nullptr;
    DGE_CUDA(ctx, dge_malloc(ctx, &d_map, (size_t)c->n_ids));
    if (dge_malloc(ctx, &d_bad, 1) != cudaSuccess) { dge_free(ctx, d_map); return dge_fail(ctx, DGE_E_CUDA, "dge_corpus_relabel: cudaMalloc"); }
    cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream);
    if (c->n_ids) cudaMemcpyAsync(d_map, id_map, sizeof(int32_t) * (size_t)c->n_ids, cudaMemcpyHostToDevice, ctx->stream);
    if (c->n * c->L > 0) {
        k_relabel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(c->tok, c->n, c->L, d_map, c->n_ids, new_n_ids, position_stride, d_bad);
        ctx->launches++;
    }
    int h_bad = 0;
    cudaError_t e = cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    dge_free(ctx, d_map); dge_free(ctx, d_bad);
    if (e != cudaSuccess) return dge_fail(ctx, DGE_E_CUDA, std::string("dge_corpus_relabel: ") + cudaGetErrorString(e));
    if (h_bad) return dge_fail(ctx, DGE_E_INVALID, "dge_corpus_relabel: mapped id outside [0, new_n_ids) (corpus is now partially relabelled)");
    c->n_ids = new_n_ids;
    return DGE_OK;
}

int dge_walk(const dge_graph *g, int64_t n_walks, int64_t first_walk_id, int32_t L, uint64_t seed, int sampler,
             dge_corpus **out) {
    if (!g) return dge_fail(nullptr, DGE_E_INVALID, "dge_walk: graph is NULL");
    dge_ctx *ctx = g->ctx;
    if (!out) return dge_fail(ctx, DGE_E_INVALID, "dge_walk: out is NULL");
    *out = nullptr;
    if (n_walks < 0 || L < 0 || first_walk_id < 0) return dge_fail(ctx, DGE_E_INVALID, "dge_walk: negative size");
    if (sampler != DGE_SAMPLER_ALIAS && sampler != DGE_SAMPLER_CDF)
        return dge_fail(ctx, DGE_E_INVALID, "dge_walk: unknown sampler");
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    dge_corpus *c = dge_new_handle<dge_corpus>(ctx);
    c->ctx = ctx; c->n = n_walks; c->L = L; c->n_ids = g->nv;
    cudaError_t e = dge_malloc(ctx, &c->tok, (size_t)n_walks * (size_t)L);
    if (e != cudaSuccess) {
        dge_delete_handle(c);
        return dge_fail(ctx, DGE_E_CUDA, std::string("dge_walk: cudaMalloc tokens: ") + cudaGetErrorString(e));
    }
    if (n_walks > 0 && L > 0) {
        dge_phase_timer t(ctx, "walk");
        const int T = 256;
        unsigned grid = (unsigned)((n_walks + T - 1) / T);
        if (sampler == DGE_SAMPLER_ALIAS)
            k_walk_alias<<<grid, T, 0, ctx->stream>>>(g->rec, g->srec, g->ns, n_walks, first_walk_id, L, seed, c->tok);
        else
            k_walk_cdf<<<grid, T, 0, ctx->stream>>>(g->row_ptr, g->col, g->w, g->out_degree, g->sources, g->ns, g->sws,
                                                    n_walks, first_walk_id, L, seed, c->tok);
        ctx->launches++;
        t.stop();
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            dge_free(ctx, c->tok);
            dge_delete_handle(c);
            return dge_fail(ctx, DGE_E_CUDA, std::string("dge_walk: ") + cudaGetErrorString(e));
        }
    }
    *out = c;
    return DGE_OK;
}

int dge_corpus_from_tokens(dge_ctx *ctx, const int32_t *tokens, int64_t n_walks, int32_t L, int32_t n_ids,
                           dge_corpus **out) {
    if (!ctx) return dge_fail(nullptr, DGE_E_INVALID, "dge_corpus_from_tokens: ctx is NULL");
    if (!out) return dge_fail(ctx, DGE_E_INVALID, "dge_corpus_from_tokens: out is NULL");
    *out = nullptr;
    if (n_walks < 0 || L < 0 || n_ids < 0 || (n_walks * L > 0 && !tokens))
        return dge_fail(ctx, DGE_E_INVALID, "dge_corpus_from_tokens: bad arguments");
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    dge_corpus *c = dge_new_handle<dge_corpus>(ctx);
    c->ctx = ctx; c->n = n_walks; c->L = L; c->n_ids = n_ids;
    size_t total = (size_t)n_walks * (size_t)L;
    int32_t *stage = nullptr;
    unsigned long long *d_cnt = nullptr;
    int rc = DGE_OK;
    auto fail = [&](const std::string &m, int code) {
        dge_free(ctx, c->tok); dge_free(ctx, stage); dge_free(ctx, d_cnt);
        dge_delete_handle(c);
        return dge_fail(ctx, code, m);
    };
    if (dge_malloc(ctx, &c->tok, total) != cudaSuccess || dge_malloc(ctx, &stage, total) != cudaSuccess ||
        dge_malloc(ctx, &d_cnt, 2) != cudaSuccess)
        return fail("dge_corpus_from_tokens: cudaMalloc failed", DGE_E_CUDA);
    if (total) {
        dge_phase_timer t(ctx, "tokens_h2d");
        cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), ctx->stream);
        cudaMemcpyAsync(stage, tokens, total * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
        dim3 grid((unsigned)((n_walks + 31) / 32), (unsigned)((L + 31) / 32)), block(32, 8);
        k_tokens_to_pos_major<<<grid, block, 0, ctx->stream>>>(stage, n_walks, L, c->tok);
        int *d_bad = (int *)(d_cnt + 1);
        k_count_tokens<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(c->tok, (int64_t)total, n_ids, d_cnt, d_bad);
        ctx->launches += 2;
        t.stop();
        int h_bad = 0;
        cudaMemcpy(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(std::string("dge_corpus_from_tokens: ") + cudaGetErrorString(e), DGE_E_CUDA);
        if (h_bad) return fail("dge_corpus_from_tokens: token id outside [-1, n_ids)", DGE_E_INVALID);
    }
    dge_free(ctx, stage); dge_free(ctx, d_cnt);
    *out = c;
    return rc;
}

int dge_corpus_shape(const dge_corpus *c, int64_t *n_walks, int32_t *L, int32_t *n_ids) {
    if (!c) return dge_fail(nullptr, DGE_E_INVALID, "dge_corpus_shape: corpus is NULL");
    if (n_walks) *n_walks = c->n;
    if (L) *L = c->L;
    if (n_ids) *n_ids = c->n_ids;
    return DGE_OK;
}

int dge_corpus_tokens(const dge_corpus *c, int32_t *tokens) {
    if (!c) return dge_fail(nullptr, DGE_E_INVALID, "dge_corpus_tokens: corpus is NULL");
    dge_ctx *ctx = c->ctx;
    size_t total = (size_t)c->n * (size_t)c->L;
    if (total == 0) return DGE_OK;
    if (!tokens) return dge_fail(ctx, DGE_E_INVALID, "dge_corpus_tokens: tokens is NULL");
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    int32_t *stage = nullptr;
    DGE_CUDA(ctx, dge_malloc(ctx, &stage, total));
    dge_phase_timer t(ctx, "tokens_d2h");
    dim3 grid((unsigned)((c->n + 31) / 32), (unsigned)((c->L + 31) / 32)), block(32, 8);
    k_tokens_to_walk_major<<<grid, block, 0, ctx->stream>>>(c->tok, c->n, c->L, stage);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(tokens, stage, total * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    t.stop();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    dge_free(ctx, stage);
    if (e != cudaSuccess) return dge_fail(ctx, DGE_E_CUDA, std::string("dge_corpus_tokens: ") + cudaGetErrorString(e));
    return DGE_OK;
}

int dge_corpus_tokens_u16(const dge_corpus *c, uint16_t *tokens) {
    if (!c) return dge_fail(nullptr, DGE_E_INVALID, "dge_corpus_tokens_u16: corpus is NULL");
    dge_ctx *ctx = c->ctx;
    if (c->n_ids > 65535) return dge_fail(ctx, DGE_E_LIMIT, "dge_corpus_tokens_u16: the id space does not fit 16 bits (0xFFFF is the padding)");
    size_t total = (size_t)c->n * (size_t)c->L;
    if (total == 0) return DGE_OK;
    if (!tokens) return dge_fail(ctx, DGE_E_INVALID, "dge_corpus_tokens_u16: tokens is NULL");
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    uint16_t *stage = nullptr;
    DGE_CUDA(ctx, dge_malloc(ctx, &stage, total));
    dge_phase_timer t(ctx, "tokens_d2h");
    dim3 grid((unsigned)((c->n + 31) / 32), (unsigned)((c->L + 31) / 32)), block(32, 8);
    k_tokens_to_walk_major_u16<<<grid, block, 0, ctx->stream>>>(c->tok, c->n, c->L, stage);
    ctx->launches++;
    cudaError_t e = cudaMemcpyAsync(tokens, stage, total * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream);
    t.stop();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    dge_free(ctx, stage);
    if (e != cudaSuccess) return dge_fail(ctx, DGE_E_CUDA, std::string("dge_corpus_tokens_u16: ") + cudaGetErrorString(e));
    return DGE_OK;
}

int dge_corpus_count_tokens(const dge_corpus *c, int64_t *n_tokens) {
    if (!c || !n_tokens) return dge_fail(c ? c->ctx : nullptr, DGE_E_INVALID, "dge_corpus_count_tokens: NULL argument");
    dge_ctx *ctx = c->ctx;
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long *d_cnt = nullptr;
    DGE_CUDA(ctx, dge_malloc(ctx, &d_cnt, 2));
    cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), ctx->stream);
    int64_t total = c->n * (int64_t)c->L;
    if (total) {
        k_count_tokens<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(c->tok, total, c->n_ids, d_cnt, (int *)(d_cnt + 1));
        ctx->launches++;
    }
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpyAsync(&h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dge_free(ctx, d_cnt);
    if (e != cudaSuccess) return dge_fail(ctx, DGE_E_CUDA, std::string("dge_corpus_count_tokens: ") + cudaGetErrorString(e));
    *n_tokens = (int64_t)h;
    return DGE_OK;
}

void dge_corpus_free(dge_corpus *c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    dge_free(c->ctx, c->tok);
    dge_delete_handle(c);
}

} // extern "C"
