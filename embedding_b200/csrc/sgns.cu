// sgns.cu -- stage 2 on device: skip-gram with negative sampling over the walk corpus.
//
// Stands under DeepWalk.learnEmbedding (DeepWalk.java:32-83): Word2Vec.Builder() :73-76 and
// w2v.fit() :79, whose arithmetic is DL4J 0.7.2 / libnd4j (external, un-vendored).  The kernel
// follows the published word2vec skip-gram update with the DL4J parameterisation listed in
// SURVEY.md 8(a) A14 and restated in oracle/sgns_oracle.c (same pair / negative enumeration,
// same per-sentence RNG and learning-rate schedule, so a sequential schedule reproduces the
// oracle to fp32 tolerance).  Pure negative sampling (north_star); no hierarchical softmax.
//
// Layout: sgns_common.cuh (kernel arguments, draw definitions, vocabulary / corpus preparation kernels, device helpers),
// sgns_kernels_items.cuh (kernel A: the oracle's exact order on the device; kernels B-E, C': the round-1 item kernels),
// sgns_kernels_sentence.cuh (kernels F-J: sentence-resident; F is what the automatic schedule runs), and this file: table
// statistics, kernel selection, the schedule (sentences in flight, write-through words, sentence counter), the
// data-parallel rounds and the C ABI.  Rows are read and updated through L2 (ld.global.cg, red.global.add.v4.f32): L2 is
// the coherence point.  No tensor cores: the work is K+1 dot products of length dim per pair, not a dense contraction.
#include "sgns_common.cuh"
#include "sgns_kernels_items.cuh"
#include "sgns_kernels_sentence.cuh"

// Placement probe.  The tables of the CA / tract workloads are a few MB and L2-resident, their traffic is concentrated on the
// rows of the frequent words, and how those rows fall on the L2 slices depends on where the tables start: the same launch
// took 987 ms at one base address and 1 018 / 1 040 / 1 148 / 1 239 / 1 401 ms at others (scripts/spread_probe.py,
// profiles/r2s36_spread_probe_before.json).  dge_sgns_train therefore allocates small tables with some slack, runs this
// probe -- random rows drawn from the unigram^0.75 table, one syn0 row loaded and two syn1neg rows loaded and reduced
// (with zeros) per step, from a full GPU of warps -- at a few candidate offsets, and trains at the fastest one.
#define SGNS_PLACE_CANDIDATES 24
#define SGNS_PLACE_STEP_FLOATS 9216            // 36 KB between candidates
#define SGNS_PLACE_MAX_BYTES (8u << 20)        // only tables this small are placed
__global__ void __launch_bounds__(640, 1)
k_place_probe(float *syn0, float *syn1neg, const int32_t *__restrict__ neg_table, uint32_t tsize, int32_t stride, int32_t n4, int iters, uint32_t seed) {
    const int lane = threadIdx.x & 7;
    const bool live = lane < n4;
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    uint64_t x = mix64(0x9E3779B97F4A7C15ULL * (uint64_t)(g + 1) + seed);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
        x = x * LCG_MUL + LCG_ADD;
        const uint32_t i0 = (uint32_t)(((x >> 16) & 0xFFFFFFFFULL) * tsize >> 32);
        const uint32_t i1 = (uint32_t)(((x >> 24) & 0xFFFFFFFFULL) * tsize >> 32);
        const uint32_t i2 = (uint32_t)(((x >> 32) & 0xFFFFFFFFULL) * tsize >> 32);
        const int32_t r0 = neg_table[i0], r1 = neg_table[i1], r2 = neg_table[i2];
        if (live) {
            float4 *p0 = reinterpret_cast<float4 *>(syn0 + (int64_t)r0 * stride) + lane;
            float4 *p1 = reinterpret_cast<float4 *>(syn1neg + (int64_t)r1 * stride) + lane;
            float4 *p2 = reinterpret_cast<float4 *>(syn1neg + (int64_t)r2 * stride) + lane;
            const float4 a = __ldcg(p0), b = __ldcg(p1), c = __ldcg(p2);
            acc += a.x + b.x + c.x;
            red_add4(p1, zero4);
            red_add4(p2, zero4);
            if ((it & 7) == 0) red_add4(p0, zero4);
        }
    }
    if (acc == 1234.5678f) syn0[0] = acc; // keeps the loads alive
}

// dge_model_stats: one warp per row of each table; acc[0] += |syn0 row|, acc[1] = max |element| (non-negative doubles
// order like their bit patterns), bad += non-finite elements
__global__ void k_model_stats(const float *__restrict__ syn0, const float *__restrict__ syn1neg, int32_t V, int32_t dim,
                              int32_t stride, double *acc, unsigned long long *bad) {
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int64_t r = warp; r < 2 * (int64_t)V; r += n_warps) {
        const float *row = (r < V ? syn0 + r * stride : syn1neg + (r - V) * stride);
        float ss = 0.f, mx = 0.f;
        unsigned nb = 0;
        for (int d = lane; d < dim; d += 32) {
            const float x = __ldcg(row + d);
            if (isfinite(x)) { ss += x * x; mx = fmaxf(mx, fabsf(x)); }
            else nb++;
        }
        for (int o = 16; o; o >>= 1) {
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            nb += __shfl_xor_sync(0xffffffffu, nb, o);
        }
        if (lane == 0) {
            if (r < V) atomicAdd(&acc[0], (double)sqrtf(ss));
            atomicMax(reinterpret_cast<unsigned long long *>(&acc[1]), (unsigned long long)__double_as_longlong((double)mx));
            if (nb) atomicAdd(bad, (unsigned long long)nb);
        }
    }
}

typedef void (*sgns_kernel_t)(const sgns_args);
struct sgns_variant { int G_seq, VPL_seq, G_items, VPL_items; sgns_kernel_t seq, items; int items_code; };
// items_code (reported as phase "sgns_kernel"): 1 k_sgns_items, 2 k_sgns_items_v2, 3 k_sgns_items_g4, 4 k_sgns_items_tp,
// 5 k_sgns_items_v3 (experimental); 0 k_sgns_seq

// Kernel A (exact order): rows of up to 8 float4 slots are held by ONE thread; wider rows give each lane of a
// 16- or 32-lane group one 128-bit slot (2 or 4 for rows wider than 32 slots).
// Kernel B (items): groups of 8 / 16 / 32 lanes, one slot per lane (2 or 4 beyond 32 slots).
static bool pick_variant(int n4, int negative, bool narrow_groups, bool target_parallel, bool staged_rows, bool plain_stores, int blk, bool smem_neg,
                         bool sentence_resident, bool block_sentence, int block_threads, bool pipelined, int pair_warps, bool duo, int prefetch_f, sgns_variant *out) {
    if (n4 > 128) return false;
    sgns_kernel_t seq = nullptr, items = nullptr;
    int Gs = 1, Vs = 1;
    switch (n4 <= 8 ? n4 : (n4 <= 16 ? 16 : (n4 <= 32 ? 32 : (n4 <= 64 ? 64 : 128)))) {
        case 1: seq = k_sgns_seq<1, 1>; Vs = 1; break;
        case 2: seq = k_sgns_seq<1, 2>; Vs = 2; break;
        case 3: seq = k_sgns_seq<1, 3>; Vs = 3; break;
        case 4: seq = k_sgns_seq<1, 4>; Vs = 4; break;
        case 5: seq = k_sgns_seq<1, 5>; Vs = 5; break;
        case 6: seq = k_sgns_seq<1, 6>; Vs = 6; break;
        case 7: seq = k_sgns_seq<1, 7>; Vs = 7; break;
        case 8: seq = k_sgns_seq<1, 8>; Vs = 8; break;
        case 16: seq = k_sgns_seq<16, 1>; Gs = 16; break;
        case 32: seq = k_sgns_seq<32, 1>; Gs = 32; break;
        case 64: seq = k_sgns_seq<32, 2>; Gs = 32; Vs = 2; break;
        default: seq = k_sgns_seq<32, 4>; Gs = 32; Vs = 4; break;
    }
    int Gi, Vi = 1;
    const bool multi = negative > SGNS_CH; // more than one 5-wide chunk of negatives per pair
    // 4-lane groups only for D <= 16 (+54 % at D = 16; at D = 20 / 32 the two-slot build measured -4 % / +5 %, the
    // reductions being the limit either way: profiles/r1s12_sgns_narrow_ab.txt) and only when the staleness bound
    // still lets them fill the GPU (with few items in flight, wider groups mean more warps to hide latency with)
    int code = 2;
    // target-parallel groups (kernel E) for narrow rows when the staleness bound leaves the GPU latency-bound
    if (n4 <= 4 && negative <= 7 && target_parallel) {
        code = 4;
        if (n4 == 1) { Gi = 8; items = k_sgns_items_tp<1>; }
        else if (n4 == 2) { Gi = 16; items = k_sgns_items_tp<2>; }
        else { Gi = 32; items = k_sgns_items_tp<4>; }
    }
    else if (n4 <= 4 && narrow_groups) { Gi = 4; code = 3; items = multi ? k_sgns_items_g4<1, true> : k_sgns_items_g4<1, false>; }
    else if (n4 <= 32 && staged_rows) { // experimental: rows of a unit staged in shared memory by cp.async (kernel C')
        code = 5;
        if (n4 <= 8) {
            Gi = 8;
            if (blk == 6) items = multi ? k_sgns_items_v3<8, true, 6> : k_sgns_items_v3<8, false, 6>;
            else if (blk == 7) items = multi ? k_sgns_items_v3<8, true, 7> : k_sgns_items_v3<8, false, 7>;
            else items = multi ? k_sgns_items_v3<8, true, 5> : k_sgns_items_v3<8, false, 5>;
        }
        else if (n4 <= 16) { Gi = 16; items = multi ? k_sgns_items_v3<16, true, 5> : k_sgns_items_v3<16, false, 5>; }
        else { Gi = 32; items = multi ? k_sgns_items_v3<32, true, 5> : k_sgns_items_v3<32, false, 5>; }
    }
    else if (n4 <= 32 && plain_stores) { // experiment: atomic-free row stores (lost updates allowed)
        code = 6;
        if (n4 <= 8) { Gi = 8; items = multi ? k_sgns_items_v2<8, true, 1> : k_sgns_items_v2<8, false, 1>; }
        else if (n4 <= 16) { Gi = 16; items = multi ? k_sgns_items_v2<16, true, 1> : k_sgns_items_v2<16, false, 1>; }
        else { Gi = 32; items = multi ? k_sgns_items_v2<32, true, 1> : k_sgns_items_v2<32, false, 1>; }
    }
    else if (n4 <= 8 && sentence_resident && block_sentence && duo && negative <= SGNS_CH) { // kernel J: critical + helper warps
        code = 12; Gi = 8;
        items = block_threads <= 192 ? k_sgns_duo<384> : k_sgns_duo<512>;
    }
    else if (n4 <= 8 && sentence_resident && block_sentence && pair_warps > 0 && negative <= 7) { // kernel I: a warp per pair, the round's pairs handed out dynamically
        code = 11; Gi = 32;
        items = pair_warps <= 8 ? k_sgns_wave<256> : (pair_warps <= 12 ? k_sgns_wave<384> : k_sgns_wave<512>);
    }
    else if (n4 <= 8 && sentence_resident && block_sentence && pipelined && negative <= SGNS_CH && block_threads <= 192) { // kernel H
        code = 10; Gi = 8; items = k_sgns_pipe<192>;
    }
    else if (n4 <= 32 && sentence_resident && block_sentence) { // kernel G: a block owns a sentence, one lane group per centre position
        code = 9;
        if (n4 <= 8 && block_threads <= 192) { Gi = 8; items = multi ? k_sgns_block<8, true, 192> : k_sgns_block<8, false, 192>; } // 170 registers
        else if (n4 <= 8) { Gi = 8; items = multi ? k_sgns_block<8, true, 256> : k_sgns_block<8, false, 256>; }
        else if (n4 <= 16) { Gi = 16; items = multi ? k_sgns_block<16, true, 256> : k_sgns_block<16, false, 256>; }
        else { Gi = 32; items = multi ? k_sgns_block<32, true, 256> : k_sgns_block<32, false, 256>; }
    }
    else if (n4 <= 32 && sentence_resident) { // kernel F: a warp owns a sentence (intra-sentence updates in sequence)
        code = 8;
        if (n4 <= 8) { Gi = 8; items = multi ? k_sgns_sent<8, true, 0> : (prefetch_f == 2 ? k_sgns_sent<8, false, 2> : (prefetch_f == 1 ? k_sgns_sent<8, false, 1> : k_sgns_sent<8, false, 0>)); }
        else if (n4 <= 16) { Gi = 16; items = multi ? k_sgns_sent<16, true, 0> : k_sgns_sent<16, false, 0>; }
        else { Gi = 32; items = multi ? k_sgns_sent<32, true, 0> : k_sgns_sent<32, false, 0>; }
    }
    else if (n4 <= 8 && smem_neg) { Gi = 8; code = 7; items = multi ? k_sgns_items_v2<8, true, 2> : k_sgns_items_v2<8, false, 2>; }
    else if (n4 <= 8) { Gi = 8; items = multi ? k_sgns_items_v2<8, true, 0> : k_sgns_items_v2<8, false, 0>; }
    else if (n4 <= 16) { Gi = 16; items = multi ? k_sgns_items_v2<16, true, 0> : k_sgns_items_v2<16, false, 0>; }
    else if (n4 <= 32) { Gi = 32; items = multi ? k_sgns_items_v2<32, true, 0> : k_sgns_items_v2<32, false, 0>; }
    else if (n4 <= 64) { Gi = 32; Vi = 2; code = 1; items = k_sgns_items<32, 2>; }
    else { Gi = 32; Vi = 4; code = 1; items = k_sgns_items<32, 4>; }
    out->G_seq = Gs; out->VPL_seq = Vs; out->seq = seq;
    out->G_items = Gi; out->VPL_items = Vi; out->items = items; out->items_code = code;
    return true;
}

static void model_release(dge_model *m) {
    if (!m) return;
    dge_free(m->ctx, m->syn0_alloc ? m->syn0_alloc : m->syn0); dge_free(m->ctx, m->syn1neg_alloc ? m->syn1neg_alloc : m->syn1neg);
    dge_free(m->ctx, m->id_of_word);
    dge_delete_handle(m);
}

// device temporaries of one call: released (stream-ordered) on every way out
struct sgns_scratch {
    dge_ctx *ctx;
    std::vector<void *> ptrs;
    explicit sgns_scratch(dge_ctx *c) : ctx(c) {}
    ~sgns_scratch() { for (void *q : ptrs) dge_free(ctx, q); }
    template <typename T> cudaError_t get(T **q, size_t n) {
        cudaError_t e = dge_malloc(ctx, q, n);
        if (e == cudaSuccess) ptrs.push_back(*q); else *q = nullptr;
        return e;
    }
};

static int sgns_check_params(dge_ctx *ctx, const dge_corpus *const *corpora, int32_t n_corpora, const dge_sgns_params *p) {
    if (!corpora || !p || n_corpora < 1 || n_corpora > SGNS_MAX_CORPORA)
        return dge_fail(ctx, DGE_E_INVALID, "dge_sgns_train: need 1..4 corpora and params");
    if (p->negative > SGNS_MAX_NEG) return dge_fail(ctx, DGE_E_LIMIT, "dge_sgns_train: negative must be <= 32");
    if (p->dim < 1 || p->window < 1 || p->negative < 0 || p->epochs < 1 || p->neg_table_size < 1 ||
        p->exp_table_size < 2 || p->min_count < 0 || p->concurrency < 0 || p->sync_rounds < 0 ||
        p->combine < DGE_COMBINE_DEFAULT || p->combine > DGE_COMBINE_SUM || p->transport < DGE_TRANSPORT_AUTO ||
        p->transport > DGE_TRANSPORT_NCCL || (p->schedule != DGE_SCHEDULE_ITEMS && p->schedule != DGE_SCHEDULE_SENTENCE))
        return dge_fail(ctx, DGE_E_INVALID, "dge_sgns_train: invalid hyper-parameter");
    if ((p->dim + 3) / 4 > 128) return dge_fail(ctx, DGE_E_LIMIT, "dge_sgns_train: dim must be <= 512");
    if (p->neg_table_size >= (1 << 30)) return dge_fail(ctx, DGE_E_LIMIT, "dge_sgns_train: neg_table_size must be < 2^30");
    const int32_t n_ids = corpora[0] ? corpora[0]->n_ids : 0;
    for (int i = 0; i < n_corpora; i++) {
        if (!corpora[i]) return dge_fail(ctx, DGE_E_INVALID, "dge_sgns_train: NULL corpus");
        if (corpora[i]->ctx != ctx) return dge_fail(ctx, DGE_E_INVALID, "dge_sgns_train: corpus belongs to another ctx");
        if (corpora[i]->n_ids != n_ids) return dge_fail(ctx, DGE_E_INVALID, "dge_sgns_train: corpora have different id spaces");
    }
    return DGE_OK;
}

extern "C" {

void dge_sgns_default_params(dge_sgns_params *p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->dim = 20; p->window = 8; p->negative = 5; p->min_count = 2; p->epochs = 1;
    p->neg_table_size = 100000; p->exp_table_size = 1000; p->concurrency = 0; p->schedule = DGE_SCHEDULE_ITEMS;
    p->sync_rounds = 0; p->combine = DGE_COMBINE_DEFAULT; p->transport = DGE_TRANSPORT_AUTO; p->flags = 0;
    p->lr = 0.025f; p->min_lr = 1e-4f; p->seed = 1;
}

int dge_sgns_train(dge_ctx *ctx, const dge_corpus *const *corpora, int32_t n_corpora, const dge_sgns_params *p,
                   dge_model **out) {
    if (!ctx) return dge_fail(nullptr, DGE_E_INVALID, "dge_sgns_train: ctx is NULL");
    const bool multi = ctx->comm != nullptr && ctx->world > 1;
    int local = DGE_OK;
    if (!out) local = dge_fail(ctx, DGE_E_INVALID, "dge_sgns_train: out is NULL");
    else *out = nullptr;
    if (local == DGE_OK) local = sgns_check_params(ctx, corpora, n_corpora, p);
    if (!multi && local != DGE_OK) return local;
    cudaSetDevice(ctx->device);
    int32_t n_ids = 0, Lmax = 0;
    int64_t n_sent = 0;
    if (local == DGE_OK) {
        n_ids = corpora[0]->n_ids;
        for (int i = 0; i < n_corpora; i++) { Lmax = std::max(Lmax, corpora[i]->L); n_sent += corpora[i]->n; }
    }
    // ---- data-parallel: the call is COLLECTIVE.  Every rank learns every rank's status, parameters and shard size
    // before anything else happens, so that a bad argument or a disagreement makes ALL ranks return the same error
    // (instead of leaving the others hung in the next collective), and the global sentence order is known.
    int64_t s_off = 0, n_global = n_sent, max_sent = n_sent;
    if (multi) {
        enum { NF = 20 };
        unsigned long long mine[NF];
        memset(mine, 0, sizeof(mine));
        mine[0] = (unsigned long long)(-(long long)local);
        if (local == DGE_OK) {
            uint32_t lrb, mlrb;
            memcpy(&lrb, &p->lr, 4); memcpy(&mlrb, &p->min_lr, 4);
            const unsigned long long f[] = {(unsigned long long)n_ids, (unsigned long long)Lmax, (unsigned long long)p->dim,
                (unsigned long long)p->window, (unsigned long long)p->negative, (unsigned long long)p->min_count,
                (unsigned long long)p->epochs, (unsigned long long)p->neg_table_size, (unsigned long long)p->exp_table_size,
                (unsigned long long)p->concurrency, (unsigned long long)p->schedule, (unsigned long long)p->sync_rounds,
                (unsigned long long)p->combine, (unsigned long long)p->transport, (unsigned long long)p->flags, lrb, mlrb, p->seed};
            for (int i = 0; i < 18; i++) mine[1 + i] = f[i];
            mine[19] = (unsigned long long)n_sent;
        }
        std::vector<unsigned long long> all((size_t)ctx->world * NF);
        int rc = dge_comm_allgather_u64(ctx, mine, NF, all.data());
        if (rc != DGE_OK) return rc;
        for (int r = 0; r < ctx->world; r++)
            if (all[(size_t)r * NF] != 0) {
                if (local != DGE_OK) return local;
                return dge_fail(ctx, -(int)all[(size_t)r * NF], "dge_sgns_train: rank " + std::to_string(r) + " rejected its arguments; every rank returns");
            }
        static const char *names[] = {"n_ids", "walk length", "dim", "window", "negative", "min_count", "epochs", "neg_table_size",
                                      "exp_table_size", "concurrency", "schedule", "sync_rounds", "combine", "transport", "flags", "lr", "min_lr", "seed"};
        for (int r = 1; r < ctx->world; r++)
            for (int i = 0; i < 18; i++)
                if (all[(size_t)r * NF + 1 + i] != all[1 + i])
                    return dge_fail(ctx, DGE_E_INVALID, std::string("dge_sgns_train: rank ") + std::to_string(r) + " disagrees with rank 0 on " + names[i] +
                                                            " (data-parallel training needs identical parameters on every rank)");
        n_global = 0; max_sent = 0;
        for (int r = 0; r < ctx->world; r++) {
            const int64_t ns = (int64_t)all[(size_t)r * NF + 19];
            if (r < ctx->rank) s_off += ns;
            n_global += ns;
            max_sent = std::max(max_sent, ns);
        }
    }
    const int32_t n4 = (p->dim + 3) / 4;        // float4 slots that carry data (zero-padded to whole slots)
    const int32_t stride = ((p->dim + 7) / 8) * 8; // row pitch in floats: rows start on 32-byte sector boundaries
    const int dbg = (int)p->flags;              // DGE_SGNS_F_* (dge.h)
    sgns_variant var;
    cudaStream_t st = ctx->stream;
    sgns_scratch tmp(ctx);

    // ---- vocabulary: device histogram, host ranking (descending count, ties ascending id)
    dge_phase_timer t_vocab(ctx, "vocab");
    unsigned long long *d_cnt = nullptr;
    local = tmp.get(&d_cnt, (size_t)n_ids + 2) == cudaSuccess ? DGE_OK : dge_fail(ctx, DGE_E_CUDA, "dge_sgns_train: cudaMalloc of the histogram failed");
    if (multi) local = dge_comm_agree(ctx, local, "dge_sgns_train (histogram)");
    if (local != DGE_OK) return local;
    cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * ((size_t)n_ids + 2), st);
    for (int i = 0; i < n_corpora; i++) {
        int64_t total = corpora[i]->n * (int64_t)corpora[i]->L;
        if (total) {
            k_hist<<<ctx->sm_count * 8, 256, 0, st>>>(corpora[i]->tok, total, d_cnt);
            ctx->launches++;
        }
    }
    // multi-GPU: every rank holds a shard of the corpus; the vocabulary is built from the global counts so that all
    // ranks index the same words identically
    if (multi) {
        int rc = dge_comm_allreduce_sum_u64(ctx, d_cnt, (size_t)n_ids);
        if (rc != DGE_OK) return rc;
    }
    // ---- ranking.  Large id spaces (the 2.4M-word synthetic vocabularies): 64-bit keys sorted on the device (CUB radix sort,
    // ~1 ms instead of ~80 ms of std::sort on every rank); small ones: on the host.  `cs` = counts in word order.
    std::vector<int32_t> order;
    std::vector<unsigned long long> cs;
    cudaError_t ce = cudaSuccess;
    bool ranked = false;
    if (n_ids >= (1 << 16)) {
        unsigned long long *d_keys = nullptr, *d_sorted = nullptr, *d_nv = nullptr;
        void *d_tmp = nullptr;
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_keys, d_sorted, n_ids, 0, 64, st);
        unsigned char *d_tmpb = nullptr;
        if (tmp.get(&d_keys, (size_t)n_ids) == cudaSuccess && tmp.get(&d_sorted, (size_t)n_ids) == cudaSuccess && tmp.get(&d_nv, 2) == cudaSuccess &&
            tmp.get(&d_tmpb, tmp_bytes) == cudaSuccess) {
            d_tmp = d_tmpb;
            cudaMemsetAsync(d_nv, 0, 2 * sizeof(unsigned long long), st);
            k_vocab_keys<<<(unsigned)((n_ids + 255) / 256), 256, 0, st>>>(d_cnt, n_ids, (unsigned long long)p->min_count, d_keys, d_nv, (int *)(d_nv + 1));
            cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_keys, d_sorted, n_ids, 0, 64, st);
            ctx->launches += 2;
            unsigned long long h_nv[2] = {0, 0};
            ce = cudaMemcpyAsync(h_nv, d_nv, sizeof(h_nv), cudaMemcpyDeviceToHost, st);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
            if (ce == cudaSuccess && (int)h_nv[1] == 0) { // every count fits 32 bits
                const size_t Vn = (size_t)h_nv[0];
                std::vector<unsigned long long> keys(Vn ? Vn : 1);
                if (Vn) ce = cudaMemcpyAsync(keys.data(), d_sorted, Vn * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
                if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
                if (ce == cudaSuccess) {
                    order.resize(Vn); cs.resize(Vn);
                    for (size_t k = 0; k < Vn; k++) { order[k] = (int32_t)(uint32_t)keys[k]; cs[k] = 0xFFFFFFFFULL - (keys[k] >> 32); }
                    ranked = true;
                }
            }
        } else cudaGetLastError();
    }
    if (!ranked && ce == cudaSuccess) {
        std::vector<unsigned long long> cnt((size_t)n_ids + 2);
        ce = cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(unsigned long long) * ((size_t)n_ids + 1), cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce == cudaSuccess) {
            // descending count, ties by ascending id: one 64-bit key per word, (2^32 - 1 - count) in the high half and the id in the low half
            std::vector<uint64_t> keys;
            keys.reserve(n_ids);
            bool small_counts = true;
            for (int32_t i = 0; i < n_ids; i++)
                if (cnt[i] > 0 && cnt[i] >= (unsigned long long)p->min_count) {
                    if (cnt[i] > 0xFFFFFFFFULL) small_counts = false;
                    keys.push_back(((0xFFFFFFFFULL - (cnt[i] & 0xFFFFFFFFULL)) << 32) | (uint32_t)i);
                }
            order.resize(keys.size());
            if (small_counts) {
                std::sort(keys.begin(), keys.end());
                for (size_t k = 0; k < keys.size(); k++) order[k] = (int32_t)(uint32_t)keys[k];
            } else { // counts beyond 32 bits (> 4e9 occurrences of one token): the plain comparator
                for (size_t k = 0; k < keys.size(); k++) order[k] = (int32_t)(uint32_t)keys[k];
                std::sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
                    if (cnt[x] != cnt[y]) return cnt[x] > cnt[y];
                    return x < y;
                });
            }
            cs.resize(order.size());
            for (size_t k = 0; k < order.size(); k++) cs[k] = cnt[order[k]];
        }
    }
    local = ce == cudaSuccess ? DGE_OK : dge_fail(ctx, DGE_E_CUDA, std::string("dge_sgns_train: histogram: ") + cudaGetErrorString(ce));
    if (multi) local = dge_comm_agree(ctx, local, "dge_sgns_train (histogram read-back)");
    if (local != DGE_OK) return local;
    const int32_t V = (int32_t)order.size();
    std::vector<int32_t> word_of_id((size_t)n_ids + 1, -1);
    for (int32_t wd = 0; wd < V; wd++) word_of_id[order[wd]] = wd;
    // unigram^0.75 table (word2vec.c InitUnigramTable / DL4J makeTable), identical to ora_neg_table
    std::vector<int32_t> table((size_t)p->neg_table_size, 0);
    if (V > 0) {
        // pow(count, 0.75) once per run of equal counts (the words are sorted by count: the long tail shares a few values);
        // the sum itself stays the oracle's sequential left-to-right double sum
        std::vector<double> pw((size_t)V);
        {
            unsigned long long prev = ~0ULL;
            double prev_pow = 0.0;
            for (int32_t wd = 0; wd < V; wd++) {
                if (cs[wd] != prev) { prev = cs[wd]; prev_pow = pow((double)prev, 0.75); }
                pw[wd] = prev_pow;
            }
        }
        double pow_sum = 0;
        for (int32_t wd = 0; wd < V; wd++) pow_sum += pw[wd];
        int32_t wi = 0;
        double d1 = pw[0] / pow_sum;
        for (int32_t i = 0; i < p->neg_table_size; i++) {
            table[i] = wi;
            if ((double)i / (double)p->neg_table_size > d1) {
                if (wi < V - 1) wi++;
                d1 += pw[wi] / pow_sum;
            }
        }
    }
    std::vector<float> exp_table((size_t)p->exp_table_size);
    for (int32_t i = 0; i < p->exp_table_size; i++) {
        double e = exp(((double)i / (double)p->exp_table_size * 2.0 - 1.0) * (double)SGNS_MAX_EXP);
        exp_table[i] = (float)(e / (e + 1.0));
    }
    t_vocab.stop();

    dge_model *m = dge_new_handle<dge_model>(ctx);
    m->ctx = ctx; m->V = V; m->dim = p->dim; m->stride = stride;
    int32_t *d_word_of_id = nullptr, *d_table = nullptr, *d_wtok = nullptr;
    float *d_exp = nullptr;
    unsigned long long *d_pairs = nullptr;
    const size_t nel = (size_t)(V ? V : 1) * (size_t)stride;
    const bool train = V > 0 && (n_sent > 0 || multi);
    // A data-parallel run trains in the rank's replica ARENA (cudaMalloc, mapped by the other ranks over NVLink, kept in the
    // ctx across calls); the model gets its own copy of the result at the end.  A single-GPU run trains in the model's tables.
    float *t0 = nullptr, *t1 = nullptr;
    if (multi) {
        float *arena = nullptr;
        // sized for the whole id space (V <= n_ids): the vocabulary of the next call (another walk seed) never makes the arena
        // grow, and a grown arena means every rank re-opens every peer's mapping (8 GPUs: ~400 ms, profiles/r2s33_bench_n8.json)
        const size_t nel_cap = (size_t)std::max<int64_t>(V ? V : 1, n_ids) * (size_t)stride;
        const int rc = dge_dp_arena(ctx, 2 * nel_cap * sizeof(float), &arena);   // collective
        if (rc != DGE_OK) { model_release(m); return rc; }
        t0 = arena; t1 = arena + nel;
    }
    const bool place = !multi && train && nel * sizeof(float) <= SGNS_PLACE_MAX_BYTES && SGNS_PLACE_CANDIDATES > 1;
    const size_t slack = place ? (size_t)SGNS_PLACE_CANDIDATES * SGNS_PLACE_STEP_FLOATS : 0;
    bool ok = dge_malloc(ctx, &m->syn0, nel + slack) == cudaSuccess && dge_malloc(ctx, &m->syn1neg, nel + slack) == cudaSuccess;
    m->syn0_alloc = m->syn0; m->syn1neg_alloc = m->syn1neg;
    if (!multi) { t0 = m->syn0; t1 = m->syn1neg; }
    ok = ok && dge_malloc(ctx, &m->id_of_word, (size_t)V) == cudaSuccess && tmp.get(&d_word_of_id, (size_t)n_ids) == cudaSuccess &&
         tmp.get(&d_table, (size_t)p->neg_table_size) == cudaSuccess && tmp.get(&d_exp, (size_t)p->exp_table_size) == cudaSuccess &&
         tmp.get(&d_pairs, 2) == cudaSuccess && (!train || tmp.get(&d_wtok, (size_t)n_sent * (size_t)Lmax) == cudaSuccess);
    local = ok ? DGE_OK : dge_fail(ctx, DGE_E_CUDA, std::string("dge_sgns_train: cudaMalloc failed (") + cudaGetErrorString(cudaGetLastError()) + ")");
    if (multi) local = dge_comm_agree(ctx, local, "dge_sgns_train (tables)");
    if (local != DGE_OK) { model_release(m); return local; }
    ctx->phase_ms["sgns_placement"] = 0.f; ctx->phase_ms["sgns_placement_best_us"] = 0.f; ctx->phase_ms["sgns_placement_worst_us"] = 0.f;
    if (place) { // where inside the allocations do the tables train fastest?  (before anything is written to them)
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (cudaMemcpyAsync(d_table, table.data(), sizeof(int32_t) * table.size(), cudaMemcpyHostToDevice, st) == cudaSuccess &&
            cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) {
            cudaMemsetAsync(m->syn0_alloc, 0, (nel + slack) * sizeof(float), st);
            cudaMemsetAsync(m->syn1neg_alloc, 0, (nel + slack) * sizeof(float), st);
            float best = 1e30f, worst = 0.f;
            int best_k = 0;
            k_place_probe<<<ctx->sm_count, 640, 0, st>>>(m->syn0_alloc, m->syn1neg_alloc, d_table, (uint32_t)p->neg_table_size, stride, n4, 32, 0u); // warm-up
            for (int k = 0; k < SGNS_PLACE_CANDIDATES; k++) {
                float ms_k = 1e30f;
                for (int rep = 0; rep < 2; rep++) {
                    float ms = 0.f;
                    cudaEventRecord(e0, st);
                    k_place_probe<<<ctx->sm_count, 640, 0, st>>>(m->syn0_alloc + (size_t)k * SGNS_PLACE_STEP_FLOATS, m->syn1neg_alloc + (size_t)k * SGNS_PLACE_STEP_FLOATS,
                                                                d_table, (uint32_t)p->neg_table_size, stride, n4, 128, (uint32_t)(17 * rep + 1));
                    cudaEventRecord(e1, st);
                    if (cudaEventSynchronize(e1) == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) ms_k = std::min(ms_k, ms);
                }
                ctx->launches += 2;
                if (ms_k < best) { best = ms_k; best_k = k; }
                if (ms_k < 1e29f) worst = std::max(worst, ms_k);
            }
            if (cudaGetLastError() == cudaSuccess && best < 1e29f) {
                m->syn0 = m->syn0_alloc + (size_t)best_k * SGNS_PLACE_STEP_FLOATS;
                m->syn1neg = m->syn1neg_alloc + (size_t)best_k * SGNS_PLACE_STEP_FLOATS;
                t0 = m->syn0; t1 = m->syn1neg;
                ctx->phase_ms["sgns_placement"] = (float)best_k; ctx->phase_ms["sgns_placement_best_us"] = best * 1e3f; ctx->phase_ms["sgns_placement_worst_us"] = worst * 1e3f;
            }
        } else cudaGetLastError();
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
    cudaMemsetAsync(t0, 0, nel * sizeof(float), st);
    cudaMemsetAsync(t1, 0, nel * sizeof(float), st);
    cudaMemsetAsync(d_pairs, 0, 2 * sizeof(unsigned long long), st);
    if (V) cudaMemcpyAsync(m->id_of_word, order.data(), sizeof(int32_t) * (size_t)V, cudaMemcpyHostToDevice, st);
    if (n_ids) cudaMemcpyAsync(d_word_of_id, word_of_id.data(), sizeof(int32_t) * (size_t)n_ids, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_table, table.data(), sizeof(int32_t) * table.size(), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_exp, exp_table.data(), sizeof(float) * exp_table.size(), cudaMemcpyHostToDevice, st);
    if (V) {
        int64_t total = (int64_t)V * p->dim;
        k_init_syn0<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(t0, V, p->dim, stride, p->seed);
        ctx->launches++;
    }
    // where the tables landed (bits 8-19 and 20-39 of their addresses): a diagnostic for the run-to-run spread of the kernel time
    ctx->phase_ms["sgns_syn0_addr_lo"] = (float)(((uintptr_t)t0 >> 8) & 0xFFF); ctx->phase_ms["sgns_syn0_addr_hi"] = (float)(((uintptr_t)t0 >> 20) & 0xFFFFF);
    ctx->phase_ms["sgns_syn1_addr_lo"] = (float)(((uintptr_t)t1 >> 8) & 0xFFF); ctx->phase_ms["sgns_syn1_addr_hi"] = (float)(((uintptr_t)t1 >> 20) & 0xFFFFF);
    ctx->phase_ms["sgns_rounds"] = 0.f; ctx->phase_ms["sgns_sync"] = 0.f; ctx->phase_ms["sgns_transport"] = 0.f; ctx->phase_ms["sgns_dp_setup"] = 0.f;
    ctx->phase_ms["compact"] = 0.f; ctx->phase_ms["sgns"] = 0.f;

    if (train) {
        // ---- compacted corpus in vocabulary indices (position-major, all corpora concatenated)
        dge_phase_timer t_prep(ctx, "compact");
        int64_t first = 0;
        for (int i = 0; i < n_corpora; i++) {
            if (corpora[i]->n > 0) {
                k_compact<<<(unsigned)((corpora[i]->n + 255) / 256), 256, 0, st>>>(corpora[i]->tok, corpora[i]->n, corpora[i]->L,
                                                                                 d_word_of_id, d_wtok, n_sent, first, Lmax, d_pairs + 1);
                ctx->launches++;
            }
            first += corpora[i]->n;
        }
        t_prep.stop();

        sgns_args a;
        memset(&a, 0, sizeof(a));
        a.wtok = d_wtok; a.n_sent = n_sent; a.s_off = s_off; a.n_global = std::max<int64_t>(1, n_global);
        a.neg_table = d_table; a.exp_table = d_exp;
        a.syn0 = t0; a.syn1neg = t1;
        a.V = V; a.dim = p->dim; a.stride = stride; a.n4 = n4; a.window = p->window; a.negative = p->negative; a.epochs = p->epochs;
        a.neg_table_size = p->neg_table_size; a.exp_table_size = p->exp_table_size; a.Lmax = Lmax;
        a.lr = p->lr; a.min_lr = p->min_lr; a.seed = p->seed; a.pairs = d_pairs;
        a.dbg = dbg;

        // ---- schedule
        //  concurrency 1            : kernel A, one group: the oracle's sequential order (parity tests)
        //  schedule SENTENCE        : kernel A, Hogwild with plain stores, `concurrency` sentences in flight (0 = fill)
        //  schedule ITEMS (default) : kernel B, (sentence, centre) items with L2 reductions; in flight:
        //                             concurrency * Lmax items, or (auto) min(full GPU, 8 * V / (negative + 1)) so
        //                             that a row sees at most ~8 concurrent stale updates
        const bool sequential = (p->concurrency == 1 && !(dbg & (2048 | 4))) || p->schedule == DGE_SCHEDULE_SENTENCE;
        int auto_wt_warps = 0, auto_wt_hot = 0;   // kernel F with write-through words chosen automatically: warps per SM, words
        {   // kernel variant; 4-lane groups need (sentences in flight allowed) >= what fills the GPU with them
            const int64_t allowed = p->concurrency > 0 ? (int64_t)p->concurrency * Lmax : (int64_t)SGNS_STALE_BOUND * V / (p->negative + 1);
            const bool narrow = (dbg & 32) || ((dbg & 65536) && !(dbg & 2) && allowed >= (int64_t)ctx->sm_count * 4 * 32);
            // fewer pairs in flight than the 8-lane kernel needs to fill the GPU (5 blocks x 16 groups per SM): latency-bound
            const bool tp = (dbg & 64) || ((dbg & 65536) && !(dbg & (128 | 32)) && allowed < (int64_t)ctx->sm_count * 5 * 16);
            // negative table in shared memory (16-bit entries): vocabularies below 65 536 words and a table that fits beside the rest
            const bool smem_neg = (dbg & 1024) != 0 && V <= 65535 && p->neg_table_size <= 100000;
            // Kernels F / G (sentence-resident: intra-sentence updates in the reference's order) are the rule for rows of up to
            // 32 slots: G (a block per sentence) for narrow rows when one lane group per position fits a block, F (a warp per
            // sentence) otherwise.  The item kernels B-E, which put the centres of one sentence in flight at once and drift
            // from the oracle, remain behind flags for A/B measurements, and kernel B for rows beyond 32 slots.
            const bool forced_other = (dbg & (32 | 64 | 256 | 512 | 1024 | 65536)) != 0;
            const bool sent = !forced_other;
            // kernel G needs one lane group per position of the longest sentence in a block of at most 256 threads
            const int G_of = n4 <= 8 ? 8 : (n4 <= 16 ? 16 : 32);
            const bool blk_fits = ((Lmax + 32 / G_of - 1) / (32 / G_of)) * 32 <= 256;
            bool force_warp_per_sentence = (dbg & 2048) != 0 && !(dbg & 4);
            // ---- automatic choice between kernel G (a block per sentence, hub-bounded) and kernel F with write-through words
            if (!dbg && p->concurrency == 0 && !sequential && !multi && blk_fits && n4 <= 8 && V > 0) {
                const double n_sents = (double)std::max<int64_t>(1, n_global);
                const int64_t g_sent = std::min<int64_t>(2 * (int64_t)ctx->sm_count, std::max<int64_t>(1, (int64_t)(SGNS_HUB_BOUND / std::max((double)cs[0] / n_sents, 1e-9))));
                int warps = (int)std::min<int64_t>(20, (int64_t)V / ctx->sm_count);
                int hot = 0;
                while (warps >= 1) { // the fewest write-through words (a power of two) that satisfy the hub bound at this many sentences
                    const double n_f = (double)warps * ctx->sm_count;
                    hot = 0;
                    while (hot < V && n_f * (double)cs[hot] / n_sents > SGNS_HUB_BOUND) hot = hot ? hot * 2 : 1;
                    if (hot >= V || hot <= SGNS_WT_MAX_WORDS) break; // every word, or few enough
                    warps--;
                }
                if (warps >= 1 && (double)warps * ctx->sm_count >= SGNS_F_OVER_G_SENTENCES * (double)g_sent &&
                    (int64_t)warps * ctx->sm_count <= std::max<int64_t>(1, n_sent)) {
                    force_warp_per_sentence = true;
                    auto_wt_warps = warps;
                    auto_wt_hot = std::min(hot, V);
                }
            }
            const bool blk = blk_fits && n4 <= 8 && !(dbg & 8) && !force_warp_per_sentence;
            // kernel I (DGE_SGNS_F_PAIR_WARPS): 8 warps per block unless bits 12-15 of the flags name another count (4 .. 16)
            // kernel F with the rows of the next unit requested ahead (DGE_SGNS_F_ROW_PREFETCH; blocks of at most 12 warps)
            // ... or into shared memory (DGE_SGNS_F_ROW_PREFETCH_SMEM: cp.async, the block keeps its 20 warps)
            const int prefetch_f = (p->negative <= SGNS_CH && n4 <= 8) ? ((dbg & (1 << 26)) ? 2 : ((dbg & (1 << 24)) ? 1 : 0)) : 0;
            const int pw_req = (dbg >> 12) & 15;
            const int pair_warps = ((dbg & 262144) && Lmax <= 32 && p->negative <= 7) ? (pw_req >= 4 ? pw_req : 8) : 0;
            pick_variant(n4, p->negative, narrow, tp, (dbg & 256) != 0, (dbg & 512) != 0, (dbg >> 12) & 15, smem_neg, sent && !forced_other, blk,
                         ((Lmax + 32 / G_of - 1) / (32 / G_of)) * 32, (dbg & 131072) != 0, pair_warps, (dbg & 524288) != 0 && Lmax <= 32, prefetch_f, &var); // n4 <= 128 was checked
        }
        sgns_kernel_t fn = sequential ? var.seq : var.items;
        const int G = sequential ? var.G_seq : var.G_items;
        uint64_t la = 1, lc = 0;
        for (int k = 0; k < SGNS_MAX_NEG; k++) { // state after k+1 steps of x -> x*MUL + ADD
            la = la * LCG_MUL; lc = lc * LCG_MUL + LCG_ADD;
            a.lcg_a[k] = la; a.lcg_c[k] = lc;
        }
        const bool pipe_kernel = !sequential && var.items_code == 10;    // kernel H: kernel G with the block's sentences pipelined
        const bool wave_kernel = !sequential && var.items_code == 11;    // kernel I: a warp per pair
        const bool duo_kernel = !sequential && var.items_code == 12;     // kernel J: critical + helper warps
        const bool block_kernel = !sequential && (var.items_code == 9 || pipe_kernel || wave_kernel || duo_kernel);    // kernel G: a.n_groups counts BLOCKS (sentences in flight)
        // sentences in flight the hottest word allows (global counts and sentences in a data-parallel run)
        const double hub_p = std::min(1.0, (double)cs[0] / (double)std::max<int64_t>(1, n_global));
        const int64_t hub_sentences = std::max<int64_t>(1, (int64_t)(SGNS_HUB_BOUND / std::max(hub_p, 1e-9)));
        ctx->phase_ms["sgns_hub_bound"] = (float)hub_sentences;
        const bool sent_kernel = !sequential && (var.items_code == 8 || block_kernel);   // kernel F: a.n_groups counts WARPS (sentences in flight)
        // negative table in shared memory as increment bitmap + per-word prefix (kernel F, narrow rows): exact iff the table never
        // grows by more than one word per slot, which its construction guarantees; checked all the same
        uint32_t *d_negbits = nullptr;
        const int nwords = (p->neg_table_size + 31) / 32;
        if ((sent_kernel && n4 <= 8) || block_kernel) {
            std::vector<uint32_t> nb((size_t)2 * nwords, 0u);
            bool exact = true;
            for (int32_t i = 0; i < p->neg_table_size; i++) {
                if ((i & 31) == 0) nb[(size_t)nwords + (i >> 5)] = (uint32_t)table[i];
                else {
                    const int32_t inc = table[i] - table[i - 1];
                    if (inc == 1) nb[i >> 5] |= 1u << (i & 31);
                    else if (inc != 0) exact = false;
                }
            }
            if (exact && tmp.get(&d_negbits, (size_t)2 * nwords) == cudaSuccess)
                cudaMemcpyAsync(d_negbits, nb.data(), sizeof(uint32_t) * nb.size(), cudaMemcpyHostToDevice, st);
            else d_negbits = nullptr;
            cudaStreamSynchronize(st);   // nb goes out of scope
        }
        a.neg_bits = d_negbits;
        const bool big_block = !sequential && (var.items_code == 7 || (sent_kernel && !block_kernel && n4 <= 8 && !(dbg & 16))); // one 640-thread block per SM
        int threads = big_block ? 640 : 128;
        // kernel F: bits 12-15 of the flags name the warps of a block (A/B runs: one block per SM with `concurrency` = SMs x warps)
        if (sent_kernel && !block_kernel && ((dbg >> 12) & 15)) threads = 32 * ((dbg >> 12) & 15);
        if (sent_kernel && !block_kernel && auto_wt_warps) threads = 32 * auto_wt_warps;   // one block per SM
        if (sent_kernel && !block_kernel && (dbg & (1 << 24)) && !(dbg & (1 << 26)) && p->negative <= SGNS_CH && n4 <= 8 && threads > 384) threads = 384;
        if (block_kernel) threads = ((Lmax + 32 / G - 1) / (32 / G)) * 32;   // one lane group per position of the longest sentence
        if (wave_kernel) { const int pw = (dbg >> 12) & 15; threads = 32 * (pw >= 4 ? pw : 8); }
        if (duo_kernel) threads *= 2;   // as many helper warps as critical warps
        // kernel J: stages of the row ring, as many (up to 4) as leave room for two blocks per SM
        auto duo_smem = [&](int stages) {
            return 16 * ((size_t)stages * Lmax * (SGNS_CH + 1) * n4 + 3 * (size_t)Lmax * n4) + sizeof(float) * (2 * (size_t)Lmax * 8 + (size_t)p->exp_table_size) +
                   sizeof(int32_t) * 5 * (size_t)Lmax + (d_negbits ? sizeof(uint32_t) * (size_t)2 * nwords : 0) +
                   sizeof(int32_t) * (size_t)Lmax * (size_t)Lmax * (size_t)std::max(1, p->negative);
        };
        a.stages = 4;
        // kernel F: write-through words (bits 20-23 of the flags: v -> the 2^(v-1) most frequent words; 0 = none)
        a.hot = 0;
        if (!sequential && var.items_code == 8 && ((dbg >> 20) & 15)) a.hot = 1 << (((dbg >> 20) & 15) - 1);
        if (!sequential && var.items_code == 8 && auto_wt_warps) a.hot = auto_wt_hot;
        ctx->phase_ms["sgns_write_through"] = (float)a.hot;
        // kernel F: sentences handed out from a counter (DGE_SGNS_F_DYNAMIC, and whenever the write-through schedule is automatic)
        unsigned long long *d_next = nullptr;
        if (!sequential && var.items_code == 8 && !(dbg & 8) && ((dbg & (1 << 25)) || auto_wt_warps)) {
            if (tmp.get(&d_next, 1) != cudaSuccess) { d_next = nullptr; cudaGetLastError(); }
        }
        a.next = d_next;
        if (duo_kernel) {
            const int req = (dbg >> 12) & 15;   // bits 12-15 of the flags: a fixed stage count (2 .. 4) for A/B runs
            if (req >= 2 && req <= 4) a.stages = req;
            else while (a.stages > 2 && 2 * (duo_smem(a.stages) + 1024) > (size_t)227 * 1024) a.stages--;
        }
        int gpb = threads / G;
        // dynamic shared memory: the sigmoid table, plus (pipelined item kernel) one staged sentence per group
        auto smem_for = [&](int thr) {
            if (duo_kernel) return duo_smem(a.stages);
            if (wave_kernel)
                return sizeof(float) * (3 * (size_t)Lmax * 32 + (size_t)p->exp_table_size) + sizeof(int32_t) * 6 * (size_t)Lmax +
                       (d_negbits ? sizeof(uint32_t) * (size_t)2 * nwords : 0) + sizeof(int32_t) * (size_t)Lmax * (size_t)Lmax * (size_t)std::max(1, p->negative) +
                       2 * (size_t)Lmax * (size_t)Lmax;
            if (pipe_kernel)
                return 3 * ((size_t)Lmax * (size_t)n4 * 16 + sizeof(int32_t) * (size_t)Lmax + sizeof(int32_t) * (size_t)Lmax * (size_t)Lmax * (size_t)std::max(1, p->negative)) +
                       sizeof(float) * (size_t)p->exp_table_size + 24 * sizeof(int32_t) + (d_negbits ? sizeof(uint32_t) * (size_t)2 * nwords : 0);
            if (block_kernel)
                return (size_t)Lmax * (size_t)n4 * 16 + sizeof(float) * (size_t)p->exp_table_size + sizeof(int32_t) * (size_t)Lmax +
                       (d_negbits ? sizeof(uint32_t) * (size_t)2 * nwords : 0) + sizeof(int32_t) * (size_t)Lmax * (size_t)Lmax * (size_t)std::max(1, p->negative);
            if (sent_kernel)
                return (size_t)(thr / 32) * (size_t)Lmax * (size_t)n4 * 16 + sizeof(float) * (size_t)p->exp_table_size +
                       sizeof(int32_t) * 2 * (size_t)(thr / 32) * (size_t)Lmax + (d_negbits ? sizeof(uint32_t) * (size_t)2 * nwords : 0) +
                       (((dbg & (1 << 26)) && p->negative <= SGNS_CH && n4 <= 8) ? (size_t)(thr / 32) * 2 * (SGNS_CH + 2) * 32 * 16 : 0);
            return sizeof(float) * (size_t)p->exp_table_size + (sequential ? 0 : sizeof(int32_t) * (size_t)(thr / G) * (size_t)Lmax) +
                   (!sequential && var.items_code == 5 ? (size_t)thr * 2 * (SGNS_CH + 1) * 16 : 0) + // kernel C': two row stages per lane
                   (var.items_code == 7 ? (((size_t)p->neg_table_size * 2 + 15) / 16) * 16 : 0);
        };
        size_t smem = smem_for(threads);
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem);
        if (per_sm < 1) per_sm = 1;
        const int GPW = 32 / G;
        const int64_t full_groups = (int64_t)ctx->sm_count * per_sm * gpb;
        const int64_t units = sequential ? std::max<int64_t>(1, n_sent) : std::max<int64_t>(1, n_sent * (int64_t)Lmax);
        int64_t want;
        if (sent_kernel && !block_kernel && auto_wt_warps) want = (int64_t)auto_wt_warps * ctx->sm_count * GPW;
        else if (p->concurrency > 0) want = sequential ? (int64_t)p->concurrency : (int64_t)p->concurrency * (sent_kernel ? GPW : Lmax);
        else if (sequential) want = full_groups;
        else want = std::min<int64_t>(full_groups, std::max<int64_t>(gpb, (int64_t)SGNS_STALE_BOUND * V / (p->negative + 1)));
        // kernel F on wider rows (a warp owns a sentence, one pair at a time): the same schedule -- sentence counter, write-through
        // words chosen for the sentences in flight (at most one per vocabulary word) -- instead of the hub bound on the
        // sentences in flight; also on the ranks of a data-parallel run (global counts, this rank's sentences in flight)
        const bool auto_wide = !dbg && p->concurrency == 0 && sent_kernel && !block_kernel && var.items_code == 8 && n4 > 8 && !auto_wt_warps;
        if (auto_wide) want = std::min<int64_t>(want, std::max<int64_t>(1, (int64_t)V) * GPW);
        if (sent_kernel && p->concurrency == 0 && !auto_wt_warps && !auto_wide) want = std::min<int64_t>(want, hub_sentences * GPW);   // kernel F: a warp (GPW groups) per sentence
        if (!sequential && (dbg & 8)) want = 1; // one warp, one item at a time, strictly in corpus order (arithmetic check against the oracle)
        want = std::max<int64_t>(1, std::min(want, sent_kernel ? std::max<int64_t>(1, n_sent) * GPW : units));
        while (!big_block && !block_kernel && threads > 32 && threads > G && want < (int64_t)ctx->sm_count * gpb) { threads >>= 1; gpb = threads / G; }
        if (sequential && want < gpb) { gpb = (int)want; threads = gpb * G; } // kernel B keeps whole warps
        int blocks = (int)((want + gpb - 1) / gpb);
        a.n_groups = (int64_t)blocks * gpb;
        if (!sequential && (dbg & 8)) a.n_groups = 1; // the single warp advances one item at a time
        if (sent_kernel && !block_kernel) { // groups in flight -> warps (= sentences) in flight
            a.n_groups = std::max<int64_t>(1, ((int64_t)blocks * gpb) / GPW);
            if (dbg & 8) { a.n_groups = 1; blocks = 1; threads = 32; }
        }
        if (block_kernel) { // sentences in flight = blocks: `concurrency`, or what fills the GPU, or the staleness bound (pairs in flight / Lmax)
            const int64_t full_blocks = (int64_t)ctx->sm_count * ((wave_kernel || duo_kernel) ? std::min(per_sm, 2) : per_sm);
            int64_t wb = p->concurrency > 0 ? p->concurrency : std::min<int64_t>(full_blocks, hub_sentences);
            wb = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(wb, full_blocks), std::max<int64_t>(1, n_sent)));
            if ((wave_kernel || duo_kernel) && p->concurrency == 0 && wb > ctx->sm_count) wb -= wb % ctx->sm_count;   // the same number of sentences on every SM
            blocks = (int)wb;
            a.n_groups = blocks;
        }
        ctx->phase_ms["sgns_groups"] = (float)a.n_groups;
        if (auto_wide) {
            const double n_sents = (double)std::max<int64_t>(1, n_global), n_f = (double)a.n_groups;
            int hot = 0;
            while (hot < V && hot <= SGNS_WT_MAX_WORDS && n_f * (double)cs[hot] / n_sents > SGNS_HUB_BOUND) hot = hot ? hot * 2 : 1;
            a.hot = std::min(hot, V);
            ctx->phase_ms["sgns_write_through"] = (float)a.hot;
            if (!d_next && tmp.get(&d_next, 1) != cudaSuccess) { d_next = nullptr; cudaGetLastError(); }
            a.next = d_next;
        }
        ctx->phase_ms["sgns_kernel"] = (float)(sequential ? 0 : var.items_code);
        // ---- launches.  Single GPU: one launch over all epochs and sentences.  Data-parallel (the ctx has a
        // communicator; or sync_rounds > 0, where the exchange is the identity): each epoch is cut into `rounds` slices
        // of the local sentences; after every slice the replicas are recombined from the per-rank deltas
        // (comm.cu dge_dp_exchange: one peer-memory kernel over NVLink, or NCCL all-reduces; rule p->combine).
        int rounds = 1;
        if (multi || p->sync_rounds > 0) {
            rounds = p->sync_rounds > 0 ? p->sync_rounds : (int)std::max<int64_t>(8, (max_sent + (1 << 19) - 1) >> 19);
            rounds = (int)std::min<int64_t>(rounds, std::max<int64_t>(1, max_sent));
        }
        dge_dp *dp = nullptr;
        ctx->phase_ms["sgns_dp_setup"] = 0.f;
        if (multi) {
            dge_phase_timer t_dp(ctx, "sgns_dp_setup");   // peer mapping of the replicas (cudaIpc), base slices
            const int rc = dge_dp_begin(ctx, t0, t1, V, stride, n4, p->combine == DGE_COMBINE_DEFAULT ? DGE_COMBINE_ALIGNED : p->combine,
                                        p->transport, &dp);
            if (rc != DGE_OK) { model_release(m); return rc; }   // collective: every rank takes this way out
            t_dp.stop();
        }
        ctx->phase_ms["sgns_rounds"] = (float)rounds;
        int rc = DGE_OK, any_error = 0;
        dge_phase_timer t_sgns(ctx, "sgns");
        smem = smem_for(threads);
        if (rounds == 1 && !multi) {
            a.ep_lo = 0; a.ep_hi = p->epochs; a.s_lo = 0; a.s_hi = n_sent;
            if (d_next) cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), st);
            fn<<<blocks, threads, smem, st>>>(a);
            ctx->launches++;
        } else {
            for (int ep = 0; ep < p->epochs && rc == DGE_OK && !any_error; ep++) {
                for (int r = 0; r < rounds && rc == DGE_OK && !any_error; r++) {
                    a.ep_lo = ep; a.ep_hi = ep + 1;
                    a.s_lo = n_sent * r / rounds; a.s_hi = n_sent * (r + 1) / rounds;
                    int launch_err = 0;
                    if (a.s_hi > a.s_lo) {
                        if (d_next) cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), st);
                        fn<<<blocks, threads, smem, st>>>(a);
                        ctx->launches++;
                        launch_err = cudaGetLastError() != cudaSuccess;
                    }
                    // every exchange agrees on the error status (a rank in trouble raises it, all ranks stop together);
                    // the host waits for it every round: the rounds are tens of milliseconds long
                    if (dp) rc = dge_dp_exchange(dp, launch_err, true, &any_error);
                    else if (launch_err) any_error = 1;
                }
            }
        }
        t_sgns.stop();
        if (dp) { ctx->phase_ms["sgns_sync"] = dge_dp_ms(dp); dge_dp_end(dp); }
        if (rc != DGE_OK) { model_release(m); return rc; }
        ce = cudaGetLastError();
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        local = (ce != cudaSuccess || any_error)
                    ? dge_fail(ctx, DGE_E_CUDA, std::string("dge_sgns_train: kernel: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "a rank reported a launch failure"))
                    : DGE_OK;
        if (multi) local = dge_comm_agree(ctx, local, "dge_sgns_train (training)");
        if (local != DGE_OK) { model_release(m); return local; }
        unsigned long long h_pairs[2] = {0, 0};
        cudaMemcpy(h_pairs, d_pairs, sizeof(h_pairs), cudaMemcpyDeviceToHost);
        m->pairs = (int64_t)h_pairs[0];
        m->words = (int64_t)h_pairs[1];
    }
    if (multi) { // the model's own copy of the trained replica (the arena is reused by the next call)
        cudaMemcpyAsync(m->syn0, t0, nel * sizeof(float), cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(m->syn1neg, t1, nel * sizeof(float), cudaMemcpyDeviceToDevice, st);
    }
    ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) { model_release(m); return dge_fail(ctx, DGE_E_CUDA, std::string("dge_sgns_train: ") + cudaGetErrorString(ce)); }
    *out = m;
    return DGE_OK;
}

int dge_model_shape(const dge_model *m, int32_t *V, int32_t *dim, int64_t *pairs) {
    if (!m) return dge_fail(nullptr, DGE_E_INVALID, "dge_model_shape: model is NULL");
    if (V) *V = m->V;
    if (dim) *dim = m->dim;
    if (pairs) *pairs = m->pairs;
    return DGE_OK;
}

int dge_model_vectors(const dge_model *m, float *syn0, float *syn1neg, int32_t *id_of_word) {
    if (!m) return dge_fail(nullptr, DGE_E_INVALID, "dge_model_vectors: model is NULL");
    dge_ctx *ctx = m->ctx;
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (m->V == 0) return DGE_OK;
    dge_phase_timer t(ctx, "vectors_d2h");
    size_t row = sizeof(float) * (size_t)m->dim, pitch = sizeof(float) * (size_t)m->stride;
    if (syn0) DGE_CUDA(ctx, cudaMemcpy2DAsync(syn0, row, m->syn0, pitch, row, (size_t)m->V, cudaMemcpyDeviceToHost, ctx->stream));
    if (syn1neg) DGE_CUDA(ctx, cudaMemcpy2DAsync(syn1neg, row, m->syn1neg, pitch, row, (size_t)m->V, cudaMemcpyDeviceToHost, ctx->stream));
    if (id_of_word) DGE_CUDA(ctx, cudaMemcpyAsync(id_of_word, m->id_of_word, sizeof(int32_t) * (size_t)m->V, cudaMemcpyDeviceToHost, ctx->stream));
    t.stop();
    DGE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DGE_OK;
}

int dge_model_stats(const dge_model *m, double *mean_row_norm, double *max_abs, int64_t *n_nonfinite) {
    if (!m) return dge_fail(nullptr, DGE_E_INVALID, "dge_model_stats: model is NULL");
    dge_ctx *ctx = m->ctx;
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    double h[3] = {0.0, 0.0, 0.0}; // sum of row norms, max |element|, non-finite count (as raw bits of an u64)
    if (m->V > 0) {
        double *d = nullptr;
        DGE_CUDA(ctx, dge_malloc(ctx, &d, 3));
        cudaMemsetAsync(d, 0, 3 * sizeof(double), ctx->stream);
        k_model_stats<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(m->syn0, m->syn1neg, m->V, m->dim, m->stride, d,
                                                                 reinterpret_cast<unsigned long long *>(d + 2));
        ctx->launches++;
        cudaError_t e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        dge_free(ctx, d);
        if (e != cudaSuccess) return dge_fail(ctx, DGE_E_CUDA, std::string("dge_model_stats: ") + cudaGetErrorString(e));
    }
    unsigned long long nb;
    memcpy(&nb, &h[2], sizeof(nb));
    if (mean_row_norm) *mean_row_norm = m->V > 0 ? h[0] / (double)m->V : 0.0;
    if (max_abs) *max_abs = h[1];
    if (n_nonfinite) *n_nonfinite = (int64_t)nb;
    return DGE_OK;
}

void dge_model_free(dge_model *m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    model_release(m);
}

} // extern "C"
