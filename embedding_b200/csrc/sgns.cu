// sgns.cu -- stage 2 on device: skip-gram with negative sampling over the walk corpus.
//
// Stands under DeepWalk.learnEmbedding (DeepWalk.java:32-83): Word2Vec.Builder() :73-76 and
// w2v.fit() :79, whose arithmetic is DL4J 0.7.2 / libnd4j (external, un-vendored).  The kernel
// follows the published word2vec skip-gram update with the DL4J parameterisation listed in
// SURVEY.md 8(a) A14 and restated in oracle/sgns_oracle.c (same pair / negative enumeration,
// same per-sentence RNG and learning-rate schedule, so a sequential schedule reproduces the
// oracle to fp32 tolerance).  Pure negative sampling (north_star); no hierarchical softmax.
//
// Design: a GROUP of G lanes owns one sentence at a time (G=1 for dim<=32: a row is 1-8 float4
// held in registers; G=16/32 for wide rows: one coalesced 128-bit load per lane, shuffle-reduced
// dot product).  Rows are read and written through L2 (ld/st.global.cg): L2 is the coherence
// point of the Hogwild-style, atomic-free updates.  No tensor cores: the work is K+1 dot
// products of length dim per pair, not a dense contraction.
#include "dge_internal.cuh"
#include <algorithm>
#include <cmath>

#define SGNS_MAX_CORPORA 4
#define SGNS_MAX_EXP 6.0f
#define LCG_MUL 25214903917ULL
#define LCG_ADD 11ULL

struct sgns_corpus_view {
    const int32_t *tok; // position-major [L][n]
    int64_t n;
    int64_t first;      // global index of its first sentence
    int32_t L;
};

struct sgns_args {
    sgns_corpus_view cv[SGNS_MAX_CORPORA];
    int32_t n_corpora;
    int64_t n_sent;
    const int32_t *word_of_id;
    const int32_t *neg_table;
    const float *exp_table;
    float *syn0, *syn1neg;
    int32_t V, dim, stride, window, negative, epochs, neg_table_size, exp_table_size, Lmax;
    float lr, min_lr;
    uint64_t seed;
    unsigned long long *pairs;
    int64_t n_groups;
};

__host__ __device__ static inline uint64_t sgns_sentence_rng(uint64_t seed, int32_t epoch, int64_t sentence) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(sentence + 1) + 0xD1B54A32D192ED03ULL * (uint64_t)epoch;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return z & 0x7FFFFFFFFFFFFFFFULL;
}
__device__ static inline uint64_t lcg_abs(uint64_t r) { // Math.abs(r * 25214903917L + 11) on a Java long
    int64_t x = (int64_t)(r * LCG_MUL + LCG_ADD);
    return (uint64_t)(x < 0 ? -x : x);
}

__global__ void k_hist(const int32_t *__restrict__ tok, int64_t total, unsigned int *__restrict__ cnt) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int32_t t = tok[i];
        if (t >= 0) atomicAdd(&cnt[t], 1u);
    }
}

// syn0 = (U[0,1) - 0.5) / dim from Philox(seed); same element stream as ora_init_syn0
__global__ void k_init_syn0(float *__restrict__ syn0, int32_t V, int32_t dim, int32_t stride, uint64_t seed) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t total = (int64_t)V * dim;
    if (e >= total) return;
    uint32_t r[4];
    uint64_t blk = (uint64_t)e >> 2;
    dge_philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), 0x5347u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    float u = (float)(r[e & 3] >> 8) * 0x1.0p-24f;
    int64_t row = e / dim;
    int32_t c = (int32_t)(e - row * dim);
    syn0[row * stride + c] = __fdiv_rn(u - 0.5f, (float)dim);
}

template <int G>
__device__ __forceinline__ float group_sum(float v, unsigned gmask) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
}

template <int G, int VPL>
__global__ void __launch_bounds__(128)
k_sgns(const sgns_args a) {
    extern __shared__ int32_t smem[];
    float *s_exp = reinterpret_cast<float *>(smem);
    int32_t *s_sent = smem + a.exp_table_size;
    const int gpb = blockDim.x / G;
    const int gl = threadIdx.x / G;
    const int lane = threadIdx.x % G;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
    for (int i = threadIdx.x; i < a.exp_table_size; i += blockDim.x) s_exp[i] = a.exp_table[i];
    __syncthreads();

    const int64_t gid = (int64_t)blockIdx.x * gpb + gl;
    const int n4 = a.stride >> 2;
    const int E = a.exp_table_size;
    const float idx_scale = (float)E / SGNS_MAX_EXP / 2.0f;
    const int win = a.window;
    unsigned long long pairs = 0;
    // float4 slots of this lane: q = lane + v*G, active when q < n4
    for (int ep = 0; ep < a.epochs; ep++) {
        for (int64_t s = gid; s < a.n_sent; s += a.n_groups) {
            // ---- locate the sentence and compact it to vocabulary indices
            int ci = 0;
            while (ci + 1 < a.n_corpora && s >= a.cv[ci + 1].first) ci++;
            const sgns_corpus_view cv = a.cv[ci];
            const int64_t sl = s - cv.first;
            int n = 0;
            for (int j = 0; j < cv.L; j++) {
                int32_t id = cv.tok[(int64_t)j * cv.n + sl];
                if (id < 0) continue;
                int32_t wd = a.word_of_id[id];
                if (wd < 0) continue;
                if (lane == 0) s_sent[n * gpb + gl] = wd;
                n++;
            }
            if (G > 1) __syncwarp(gmask);
            double progress = (double)((int64_t)ep * a.n_sent + s) / (double)((int64_t)a.epochs * a.n_sent);
            float alpha = a.lr * (float)(1.0 - progress);
            if (alpha < a.min_lr) alpha = a.min_lr;
            uint64_t r = sgns_sentence_rng(a.seed, ep, s);
            for (int i = 0; i < n; i++) {
                r = lcg_abs(r);
                const int b = (int32_t)(uint32_t)r % win;
                const int32_t w1 = s_sent[i * gpb + gl];
                const int end = win * 2 + 1 - b;
                for (int aa = b; aa < end; aa++) {
                    if (aa == win) continue;
                    const int c = i - win + aa;
                    if (c < 0 || c >= n) continue;
                    const int32_t last = s_sent[c * gpb + gl];
                    if (last == w1) continue;
                    uint64_t ns = r;
                    r = lcg_abs(r);
                    pairs++;
                    // ---- one (centre w1, context last) update
                    float4 v0[VPL], neu[VPL];
                    float4 *p0 = reinterpret_cast<float4 *>(a.syn0 + (int64_t)last * a.stride);
#pragma unroll
                    for (int v = 0; v < VPL; v++) {
                        int q = lane + v * G;
                        v0[v] = q < n4 ? __ldcg(p0 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                        neu[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    for (int k = 0; k < a.negative + 1; k++) {
                        int32_t target;
                        float label;
                        if (k == 0) { target = w1; label = 1.f; }
                        else {
                            if (a.V < 2) break;
                            ns = ns * LCG_MUL + LCG_ADD;
                            target = a.neg_table[(ns >> 16) % (uint64_t)a.neg_table_size];
                            if (target <= 0 || target >= a.V) target = (int32_t)(ns % (uint64_t)(a.V - 1)) + 1;
                            if (target == w1) continue;
                            label = 0.f;
                        }
                        float4 *p1 = reinterpret_cast<float4 *>(a.syn1neg + (int64_t)target * a.stride);
                        float4 v1[VPL];
                        float dot = 0.f;
#pragma unroll
                        for (int v = 0; v < VPL; v++) {
                            int q = lane + v * G;
                            v1[v] = q < n4 ? __ldcg(p1 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                            dot += v0[v].x * v1[v].x + v0[v].y * v1[v].y + v0[v].z * v1[v].z + v0[v].w * v1[v].w;
                        }
                        dot = group_sum<G>(dot, gmask);
                        float g;
                        if (dot > SGNS_MAX_EXP) g = (label - 1.f) * alpha;
                        else if (dot < -SGNS_MAX_EXP) g = (label - 0.f) * alpha;
                        else {
                            int idx = (int)((dot + SGNS_MAX_EXP) * idx_scale);
                            if (idx < 0 || idx >= E) continue;
                            g = (label - s_exp[idx]) * alpha;
                        }
#pragma unroll
                        for (int v = 0; v < VPL; v++) {
                            int q = lane + v * G;
                            neu[v].x += g * v1[v].x; neu[v].y += g * v1[v].y; neu[v].z += g * v1[v].z; neu[v].w += g * v1[v].w;
                            v1[v].x += g * v0[v].x; v1[v].y += g * v0[v].y; v1[v].z += g * v0[v].z; v1[v].w += g * v0[v].w;
                            if (q < n4) __stcg(p1 + q, v1[v]);
                        }
                    }
#pragma unroll
                    for (int v = 0; v < VPL; v++) {
                        int q = lane + v * G;
                        v0[v].x += neu[v].x; v0[v].y += neu[v].y; v0[v].z += neu[v].z; v0[v].w += neu[v].w;
                        if (q < n4) __stcg(p0 + q, v0[v]);
                    }
                }
            }
            if (G > 1) __syncwarp(gmask); // s_sent is reused by the next sentence
        }
    }
    if (lane == 0 && pairs) atomicAdd(a.pairs, pairs);
}

typedef void (*sgns_kernel_t)(const sgns_args);
struct sgns_variant { int G, VPL; sgns_kernel_t fn; };

static bool pick_variant(int n4, sgns_variant *out) {
    static const sgns_variant table[] = {
        {1, 1, k_sgns<1, 1>}, {1, 2, k_sgns<1, 2>}, {1, 3, k_sgns<1, 3>}, {1, 4, k_sgns<1, 4>},
        {1, 5, k_sgns<1, 5>}, {1, 6, k_sgns<1, 6>}, {1, 7, k_sgns<1, 7>}, {1, 8, k_sgns<1, 8>},
        {16, 1, k_sgns<16, 1>}, {32, 1, k_sgns<32, 1>}, {32, 2, k_sgns<32, 2>}, {32, 4, k_sgns<32, 4>},
    };
    int G, VPL;
    if (n4 <= 8) { G = 1; VPL = n4; }
    else if (n4 <= 16) { G = 16; VPL = 1; }
    else if (n4 <= 32) { G = 32; VPL = 1; }
    else if (n4 <= 64) { G = 32; VPL = 2; }
    else if (n4 <= 128) { G = 32; VPL = 4; }
    else return false;
    for (const auto &t : table)
        if (t.G == G && t.VPL == VPL) { *out = t; return true; }
    return false;
}

static void model_release(dge_model *m) {
    if (!m) return;
    cudaFree(m->syn0); cudaFree(m->syn1neg); cudaFree(m->id_of_word);
    delete m;
}

extern "C" {

void dge_sgns_default_params(dge_sgns_params *p) {
    if (!p) return;
    p->dim = 20; p->window = 8; p->negative = 5; p->min_count = 2; p->epochs = 1;
    p->neg_table_size = 100000; p->exp_table_size = 1000; p->concurrency = 0;
    p->lr = 0.025f; p->min_lr = 1e-4f; p->seed = 1;
}

int dge_sgns_train(dge_ctx *ctx, const dge_corpus *const *corpora, int32_t n_corpora, const dge_sgns_params *p,
                   dge_model **out) {
    if (!ctx) return dge_fail(nullptr, DGE_E_INVALID, "dge_sgns_train: ctx is NULL");
    if (!out) return dge_fail(ctx, DGE_E_INVALID, "dge_sgns_train: out is NULL");
    *out = nullptr;
    if (!corpora || !p || n_corpora < 1 || n_corpora > SGNS_MAX_CORPORA)
        return dge_fail(ctx, DGE_E_INVALID, "dge_sgns_train: need 1..4 corpora and params");
    if (p->dim < 1 || p->window < 1 || p->negative < 0 || p->epochs < 1 || p->neg_table_size < 1 ||
        p->exp_table_size < 2 || p->min_count < 0 || p->concurrency < 0)
        return dge_fail(ctx, DGE_E_INVALID, "dge_sgns_train: invalid hyper-parameter");
    int32_t n_ids = corpora[0] ? corpora[0]->n_ids : 0;
    int32_t Lmax = 0;
    int64_t n_sent = 0;
    for (int i = 0; i < n_corpora; i++) {
        if (!corpora[i]) return dge_fail(ctx, DGE_E_INVALID, "dge_sgns_train: NULL corpus");
        if (corpora[i]->ctx != ctx) return dge_fail(ctx, DGE_E_INVALID, "dge_sgns_train: corpus belongs to another ctx");
        if (corpora[i]->n_ids != n_ids) return dge_fail(ctx, DGE_E_INVALID, "dge_sgns_train: corpora have different id spaces");
        Lmax = std::max(Lmax, corpora[i]->L);
        n_sent += corpora[i]->n;
    }
    int32_t stride = (p->dim + 3) & ~3;
    sgns_variant var;
    if (!pick_variant(stride / 4, &var)) return dge_fail(ctx, DGE_E_LIMIT, "dge_sgns_train: dim must be <= 512");
    DGE_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;

    // ---- vocabulary: device histogram, host ranking (descending count, ties ascending id)
    dge_phase_timer t_vocab(ctx, "vocab");
    unsigned int *d_cnt = nullptr;
    DGE_CUDA(ctx, dge_malloc(&d_cnt, (size_t)n_ids));
    cudaMemsetAsync(d_cnt, 0, sizeof(unsigned int) * (size_t)(n_ids ? n_ids : 1), st);
    for (int i = 0; i < n_corpora; i++) {
        int64_t total = corpora[i]->n * (int64_t)corpora[i]->L;
        if (total) {
            k_hist<<<ctx->sm_count * 8, 256, 0, st>>>(corpora[i]->tok, total, d_cnt);
            ctx->launches++;
        }
    }
    std::vector<unsigned int> cnt((size_t)n_ids + 1);
    cudaError_t ce = cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(unsigned int) * (size_t)n_ids, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(d_cnt);
    if (ce != cudaSuccess) return dge_fail(ctx, DGE_E_CUDA, std::string("dge_sgns_train: histogram: ") + cudaGetErrorString(ce));
    std::vector<int32_t> order;
    order.reserve(n_ids);
    for (int32_t i = 0; i < n_ids; i++)
        if (cnt[i] > 0 && (int64_t)cnt[i] >= p->min_count) order.push_back(i);
    std::sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
        if (cnt[x] != cnt[y]) return cnt[x] > cnt[y];
        return x < y;
    });
    const int32_t V = (int32_t)order.size();
    std::vector<int32_t> word_of_id((size_t)n_ids + 1, -1);
    for (int32_t wd = 0; wd < V; wd++) word_of_id[order[wd]] = wd;
    // unigram^0.75 table (word2vec.c InitUnigramTable / DL4J makeTable), identical to ora_neg_table
    std::vector<int32_t> table((size_t)p->neg_table_size, 0);
    if (V > 0) {
        double pow_sum = 0;
        for (int32_t wd = 0; wd < V; wd++) pow_sum += pow((double)cnt[order[wd]], 0.75);
        int32_t wi = 0;
        double d1 = pow((double)cnt[order[0]], 0.75) / pow_sum;
        for (int32_t i = 0; i < p->neg_table_size; i++) {
            table[i] = wi;
            if ((double)i / (double)p->neg_table_size > d1) {
                if (wi < V - 1) wi++;
                d1 += pow((double)cnt[order[wi]], 0.75) / pow_sum;
            }
        }
    }
    std::vector<float> exp_table((size_t)p->exp_table_size);
    for (int32_t i = 0; i < p->exp_table_size; i++) {
        double e = exp(((double)i / (double)p->exp_table_size * 2.0 - 1.0) * (double)SGNS_MAX_EXP);
        exp_table[i] = (float)(e / (e + 1.0));
    }
    t_vocab.stop();

    dge_model *m = new dge_model();
    m->ctx = ctx; m->V = V; m->dim = p->dim; m->stride = stride;
    int32_t *d_word_of_id = nullptr, *d_table = nullptr;
    float *d_exp = nullptr;
    unsigned long long *d_pairs = nullptr;
    auto cleanup = [&]() { cudaFree(d_word_of_id); cudaFree(d_table); cudaFree(d_exp); cudaFree(d_pairs); };
    auto fail = [&](const std::string &msg) {
        cleanup();
        model_release(m);
        return dge_fail(ctx, DGE_E_CUDA, msg);
    };
    size_t nel = (size_t)(V ? V : 1) * (size_t)stride;
    if (dge_malloc(&m->syn0, nel) != cudaSuccess || dge_malloc(&m->syn1neg, nel) != cudaSuccess ||
        dge_malloc(&m->id_of_word, (size_t)V) != cudaSuccess || dge_malloc(&d_word_of_id, (size_t)n_ids) != cudaSuccess ||
        dge_malloc(&d_table, (size_t)p->neg_table_size) != cudaSuccess ||
        dge_malloc(&d_exp, (size_t)p->exp_table_size) != cudaSuccess || dge_malloc(&d_pairs, 1) != cudaSuccess)
        return fail("dge_sgns_train: cudaMalloc failed");
    cudaMemsetAsync(m->syn0, 0, nel * sizeof(float), st);
    cudaMemsetAsync(m->syn1neg, 0, nel * sizeof(float), st);
    cudaMemsetAsync(d_pairs, 0, sizeof(unsigned long long), st);
    if (V) cudaMemcpyAsync(m->id_of_word, order.data(), sizeof(int32_t) * (size_t)V, cudaMemcpyHostToDevice, st);
    if (n_ids) cudaMemcpyAsync(d_word_of_id, word_of_id.data(), sizeof(int32_t) * (size_t)n_ids, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_table, table.data(), sizeof(int32_t) * table.size(), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_exp, exp_table.data(), sizeof(float) * exp_table.size(), cudaMemcpyHostToDevice, st);
    if (V) {
        int64_t total = (int64_t)V * p->dim;
        k_init_syn0<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(m->syn0, V, p->dim, stride, p->seed);
        ctx->launches++;
    }

    if (V > 0 && n_sent > 0) {
        sgns_args a;
        memset(&a, 0, sizeof(a));
        int64_t first = 0;
        for (int i = 0; i < n_corpora; i++) {
            a.cv[i].tok = corpora[i]->tok; a.cv[i].n = corpora[i]->n; a.cv[i].L = corpora[i]->L; a.cv[i].first = first;
            first += corpora[i]->n;
        }
        a.n_corpora = n_corpora; a.n_sent = n_sent;
        a.word_of_id = d_word_of_id; a.neg_table = d_table; a.exp_table = d_exp;
        a.syn0 = m->syn0; a.syn1neg = m->syn1neg;
        a.V = V; a.dim = p->dim; a.stride = stride; a.window = p->window; a.negative = p->negative; a.epochs = p->epochs;
        a.neg_table_size = p->neg_table_size; a.exp_table_size = p->exp_table_size; a.Lmax = Lmax;
        a.lr = p->lr; a.min_lr = p->min_lr; a.seed = p->seed; a.pairs = d_pairs;

        int threads = 128;
        int gpb = threads / var.G;
        int64_t want_groups = p->concurrency;
        int blocks;
        size_t smem = 0;
        if (want_groups == 0) {
            smem = sizeof(int32_t) * ((size_t)p->exp_table_size + (size_t)Lmax * gpb);
            int per_sm = 0;
            cudaFuncSetAttribute(var.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, var.fn, threads, smem);
            if (per_sm < 1) per_sm = 1;
            blocks = ctx->sm_count * per_sm;
            int64_t need = (n_sent + gpb - 1) / gpb;
            if (blocks > need) blocks = (int)need;
        } else {
            // exactly `concurrency` sentences in flight (1 = the sequential schedule of the oracle)
            if (want_groups < gpb) { gpb = (int)want_groups; threads = gpb * var.G; }
            blocks = (int)((want_groups + gpb - 1) / gpb);
            smem = sizeof(int32_t) * ((size_t)p->exp_table_size + (size_t)Lmax * gpb);
            cudaFuncSetAttribute(var.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        a.n_groups = (int64_t)blocks * gpb;
        dge_phase_timer t_sgns(ctx, "sgns");
        var.fn<<<blocks, threads, smem, st>>>(a);
        ctx->launches++;
        t_sgns.stop();
        ce = cudaGetLastError();
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) return fail(std::string("dge_sgns_train: kernel: ") + cudaGetErrorString(ce));
        unsigned long long h_pairs = 0;
        cudaMemcpy(&h_pairs, d_pairs, sizeof(h_pairs), cudaMemcpyDeviceToHost);
        m->pairs = (int64_t)h_pairs;
    }
    ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) return fail(std::string("dge_sgns_train: ") + cudaGetErrorString(ce));
    cleanup();
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

void dge_model_free(dge_model *m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    model_release(m);
}

} // extern "C"
