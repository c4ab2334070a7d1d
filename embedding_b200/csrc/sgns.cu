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
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <cmath>
#include <cstdlib>

#define SGNS_MAX_CORPORA 4
#define SGNS_MAX_NEG 32
#define SGNS_MAX_EXP 6.0f
// automatic schedule: at most this many concurrent (stale) updates per embedding row (measured in
// profiles/quality_tract_r1.json: nDCG stays inside the oracle's seed-to-seed band up to ~8)
#define SGNS_STALE_BOUND 8
// sentence-resident kernels F / G: a sentence holds its context-row updates pending until its rows are flushed, so what
// must stay bounded is how many sentences IN FLIGHT contain the same (hottest) word: in_flight x P(sentence contains the
// most frequent word) <= SGNS_HUB_BOUND.  Calibrated on the full-size tract x 24 fixture (296 sentences in flight, the top
// word in 5.8 % of the sentences: 17 concurrent holders, kNN agreement with the oracle 0.886; 370 in flight: 0.72).
#define SGNS_HUB_BOUND 18.0
// Kernel F with WRITE-THROUGH words and a sentence COUNTER -- the default schedule for rows of up to 8 slots.
//  * Write-through: the rows of the most frequent words are re-read for every pair and their updates sent at once, so they
//    are never held pending; the hub bound then applies to the most frequent word that is NOT written through.
//  * Counter: sentences are handed out in corpus order from a device counter, so the warps sweep the corpus front together
//    whatever their speeds.  With the strided assignment a warp that runs slower (an SM sub-partition with one warp more,
//    an SM with one block more) falls behind in the corpus and in the learning-rate schedule, and the corpus' last part
//    (the spatial walks) is no longer trained last: agreement with the oracle 0.82 instead of 0.88 at 10 or 13 warps per SM,
//    row-norm collapse beyond ~V / 10 sentences in flight (profiles/r2s19, r2s20, r2s22).
// With both, the full-size tract x 24 fixture is reproduced with a FULL GPU of sentences in flight (20 warps per SM, 2 960
// sentences, 512-1 024 words written through: kNN agreement 0.887-0.892 against 0.873-0.882 between oracle runs, nDCG@5
// within 0.0004 of the oracle mean, 4.5 G pairs/s -- profiles/r2s24, r2s25), and the CA fixture (V = 1 848: every word
// written through) at 1 480-2 960 in flight (agreement 0.76-0.84 against 0.70-0.88 between oracle runs, 3.0 G pairs/s).
// Sentences in flight are kept <= V (one per vocabulary word; CA agrees better at 1 480 than at 2 960 for the same rate).
#define SGNS_WT_MAX_WORDS 2048
// measured per-sentence rates of the two kernels on narrow rows (pairs / s per sentence in flight): kernel G 9.6e6 per
// block, kernel F 2.06e6 per warp -- kernel F pays once it may hold ~4.7 x the sentences
#define SGNS_F_OVER_G_SENTENCES 4.7
#define LCG_MUL 25214903917ULL
#define LCG_ADD 11ULL

struct sgns_args {
    const int32_t *wtok;      // compacted corpus, vocabulary indices, position-major [Lmax][n_sent], -1 padded
    int64_t n_sent;
    const int32_t *neg_table;
    const float *exp_table;
    float *syn0, *syn1neg;
    int32_t V, dim, stride, n4, window, negative, epochs, neg_table_size, exp_table_size, Lmax;
    // stride = row pitch in floats, a multiple of 8 (rows start on 32-byte sector boundaries); n4 = ceil(dim/4)
    // float4 slots carry data, the pad up to the pitch is never read or written
    float lr, min_lr;
    uint64_t seed;
    unsigned long long *pairs;
    int64_t n_groups;
    int32_t ep_lo, ep_hi;     // epochs [ep_lo, ep_hi) and sentences [s_lo, s_hi) of this launch (multi-GPU rounds
    int64_t s_lo, s_hi;       // launch one slice at a time; a single-GPU run is one launch over everything)
    int64_t s_off, n_global;  // data-parallel shard: global index of local sentence 0 and the global sentence count.  RNG keys
                              // and the learning-rate schedule use GLOBAL sentence indices, so the shards of all ranks
                              // enumerate exactly the pairs and negatives of a single-GPU run over the whole corpus
    const uint32_t *neg_bits;  // kernel F: the negative table as increment bitmap [nwords] + per-word prefix [nwords], or NULL
    uint64_t lcg_a[SGNS_MAX_NEG], lcg_c[SGNS_MAX_NEG]; // (k+1)-step jump of the negative-sampling LCG
    int32_t dbg;
    int32_t stages;            // kernel J: stages of the row ring in shared memory (2 .. 4)
    int32_t hot;               // kernel F: words with index < hot (the most frequent) are write-through
    unsigned long long *next;  // kernel F: the next (epoch, sentence) of this launch to hand out, or NULL for the strided assignment
};

__host__ __device__ static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
// Same draw definitions as oracle/sgns_oracle.c: pure functions of (seed, epoch, sentence, position[, context]).
__host__ __device__ static inline uint64_t sgns_sentence_rng(uint64_t seed, int32_t epoch, int64_t sentence) {
    return mix64(seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(sentence + 1) + 0xD1B54A32D192ED03ULL * (uint64_t)epoch) &
           0x7FFFFFFFFFFFFFFFULL;
}
__host__ __device__ static inline uint64_t sgns_position_rng(uint64_t S, int32_t i) {
    return mix64(S + 0x9E3779B97F4A7C15ULL * (uint64_t)(i + 1)) & 0x7FFFFFFFFFFFFFFFULL;
}
__host__ __device__ static inline uint64_t sgns_pair_rng(uint64_t S, int32_t i, int32_t c) {
    return mix64(S ^ (0xD6E8FEB86659FD93ULL * (uint64_t)((int64_t)i * 65536 + c + 1)));
}
__device__ __forceinline__ float sgns_alpha(const sgns_args &a, int ep, int64_t s) {
    double progress = (double)((int64_t)ep * a.n_global + a.s_off + s) / (double)((int64_t)a.epochs * a.n_global);
    float alpha = a.lr * (float)(1.0 - progress);
    return alpha < a.min_lr ? a.min_lr : alpha;
}

__global__ void k_hist(const int32_t *__restrict__ tok, int64_t total, unsigned long long *__restrict__ cnt) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int32_t t = tok[i];
        if (t >= 0) atomicAdd(&cnt[t], 1ULL);
    }
}

// ranking keys of the vocabulary: descending count, ties by ascending id = ascending order of ((2^32 - 1 - count) << 32 | id);
// ids below min_count sort last (all-ones key).  *big is set when a count does not fit 32 bits (the host path ranks then).
__global__ void k_vocab_keys(const unsigned long long *__restrict__ cnt, int32_t n_ids, unsigned long long min_count,
                             unsigned long long *__restrict__ keys, unsigned long long *n_valid, int *big) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    if (i < n_ids) {
        const unsigned long long c = cnt[i];
        const bool ok = c > 0 && c >= min_count;
        if (c > 0xFFFFFFFFULL) *big = 1;
        keys[i] = ok ? (((0xFFFFFFFFULL - (c & 0xFFFFFFFFULL)) << 32) | (uint32_t)i) : ~0ULL;
        v = ok;
    }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(n_valid, v);
}

// corpus ids -> vocabulary indices, dropping padding and out-of-vocabulary tokens (DL4J removes words below
// minWordFrequency from the sentence before windowing); thread per sentence, position-major on both sides.
__global__ void k_compact(const int32_t *__restrict__ tok, int64_t n, int32_t L, const int32_t *__restrict__ word_of_id,
                          int32_t *__restrict__ wtok, int64_t n_total, int64_t first, int32_t Lmax,
                          unsigned long long *words) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned long long c = 0;
    if (s < n) {
        int cnt = 0;
        for (int j = 0; j < L; j++) {
            int32_t id = tok[(int64_t)j * n + s];
            int32_t wd = id >= 0 ? word_of_id[id] : -1;
            if (wd >= 0) { wtok[(int64_t)cnt * n_total + first + s] = wd; cnt++; }
        }
        c = cnt;
        for (; cnt < Lmax; cnt++) wtok[(int64_t)cnt * n_total + first + s] = -1;
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(words, c);
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
__device__ __forceinline__ float dot4(const float4 &a, const float4 &b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ void axpy4(float4 &y, float g, const float4 &x) { y.x += g * x.x; y.y += g * x.y; y.z += g * x.z; y.w += g * x.w; }
__device__ __forceinline__ float4 scale4(float g, const float4 &x) { return make_float4(g * x.x, g * x.y, g * x.z, g * x.w); }
// 128-bit reduction at L2: no lost update, no return value
__device__ __forceinline__ void red_add4(float4 *p, const float4 &v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// Predicated 128-bit L2 load: lanes with pred == false issue NO request and keep zeros.  Written as one PTX
// statement so that no "else" move depends on the load (which would make ptxas wait for each load before
// issuing the next); consecutive calls stay back to back and keep K+1 rows in flight per lane.
__device__ __forceinline__ float4 ldcg4_if(const float4 *p, bool pred) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+f"(r.x), "+f"(r.y), "+f"(r.z), "+f"(r.w)
                 : "l"(p), "r"((int)pred));
    return r;
}
// gradient scale of one (input, target) dot product: libnd4j NegativeSampling aggregate with the expTable sigmoid
__device__ __forceinline__ bool sgns_g(float dot, float label, float alpha, const float *s_exp, int E, float idx_scale, float &g) {
    if (dot > SGNS_MAX_EXP) g = (label - 1.f) * alpha;
    else if (dot < -SGNS_MAX_EXP) g = (label - 0.f) * alpha;
    else {
        int idx = (int)((dot + SGNS_MAX_EXP) * idx_scale);
        if (idx < 0 || idx >= E) return false;
        g = (label - s_exp[idx]) * alpha;
    }
    return true;
}
__device__ __forceinline__ int32_t sgns_negative(uint64_t &ns, const sgns_args &a) {
    ns = ns * LCG_MUL + LCG_ADD;
    int32_t t = a.neg_table[(ns >> 16) % (uint64_t)a.neg_table_size];
    if (t <= 0 || t >= a.V) t = (int32_t)(ns % (uint64_t)(a.V - 1)) + 1;
    return t;
}

// ---------------------------------------------------------------------------------------------------------
// Kernel A: sentence per group, pairs and targets strictly in the oracle's order, plain (atomic-free) row
// stores.  With concurrency 1 it reproduces oracle/sgns_oracle.c to fp32 tolerance; with many groups it is
// the classic Hogwild schedule (use it when the vocabulary is much larger than the sentences in flight).
template <int G, int VPL>
__global__ void __launch_bounds__(128)
k_sgns_seq(const sgns_args a) {
    extern __shared__ int32_t smem[];
    float *s_exp = reinterpret_cast<float *>(smem);
    const int gpb = blockDim.x / G;
    const int gl = threadIdx.x / G;
    const int lane = threadIdx.x % G;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
    for (int i = threadIdx.x; i < a.exp_table_size; i += blockDim.x) s_exp[i] = a.exp_table[i];
    __syncthreads();
    const int64_t gid = (int64_t)blockIdx.x * gpb + gl;
    const int n4 = a.n4;
    const int E = a.exp_table_size;
    const float idx_scale = (float)E / SGNS_MAX_EXP / 2.0f;
    const int win = a.window;
    const int64_t N = a.n_sent;
    unsigned long long pairs = 0;
    for (int ep = a.ep_lo; ep < a.ep_hi; ep++) {
        for (int64_t s = a.s_lo + gid; s < a.s_hi; s += a.n_groups) {
            int n = 0;
            while (n < a.Lmax && a.wtok[(int64_t)n * N + s] >= 0) n++;
            const float alpha = sgns_alpha(a, ep, s);
            const uint64_t S = sgns_sentence_rng(a.seed, ep, s + a.s_off);
            for (int i = 0; i < n; i++) {
                const int b = (int32_t)(uint32_t)sgns_position_rng(S, i) % win;
                const int32_t w1 = a.wtok[(int64_t)i * N + s];
                const int end = win * 2 + 1 - b;
                for (int aa = b; aa < end; aa++) {
                    if (aa == win) continue;
                    const int c = i - win + aa;
                    if (c < 0 || c >= n) continue;
                    const int32_t last = a.wtok[(int64_t)c * N + s];
                    if (last == w1) continue;
                    uint64_t ns = sgns_pair_rng(S, i, c);
                    pairs++;
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
                            target = sgns_negative(ns, a);
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
                            dot += dot4(v0[v], v1[v]);
                        }
                        dot = group_sum<G>(dot, gmask);
                        float g;
                        if (!sgns_g(dot, label, alpha, s_exp, E, idx_scale, g)) continue;
#pragma unroll
                        for (int v = 0; v < VPL; v++) {
                            int q = lane + v * G;
                            axpy4(neu[v], g, v1[v]);
                            axpy4(v1[v], g, v0[v]);
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
        }
    }
    if (lane == 0 && pairs) atomicAdd(a.pairs, pairs);
}

// ---------------------------------------------------------------------------------------------------------
// Kernel B: the throughput kernel.  Work item = (sentence, centre position); a group of G lanes owns one
// item (G = 1 for rows of up to 8 float4: a thread per item, no shuffles; G = 16/32 for wide rows: one
// coalesced 128-bit slot per lane).  The item walks ALL positions c of its sentence with a uniform trip count
// and a predicate, so the lanes of a warp stay in lockstep.  The centre's output row syn1neg[w1] stays in
// registers for the whole item (read once, its delta reduced once); per pair the negative rows are fetched a
// chunk at a time before use (memory-level parallelism), and every update is a 128-bit L2 reduction
// (red.global.add.v4.f32): updates are never lost, they are only applied to slightly stale rows -- the Hogwild
// contract without its failure mode on small vocabularies (DESIGN.md "SGNS schedule").
// Items are taken in corpus order by a grid-stride loop, so n_groups bounds the sentences in flight.

// x mod m for x < 2^48, m < 2^30, exact: one double multiply + fix-up instead of a 64-bit division
__device__ __forceinline__ uint32_t mod48(uint64_t x, uint32_t m, double inv_m) {
    // q is floor(x/m) or one off (x < 2^48 is exact in a double, the product is off by < 1), so the remainder lies
    // in (-m, 2m): for m < 2^30 the low 32 bits are enough
    const uint64_t q = (uint64_t)((double)x * inv_m);
    int32_t r = (int32_t)((uint32_t)x - (uint32_t)q * m);
    if (r < 0) r += (int32_t)m;
    else if (r >= (int32_t)m) r -= (int32_t)m;
    return (uint32_t)r;
}
// full 64-bit x mod m through three 48-bit steps
__device__ __forceinline__ uint32_t mod64(uint64_t x, uint32_t m, double inv_m) {
    uint32_t r = mod48(x >> 32, m, inv_m);
    r = mod48(((uint64_t)r << 16) | ((x >> 16) & 0xFFFFu), m, inv_m);
    return mod48(((uint64_t)r << 16) | (x & 0xFFFFu), m, inv_m);
}

#define SGNS_CH 5 // negatives drawn (one per lane) and fetched ahead per chunk
template <int G, int VPL>
__global__ void __launch_bounds__(128)
k_sgns_items(const sgns_args a) {
    static_assert(G >= 8, "the item kernel draws one negative per lane: groups have at least 8 lanes");
    extern __shared__ int32_t smem[];
    float *s_exp = reinterpret_cast<float *>(smem);
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int GPW = 32 / G; // groups per warp; they run in lockstep
    const int gpw_eff = (a.dbg & 8) ? 1 : GPW; // test mode: one item at a time (n_groups = 1)
    const int lane = threadIdx.x % G;
    const int gw = (threadIdx.x & 31) / G;
    for (int i = threadIdx.x; i < a.exp_table_size; i += blockDim.x) s_exp[i] = a.exp_table[i];
    __syncthreads();
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int E = a.exp_table_size;
    const float idx_scale = (float)E / SGNS_MAX_EXP / 2.0f;
    const int win = a.window;
    const int64_t N = a.n_sent;
    const int64_t item_lo = a.s_lo * a.Lmax, n_items = a.s_hi * a.Lmax; // items [item_lo, n_items) of this launch
    const double inv_total = 1.0 / (double)((int64_t)a.epochs * a.n_global);
    const uint32_t tsize = (uint32_t)a.neg_table_size, vm1 = (uint32_t)(a.V > 1 ? a.V - 1 : 1);
    const double inv_tsize = 1.0 / (double)tsize, inv_vm1 = 1.0 / (double)vm1;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int K = a.V >= 2 ? a.negative : 0;
    // Lane slots: slot q = lane + v*G holds floats 4q..4q+3 of a row.  A lane without a slot re-reads slot 0 (same
    // sector, no extra traffic) and its dot-product term is dropped; invalid work is cancelled through g = 0.
    // Loaded values are never masked or predicated: that makes ptxas consume each load before issuing the next,
    // whereas plain back-to-back loads keep K+1 rows in flight per lane.
    const int n4 = a.n4;
    int slot[VPL];
    bool live[VPL];
#pragma unroll
    for (int v = 0; v < VPL; v++) { live[v] = lane + v * G < n4; slot[v] = live[v] ? lane + v * G : 0; }
    unsigned long long pairs = 0;
    for (int ep = a.ep_lo; ep < a.ep_hi; ep++) {
        for (int64_t base = item_lo + warp_id * gpw_eff; base < n_items; base += a.n_groups) { // warp-uniform trip count
            const int64_t item = base + gw;
            bool valid = item < n_items && gw < gpw_eff;
            const int64_t s = valid ? item / a.Lmax : 0;
            const int i = valid ? (int)(item - s * a.Lmax) : 0;
            const int32_t w1 = a.wtok[(int64_t)i * N + s]; // (s, i) = (0, 0) when the item is out of range: in bounds
            valid = valid && w1 >= 0;
            if (!__any_sync(FULL, valid)) continue;
            float alpha = a.lr * (float)(1.0 - (double)((int64_t)ep * a.n_global + a.s_off + s) * inv_total);
            if (alpha < a.min_lr) alpha = a.min_lr;
            const uint64_t S = sgns_sentence_rng(a.seed, ep, s + a.s_off);
            const int b = (int32_t)(uint32_t)sgns_position_rng(S, i) % win;
            const int lo = i - win + b, hi = i + win - b; // inclusive context range (SkipGram.skipGram)
            float4 *pw = reinterpret_cast<float4 *>(a.syn1neg + (int64_t)(valid ? w1 : 0) * a.stride);
            float4 cur[VPL], d1[VPL]; // current value and accumulated delta of syn1neg[w1]
#pragma unroll
            for (int v = 0; v < VPL; v++) { cur[v] = __ldcg(pw + slot[v]); d1[v] = zero4; }
            for (int c = 0; c < a.Lmax; c++) {
                const int32_t last = a.wtok[(int64_t)c * N + s];
                const bool act = valid && c >= lo && c <= hi && c != i && last >= 0 && last != w1;
                if (!__any_sync(FULL, act)) continue;
                const uint64_t ns0 = sgns_pair_rng(S, i, c);
                pairs += act;
                float4 *p0 = reinterpret_cast<float4 *>(a.syn0 + (int64_t)(act ? last : 0) * a.stride);
                float4 v0[VPL], neu[VPL];
#pragma unroll
                for (int v = 0; v < VPL; v++) { v0[v] = __ldcg(p0 + slot[v]); neu[v] = zero4; }
                // negatives of the first chunk: lane k draws negative k (the LCG is affine: state k+1 = A_k*ns0 + C_k)
                int32_t mine = -1;
                if (lane < SGNS_CH && lane < K && act) {
                    const uint64_t nsk = a.lcg_a[lane] * ns0 + a.lcg_c[lane];
                    int32_t t = a.neg_table[mod48(nsk >> 16, tsize, inv_tsize)];
                    if (t <= 0 || t >= a.V) t = (int32_t)mod64(nsk, vm1, inv_vm1) + 1;
                    if (t != w1) mine = t;
                }
                { // positive target: the item's private, always-current copy of syn1neg[w1]
                    float dot = 0.f;
#pragma unroll
                    for (int v = 0; v < VPL; v++) dot += live[v] ? dot4(v0[v], cur[v]) : 0.f;
                    dot = group_sum<G>(dot, FULL);
                    float g = 0.f;
                    if (!(sgns_g(dot, 1.f, alpha, s_exp, E, idx_scale, g) && act)) g = 0.f;
                    {
#pragma unroll
                        for (int v = 0; v < VPL; v++) {
                            axpy4(neu[v], g, cur[v]);
                            axpy4(d1[v], g, v0[v]);
                            axpy4(cur[v], g, v0[v]);
                        }
                    }
                }
                for (int k0 = 0; k0 < K; k0 += SGNS_CH) {
                    if (k0 > 0) { // further chunks (negative > 5)
                        mine = -1;
                        if (lane < SGNS_CH && k0 + lane < K && act) {
                            const uint64_t nsk = a.lcg_a[k0 + lane] * ns0 + a.lcg_c[k0 + lane];
                            int32_t t = a.neg_table[mod48(nsk >> 16, tsize, inv_tsize)];
                            if (t <= 0 || t >= a.V) t = (int32_t)mod64(nsk, vm1, inv_vm1) + 1;
                            if (t != w1) mine = t;
                        }
                    }
                    int32_t tg[SGNS_CH];
                    float4 vk[SGNS_CH][VPL];
#pragma unroll
                    for (int k = 0; k < SGNS_CH; k++) tg[k] = __shfl_sync(FULL, mine, k, G);
#pragma unroll
                    for (int k = 0; k < SGNS_CH; k++) {
                        const float4 *pk = reinterpret_cast<const float4 *>(a.syn1neg + (int64_t)(tg[k] < 0 ? 0 : tg[k]) * a.stride);
#pragma unroll
                        for (int v = 0; v < VPL; v++) vk[k][v] = __ldcg(pk + slot[v]);
                    }
#pragma unroll
                    for (int k = 0; k < SGNS_CH; k++) {
                        float dot = 0.f;
#pragma unroll
                        for (int v = 0; v < VPL; v++) dot += live[v] ? dot4(v0[v], vk[k][v]) : 0.f;
                        dot = group_sum<G>(dot, FULL);
                        float g = 0.f;
                        const bool upd = sgns_g(dot, 0.f, alpha, s_exp, E, idx_scale, g) && tg[k] >= 0;
                        if (!upd) g = 0.f;
#pragma unroll
                        for (int v = 0; v < VPL; v++) axpy4(neu[v], g, vk[k][v]);
                        if (upd) {
                            float4 *pk = reinterpret_cast<float4 *>(a.syn1neg + (int64_t)tg[k] * a.stride);
#pragma unroll
                            for (int v = 0; v < VPL; v++)
                                if (live[v]) red_add4(pk + slot[v], scale4(g, v0[v]));
                        }
                    }
                }
                if (act) {
#pragma unroll
                    for (int v = 0; v < VPL; v++)
                        if (live[v]) red_add4(p0 + slot[v], neu[v]);
                }
            }
            if (valid) {
#pragma unroll
                for (int v = 0; v < VPL; v++)
                    if (live[v]) red_add4(pw + slot[v], d1[v]);
            }
        }
    }
    if (lane == 0 && pairs) atomicAdd(a.pairs, pairs);
}

// ---------------------------------------------------------------------------------------------------------
// Kernel C: the item kernel for rows of up to 32 float4 slots (one slot per lane), rebuilt around its measured
// limit.  ncu on the tract x 24 workload (profiles/r1s3_sgns_tract24.json) showed its predecessor issue-bound (58 %
// of the issue slots busy, no memory stall): ~600 warp instructions per 4 pairs, most of them integer / control
// overhead.  Same work decomposition as kernel B, same draws, same arithmetic per pair; what changed:
//   * the K+1 dot products of a pair are reduced with ONE transposed butterfly (7 shuffles for up to 8 values over
//     8 lanes, lane L ends with the total of value L) instead of K+1 separate butterflies (3 shuffles each);
//   * lane L alone turns total L into its gradient scale g_L (one branch-free sigmoid-table lookup per lane instead
//     of K+1 per lane) and the six g are broadcast back;
//   * the per-pair hash of the negative stream is computed for G context positions at once (lane l: position
//     c0 + l) and broadcast per pair, instead of G times redundantly per pair;
//   * row addresses are 32-bit slot offsets from a per-lane base pointer (one IMAD.WIDE each);
//   * only the negative-table lookups run one unit ahead; the rows of a unit are requested and consumed in the
//     same unit, which fits 96 registers => 5 blocks per SM, and the extra resident warps hide the L2 latency
//     better than a second row buffer did (profiles/r1s6_sgns_builds.txt);
//   * negatives > 5 are further 5-wide chunks (units) of the same pair (MULTI) instead of a serial tail;
//   * a reduction whose g is exactly 0 (saturated sigmoid) is not sent.
// Rows sit on a sector-aligned pitch (args.stride, multiple of 8 floats), so a row of D floats touches
// ceil(D/8) sectors instead of one more on every other row.
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src, int width) {
    uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src, width);
    uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src, width);
    return ((uint64_t)hi << 32) | lo;
}
// address of float4 slot `base` (a per-lane pointer into row 0) in row `row`: one 32 x 32 + 64-bit multiply-add
__device__ __forceinline__ uint64_t row_addr(const char *base, uint32_t row, uint32_t pitch) {
    uint64_t p;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(p) : "r"(row), "r"(pitch), "l"(base));
    return p;
}
// predicated 128-bit L2 reduction / load as single PTX statements (no branch around them).  The load keeps the
// previous register contents where pred is false: the callers make stale (finite) values harmless through g = 0.
__device__ __forceinline__ void red_add4_if(uint64_t p, const float4 &v, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void ldcg4_into(float4 &r, uint64_t p, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+f"(r.x), "+f"(r.y), "+f"(r.z), "+f"(r.w)
                 : "l"(p), "r"((int)pred));
}
// predicated 16-byte cp.async (LDGSTS, L2 only) and its group bookkeeping
__device__ __forceinline__ void cp_async16_if(uint32_t smem_addr, uint64_t gptr, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                 ::"r"(smem_addr), "l"(gptr), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// After this rebuild the kernel runs at ~2/3 of what the memory system itself delivers for its access pattern
// (random 80-byte-row 128-bit loads + reductions, scripts/red_microbench.cu, profiles/r1s7_red_microbench.txt):
// the reductions, not the instruction stream, are the limit now.
// PLAIN = true is the atomic-free build north_star's wording asks for ("Hogwild-style atomic-free row updates"): every
// row update is a plain 128-bit store of (row as loaded + its update) instead of an L2 reduction, so an update that
// lands between a group's load and its store is LOST (classic Hogwild).  Selected only by DGE_SGNS_F_PLAIN_STORES
// (A/B: throughput and downstream metric against the reduction build); never the default.
__device__ __forceinline__ void stcg4_if(uint64_t p, const float4 &v, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p st.global.cg.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)pred) : "memory");
}
__device__ __forceinline__ float4 add4(const float4 &a, const float4 &b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// MODE 2 (DGE_SGNS_F_SMEM_NEG_TABLE; vocabularies below 65 536 words): ONE block of 640 threads per SM instead of five of
// 128, and the unigram^0.75 negative table lives in its shared memory as 16-bit entries (100 000 x 2 bytes), so the five
// table lookups of a pair are LDS instead of five scattered 4-byte global loads through the same LSU path the row loads
// and reductions need.
template <int G, bool MULTI, int MODE>
__global__ void __launch_bounds__(MODE == 2 ? 640 : 128, MODE == 2 ? 1 : 5)
k_sgns_items_v2(const sgns_args a) {
    static_assert(G == 8 || G == 16 || G == 32, "lane groups of 8, 16 or 32");
    constexpr bool PLAIN = MODE == 1;
    constexpr bool SMEM_NEG = MODE == 2;
    extern __shared__ int32_t smem[];
    float *s_exp = reinterpret_cast<float *>(smem);
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int GPW = 32 / G;
    constexpr bool MERGE_SYN0 = GPW > 1; // sum the syn0[last] updates of the warp's groups before reducing them (+6 % at G = 8)
    const int gpw_eff = (a.dbg & 8) ? 1 : GPW; // test mode: one item at a time (n_groups = 1)
    const int lane = threadIdx.x % G;
    const int gw = (threadIdx.x & 31) / G;
    int32_t *mytok = smem + a.exp_table_size + (threadIdx.x / G) * a.Lmax;
    uint16_t *s_neg = reinterpret_cast<uint16_t *>(smem + a.exp_table_size + (blockDim.x / G) * a.Lmax);
    for (int i = threadIdx.x; i < a.exp_table_size; i += blockDim.x) s_exp[i] = a.exp_table[i];
    if (SMEM_NEG)
        for (int i = threadIdx.x; i < a.neg_table_size; i += blockDim.x) s_neg[i] = (uint16_t)a.neg_table[i]; // V <= 65535 (host)
    __syncthreads();
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int E = a.exp_table_size;
    const float idx_scale = (float)E / SGNS_MAX_EXP / 2.0f;
    const int win = a.window;
    const int64_t N = a.n_sent;
    const int Lmax = a.Lmax;
    const int64_t item_lo = a.s_lo * Lmax, n_items = a.s_hi * Lmax; // items [item_lo, n_items) of this launch
    const double inv_total = 1.0 / (double)((int64_t)a.epochs * a.n_global);
    const uint32_t tsize = (uint32_t)a.neg_table_size, vm1 = (uint32_t)(a.V > 1 ? a.V - 1 : 1);
    const double inv_tsize = 1.0 / (double)tsize, inv_vm1 = 1.0 / (double)vm1;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int K = a.V >= 2 ? a.negative : 0;
    const int NCH = MULTI ? max(1, (K + SGNS_CH - 1) / SGNS_CH) : 1; // 5-wide chunks of negatives per pair
    const bool live = lane < a.n4;
    const uint32_t pitch = (uint32_t)a.stride * 4u;           // bytes; V * pitch < 2^32 * 16 is checked by the host
    const char *base0 = reinterpret_cast<const char *>(a.syn0) + (live ? lane : 0) * 16;
    const char *base1 = reinterpret_cast<const char *>(a.syn1neg) + (live ? lane : 0) * 16;
    // value index owned by this lane after the transposed reduction: negatives 0..4 of the chunk, 5 = positive
    const int L8 = lane & 7;
    const bool up4 = (lane & 4) != 0, up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
    const float my_label = L8 == SGNS_CH ? 1.f : 0.f;
    unsigned long long pairs = 0;

    struct stage_t { int32_t last; bool act; uint64_t nsk; int32_t traw; int j; };
    struct stage_r { int32_t last; bool act; int j; int32_t mine; int32_t tg[SGNS_CH]; float4 row[SGNS_CH]; float4 v0; };

    for (int ep = a.ep_lo; ep < a.ep_hi; ep++) {
        for (int64_t base = item_lo + warp_id * gpw_eff; base < n_items; base += a.n_groups) { // warp-uniform trip count
            const int64_t item = base + gw;
            bool valid = item < n_items && gw < gpw_eff;
            const int64_t s = valid ? item / Lmax : 0;
            const int i = valid ? (int)(item - s * Lmax) : 0;
            // all groups of the warp on one sentence (the rule; not at the tail or in the one-item test mode): they
            // share every context row syn0[last], whose K+1-target updates are then summed in the warp and reduced once
            const long long s_first = __shfl_sync(FULL, (long long)s, 0); // every lane takes part (no short-circuit)
            const bool same_s = MERGE_SYN0 && __all_sync(FULL, valid && (long long)s == s_first);
            __syncwarp();
            int n_tok = 0; // tokens of the (compacted) sentence
            for (int j = lane; j < Lmax; j += G) { const int32_t tk = a.wtok[(int64_t)j * N + s]; mytok[j] = tk; n_tok += tk >= 0; }
            __syncwarp();
#pragma unroll
            for (int o = G >> 1; o > 0; o >>= 1) n_tok += __shfl_xor_sync(FULL, n_tok, o);
            const int32_t w1 = mytok[i];
            valid = valid && w1 >= 0;
            if (!__any_sync(FULL, valid)) continue;
            float alpha = a.lr * (float)(1.0 - (double)((int64_t)ep * a.n_global + a.s_off + s) * inv_total);
            if (alpha < a.min_lr) alpha = a.min_lr;
            const float g_hi = (my_label - 1.f) * alpha, g_lo = my_label * alpha; // saturated sigmoid: dot > 6, dot < -6
            const uint64_t S = sgns_sentence_rng(a.seed, ep, s + a.s_off);
            const int b = (int32_t)(uint32_t)sgns_position_rng(S, i) % win;
            // inclusive context range (SkipGram.skipGram); an invalid item gets the empty range
            const int lo = valid ? i - win + b : 1, hi = valid ? i + win - b : 0;
            // context positions any group of the warp can pair with: units outside [c_min, c_max] are skipped
            const int c_min = __reduce_min_sync(FULL, valid ? max(lo, 0) : Lmax);
            const int c_max = __reduce_max_sync(FULL, valid ? min(hi, n_tok - 1) : -1);
            if (c_max < c_min) continue;
            float4 cur = zero4, d1 = zero4, neu = zero4, v0p = zero4;
            ldcg4_into(cur, row_addr(base1, (uint32_t)w1, pitch), valid && live); // private copy of syn1neg[w1]
            int npairs = 0;
            int cT = c_min, jT = 0; // (context position, chunk) of the next unit the T stage hands out
            uint64_t hc = 0;        // pair hash of context position hcb * G + lane
            int hcb = -1;

            auto stageT = [&]() { // which (pair, chunk) comes next; request its negatives' table entries
                stage_t t;
                t.j = jT;
                t.last = cT < Lmax ? mytok[cT] : -1;
                t.act = cT >= lo && cT <= hi && cT != i && t.last >= 0 && t.last != w1;
                if (cT / G != hcb) { hcb = cT / G; hc = sgns_pair_rng(S, i, hcb * G + lane); } // warp-uniform condition
                const uint64_t ns0 = shfl64(hc, cT & (G - 1), G);
                const int kk = jT * SGNS_CH + lane;      // this lane's negative of the pair (lanes 0..4 draw)
                const bool drawer = lane < SGNS_CH && kk < K;
                const int kc = drawer ? kk : 0;
                t.nsk = a.lcg_a[kc] * ns0 + a.lcg_c[kc]; // the LCG is affine: state after kk+1 steps
                t.traw = -2;                             // "draws nothing"
                if (drawer && t.act) t.traw = SMEM_NEG ? (int32_t)s_neg[mod48(t.nsk >> 16, tsize, inv_tsize)] : a.neg_table[mod48(t.nsk >> 16, tsize, inv_tsize)];
                if (MULTI) { if (++jT == NCH) { jT = 0; cT++; } }
                else cT++;
                return t;
            };
            auto stageR = [&](const stage_t &t, stage_r &r) { // resolve the negatives, request all rows of the unit
                r.last = t.last; r.act = t.act; r.j = t.j;
                int32_t tt = t.traw;
                const bool redraw = tt != -2 && (tt <= 0 || tt >= a.V); // DL4J: target = r % (V-1) + 1
                if (__any_sync(FULL, redraw)) {
                    if (redraw) tt = (int32_t)mod64(t.nsk, vm1, inv_vm1) + 1;
                }
                r.mine = (tt != -2 && tt != w1) ? tt : -1;
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) r.tg[k] = __shfl_sync(FULL, r.mine, k, G);
                if (!MULTI || t.j == 0) ldcg4_into(r.v0, row_addr(base0, (uint32_t)t.last, pitch), t.act && live);
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) ldcg4_into(r.row[k], row_addr(base1, (uint32_t)r.tg[k], pitch), r.tg[k] >= 0 && live);
            };
            auto compute = [&](const stage_r &r) {
                if (!__any_sync(FULL, r.act)) return;
                const bool first = !MULTI || r.j == 0;
                if (first) { npairs += r.act; neu = zero4; }
                if (MULTI && first) v0p = r.v0;
                const float4 v0 = MULTI ? v0p : r.v0;
                // ---- K+1 dot products, transposed reduction: lane L8 ends with the group total of value L8
                float d0 = dot4(v0, r.row[0]), d1v = dot4(v0, r.row[1]), d2 = dot4(v0, r.row[2]), d3 = dot4(v0, r.row[3]);
                float d4 = dot4(v0, r.row[4]), d5 = first ? dot4(v0, cur) : 0.f;
                // offset 4: lanes with bit 2 clear keep values 0..3, the others keep 4..7 (6, 7 are empty)
                float e0 = (up4 ? d4 : d0) + __shfl_xor_sync(FULL, up4 ? d0 : d4, 4);
                float e1 = (up4 ? d5 : d1v) + __shfl_xor_sync(FULL, up4 ? d1v : d5, 4);
                float e2 = (up4 ? 0.f : d2) + __shfl_xor_sync(FULL, up4 ? d2 : 0.f, 4);
                float e3 = (up4 ? 0.f : d3) + __shfl_xor_sync(FULL, up4 ? d3 : 0.f, 4);
                // offset 2: bit 1 clear keeps the lower two of its four
                float f0 = (up2 ? e2 : e0) + __shfl_xor_sync(FULL, up2 ? e0 : e2, 2);
                float f1 = (up2 ? e3 : e1) + __shfl_xor_sync(FULL, up2 ? e1 : e3, 2);
                // offset 1
                float tot = (up1 ? f1 : f0) + __shfl_xor_sync(FULL, up1 ? f0 : f1, 1);
                if (G >= 16) tot += __shfl_xor_sync(FULL, tot, 8);
                if (G >= 32) tot += __shfl_xor_sync(FULL, tot, 16);
                // ---- lane L8 owns target L8: its gradient scale (libnd4j NegativeSampling, expTable sigmoid)
                float g;
                {
                    const float f = (tot + SGNS_MAX_EXP) * idx_scale;
                    const int idx = (int)f;
                    const float sg = s_exp[min(max(idx, 0), E - 1)];
                    g = (my_label - sg) * alpha;
                    if (idx < 0 || idx >= E) g = 0.f;          // table index out of range: the aggregate skips the target
                    if (tot > SGNS_MAX_EXP) g = g_hi;
                    else if (tot < -SGNS_MAX_EXP) g = g_lo;
                    const bool mine_ok = L8 < SGNS_CH ? r.mine >= 0 : (L8 == SGNS_CH && r.act && first);
                    if (!mine_ok) g = 0.f;                      // lanes >= 8 of a wide group are never read
                }
                float gk[SGNS_CH + 1];
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) gk[k] = __shfl_sync(FULL, g, k, G);
                gk[SGNS_CH] = first ? __shfl_sync(FULL, g, SGNS_CH, G) : 0.f;
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) {
                    axpy4(neu, gk[k], r.row[k]);
                    if (PLAIN) { float4 nr = r.row[k]; axpy4(nr, gk[k], v0); stcg4_if(row_addr(base1, (uint32_t)r.tg[k], pitch), nr, gk[k] != 0.f && live && !(a.dbg & 1)); }
                    else red_add4_if(row_addr(base1, (uint32_t)r.tg[k], pitch), scale4(gk[k], v0), gk[k] != 0.f && live && !(a.dbg & 1));
                }
                if (first) { // positive target: the item's private, always-current copy of syn1neg[w1]
                    axpy4(neu, gk[SGNS_CH], cur);
                    axpy4(d1, gk[SGNS_CH], v0);
                    axpy4(cur, gk[SGNS_CH], v0);
                }
                if (!MULTI || r.j == NCH - 1) { // the pair is complete: syn0[last] += neu
                    if (same_s) { // one row for the whole warp (inactive groups carry neu = 0)
                        float4 ns = neu;
#pragma unroll
                        for (int o = G; o < 32; o <<= 1) {
                            ns.x += __shfl_xor_sync(FULL, ns.x, o); ns.y += __shfl_xor_sync(FULL, ns.y, o);
                            ns.z += __shfl_xor_sync(FULL, ns.z, o); ns.w += __shfl_xor_sync(FULL, ns.w, o);
                        }
                        if (PLAIN) { // the first active group holds a valid copy of the row and stores row + sum
                            const unsigned am = __ballot_sync(FULL, r.act);
                            const int first_gw = am ? (__ffs(am) - 1) / G : -1;
                            stcg4_if(row_addr(base0, (uint32_t)r.last, pitch), add4(v0, ns), gw == first_gw && live && !(a.dbg & 1));
                        } else
                        red_add4_if(row_addr(base0, (uint32_t)r.last, pitch), ns, gw == 0 && live && !(a.dbg & 1));
                    } else {
                        if (PLAIN) stcg4_if(row_addr(base0, (uint32_t)r.last, pitch), add4(v0, neu), r.act && live && !(a.dbg & 1));
                        else red_add4_if(row_addr(base0, (uint32_t)r.last, pitch), neu, r.act && live && !(a.dbg & 1));
                    }
                }
            };

            const int U = (c_max - c_min + 1) * NCH;
            stage_r rA; // rows not (re)loaded keep stale finite values, cancelled by g = 0; start from zeros
            rA.v0 = zero4;
#pragma unroll
            for (int k = 0; k < SGNS_CH; k++) rA.row[k] = zero4;
            stage_t t1 = stageT();
            for (int u = 0; u < U; u++) {
                stageR(t1, rA);   // rows of unit u
                t1 = stageT();    // table lookups of unit u+1 (independent work while the rows arrive)
                compute(rA);
            }
            if (PLAIN) { // re-read the row and store row + the item's accumulated delta (a short load-to-store window)
                float4 now = zero4;
                ldcg4_into(now, row_addr(base1, (uint32_t)w1, pitch), valid && live);
                stcg4_if(row_addr(base1, (uint32_t)w1, pitch), add4(now, d1), valid && live && !(a.dbg & 1));
            } else
            red_add4_if(row_addr(base1, (uint32_t)w1, pitch), d1, valid && live && !(a.dbg & 1));
            pairs += (unsigned)npairs;
        }
    }
    if (lane == 0 && pairs) atomicAdd(a.pairs, pairs);
}

// ---------------------------------------------------------------------------------------------------------
// Kernel F: the SENTENCE-RESIDENT item kernel -- what the reference's semantics need on a GPU.
//
// Kernels B-E hand the <= 24 centre positions of ONE sentence to different lane groups that run at the same time, so
// the ~23 updates a sentence makes to each of its context rows syn0[last] (one per centre) are all computed from
// (nearly) the same stale value and summed: the diminishing steps of word2vec's sequential loop -- the second centre sees
// the row the first one already moved -- are lost, and the embedding drifts systematically (at the full bench size: row
// norms 2.5 instead of the oracle's 2.3, and only 0.66 of the oracle's 10 nearest neighbours recovered even with just 8
// sentences in flight, while oracle runs with different seeds agree to 0.88: profiles/r2s4_fullsize_staleness_v2.json).
//
// Here a WARP owns a sentence for all its centres.  The warp's lane groups take the centres in batches (4 at G = 8),
// walking the context positions STAGGERED (group g works on context position c - g), so that the pairs in flight in a
// warp never share a row: they are a legitimate sequential order of the sentence's pairs.  What the warp has added to
// the sentence's context rows lives in a per-warp shared-memory DELTA cache: a pair reads syn0[last] fresh from L2
// (other sentences' updates) plus the warp's own pending delta, adds its neu1e to the cache, and the cache is flushed
// to L2 with 128-bit reductions after every batch of centres (+ a fence, so the next batch reads them back).  The
// centre's output row syn1neg[w1] stays private in registers for the item, as before.  L2 traffic per pair is what
// kernel C had (one context-row load, K negative-row loads, K reductions, 1/23 flush); the negative table is read from
// shared memory (exact bitmap + prefix form of the unigram^0.75 table, 25 KB for 100 000 slots instead of 400 KB in L2).
__device__ __forceinline__ float sgns_g_lane(float tot, float label, float alpha, float g_hi, float g_lo, const float *s_exp,
                                             int E, float idx_scale);
__device__ __forceinline__ int32_t neg_lookup(const uint32_t *__restrict__ s_bits, const uint32_t *__restrict__ s_pref, uint32_t idx) {
    // table[idx] = table[32 w] + number of increments in slots 32 w + 1 .. idx (the table never grows by more than one per slot)
    const uint32_t w = idx >> 5, j = idx & 31u;
    return (int32_t)(s_pref[w] + __popc(s_bits[w] & ((2u << j) - 2u)));
}

// PF = true (narrow rows, K <= 5, at most 12 warps per block): the rows of unit u + 1 are requested before unit u is computed
// (two row buffers in registers).  A write-through row that this warp updated in unit u is then missing that update in the
// copy requested before it: a context row's last update stays in the warp's cache, tagged with its unit, and is added by
// the reader of the next unit only; a centre adds its own last update from a register.
// PF = 2: the same, with the requested rows landing in shared memory (cp.async, two stages of 7 rows per lane) instead of
// registers, so the block keeps its 20 warps; a lane reads back exactly the slots it copied itself.
template <int G, bool MULTI, int PF>
__global__ void __launch_bounds__(PF == 1 ? 384 : 640, 1)
k_sgns_sent(const sgns_args a) {
    static_assert(!(PF && MULTI), "the prefetching build handles one chunk of negatives per pair");
    static_assert(G == 8 || G == 16 || G == 32, "lane groups of 8, 16 or 32");
    extern __shared__ __align__(16) int32_t smem_f[];
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int GPW = 32 / G;
    const int warps_per_block = blockDim.x >> 5, wib = threadIdx.x >> 5;
    const int n4 = a.n4, Lmax = a.Lmax;
    const int nwords = (a.neg_table_size + 31) >> 5;
    // shared memory: [delta cache of every warp: Lmax x n4 float4][sigmoid table][tokens of every warp][negative table bits | prefixes]
    float4 *my_delta = reinterpret_cast<float4 *>(smem_f) + (size_t)wib * Lmax * n4;
    float4 *stage_all = reinterpret_cast<float4 *>(smem_f) + (size_t)warps_per_block * Lmax * n4; // PF == 2: [warp][2 stages][7 rows][32 lanes]
    constexpr int SROWS = SGNS_CH + 2;
    float4 *my_stage = stage_all + (size_t)wib * 2 * SROWS * 32 + (threadIdx.x & 31);
    float *s_exp = reinterpret_cast<float *>(stage_all + (PF == 2 ? (size_t)warps_per_block * 2 * SROWS * 32 : 0));
    int32_t *mytok = reinterpret_cast<int32_t *>(s_exp + a.exp_table_size) + wib * Lmax;
    int32_t *my_tag = reinterpret_cast<int32_t *>(s_exp + a.exp_table_size) + (warps_per_block + wib) * Lmax; // PF: unit of a write-through row's cached update
    uint32_t *s_bits = reinterpret_cast<uint32_t *>(s_exp + a.exp_table_size) + 2 * warps_per_block * Lmax;
    uint32_t *s_pref = s_bits + nwords;
    const bool smem_neg = a.neg_bits != nullptr;
    for (int i = threadIdx.x; i < a.exp_table_size; i += blockDim.x) s_exp[i] = a.exp_table[i];
    if (PF == 2) // the stages only ever hold table rows afterwards (a slot that is not copied keeps an older row: finite)
        for (int i = threadIdx.x; i < warps_per_block * 2 * SROWS * 32; i += blockDim.x) stage_all[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (smem_neg)
        for (int i = threadIdx.x; i < 2 * nwords; i += blockDim.x) s_bits[i] = a.neg_bits[i];
    __syncthreads();
    const int gpw_eff = (a.dbg & 8) ? 1 : GPW; // test mode: one group, i.e. the oracle's exact pair order
    const int lane = threadIdx.x % G, wl = threadIdx.x & 31;
    const int gw = wl / G;
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int E = a.exp_table_size;
    const float idx_scale = (float)E / SGNS_MAX_EXP / 2.0f;
    const int win = a.window;
    const int64_t N = a.n_sent;
    const double inv_total = 1.0 / (double)((int64_t)a.epochs * a.n_global);
    const uint32_t tsize = (uint32_t)a.neg_table_size, vm1 = (uint32_t)(a.V > 1 ? a.V - 1 : 1);
    const double inv_tsize = 1.0 / (double)tsize, inv_vm1 = 1.0 / (double)vm1;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int K = a.V >= 2 ? a.negative : 0;
    const int NCH = MULTI ? max(1, (K + SGNS_CH - 1) / SGNS_CH) : 1;
    const bool live = lane < n4;
    const uint32_t pitch = (uint32_t)a.stride * 4u;
    const char *base0 = reinterpret_cast<const char *>(a.syn0) + (live ? lane : 0) * 16;
    const char *base1 = reinterpret_cast<const char *>(a.syn1neg) + (live ? lane : 0) * 16;
    const int L8 = lane & 7;
    const bool up4 = (lane & 4) != 0, up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
    const float my_label = L8 == SGNS_CH ? 1.f : 0.f;
    const bool reds_on = !(a.dbg & 1);
    unsigned long long pairs = 0;

    struct stage_t { int32_t last; bool act; uint64_t nsk; int32_t traw; int j; int c; };
    struct stage_r { int32_t last; bool act; int j; int c; int32_t mine; int32_t tg[SGNS_CH]; float4 row[SGNS_CH]; float4 v0; float4 cur; };

    // Sentences are handed out in corpus order, either strided (warp w takes w, w + n_groups, ...) or -- a.next != NULL -- from
    // a counter: then the warps sweep the corpus front together whatever their speeds (a strided warp that runs slower, e.g.
    // on an SM sub-partition with one warp more, falls behind in the corpus and in the learning-rate schedule, and the
    // corpus' last part -- the spatial walks -- is no longer trained last: the full-size agreement with the oracle drops
    // from 0.88 to 0.82 with 10 or 13 warps per SM, profiles/r2s19 / r2s22).
    const int64_t ns_launch = a.s_hi - a.s_lo;
    const unsigned long long total_launch = (unsigned long long)(a.ep_hi - a.ep_lo) * (unsigned long long)ns_launch;
    unsigned long long it = (unsigned long long)warp_id;
    if (warp_id >= a.n_groups) it = total_launch; // (a block's spare warps)
    for (;; it += (unsigned long long)a.n_groups) {
        {
            if (a.next) {
                unsigned long long nx = 0;
                if (wl == 0) nx = atomicAdd(a.next, 1ULL);
                it = shfl64(nx, 0, 32);
            }
            if (it >= total_launch) break;
            const int ep = a.ep_lo + (int)(it / (unsigned long long)ns_launch);
            const int64_t s = a.s_lo + (int64_t)(it % (unsigned long long)ns_launch);
            __syncwarp();
            int n_tok = 0;
            for (int j = wl; j < Lmax; j += 32) { const int32_t tk = a.wtok[(int64_t)j * N + s]; mytok[j] = tk; n_tok += tk >= 0; }
            for (int q = wl; q < Lmax * n4; q += 32) my_delta[q] = zero4;
            if (PF) for (int j = wl; j < Lmax; j += 32) my_tag[j] = -2;
            __syncwarp();
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) n_tok += __shfl_xor_sync(FULL, n_tok, o);
            if (n_tok < 2) continue;
            float alpha = a.lr * (float)(1.0 - (double)((int64_t)ep * a.n_global + a.s_off + s) * inv_total);
            if (alpha < a.min_lr) alpha = a.min_lr;
            const float g_hi = (my_label - 1.f) * alpha, g_lo = my_label * alpha;
            const uint64_t S = sgns_sentence_rng(a.seed, ep, s + a.s_off);
            int npairs = 0;
            for (int i0 = 0; i0 < n_tok; i0 += gpw_eff) { // a batch of centres: one per lane group
                const int i = i0 + gw;
                const bool valid = gw < gpw_eff && i < n_tok;
                const int32_t w1 = valid ? mytok[i] : 0;
                const int b = (int32_t)(uint32_t)sgns_position_rng(S, valid ? i : 0) % win;
                const int lo = valid ? i - win + b : 1, hi = valid ? i + win - b : 0; // inclusive context range; empty if invalid
                const int c_min = __reduce_min_sync(FULL, valid ? max(lo, 0) : Lmax);
                const int c_max = __reduce_max_sync(FULL, valid ? min(hi, n_tok - 1) : -1);
                if (c_max < c_min) continue;
                float4 cur = zero4, d1 = zero4, neu = zero4, v0p = zero4;
                ldcg4_into(cur, row_addr(base1, (uint32_t)w1, pitch), valid && live); // private copy of syn1neg[w1]
                // write-through words (index < a.hot: the most frequent ones): their rows are re-read for every pair and
                // their updates sent at once instead of staying pending for the batch (see the schedule in dge_sgns_train)
                const bool hot_w1 = w1 < a.hot;
                float4 upd_last = zero4; // PF: what this centre sent to its write-through output row in the previous unit
                // unit u of the batch: group g works on context position c_min + u - g (staggered: no two groups on one row)
                int uT = 0, jT = 0;
                uint64_t hc = 0;
                int hcb = -1;

                auto stageT = [&]() {
                    stage_t t;
                    t.j = jT;
                    t.c = c_min + uT - gw;
                    const bool in_row = t.c >= 0 && t.c < Lmax;
                    t.last = in_row ? mytok[t.c] : -1;
                    t.act = valid && in_row && t.c >= lo && t.c <= hi && t.c != i && t.last >= 0 && t.last != w1;
                    const int cc = in_row ? t.c : 0;
                    if (cc / G != hcb) { hcb = cc / G; hc = sgns_pair_rng(S, i, hcb * G + lane); } // per group
                    const uint64_t ns0 = shfl64(hc, cc & (G - 1), G);
                    const int kk = jT * SGNS_CH + lane;
                    const bool drawer = lane < SGNS_CH && kk < K;
                    const int kc = drawer ? kk : 0;
                    t.nsk = a.lcg_a[kc] * ns0 + a.lcg_c[kc];
                    t.traw = -2;
                    if (drawer && t.act) {
                        const uint32_t idx = mod48(t.nsk >> 16, tsize, inv_tsize);
                        t.traw = smem_neg ? neg_lookup(s_bits, s_pref, idx) : a.neg_table[idx];
                    }
                    if (MULTI) { if (++jT == NCH) { jT = 0; uT++; } }
                    else uT++;
                    return t;
                };
                const uint32_t my_stage_s = (uint32_t)__cvta_generic_to_shared(my_stage);
                auto stageR = [&](const stage_t &t, stage_r &r, int sidx) {
                    r.last = t.last; r.act = t.act; r.j = t.j; r.c = t.c;
                    int32_t tt = t.traw;
                    const bool redraw = tt != -2 && (tt <= 0 || tt >= a.V);
                    if (__any_sync(FULL, redraw)) {
                        if (redraw) tt = (int32_t)mod64(t.nsk, vm1, inv_vm1) + 1;
                    }
                    r.mine = (tt != -2 && tt != w1) ? tt : -1;
#pragma unroll
                    for (int k = 0; k < SGNS_CH; k++) r.tg[k] = __shfl_sync(FULL, r.mine, k, G);
                    if (PF == 2) {
                        const uint32_t dst = my_stage_s + (uint32_t)(sidx * SROWS * 32 * 16);
                        cp_async16_if(dst, row_addr(base0, (uint32_t)t.last, pitch), t.act && live);
#pragma unroll
                        for (int k = 0; k < SGNS_CH; k++) cp_async16_if(dst + (uint32_t)((k + 1) * 32 * 16), row_addr(base1, (uint32_t)r.tg[k], pitch), r.tg[k] >= 0 && live);
                        cp_async16_if(dst + (uint32_t)((SGNS_CH + 1) * 32 * 16), row_addr(base1, (uint32_t)w1, pitch), t.act && live && hot_w1);
                        cp_async_commit();
                        return;
                    }
                    if (!MULTI || t.j == 0) ldcg4_into(r.v0, row_addr(base0, (uint32_t)t.last, pitch), t.act && live);
#pragma unroll
                    for (int k = 0; k < SGNS_CH; k++) ldcg4_into(r.row[k], row_addr(base1, (uint32_t)r.tg[k], pitch), r.tg[k] >= 0 && live);
                    // a write-through centre: its output row as L2 has it now (this lane's own earlier reductions included)
                    if (!MULTI || t.j == 0) ldcg4_into(PF ? r.cur : cur, row_addr(base1, (uint32_t)w1, pitch), t.act && live && hot_w1);
                };
                auto compute = [&](stage_r &r, int u, int sidx) {
                    const float4 upd_prev = upd_last;
                    upd_last = zero4;
                    if (PF == 2) cp_async_wait<1>(); // everything but the newest group (the next unit's rows) has landed
                    if (!__any_sync(FULL, r.act)) return;
                    if (PF == 2) { // this lane's slots of the unit's rows
                        const float4 *sp = my_stage + sidx * SROWS * 32;
                        r.v0 = sp[0];
#pragma unroll
                        for (int k = 0; k < SGNS_CH; k++) r.row[k] = sp[(k + 1) * 32];
                        r.cur = sp[(SGNS_CH + 1) * 32];
                    }
                    const bool first = !MULTI || r.j == 0;
                    if (first) {
                        npairs += r.act;
                        neu = zero4;
                        // the row as this sentence sees it: L2's value + what this warp has added since its last flush
                        v0p = r.v0;
                        if (r.act && live) {
                            // PF, write-through row: the cached update counts only if it was made in the unit just before (it
                            // is in every copy requested later)
                            if (!PF || r.last >= a.hot || my_tag[r.c] == u - 1) {
                                const float4 dl = my_delta[r.c * n4 + lane]; v0p.x += dl.x; v0p.y += dl.y; v0p.z += dl.z; v0p.w += dl.w;
                            }
                        }
                        if (PF && hot_w1 && r.act) cur = add4(r.cur, upd_prev); // requested before the previous unit's update left
                    }
                    const float4 v0 = v0p;
                    float d0 = dot4(v0, r.row[0]), d1v = dot4(v0, r.row[1]), d2 = dot4(v0, r.row[2]), d3 = dot4(v0, r.row[3]);
                    float d4 = dot4(v0, r.row[4]), d5 = first ? dot4(v0, cur) : 0.f;
                    float e0 = (up4 ? d4 : d0) + __shfl_xor_sync(FULL, up4 ? d0 : d4, 4);
                    float e1 = (up4 ? d5 : d1v) + __shfl_xor_sync(FULL, up4 ? d1v : d5, 4);
                    float e2 = (up4 ? 0.f : d2) + __shfl_xor_sync(FULL, up4 ? d2 : 0.f, 4);
                    float e3 = (up4 ? 0.f : d3) + __shfl_xor_sync(FULL, up4 ? d3 : 0.f, 4);
                    float f0 = (up2 ? e2 : e0) + __shfl_xor_sync(FULL, up2 ? e0 : e2, 2);
                    float f1 = (up2 ? e3 : e1) + __shfl_xor_sync(FULL, up2 ? e1 : e3, 2);
                    float tot = (up1 ? f1 : f0) + __shfl_xor_sync(FULL, up1 ? f0 : f1, 1);
                    if (G >= 16) tot += __shfl_xor_sync(FULL, tot, 8);
                    if (G >= 32) tot += __shfl_xor_sync(FULL, tot, 16);
                    float g = sgns_g_lane(tot, my_label, alpha, g_hi, g_lo, s_exp, E, idx_scale);
                    {
                        const bool mine_ok = L8 < SGNS_CH ? r.mine >= 0 : (L8 == SGNS_CH && r.act && first);
                        if (!mine_ok) g = 0.f;
                    }
                    float gk[SGNS_CH + 1];
#pragma unroll
                    for (int k = 0; k < SGNS_CH; k++) gk[k] = __shfl_sync(FULL, g, k, G);
                    gk[SGNS_CH] = first ? __shfl_sync(FULL, g, SGNS_CH, G) : 0.f;
#pragma unroll
                    for (int k = 0; k < SGNS_CH; k++) {
                        axpy4(neu, gk[k], r.row[k]);
                        red_add4_if(row_addr(base1, (uint32_t)r.tg[k], pitch), scale4(gk[k], v0), gk[k] != 0.f && live && reds_on);
                    }
                    if (first) {
                        axpy4(neu, gk[SGNS_CH], cur);
                        if (hot_w1) {
                            const float4 upd = scale4(gk[SGNS_CH], v0);
                            red_add4_if(row_addr(base1, (uint32_t)w1, pitch), upd, gk[SGNS_CH] != 0.f && live && reds_on);
                            if (PF && reds_on) upd_last = upd;
                        }
                        else { axpy4(d1, gk[SGNS_CH], v0); axpy4(cur, gk[SGNS_CH], v0); }
                    }
                    if ((!MULTI || r.j == NCH - 1) && r.act && live) { // the pair is complete: syn0[last] += neu
                        if (r.last < a.hot) { // write-through word: sent at once, the next pair on this row reads it back from L2
                            if (reds_on) red_add4(reinterpret_cast<float4 *>(a.syn0 + (int64_t)r.last * a.stride) + lane, neu);
                            if (PF && reds_on) { my_delta[r.c * n4 + lane] = neu; if (lane == 0) my_tag[r.c] = u; }
                        } else { // kept in the warp's cache until the batch is flushed
                            float4 dl = my_delta[r.c * n4 + lane];
                            dl.x += neu.x; dl.y += neu.y; dl.z += neu.z; dl.w += neu.w;
                            my_delta[r.c * n4 + lane] = dl;
                        }
                    }
                };

                const int U = (c_max - c_min + 1 + (gpw_eff - 1)) * NCH;
                stage_r rA;
                rA.v0 = rA.cur = zero4;
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) rA.row[k] = zero4;
                stage_t t1 = stageT();
                if (PF) {
                    stage_r rB;
                    rB.v0 = rB.cur = zero4;
#pragma unroll
                    for (int k = 0; k < SGNS_CH; k++) rB.row[k] = zero4;
                    stageR(t1, rA, 0);
                    t1 = stageT();
                    for (int u = 0; u < U; u += 2) {
                        stageR(t1, rB, 1); // the rows of unit u + 1 (nothing is requested past the end: act is false there)
                        t1 = stageT();
                        compute(rA, u, 0);
                        __syncwarp(); // the cache rows written in this unit are read by other groups in later units
                        if (u + 1 < U) {
                            stageR(t1, rA, 0);
                            t1 = stageT();
                            compute(rB, u + 1, 1);
                            __syncwarp();
                        }
                    }
                    if (PF == 2) cp_async_wait<0>(); // no copy may land in a stage the next batch is already filling
                } else
                for (int u = 0; u < U; u++) {
                    stageR(t1, rA, 0);
                    t1 = stageT();
                    compute(rA, u, 0);
                    __syncwarp(); // the cache rows written in this unit are read by other groups in later units
                }
                red_add4_if(row_addr(base1, (uint32_t)w1, pitch), d1, valid && live && reds_on && !hot_w1);
                // flush the warp's pending context-row updates: one 128-bit reduction per slot that moved
                for (int q = wl; q < n_tok * n4; q += 32) {
                    const float4 dl = my_delta[q];
                    const int row = q / n4, slot = q - row * n4;
                    if (dl.x != 0.f || dl.y != 0.f || dl.z != 0.f || dl.w != 0.f) {
                        // (PF: a write-through row's entry is the copy of an update that has been sent already)
                        if (reds_on && !(PF && mytok[row] < a.hot)) red_add4(reinterpret_cast<float4 *>(a.syn0 + (int64_t)mytok[row] * a.stride) + slot, dl);
                        my_delta[q] = zero4;
                    }
                }
                if (PF) for (int j = wl; j < Lmax; j += 32) my_tag[j] = -2; // units are counted per batch
                __threadfence(); // the next batch re-reads these rows from L2
                __syncwarp();
            }
            pairs += (unsigned)npairs;
        }
    }
    if ((threadIdx.x % G) == 0 && pairs) atomicAdd(a.pairs, pairs);
}

// ---------------------------------------------------------------------------------------------------------
// Kernel G: kernel F's semantics with the parallelism INSIDE the sentence.  Parity bounds the number of sentences in
// flight (a few hundred on a 19 K-word vocabulary: profiles/r2s5_fullsize_staleness_kernelF.json), and one warp per
// sentence then leaves the GPU nearly empty.  A BLOCK owns a sentence, one lane group per centre position (6 warps for
// 24 positions at G = 8), and the pairs run as a WAVEFRONT: in round u the group of centre i takes context u - i.
// Two pairs of a sentence conflict only if they share the centre (its output row syn1neg[w1], private to the group) or
// the context (its row syn0[last]); the wavefront keeps both relative orders of word2vec's centre-major loop -- every
// centre sees its contexts in ascending order, every context its centres in ascending order -- so the schedule is
// conflict-equivalent to the sequential loop (2 n - 3 rounds is the shortest such schedule: the chain (0,1) ... (0,n-1),
// (1,n-1) ... (n-1,n-2) must stay in order).  A first version walked the contexts round-robin ((i + r) mod n: n - 1
// rounds, every group busy); it is a valid order too but not the reference's, and its embedding agreed with the
// oracle's only to 0.81 where kernel F reaches 0.92 (profiles/r2s6_fullsize_staleness_kernelG_roundrobin.json).
// Every pair reads syn0[last] fresh from L2 plus the block's own pending delta from shared memory; a context row is
// flushed (one 128-bit reduction per slot) in the round after its last centre, a centre's output-row delta after its
// last context, so nothing stays pending longer than ~n rounds.
template <int G, bool MULTI, int MAXT>
__global__ void __launch_bounds__(MAXT, 2)
k_sgns_block(const sgns_args a) {
    static_assert(G == 8 || G == 16 || G == 32, "lane groups of 8, 16 or 32");
    extern __shared__ __align__(16) int32_t smem_g[];
    constexpr unsigned FULL = 0xffffffffu;
    const int n4 = a.n4, Lmax = a.Lmax;
    const int nwords = (a.neg_table_size + 31) >> 5;
    float4 *delta = reinterpret_cast<float4 *>(smem_g);                       // [Lmax][n4]
    float *s_exp = reinterpret_cast<float *>(delta + (size_t)Lmax * n4);
    int32_t *tok = reinterpret_cast<int32_t *>(s_exp + a.exp_table_size);
    uint32_t *s_bits = reinterpret_cast<uint32_t *>(tok + Lmax);
    uint32_t *s_pref = s_bits + nwords;
    const bool smem_neg = a.neg_bits != nullptr;
    // the negatives of every (centre, context) pair of the sentence, drawn by all threads before the rounds start:
    // [Lmax][Lmax][K] vocabulary indices, -1 = none (a draw that hit the centre itself is skipped, as in the oracle)
    int32_t *s_tg = reinterpret_cast<int32_t *>(s_bits + (smem_neg ? 2 * nwords : 0));
    for (int q = threadIdx.x; q < a.exp_table_size; q += blockDim.x) s_exp[q] = a.exp_table[q];
    if (smem_neg)
        for (int q = threadIdx.x; q < 2 * nwords; q += blockDim.x) s_bits[q] = a.neg_bits[q];
    const int lane = threadIdx.x % G;
    const int i = threadIdx.x / G;          // this lane group's centre position, for every sentence of the block
    const int E = a.exp_table_size;
    const float idx_scale = (float)E / SGNS_MAX_EXP / 2.0f;
    const int win = a.window;
    const int64_t N = a.n_sent;
    const double inv_total = 1.0 / (double)((int64_t)a.epochs * a.n_global);
    const uint32_t tsize = (uint32_t)a.neg_table_size, vm1 = (uint32_t)(a.V > 1 ? a.V - 1 : 1);
    const double inv_tsize = 1.0 / (double)tsize, inv_vm1 = 1.0 / (double)vm1;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int K = a.V >= 2 ? a.negative : 0;
    const int NCH = MULTI ? max(1, (K + SGNS_CH - 1) / SGNS_CH) : 1;
    const bool live = lane < n4;
    const uint32_t pitch = (uint32_t)a.stride * 4u;
    const char *base0 = reinterpret_cast<const char *>(a.syn0) + (live ? lane : 0) * 16;
    const char *base1 = reinterpret_cast<const char *>(a.syn1neg) + (live ? lane : 0) * 16;
    const int L8 = lane & 7;
    const bool up4 = (lane & 4) != 0, up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
    const float my_label = L8 == SGNS_CH ? 1.f : 0.f;
    const bool reds_on = !(a.dbg & 1);
    unsigned long long pairs = 0;

    struct stage_t { int32_t last; bool act; int j; int c; };
    struct stage_r { int32_t last; bool act; int j; int c; int32_t mine; int32_t tg[SGNS_CH]; float4 row[SGNS_CH]; float4 v0; };

    for (int ep = a.ep_lo; ep < a.ep_hi; ep++) {
        for (int64_t s = a.s_lo + blockIdx.x; s < a.s_hi; s += a.n_groups) { // n_groups = blocks = sentences in flight
            __syncthreads(); // the previous sentence's flush has read the delta cache
            int32_t tk = -1;
            if ((int)threadIdx.x < Lmax) { tk = a.wtok[(int64_t)threadIdx.x * N + s]; tok[threadIdx.x] = tk; }
            for (int q = threadIdx.x; q < Lmax * n4; q += blockDim.x) delta[q] = zero4;
            const int n_tok = __syncthreads_count(tk >= 0); // the compacted sentence: tokens first, then padding
            if (n_tok < 2) continue;
            float alpha = a.lr * (float)(1.0 - (double)((int64_t)ep * a.n_global + a.s_off + s) * inv_total);
            if (alpha < a.min_lr) alpha = a.min_lr;
            const float g_hi = (my_label - 1.f) * alpha, g_lo = my_label * alpha;
            const uint64_t S = sgns_sentence_rng(a.seed, ep, s + a.s_off);
            // ---- draw phase: K negatives for each of the n (n - 1) ordered pairs, off the rounds' critical path and on every lane
            for (int e = threadIdx.x; e < n_tok * n_tok * K; e += blockDim.x) {
                const int kq = e % K, ic = e / K;
                const int cc = ic % n_tok, ii = ic / n_tok;
                if (cc == ii) continue;
                const uint64_t nsk = a.lcg_a[kq] * sgns_pair_rng(S, ii, cc) + a.lcg_c[kq]; // the LCG is affine: state after kq + 1 steps
                const uint32_t idx = mod48(nsk >> 16, tsize, inv_tsize);
                int32_t t = smem_neg ? neg_lookup(s_bits, s_pref, idx) : a.neg_table[idx];
                if (t <= 0 || t >= a.V) t = (int32_t)mod64(nsk, vm1, inv_vm1) + 1;   // DL4J: target = r % (V - 1) + 1
                s_tg[(ii * Lmax + cc) * K + kq] = t == tok[ii] ? -1 : t;
            }
            __syncthreads();
            const bool valid = i < n_tok;
            const int32_t w1 = valid ? tok[i] : 0;
            const int b = (int32_t)(uint32_t)sgns_position_rng(S, valid ? i : 0) % win;
            const int lo = valid ? i - win + b : 1, hi = valid ? i + win - b : 0; // inclusive context range; empty if invalid
            float4 cur = zero4, d1 = zero4, neu = zero4, v0p = zero4;
            ldcg4_into(cur, row_addr(base1, (uint32_t)w1, pitch), valid && live); // private copy of syn1neg[w1]
            int npairs = 0;
            int rT = 1, jT = 0; // (round, chunk) of the next unit the T stage hands out; rounds 1 .. 2 n_tok - 3
            const int n_rounds = 2 * n_tok - 3;
            bool d1_flushed = false;

            auto stageT = [&]() {
                stage_t t;
                t.j = jT;
                const int c = rT - i;   // wavefront: centre i meets context u - i in round u
                const bool in_round = valid && rT <= n_rounds && c >= 0 && c < n_tok && c != i;
                t.c = in_round ? c : 0;
                t.last = in_round ? tok[t.c] : -1;
                t.act = in_round && t.c >= lo && t.c <= hi && t.last >= 0 && t.last != w1;
                if (MULTI) { if (++jT == NCH) { jT = 0; rT++; } }
                else rT++;
                return t;
            };
            auto stageR = [&](const stage_t &t, stage_r &r) {
                r.last = t.last; r.act = t.act; r.j = t.j; r.c = t.c;
                const int32_t *tgp = s_tg + ((i < Lmax ? i : 0) * Lmax + t.c) * K + t.j * SGNS_CH; // the pair's negatives of this chunk (broadcast reads)
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) r.tg[k] = (t.act && t.j * SGNS_CH + k < K) ? tgp[k] : -1;
                r.mine = (t.act && L8 < SGNS_CH && t.j * SGNS_CH + L8 < K) ? tgp[L8] : -1;
                if (!MULTI || t.j == 0) ldcg4_into(r.v0, row_addr(base0, (uint32_t)t.last, pitch), t.act && live);
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) ldcg4_into(r.row[k], row_addr(base1, (uint32_t)r.tg[k], pitch), r.tg[k] >= 0 && live);
            };
            auto compute = [&](const stage_r &r) {
                if (!__any_sync(FULL, r.act)) return;
                const bool first = !MULTI || r.j == 0;
                if (first) {
                    npairs += r.act;
                    neu = zero4;
                    v0p = r.v0; // L2's value + what this sentence has added to the row so far
                    if (r.act && live) { const float4 dl = delta[r.c * n4 + lane]; v0p.x += dl.x; v0p.y += dl.y; v0p.z += dl.z; v0p.w += dl.w; }
                }
                const float4 v0 = v0p;
                float d0 = dot4(v0, r.row[0]), d1v = dot4(v0, r.row[1]), d2 = dot4(v0, r.row[2]), d3 = dot4(v0, r.row[3]);
                float d4 = dot4(v0, r.row[4]), d5 = first ? dot4(v0, cur) : 0.f;
                float e0 = (up4 ? d4 : d0) + __shfl_xor_sync(FULL, up4 ? d0 : d4, 4);
                float e1 = (up4 ? d5 : d1v) + __shfl_xor_sync(FULL, up4 ? d1v : d5, 4);
                float e2 = (up4 ? 0.f : d2) + __shfl_xor_sync(FULL, up4 ? d2 : 0.f, 4);
                float e3 = (up4 ? 0.f : d3) + __shfl_xor_sync(FULL, up4 ? d3 : 0.f, 4);
                float f0 = (up2 ? e2 : e0) + __shfl_xor_sync(FULL, up2 ? e0 : e2, 2);
                float f1 = (up2 ? e3 : e1) + __shfl_xor_sync(FULL, up2 ? e1 : e3, 2);
                float tot = (up1 ? f1 : f0) + __shfl_xor_sync(FULL, up1 ? f0 : f1, 1);
                if (G >= 16) tot += __shfl_xor_sync(FULL, tot, 8);
                if (G >= 32) tot += __shfl_xor_sync(FULL, tot, 16);
                float g = sgns_g_lane(tot, my_label, alpha, g_hi, g_lo, s_exp, E, idx_scale);
                {
                    const bool mine_ok = L8 < SGNS_CH ? r.mine >= 0 : (L8 == SGNS_CH && r.act && first);
                    if (!mine_ok) g = 0.f;
                }
                float gk[SGNS_CH + 1];
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) gk[k] = __shfl_sync(FULL, g, k, G);
                gk[SGNS_CH] = first ? __shfl_sync(FULL, g, SGNS_CH, G) : 0.f;
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) {
                    axpy4(neu, gk[k], r.row[k]);
                    red_add4_if(row_addr(base1, (uint32_t)r.tg[k], pitch), scale4(gk[k], v0), gk[k] != 0.f && live && reds_on);
                }
                if (first) {
                    axpy4(neu, gk[SGNS_CH], cur);
                    axpy4(d1, gk[SGNS_CH], v0);
                    axpy4(cur, gk[SGNS_CH], v0);
                }
                if ((!MULTI || r.j == NCH - 1) && r.act && live) { // the pair is complete: syn0[last] += neu, pending in the block's cache
                    float4 dl = delta[r.c * n4 + lane];
                    dl.x += neu.x; dl.y += neu.y; dl.z += neu.z; dl.w += neu.w;
                    delta[r.c * n4 + lane] = dl;
                }
            };

            // Parity keeps the sentences in flight few (two blocks per SM), so latency is hidden INSIDE the block: the rows of
            // unit k + 1 are requested before unit k is computed (two row buffers in registers; the negative-table entries run
            // two units ahead).  What a pair reads early is only L2's copy; the sentence's own pending delta is added from
            // shared memory when the pair is computed, after the barrier.
            stage_r rA, rB;
            rA.v0 = rB.v0 = zero4;
#pragma unroll
            for (int k = 0; k < SGNS_CH; k++) rA.row[k] = rB.row[k] = zero4;
            const int last_ctx = min(hi, n_tok - 1);   // beyond it this centre has no context left
            // every thread of the block walks the same unit sequence (round u = 1 + k / NCH, chunk k % NCH): rT / jT advance identically everywhere
            auto before_compute = [&](int k) {
                __syncthreads(); // the delta rows written in the previous unit are read now (one writer per row per round)
                if (MULTI && k % NCH != 0) return;
                const int u = 1 + k / NCH;
                // context row i saw its last centre in round i + n_tok - 1 at the latest: its group sends the row's pending delta
                // now, nobody reads or writes it again in this sentence
                if (valid && live && u == i + n_tok) {
                    const float4 dl = delta[i * n4 + lane];
                    if (reds_on && (dl.x != 0.f || dl.y != 0.f || dl.z != 0.f || dl.w != 0.f))
                        red_add4(reinterpret_cast<float4 *>(a.syn0 + (int64_t)w1 * a.stride) + lane, dl);
                }
                // and the centre's own output row once its contexts are exhausted
                if (!d1_flushed && u - i > last_ctx) {
                    red_add4_if(row_addr(base1, (uint32_t)w1, pitch), d1, valid && live && reds_on);
                    d1_flushed = true;
                }
            };
            const int U = n_rounds * NCH;
            stage_t t1 = stageT();  // unit 0
            stageR(t1, rA);
            t1 = stageT();          // unit 1
            for (int k = 0; k < U; k += 2) {
                stageR(t1, rB);     // rows of unit k + 1 (nothing is requested past the end: act is false there)
                t1 = stageT();
                before_compute(k);
                compute(rA);
                if (k + 1 < U) {
                    stageR(t1, rA);
                    t1 = stageT();
                    before_compute(k + 1);
                    compute(rB);
                }
            }
            if (!d1_flushed) red_add4_if(row_addr(base1, (uint32_t)w1, pitch), d1, valid && live && reds_on);
            pairs += (unsigned)npairs;
            __syncthreads();
            // the context rows whose last centre came in the final rounds (u == i + n_tok was never reached)
            if (valid && live && i + n_tok > n_rounds) {
                const float4 dl = delta[i * n4 + lane];
                if (reds_on && (dl.x != 0.f || dl.y != 0.f || dl.z != 0.f || dl.w != 0.f))
                    red_add4(reinterpret_cast<float4 *>(a.syn0 + (int64_t)w1 * a.stride) + lane, dl);
            }
        }
    }
    if (lane == 0 && pairs) atomicAdd(a.pairs, pairs);
}

// ---------------------------------------------------------------------------------------------------------
// Kernel H: kernel G with the sentences of a block PIPELINED through the wavefront.  In kernel G half of the lane groups
// idle on average: the wavefront of a sentence fills for n rounds and drains for n rounds.  Here the groups that have
// finished their centre of sentence k start sentence k + 1 at once (sentence k + 1 enters the block n_k rounds after
// sentence k, not 2 n_k - 3), so the drain of one sentence overlaps the fill of the next and every group has a pair in
// (almost) every round.  The number of PAIRS in flight in a block is unchanged (one per lane group), every sentence
// still runs the conflict-equivalent wavefront order; what a block holds pending at any time is the second half of one
// sentence and the first half of the next.  Three sentence slots in shared memory (tokens, pending context-row deltas,
// pre-drawn negatives): sentence k + 2 is set up while k + 1 starts and k drains; a slot is reused only after every row
// of its old sentence has been flushed (start_{k+2} >= start_k + 2 n_k).
// Rows of up to 8 slots, K <= 5 negatives, sentences of up to 24 tokens (192 threads); otherwise kernel G runs.
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 2)
k_sgns_pipe(const sgns_args a) {
    constexpr int G = 8;
    extern __shared__ __align__(16) int32_t smem_h[];
    constexpr unsigned FULL = 0xffffffffu;
    const int n4 = a.n4, Lmax = a.Lmax;
    const int nwords = (a.neg_table_size + 31) >> 5;
    const int K = a.V >= 2 ? a.negative : 0; // <= 5 (host)
    float4 *delta = reinterpret_cast<float4 *>(smem_h);                         // [3][Lmax][n4]
    float *s_exp = reinterpret_cast<float *>(delta + (size_t)3 * Lmax * n4);
    int32_t *tok = reinterpret_cast<int32_t *>(s_exp + a.exp_table_size);       // [3][Lmax]
    int32_t *meta = tok + 3 * Lmax;                                              // [3][8]: n, start, alpha bits, S lo, S hi
    uint32_t *s_bits = reinterpret_cast<uint32_t *>(meta + 24);
    uint32_t *s_pref = s_bits + nwords;
    const bool smem_neg = a.neg_bits != nullptr;
    int32_t *s_tg = reinterpret_cast<int32_t *>(s_bits + (smem_neg ? 2 * nwords : 0)); // [3][Lmax][Lmax][K]
    const int tg_slot = Lmax * Lmax * (K > 0 ? K : 1);
    for (int q = threadIdx.x; q < a.exp_table_size; q += blockDim.x) s_exp[q] = a.exp_table[q];
    if (smem_neg)
        for (int q = threadIdx.x; q < 2 * nwords; q += blockDim.x) s_bits[q] = a.neg_bits[q];
    if (threadIdx.x < 24) meta[threadIdx.x] = 0;
    const int lane = threadIdx.x % G;
    const int i = threadIdx.x / G;          // this lane group's centre position in every sentence
    const int E = a.exp_table_size;
    const float idx_scale = (float)E / SGNS_MAX_EXP / 2.0f;
    const int win = a.window;
    const int64_t N = a.n_sent;
    const double inv_total = 1.0 / (double)((int64_t)a.epochs * a.n_global);
    const uint32_t tsize = (uint32_t)a.neg_table_size, vm1 = (uint32_t)(a.V > 1 ? a.V - 1 : 1);
    const double inv_tsize = 1.0 / (double)tsize, inv_vm1 = 1.0 / (double)vm1;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool live = lane < n4;
    const uint32_t pitch = (uint32_t)a.stride * 4u;
    const char *base0 = reinterpret_cast<const char *>(a.syn0) + (live ? lane : 0) * 16;
    const char *base1 = reinterpret_cast<const char *>(a.syn1neg) + (live ? lane : 0) * 16;
    const int L8 = lane & 7;
    const bool up4 = (lane & 4) != 0, up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
    const float my_label = L8 == SGNS_CH ? 1.f : 0.f;
    const bool reds_on = !(a.dbg & 1);
    unsigned long long pairs = 0;
    __syncthreads();

    struct stage_t { int32_t last; bool act; int c; int slot; };
    struct stage_r { int32_t last; bool act; int c; int slot; int32_t mine; int32_t tg[SGNS_CH]; float4 row[SGNS_CH]; float4 v0; };

    for (int ep = a.ep_lo; ep < a.ep_hi; ep++) {
        // the block's sentences: s_lo + blockIdx.x + k * n_groups.  `ns` of them have been set up; start_m1 / start_m2 and
        // n_m1 / n_m2 are the start rounds and lengths of the last two that were (uniform across the block).
        int64_t s_next = a.s_lo + blockIdx.x;
        int ns = 0, start_m1 = 0, start_m2 = 0, n_m1 = 0, n_m2 = 0, end_round = 0;
        // ---- sets up the next non-empty sentence of the block in slot ns % 3; false when the block has no sentence left
        auto setup_next = [&]() -> bool {
            while (s_next < a.s_hi) {
                const int64_t s = s_next;
                s_next += a.n_groups;
                const int slot = ns % 3;
                __syncthreads();
                int32_t tk = -1;
                if ((int)threadIdx.x < Lmax) { tk = a.wtok[(int64_t)threadIdx.x * N + s]; tok[slot * Lmax + threadIdx.x] = tk; }
                const int n = __syncthreads_count(tk >= 0);
                if (n < 2) continue; // no pair in it
                for (int q = threadIdx.x; q < Lmax * n4; q += blockDim.x) delta[slot * Lmax * n4 + q] = zero4;
                const uint64_t S = sgns_sentence_rng(a.seed, ep, s + a.s_off);
                // sentence k enters n_{k-1} rounds after sentence k - 1, and not before sentence k - 2 has flushed its last row
                // (+ 4 rounds of margin: a sentence is set up 4 rounds before its predecessor starts, so that the stages that run
                // 2-3 rounds ahead of the computation always find it)
                const int start = ns == 0 ? 0 : max(start_m1 + n_m1, ns >= 2 ? start_m2 + 2 * n_m2 + 4 : 0);
                if (threadIdx.x == 0) {
                    float alpha = a.lr * (float)(1.0 - (double)((int64_t)ep * a.n_global + a.s_off + s) * inv_total);
                    if (alpha < a.min_lr) alpha = a.min_lr;
                    meta[slot * 8 + 0] = n; meta[slot * 8 + 1] = start; meta[slot * 8 + 2] = __float_as_int(alpha);
                    meta[slot * 8 + 3] = (int32_t)(uint32_t)S; meta[slot * 8 + 4] = (int32_t)(uint32_t)(S >> 32);
                }
                const int32_t *tk_s = tok + slot * Lmax;
                for (int e = threadIdx.x; e < n * n * K; e += blockDim.x) { // the K negatives of all n (n - 1) ordered pairs
                    const int kq = e % K, ic = e / K;
                    const int cc = ic % n, ii = ic / n;
                    if (cc == ii) continue;
                    const uint64_t nsk = a.lcg_a[kq] * sgns_pair_rng(S, ii, cc) + a.lcg_c[kq];
                    const uint32_t idx = mod48(nsk >> 16, tsize, inv_tsize);
                    int32_t t = smem_neg ? neg_lookup(s_bits, s_pref, idx) : a.neg_table[idx];
                    if (t <= 0 || t >= a.V) t = (int32_t)mod64(nsk, vm1, inv_vm1) + 1;
                    s_tg[slot * tg_slot + (ii * Lmax + cc) * K + kq] = t == tk_s[ii] ? -1 : t;
                }
                __syncthreads();
                start_m2 = start_m1; n_m2 = n_m1; start_m1 = start; n_m1 = n;
                end_round = start + 2 * n; // every row of this sentence has been flushed by then
                ns++;
                return true;
            }
            return false;
        };
        // ---- where is this lane group in round U?  (slot, context position) or slot = -1
        auto locate = [&](int U, int &slot, int &c) {
            slot = -1; c = 0;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const int n = meta[q * 8 + 0], cc = U - meta[q * 8 + 1] - i;
                if (i < n && cc >= 0 && cc < n) { slot = q; c = cc; }
            }
        };
        if (!setup_next()) continue;
        bool more = setup_next();
        // per-group state of the sentence it is on
        int cur_slot = -1, lo = 1, hi = 0;
        int32_t w1 = 0;
        float alpha = 0.f;
        float4 cur = zero4, d1 = zero4, cur_next = zero4;
        int npairs = 0;
        int UT = 0; // round of the next unit the T stage hands out

        auto stageT = [&]() {
            stage_t t;
            int slot, c;
            locate(UT, slot, c);
            const bool on = slot >= 0 && c != i;
            t.slot = on ? slot : 0;
            t.c = on ? c : 0;
            t.last = on ? tok[t.slot * Lmax + t.c] : -1;
            // the window of the centre: b from the sentence key of that slot (the group may be about to change sentences)
            bool act = false;
            if (on) {
                const uint64_t S = ((uint64_t)(uint32_t)meta[t.slot * 8 + 4] << 32) | (uint32_t)meta[t.slot * 8 + 3];
                const int b = (int32_t)(uint32_t)sgns_position_rng(S, i) % win;
                act = t.c >= i - win + b && t.c <= i + win - b && t.last >= 0 && t.last != tok[t.slot * Lmax + i];
            }
            t.act = act;
            UT++;
            return t;
        };
        auto stageR = [&](const stage_t &t, stage_r &r, int U) {
            r.last = t.last; r.act = t.act; r.c = t.c; r.slot = t.slot;
            const int32_t *tgp = s_tg + t.slot * tg_slot + ((i < Lmax ? i : 0) * Lmax + t.c) * K;
#pragma unroll
            for (int k = 0; k < SGNS_CH; k++) r.tg[k] = (t.act && k < K) ? tgp[k] : -1;
            r.mine = (t.act && L8 < SGNS_CH && L8 < K) ? tgp[L8] : -1;
            ldcg4_into(r.v0, row_addr(base0, (uint32_t)t.last, pitch), t.act && live);
#pragma unroll
            for (int k = 0; k < SGNS_CH; k++) ldcg4_into(r.row[k], row_addr(base1, (uint32_t)r.tg[k], pitch), r.tg[k] >= 0 && live);
            // the group starts a new sentence in round U: its centre's output row is requested one round ahead
            int slot, c;
            locate(U, slot, c);
            if (slot >= 0 && c == 0 && slot != cur_slot)
                ldcg4_into(cur_next, row_addr(base1, (uint32_t)tok[slot * Lmax + i], pitch), live);
        };
        auto compute = [&](const stage_r &r, int U) {
            // ---- time-triggered flushes: row i of a sentence saw its last centre in round start + i + n - 1
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const int n = meta[q * 8 + 0];
                if (i < n && U == meta[q * 8 + 1] + i + n) {
                    if (live) {
                        const float4 dl = delta[(q * Lmax + i) * n4 + lane];
                        if (reds_on && (dl.x != 0.f || dl.y != 0.f || dl.z != 0.f || dl.w != 0.f))
                            red_add4(reinterpret_cast<float4 *>(a.syn0 + (int64_t)tok[q * Lmax + i] * a.stride) + lane, dl);
                    }
                    if (q == cur_slot) { // and the centre's output row: the group has left the sentence
                        red_add4_if(row_addr(base1, (uint32_t)w1, pitch), d1, live && reds_on);
                        cur_slot = -1;
                    }
                }
            }
            // ---- entering a sentence: the group's centre, its window and its private copy of syn1neg[w1]
            int slot, c;
            locate(U, slot, c);
            if (slot >= 0 && slot != cur_slot) {
                cur_slot = slot;
                w1 = tok[slot * Lmax + i];
                alpha = __int_as_float(meta[slot * 8 + 2]);
                const uint64_t S = ((uint64_t)(uint32_t)meta[slot * 8 + 4] << 32) | (uint32_t)meta[slot * 8 + 3];
                const int b = (int32_t)(uint32_t)sgns_position_rng(S, i) % win;
                lo = i - win + b; hi = i + win - b;
                cur = cur_next;
                d1 = zero4;
            }
            if (!__any_sync(FULL, r.act)) return;
            const float g_hi = (my_label - 1.f) * alpha, g_lo = my_label * alpha;
            npairs += r.act;
            float4 neu = zero4;
            float4 v0 = r.v0; // L2's value + what the sentence has added to the row so far
            if (r.act && live) { const float4 dl = delta[(r.slot * Lmax + r.c) * n4 + lane]; v0.x += dl.x; v0.y += dl.y; v0.z += dl.z; v0.w += dl.w; }
            float d0 = dot4(v0, r.row[0]), d1v = dot4(v0, r.row[1]), d2 = dot4(v0, r.row[2]), d3 = dot4(v0, r.row[3]);
            float d4 = dot4(v0, r.row[4]), d5 = dot4(v0, cur);
            float e0 = (up4 ? d4 : d0) + __shfl_xor_sync(FULL, up4 ? d0 : d4, 4);
            float e1 = (up4 ? d5 : d1v) + __shfl_xor_sync(FULL, up4 ? d1v : d5, 4);
            float e2 = (up4 ? 0.f : d2) + __shfl_xor_sync(FULL, up4 ? d2 : 0.f, 4);
            float e3 = (up4 ? 0.f : d3) + __shfl_xor_sync(FULL, up4 ? d3 : 0.f, 4);
            float f0 = (up2 ? e2 : e0) + __shfl_xor_sync(FULL, up2 ? e0 : e2, 2);
            float f1 = (up2 ? e3 : e1) + __shfl_xor_sync(FULL, up2 ? e1 : e3, 2);
            float tot = (up1 ? f1 : f0) + __shfl_xor_sync(FULL, up1 ? f0 : f1, 1);
            float g = sgns_g_lane(tot, my_label, alpha, g_hi, g_lo, s_exp, E, idx_scale);
            {
                const bool mine_ok = L8 < SGNS_CH ? r.mine >= 0 : (L8 == SGNS_CH && r.act);
                if (!mine_ok) g = 0.f;
            }
            float gk[SGNS_CH + 1];
#pragma unroll
            for (int k = 0; k <= SGNS_CH; k++) gk[k] = __shfl_sync(FULL, g, k, G);
#pragma unroll
            for (int k = 0; k < SGNS_CH; k++) {
                axpy4(neu, gk[k], r.row[k]);
                red_add4_if(row_addr(base1, (uint32_t)r.tg[k], pitch), scale4(gk[k], v0), gk[k] != 0.f && live && reds_on);
            }
            axpy4(neu, gk[SGNS_CH], cur);
            axpy4(d1, gk[SGNS_CH], v0);
            axpy4(cur, gk[SGNS_CH], v0);
            if (r.act && live) { // syn0[last] += neu, pending in the block's cache
                float4 dl = delta[(r.slot * Lmax + r.c) * n4 + lane];
                dl.x += neu.x; dl.y += neu.y; dl.z += neu.z; dl.w += neu.w;
                delta[(r.slot * Lmax + r.c) * n4 + lane] = dl;
            }
        };

        stage_r rA, rB;
        rA.v0 = rB.v0 = zero4;
#pragma unroll
        for (int k = 0; k < SGNS_CH; k++) rA.row[k] = rB.row[k] = zero4;
        stage_t t1 = stageT();   // round 0
        stageR(t1, rA, 0);
        t1 = stageT();           // round 1
        for (int U = 0; U <= end_round; U += 2) {
            // the sentence after the newest one is set up as soon as the newest has started (uniform decision)
            if (more && U + 4 >= start_m1) more = setup_next();
            stageR(t1, rB, U + 1);
            t1 = stageT();
            __syncthreads();
            compute(rA, U);
            if (more && U + 5 >= start_m1) more = setup_next();
            stageR(t1, rA, U + 2);
            t1 = stageT();
            __syncthreads();
            compute(rB, U + 1);
        }
        pairs += (unsigned)npairs;
        __syncthreads();
    }
    if (lane == 0 && pairs) atomicAdd(a.pairs, pairs);
}

// ---------------------------------------------------------------------------------------------------------
// Kernel I: kernel G's wavefront with a WARP per pair and the round's pairs handed to the block's warps dynamically.
// What the ncu capture of kernel G shows (profiles/r2s13_sgns_block_tract24.json): a sentence of 24 tokens with word2vec's
// random window has ~283 pairs in 45 rounds, 6.3 active centres per round on average -- 60 % of kernel G's warp-rounds
// carry no pair and still run the staging code (163 warp instructions per pair), the active warps run a 330-instruction
// chain per round with the K + 1 targets of a pair sequential in each lane, and 35 % of all stall samples wait at the
// round barrier for that chain.  Parity caps the sentences in flight (two blocks per SM), so the round latency is what
// sets the throughput.  Here
//   * the pairs of round u (centre i, context u - i, inside i's window) are listed per round while the negatives are
//     drawn, and warp w takes entries w, w + W, ... of the list: no warp stages or computes an empty slot;
//   * a pair is spread over the whole warp: lane = (target t = lane / 4, quarter q = lane % 4), t = 0 the centre's own
//     output row, t = 1 .. K the negatives; a lane holds float4 slots q and 4 + q of ITS target's row and of the context
//     row.  The K + 1 dot products are two shuffles deep, every lane computes its target's sigmoid itself, the negative
//     rows go out as two 128-bit reductions per lane, and neu1e = sum_t g_t row_t is a 7-shuffle transposed reduction
//     that leaves one float of the sum in every lane;
//   * the centres' output rows (their private copies and deltas) live in shared memory beside the context-row deltas,
//     because a centre is no longer tied to a lane group;
//   * the rows of a warp's next pair are requested before the current one is computed (two register sets), as in kernel G.
// Same pair / negative enumeration, same wavefront order (conflict-equivalent to the centre-major loop), same flush
// points as kernel G.  Rows of up to 8 slots, K <= 7, sentences of up to 32 tokens.
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 2)
k_sgns_wave(const sgns_args a) {
    extern __shared__ __align__(16) int32_t smem_i[];
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int RP = 32; // floats per cached row: 8 slots
    const int n4 = a.n4, Lmax = a.Lmax;
    const int nwords = (a.neg_table_size + 31) >> 5;
    const int K = a.V >= 2 ? a.negative : 0;
    const bool smem_neg = a.neg_bits != nullptr;
    float *delta = reinterpret_cast<float *>(smem_i);       // [Lmax][32] pending syn0 updates of the sentence's context rows
    float *cur = delta + Lmax * RP;                         // [Lmax][32] the centres' output rows syn1neg[w_i] as this sentence sees them
    float *d1 = cur + Lmax * RP;                            // [Lmax][32] what this sentence has added to them
    float *s_exp = d1 + Lmax * RP;
    int32_t *tok = reinterpret_cast<int32_t *>(s_exp + a.exp_table_size);
    int32_t *s_lo = tok + Lmax, *s_hi = s_lo + Lmax, *s_fr = s_hi + Lmax;
    int32_t *s_cnt = s_fr + Lmax;                           // [2 Lmax] pairs of round u
    uint32_t *s_bits = reinterpret_cast<uint32_t *>(s_cnt + 2 * Lmax);
    uint32_t *s_pref = s_bits + nwords;
    int32_t *s_tg = reinterpret_cast<int32_t *>(s_bits + (smem_neg ? 2 * nwords : 0)); // [Lmax][Lmax][K] negatives of every pair
    uint8_t *s_list = reinterpret_cast<uint8_t *>(s_tg + Lmax * Lmax * max(K, 1));    // [2 Lmax][Lmax] centres of round u
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    for (int q = tid; q < a.exp_table_size; q += blockDim.x) s_exp[q] = a.exp_table[q];
    if (smem_neg)
        for (int q = tid; q < 2 * nwords; q += blockDim.x) s_bits[q] = a.neg_bits[q];
    const int t = lane >> 2, q4 = lane & 3;
    const bool liveA = q4 < n4, liveB = 4 + q4 < n4;
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
    const int my_pos = (q4 + (h16 ? 4 : 0)) * 4 + (h8 ? 2 : 0) + (h4 ? 1 : 0); // the float of the row this lane ends up owning in the neu1e sum
    const bool pos_live = (q4 + (h16 ? 4 : 0)) < n4;
    const int E = a.exp_table_size;
    const float idx_scale = (float)E / SGNS_MAX_EXP / 2.0f;
    const int win = a.window;
    const int64_t N = a.n_sent;
    const double inv_total = 1.0 / (double)((int64_t)a.epochs * a.n_global);
    const uint32_t tsize = (uint32_t)a.neg_table_size, vm1 = (uint32_t)(a.V > 1 ? a.V - 1 : 1);
    const double inv_tsize = 1.0 / (double)tsize, inv_vm1 = 1.0 / (double)vm1;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t pitch = (uint32_t)a.stride * 4u;
    const char *base0 = reinterpret_cast<const char *>(a.syn0) + q4 * 16;
    const char *base1 = reinterpret_cast<const char *>(a.syn1neg) + q4 * 16;
    const float my_label = t == 0 ? 1.f : 0.f;
    const bool reds_on = !(a.dbg & 1);
    unsigned long long pairs = 0;

    struct item_t { int u, i, c; int32_t tg; uint64_t ra; float4 vA, vB, rA, rB; };
    item_t A, B;
    A.vA = A.vB = A.rA = A.rB = B.vA = B.vB = B.rA = B.rB = zero4; // slots that carry no data are never loaded and stay zero
    A.ra = B.ra = 0;

    for (int ep = a.ep_lo; ep < a.ep_hi; ep++) {
        int32_t tk_next = -1;
        if (tid < Lmax && a.s_lo + blockIdx.x < a.s_hi) tk_next = a.wtok[(int64_t)tid * N + a.s_lo + blockIdx.x];
        for (int64_t s = a.s_lo + blockIdx.x; s < a.s_hi; s += a.n_groups) { // n_groups = blocks = sentences in flight
            __syncthreads(); // the previous sentence's final flush has read the caches
            const int32_t tk = tk_next;
            if (tid < Lmax) {
                tok[tid] = tk;
                if (s + a.n_groups < a.s_hi) tk_next = a.wtok[(int64_t)tid * N + s + a.n_groups]; // lands during this sentence's rounds
            }
            for (int e = tid; e < Lmax * RP; e += blockDim.x) { delta[e] = 0.f; d1[e] = 0.f; }
            const int n_tok = __syncthreads_count(tid < Lmax && tk >= 0); // the compacted sentence: tokens first, then padding
            if (n_tok < 2) continue;
            float alpha = a.lr * (float)(1.0 - (double)((int64_t)ep * a.n_global + a.s_off + s) * inv_total);
            if (alpha < a.min_lr) alpha = a.min_lr;
            const float g_hi = (my_label - 1.f) * alpha, g_lo = my_label * alpha;
            const uint64_t S = sgns_sentence_rng(a.seed, ep, s + a.s_off);
            const int R = 2 * n_tok - 3;
            if (tid < n_tok) { // the centre's window (word2vec's random shrink), clamped to the sentence
                const int b = (int32_t)(uint32_t)sgns_position_rng(S, tid) % win;
                const int hi_c = min(tid + win - b, n_tok - 1);
                s_lo[tid] = max(tid - win + b, 0);
                s_hi[tid] = hi_c;
                s_fr[tid] = tid + hi_c + 1; // the round after its last context: its output-row delta is sent then
            }
            for (int e = tid; e < n_tok * 8; e += blockDim.x) { // private copies of the centres' output rows
                const int i = e >> 3, slot = e & 7;
                float4 v = zero4;
                if (slot < n4) v = __ldcg(reinterpret_cast<const float4 *>(a.syn1neg + (int64_t)tok[i] * a.stride) + slot);
                reinterpret_cast<float4 *>(cur)[i * 8 + slot] = v;
            }
            __syncthreads();
            // ---- draw phase: the K negatives of every pair inside a window; the pairs of every round
            for (int e = tid; e < n_tok * n_tok * K; e += blockDim.x) {
                const int kq = e % K, ic = e / K;
                const int cc = ic % n_tok, ii = ic / n_tok;
                if (cc == ii || cc < s_lo[ii] || cc > s_hi[ii]) continue;
                const uint64_t nsk = a.lcg_a[kq] * sgns_pair_rng(S, ii, cc) + a.lcg_c[kq]; // the LCG is affine: state after kq + 1 steps
                const uint32_t idx = mod48(nsk >> 16, tsize, inv_tsize);
                int32_t tg = smem_neg ? neg_lookup(s_bits, s_pref, idx) : a.neg_table[idx];
                if (tg <= 0 || tg >= a.V) tg = (int32_t)mod64(nsk, vm1, inv_vm1) + 1;   // DL4J: target = r % (V - 1) + 1
                s_tg[(ii * Lmax + cc) * K + kq] = tg == tok[ii] ? -1 : tg;
            }
            if (tid >= 1 && tid <= R) { // round u = tid: centre i meets context u - i
                int n = 0;
                for (int i = max(0, tid - (n_tok - 1)); i <= min(n_tok - 1, tid); i++) {
                    const int c = tid - i;
                    if (c != i && c >= s_lo[i] && c <= s_hi[i] && tok[c] != tok[i]) s_list[tid * Lmax + n++] = (uint8_t)i;
                }
                s_cnt[tid] = n;
            }
            __syncthreads();

            int ubar = 0; // rounds this warp has opened
            auto open_round = [&]() {
                asm volatile("bar.sync 0;" ::: "memory"); // what round u - 1 wrote to the caches is read in round u
                ubar++;
                if (warp == W - 1) { // the warp with the fewest pairs sends what is complete
                    const int c = ubar - n_tok; // context row c saw its last centre in round c + n_tok - 1 at the latest
                    if (c >= 0 && lane < n4) {
                        const float4 dl = reinterpret_cast<const float4 *>(delta)[c * 8 + lane];
                        if (reds_on && (dl.x != 0.f || dl.y != 0.f || dl.z != 0.f || dl.w != 0.f))
                            red_add4(reinterpret_cast<float4 *>(a.syn0 + (int64_t)tok[c] * a.stride) + lane, dl);
                    }
                    unsigned mk = __ballot_sync(FULL, lane < n_tok && s_fr[lane] == ubar);
                    while (mk) { // centres whose contexts are exhausted
                        const int i = __ffs(mk) - 1;
                        mk &= mk - 1;
                        if (lane < n4 && reds_on)
                            red_add4(reinterpret_cast<float4 *>(a.syn1neg + (int64_t)tok[i] * a.stride) + lane, reinterpret_cast<const float4 *>(d1)[i * 8 + lane]);
                    }
                }
            };
            int pu = 1, pp = warp; // the next list entry this warp has not requested yet
            auto issue = [&](item_t &r) {
                while (pu <= R && pp >= s_cnt[pu]) { pu++; pp = warp; }
                r.u = pu;
                if (pu > R) return;
                const int i = s_list[pu * Lmax + pp], c = pu - i;
                pp += W;
                r.i = i; r.c = c;
                const int32_t tg = (t >= 1 && t <= K) ? s_tg[(i * Lmax + c) * K + t - 1] : -1;
                r.tg = tg;
                const uint64_t va = row_addr(base0, (uint32_t)tok[c], pitch);
                r.ra = row_addr(base1, (uint32_t)max(tg, 0), pitch);
                ldcg4_into(r.vA, va, liveA);
                ldcg4_into(r.vB, va + 64, liveB);
                ldcg4_into(r.rA, r.ra, tg >= 0 && liveA);
                ldcg4_into(r.rB, r.ra + 64, tg >= 0 && liveB);
            };
            auto compute = [&](const item_t &r) {
                const float4 *dc = reinterpret_cast<const float4 *>(delta) + r.c * 8;
                float4 *ci = reinterpret_cast<float4 *>(cur) + r.i * 8;
                // the context row as this sentence sees it: L2's value + the sentence's pending delta
                float4 vA = add4(r.vA, dc[q4]), vB = add4(r.vB, dc[4 + q4]);
                float4 rA = r.rA, rB = r.rB;
                if (t == 0) { rA = ci[q4]; rB = ci[4 + q4]; }
                float part = dot4(vA, rA) + dot4(vB, rB);
                part += __shfl_xor_sync(FULL, part, 1);
                part += __shfl_xor_sync(FULL, part, 2);
                float g = sgns_g_lane(part, my_label, alpha, g_hi, g_lo, s_exp, E, idx_scale);
                if (!(t == 0 || r.tg >= 0)) g = 0.f;
                const float4 uA = scale4(g, vA), uB = scale4(g, vB); // the target row's update
                const bool send = t != 0 && g != 0.f && reds_on;
                red_add4_if(r.ra, uA, send && liveA);
                red_add4_if(r.ra + 64, uB, send && liveB);
                if (t == 0) { // the centre's own output row: private copy and its delta
                    float4 *di = reinterpret_cast<float4 *>(d1) + r.i * 8;
                    if (liveA) { ci[q4] = add4(rA, uA); di[q4] = add4(di[q4], uA); }
                    if (liveB) { ci[4 + q4] = add4(rB, uB); di[4 + q4] = add4(di[4 + q4], uB); }
                }
                // neu1e = sum over the targets of g_t row_t: transposed reduction over the lanes' target bits
                const float4 nA = scale4(g, rA), nB = scale4(g, rB);
                const float m0 = (h16 ? nB.x : nA.x) + __shfl_xor_sync(FULL, h16 ? nA.x : nB.x, 16);
                const float m1 = (h16 ? nB.y : nA.y) + __shfl_xor_sync(FULL, h16 ? nA.y : nB.y, 16);
                const float m2 = (h16 ? nB.z : nA.z) + __shfl_xor_sync(FULL, h16 ? nA.z : nB.z, 16);
                const float m3 = (h16 ? nB.w : nA.w) + __shfl_xor_sync(FULL, h16 ? nA.w : nB.w, 16);
                const float p0 = (h8 ? m2 : m0) + __shfl_xor_sync(FULL, h8 ? m0 : m2, 8);
                const float p1 = (h8 ? m3 : m1) + __shfl_xor_sync(FULL, h8 ? m1 : m3, 8);
                const float val = (h4 ? p1 : p0) + __shfl_xor_sync(FULL, h4 ? p0 : p1, 4);
                if (pos_live) delta[r.c * RP + my_pos] += val; // syn0[last] += neu1e, pending in the block's cache
                if (lane == 0) pairs++;
            };

            issue(A);
            while (A.u <= R) {
                issue(B); // the rows of this warp's next pair are in flight while this one is computed
                while (ubar < A.u) open_round();
                compute(A);
                if (B.u > R) break;
                issue(A);
                while (ubar < B.u) open_round();
                compute(B);
            }
            while (ubar < R) open_round();
            __syncthreads();
            // what the rounds did not send: the context rows whose last centre came in the final rounds, the last centres' rows
            for (int e = tid; e < n_tok * 8; e += blockDim.x) {
                const int row = e >> 3, slot = e & 7;
                if (slot >= n4 || !reds_on) continue;
                if (row + n_tok > R) {
                    const float4 dl = reinterpret_cast<const float4 *>(delta)[row * 8 + slot];
                    if (dl.x != 0.f || dl.y != 0.f || dl.z != 0.f || dl.w != 0.f)
                        red_add4(reinterpret_cast<float4 *>(a.syn0 + (int64_t)tok[row] * a.stride) + slot, dl);
                }
                if (s_fr[row] > R) {
                    const float4 dd = reinterpret_cast<const float4 *>(d1)[row * 8 + slot];
                    if (dd.x != 0.f || dd.y != 0.f || dd.z != 0.f || dd.w != 0.f)
                        red_add4(reinterpret_cast<float4 *>(a.syn1neg + (int64_t)tok[row] * a.stride) + slot, dd);
                }
            }
        }
    }
    if (lane == 0 && pairs) atomicAdd(a.pairs, pairs);
}

// ---------------------------------------------------------------------------------------------------------
// Kernel C': kernel C with the rows of a unit staged in SHARED MEMORY by cp.async instead of registers.
// EXPERIMENTAL (DGE_SGNS_DEBUG bit 256; never chosen by default; not yet measured on the GPU).  Motivation, from
// the ncu source page of kernel C on tract x 24 (profiles/r1_stalls_sgns15_tract24.txt): 40.6 % of all stall samples
// sit on ONE instruction, the first FMUL that consumes the rows requested earlier in the same unit -- the warps wait
// for L2.  A second row buffer in REGISTERS cost a resident block (128 registers, 4 blocks/SM) and lost 6 %
// (profiles/r1s6_sgns_builds.txt).  cp.async.cg (LDGSTS, L2 only) keeps the rows of unit u+1 in flight through the
// whole compute of unit u without holding a register: per group 2 stages x 6 rows x G slots x 16 B.  Every lane
// reads back exactly the slots it copied itself, so cp.async.wait_group is the only synchronisation needed.
__device__ __forceinline__ float4 lds4(uint32_t smem_addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(smem_addr));
    return r;
}

__device__ __forceinline__ float sgns_g_lane(float tot, float label, float alpha, float g_hi, float g_lo, const float *s_exp,
                                             int E, float idx_scale);
// BLK = resident blocks per SM the register allocation is made for (5: 96 registers, no spill to speak of; 6: 80; 7: 72
// with a few dozen bytes of spill -- more warps to hide the L2 latency with; A/B by DGE_SGNS_F_BLOCKS_*).
template <int G, bool MULTI, int BLK>
__global__ void __launch_bounds__(128, BLK)
k_sgns_items_v3(const sgns_args a) {
    static_assert(G == 8 || G == 16 || G == 32, "lane groups of 8, 16 or 32");
    extern __shared__ __align__(16) int32_t smem_v3[];
    int32_t *const smem = smem_v3;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int GPW = 32 / G;
    constexpr bool MERGE_SYN0 = GPW > 1;
    constexpr int ROWS = SGNS_CH + 1;                 // slot 0: syn0[last]; 1..5: the negatives' syn1neg rows
    constexpr int STAGE_BYTES = ROWS * G * 16;        // one unit of one group
    // dynamic shared memory: [row stages of every group][sigmoid table][staged sentence of every group]
    const int groups_per_block = blockDim.x / G;
    float *s_exp = reinterpret_cast<float *>(reinterpret_cast<char *>(smem) + (size_t)groups_per_block * 2 * STAGE_BYTES);
    const int gpw_eff = (a.dbg & 8) ? 1 : GPW;
    const int lane = threadIdx.x % G;
    const int gw = (threadIdx.x & 31) / G;
    int32_t *mytok = reinterpret_cast<int32_t *>(s_exp + a.exp_table_size) + (threadIdx.x / G) * a.Lmax;
    const uint32_t my_rows = (uint32_t)__cvta_generic_to_shared(smem) + (uint32_t)(threadIdx.x / G) * 2u * STAGE_BYTES + (uint32_t)lane * 16u;
    for (int i = threadIdx.x; i < a.exp_table_size; i += blockDim.x) s_exp[i] = a.exp_table[i];
    for (int i = threadIdx.x; i < groups_per_block * 2 * STAGE_BYTES / 4; i += blockDim.x) smem[i] = 0; // finite stale values
    __syncthreads();
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int E = a.exp_table_size;
    const float idx_scale = (float)E / SGNS_MAX_EXP / 2.0f;
    const int win = a.window;
    const int64_t N = a.n_sent;
    const int Lmax = a.Lmax;
    const int64_t item_lo = a.s_lo * Lmax, n_items = a.s_hi * Lmax;
    const double inv_total = 1.0 / (double)((int64_t)a.epochs * a.n_global);
    const uint32_t tsize = (uint32_t)a.neg_table_size, vm1 = (uint32_t)(a.V > 1 ? a.V - 1 : 1);
    const double inv_tsize = 1.0 / (double)tsize, inv_vm1 = 1.0 / (double)vm1;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int K = a.V >= 2 ? a.negative : 0;
    const int NCH = MULTI ? max(1, (K + SGNS_CH - 1) / SGNS_CH) : 1;
    const bool live = lane < a.n4;
    const uint32_t pitch = (uint32_t)a.stride * 4u;
    const char *base0 = reinterpret_cast<const char *>(a.syn0) + (live ? lane : 0) * 16;
    const char *base1 = reinterpret_cast<const char *>(a.syn1neg) + (live ? lane : 0) * 16;
    const int L8 = lane & 7;
    const bool up4 = (lane & 4) != 0, up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
    const float my_label = L8 == SGNS_CH ? 1.f : 0.f;
    const bool reds_on = !(a.dbg & 1);
    unsigned long long pairs = 0;

    struct stage_t { int32_t last; bool act; uint64_t nsk; int32_t traw; int j; };
    struct stage_r { int32_t last; bool act; int j; int32_t mine; }; // the targets are re-broadcast from `mine` where needed

    for (int ep = a.ep_lo; ep < a.ep_hi; ep++) {
        for (int64_t base = item_lo + warp_id * gpw_eff; base < n_items; base += a.n_groups) {
            const int64_t item = base + gw;
            bool valid = item < n_items && gw < gpw_eff;
            const int64_t s = valid ? item / Lmax : 0;
            const int i = valid ? (int)(item - s * Lmax) : 0;
            const long long s_first = __shfl_sync(FULL, (long long)s, 0);
            const bool same_s = MERGE_SYN0 && __all_sync(FULL, valid && (long long)s == s_first);
            __syncwarp();
            int n_tok = 0;
            for (int j = lane; j < Lmax; j += G) { const int32_t tk = a.wtok[(int64_t)j * N + s]; mytok[j] = tk; n_tok += tk >= 0; }
            __syncwarp();
#pragma unroll
            for (int o = G >> 1; o > 0; o >>= 1) n_tok += __shfl_xor_sync(FULL, n_tok, o);
            const int32_t w1 = mytok[i];
            valid = valid && w1 >= 0;
            if (!__any_sync(FULL, valid)) continue;
            float alpha = a.lr * (float)(1.0 - (double)((int64_t)ep * a.n_global + a.s_off + s) * inv_total);
            if (alpha < a.min_lr) alpha = a.min_lr;
            const float g_hi = (my_label - 1.f) * alpha, g_lo = my_label * alpha;
            const uint64_t S = sgns_sentence_rng(a.seed, ep, s + a.s_off);
            const int b = (int32_t)(uint32_t)sgns_position_rng(S, i) % win;
            const int lo = valid ? i - win + b : 1, hi = valid ? i + win - b : 0;
            const int c_min = __reduce_min_sync(FULL, valid ? max(lo, 0) : Lmax);
            const int c_max = __reduce_max_sync(FULL, valid ? min(hi, n_tok - 1) : -1);
            if (c_max < c_min) continue;
            float4 cur = zero4, d1 = zero4, neu = zero4, v0p = zero4;
            ldcg4_into(cur, row_addr(base1, (uint32_t)w1, pitch), valid && live);
            int npairs = 0;
            int cT = c_min, jT = 0;
            uint64_t hc = 0;
            int hcb = -1;

            auto stageT = [&]() {
                stage_t t;
                t.j = jT;
                t.last = cT < Lmax ? mytok[cT] : -1;
                t.act = cT >= lo && cT <= hi && cT != i && t.last >= 0 && t.last != w1;
                if (cT / G != hcb) { hcb = cT / G; hc = sgns_pair_rng(S, i, hcb * G + lane); }
                const uint64_t ns0 = shfl64(hc, cT & (G - 1), G);
                const int kk = jT * SGNS_CH + lane;
                const bool drawer = lane < SGNS_CH && kk < K;
                const int kc = drawer ? kk : 0;
                t.nsk = a.lcg_a[kc] * ns0 + a.lcg_c[kc];
                t.traw = -2;
                if (drawer && t.act) t.traw = a.neg_table[mod48(t.nsk >> 16, tsize, inv_tsize)];
                if (MULTI) { if (++jT == NCH) { jT = 0; cT++; } }
                else cT++;
                return t;
            };
            // resolve the negatives, start the asynchronous copies of all rows of the unit into stage `st`
            auto stageR = [&](const stage_t &t, stage_r &r, int st) {
                r.last = t.last; r.act = t.act; r.j = t.j;
                int32_t tt = t.traw;
                const bool redraw = tt != -2 && (tt <= 0 || tt >= a.V);
                if (__any_sync(FULL, redraw)) {
                    if (redraw) tt = (int32_t)mod64(t.nsk, vm1, inv_vm1) + 1;
                }
                r.mine = (tt != -2 && tt != w1) ? tt : -1;
                const uint32_t dst = my_rows + (uint32_t)st * STAGE_BYTES;
                if (!MULTI || t.j == 0) cp_async16_if(dst, row_addr(base0, (uint32_t)t.last, pitch), t.act && live);
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) {
                    const int32_t tg = __shfl_sync(FULL, r.mine, k, G);
                    cp_async16_if(dst + (uint32_t)(k + 1) * G * 16, row_addr(base1, (uint32_t)tg, pitch), tg >= 0 && live);
                }
                cp_async_commit();
            };
            auto compute = [&](const stage_r &r, int st) {
                if (!__any_sync(FULL, r.act)) return;
                const uint32_t src = my_rows + (uint32_t)st * STAGE_BYTES;
                const bool first = !MULTI || r.j == 0;
                if (first) { npairs += r.act; neu = zero4; }
                if (first) v0p = lds4(src); // MULTI: later chunks of the pair keep the copy (their stage slot 0 is not refilled)
                const float4 v0 = v0p;
                // rows are read from shared memory where they are used (twice: dot product, then neu1e) instead of being
                // held in registers across the reduction
                float dk[SGNS_CH];
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) dk[k] = dot4(v0, lds4(src + (uint32_t)(k + 1) * G * 16));
                const float d0 = dk[0], d1v = dk[1], d2 = dk[2], d3 = dk[3], d4 = dk[4], d5 = first ? dot4(v0, cur) : 0.f;
                float e0 = (up4 ? d4 : d0) + __shfl_xor_sync(FULL, up4 ? d0 : d4, 4);
                float e1 = (up4 ? d5 : d1v) + __shfl_xor_sync(FULL, up4 ? d1v : d5, 4);
                float e2 = (up4 ? 0.f : d2) + __shfl_xor_sync(FULL, up4 ? d2 : 0.f, 4);
                float e3 = (up4 ? 0.f : d3) + __shfl_xor_sync(FULL, up4 ? d3 : 0.f, 4);
                float f0 = (up2 ? e2 : e0) + __shfl_xor_sync(FULL, up2 ? e0 : e2, 2);
                float f1 = (up2 ? e3 : e1) + __shfl_xor_sync(FULL, up2 ? e1 : e3, 2);
                float tot = (up1 ? f1 : f0) + __shfl_xor_sync(FULL, up1 ? f0 : f1, 1);
                if (G >= 16) tot += __shfl_xor_sync(FULL, tot, 8);
                if (G >= 32) tot += __shfl_xor_sync(FULL, tot, 16);
                float g = sgns_g_lane(tot, my_label, alpha, g_hi, g_lo, s_exp, E, idx_scale);
                {
                    const bool mine_ok = L8 < SGNS_CH ? r.mine >= 0 : (L8 == SGNS_CH && r.act && first);
                    if (!mine_ok) g = 0.f;
                }
                float gk[SGNS_CH + 1];
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) gk[k] = __shfl_sync(FULL, g, k, G);
                gk[SGNS_CH] = first ? __shfl_sync(FULL, g, SGNS_CH, G) : 0.f;
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) {
                    axpy4(neu, gk[k], lds4(src + (uint32_t)(k + 1) * G * 16));
                    const int32_t tg = __shfl_sync(FULL, r.mine, k, G); // gk[k] != 0 implies tg >= 0
                    red_add4_if(row_addr(base1, (uint32_t)tg, pitch), scale4(gk[k], v0), gk[k] != 0.f && live && reds_on);
                }
                if (first) {
                    axpy4(neu, gk[SGNS_CH], cur);
                    axpy4(d1, gk[SGNS_CH], v0);
                    axpy4(cur, gk[SGNS_CH], v0);
                }
                if (!MULTI || r.j == NCH - 1) {
                    if (same_s) {
                        float4 ns = neu;
#pragma unroll
                        for (int o = G; o < 32; o <<= 1) {
                            ns.x += __shfl_xor_sync(FULL, ns.x, o); ns.y += __shfl_xor_sync(FULL, ns.y, o);
                            ns.z += __shfl_xor_sync(FULL, ns.z, o); ns.w += __shfl_xor_sync(FULL, ns.w, o);
                        }
                        red_add4_if(row_addr(base0, (uint32_t)r.last, pitch), ns, gw == 0 && live && reds_on);
                    } else {
                        red_add4_if(row_addr(base0, (uint32_t)r.last, pitch), neu, r.act && live && reds_on);
                    }
                }
            };

            const int U = (c_max - c_min + 1) * NCH;
            stage_r rA, rB;
            stage_t t1 = stageT();
            stageR(t1, rA, 0); // copies of unit 0 -> stage 0
            t1 = stageT();     // table entries of unit 1
            for (int u = 0; u < U; u += 2) {
                stageR(t1, rB, 1); // copies of unit u + 1 -> stage 1 (nothing is copied past the end: act is false there)
                t1 = stageT();
                cp_async_wait<1>(); // everything but the newest group has landed: stage 0 is readable
                compute(rA, 0);
                if (u + 1 < U) {
                    stageR(t1, rA, 0);
                    t1 = stageT();
                    cp_async_wait<1>();
                    compute(rB, 1);
                }
            }
            cp_async_wait<0>(); // no copy of this item may land in a stage the next item is already filling
            red_add4_if(row_addr(base1, (uint32_t)w1, pitch), d1, valid && live && reds_on);
            pairs += (unsigned)npairs;
        }
    }
    if (lane == 0 && pairs) atomicAdd(a.pairs, pairs);
}

// ---------------------------------------------------------------------------------------------------------
// Kernel D: the item kernel for NARROW rows (up to 8 float4 slots: D <= 32, i.e. the reference's own D = 8 and
// D = 20).  Same draws and arithmetic as kernel C; a group is 4 lanes holding VPL = 1 or 2 slots each (slot
// lane + 4v), so 8 items run in lockstep per warp instead of 4 and the per-unit overhead (pair hash, negative
// draws, shuffles, sigmoid lookups, addressing) is spread over twice as many pairs.  The reductions are the
// limit of kernel C at these sizes: fewer instructions per pair leave the LSU / L2 reduction
// path less idle.  Lane ownership after the transposed reduction (8 values over 4 lanes, 6 shuffles): lane l owns
// values 2l and 2l+1 -- negatives 0..4 of the chunk and, as value 5, the positive target; lane l therefore also
// draws negatives 2l and 2l+1.
__device__ __forceinline__ float sgns_g_lane(float tot, float label, float alpha, float g_hi, float g_lo, const float *s_exp,
                                             int E, float idx_scale) {
    const float f = (tot + SGNS_MAX_EXP) * idx_scale;
    const int idx = (int)f;
    const float sg = s_exp[min(max(idx, 0), E - 1)];
    float g = (label - sg) * alpha;
    if (idx < 0 || idx >= E) g = 0.f; // table index out of range: the aggregate skips the target
    if (tot > SGNS_MAX_EXP) g = g_hi;
    else if (tot < -SGNS_MAX_EXP) g = g_lo;
    return g;
}

template <int VPL, bool MULTI>
__global__ void __launch_bounds__(128, 4)
k_sgns_items_g4(const sgns_args a) {
    static_assert(VPL == 1 || VPL == 2, "one or two float4 slots per lane");
    constexpr int G = 4;
    extern __shared__ int32_t smem[];
    float *s_exp = reinterpret_cast<float *>(smem);
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int GPW = 32 / G;
    constexpr bool MERGE_SYN0 = false; // summing 8 groups costs 12 shuffles per pair: measured -7 % at D = 16, so off here
    const int gpw_eff = (a.dbg & 8) ? 1 : GPW; // test mode: one item at a time (n_groups = 1)
    const int lane = threadIdx.x % G;
    const int gw = (threadIdx.x & 31) / G;
    int32_t *mytok = smem + a.exp_table_size + (threadIdx.x / G) * a.Lmax;
    for (int i = threadIdx.x; i < a.exp_table_size; i += blockDim.x) s_exp[i] = a.exp_table[i];
    __syncthreads();
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int E = a.exp_table_size;
    const float idx_scale = (float)E / SGNS_MAX_EXP / 2.0f;
    const int win = a.window;
    const int64_t N = a.n_sent;
    const int Lmax = a.Lmax;
    const int64_t item_lo = a.s_lo * Lmax, n_items = a.s_hi * Lmax;
    const double inv_total = 1.0 / (double)((int64_t)a.epochs * a.n_global);
    const uint32_t tsize = (uint32_t)a.neg_table_size, vm1 = (uint32_t)(a.V > 1 ? a.V - 1 : 1);
    const double inv_tsize = 1.0 / (double)tsize, inv_vm1 = 1.0 / (double)vm1;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int K = a.V >= 2 ? a.negative : 0;
    const int NCH = MULTI ? max(1, (K + SGNS_CH - 1) / SGNS_CH) : 1;
    bool live[VPL];
#pragma unroll
    for (int v = 0; v < VPL; v++) live[v] = lane + v * G < a.n4;
    const uint32_t pitch = (uint32_t)a.stride * 4u;
    // slot v of a row sits at base + v * 64 bytes (a dead slot is never dereferenced)
    const char *base0 = reinterpret_cast<const char *>(a.syn0) + (live[0] ? lane : 0) * 16;
    const char *base1 = reinterpret_cast<const char *>(a.syn1neg) + (live[0] ? lane : 0) * 16;
    const bool up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
    // owned values: A = 2*lane (always a negative), B = 2*lane + 1 (lane 2: the positive target, lane 3: nothing)
    const int kA = 2 * lane, kB = 2 * lane + 1;
    const float labelB = kB == SGNS_CH ? 1.f : 0.f;
    unsigned long long pairs = 0;

    struct stage_t { int32_t last; bool act; uint64_t nskA, nskB; int32_t trawA, trawB; int j; };
    struct stage_r { int32_t last; bool act; int j; int32_t mineA, mineB; int32_t tg[SGNS_CH]; float4 row[SGNS_CH][VPL]; float4 v0[VPL]; };

    for (int ep = a.ep_lo; ep < a.ep_hi; ep++) {
        for (int64_t base = item_lo + warp_id * gpw_eff; base < n_items; base += a.n_groups) { // warp-uniform trip count
            const int64_t item = base + gw;
            bool valid = item < n_items && gw < gpw_eff;
            const int64_t s = valid ? item / Lmax : 0;
            const int i = valid ? (int)(item - s * Lmax) : 0;
            // all groups of the warp on one sentence (the rule; not at the tail or in the one-item test mode): they
            // share every context row syn0[last], whose K+1-target updates are then summed in the warp and reduced once
            const long long s_first = __shfl_sync(FULL, (long long)s, 0); // every lane takes part (no short-circuit)
            const bool same_s = MERGE_SYN0 && __all_sync(FULL, valid && (long long)s == s_first);
            __syncwarp();
            int n_tok = 0; // tokens of the (compacted) sentence
            for (int j = lane; j < Lmax; j += G) { const int32_t tk = a.wtok[(int64_t)j * N + s]; mytok[j] = tk; n_tok += tk >= 0; }
            __syncwarp();
#pragma unroll
            for (int o = G >> 1; o > 0; o >>= 1) n_tok += __shfl_xor_sync(FULL, n_tok, o);
            const int32_t w1 = mytok[i];
            valid = valid && w1 >= 0;
            if (!__any_sync(FULL, valid)) continue;
            float alpha = a.lr * (float)(1.0 - (double)((int64_t)ep * a.n_global + a.s_off + s) * inv_total);
            if (alpha < a.min_lr) alpha = a.min_lr;
            const float gB_hi = (labelB - 1.f) * alpha, gB_lo = labelB * alpha; // saturated sigmoid (value B)
            const float gA_hi = -alpha;                                         // value A is always a negative: label 0
            const uint64_t S = sgns_sentence_rng(a.seed, ep, s + a.s_off);
            const int b = (int32_t)(uint32_t)sgns_position_rng(S, i) % win;
            const int lo = valid ? i - win + b : 1, hi = valid ? i + win - b : 0;
            // context positions any group of the warp can pair with: units outside [c_min, c_max] are skipped
            const int c_min = __reduce_min_sync(FULL, valid ? max(lo, 0) : Lmax);
            const int c_max = __reduce_max_sync(FULL, valid ? min(hi, n_tok - 1) : -1);
            if (c_max < c_min) continue;
            float4 cur[VPL], d1[VPL], neu[VPL], v0p[VPL];
#pragma unroll
            for (int v = 0; v < VPL; v++) {
                cur[v] = d1[v] = neu[v] = v0p[v] = zero4;
                ldcg4_into(cur[v], row_addr(base1, (uint32_t)w1, pitch) + v * 64, valid && live[v]);
            }
            int npairs = 0;
            int cT = c_min, jT = 0;
            uint64_t hc = 0;
            int hcb = -1;

            auto stageT = [&]() {
                stage_t t;
                t.j = jT;
                t.last = cT < Lmax ? mytok[cT] : -1;
                t.act = cT >= lo && cT <= hi && cT != i && t.last >= 0 && t.last != w1;
                if (cT / G != hcb) { hcb = cT / G; hc = sgns_pair_rng(S, i, hcb * G + lane); } // warp-uniform condition
                const uint64_t ns0 = shfl64(hc, cT & (G - 1), G);
                const int kkA = jT * SGNS_CH + kA, kkB = jT * SGNS_CH + kB;
                const bool drawA = kA < SGNS_CH && kkA < K, drawB = kB < SGNS_CH && kkB < K;
                t.nskA = a.lcg_a[drawA ? kkA : 0] * ns0 + a.lcg_c[drawA ? kkA : 0];
                t.nskB = a.lcg_a[drawB ? kkB : 0] * ns0 + a.lcg_c[drawB ? kkB : 0];
                t.trawA = t.trawB = -2; // "draws nothing"
                if (drawA && t.act) t.trawA = a.neg_table[mod48(t.nskA >> 16, tsize, inv_tsize)];
                if (drawB && t.act) t.trawB = a.neg_table[mod48(t.nskB >> 16, tsize, inv_tsize)];
                if (MULTI) { if (++jT == NCH) { jT = 0; cT++; } }
                else cT++;
                return t;
            };
            auto stageR = [&](const stage_t &t, stage_r &r) {
                r.last = t.last; r.act = t.act; r.j = t.j;
                int32_t ta = t.trawA, tb = t.trawB;
                const bool reA = ta != -2 && (ta <= 0 || ta >= a.V), reB = tb != -2 && (tb <= 0 || tb >= a.V);
                if (__any_sync(FULL, reA || reB)) { // DL4J: target = r % (V-1) + 1
                    if (reA) ta = (int32_t)mod64(t.nskA, vm1, inv_vm1) + 1;
                    if (reB) tb = (int32_t)mod64(t.nskB, vm1, inv_vm1) + 1;
                }
                r.mineA = (ta != -2 && ta != w1) ? ta : -1;
                r.mineB = (tb != -2 && tb != w1) ? tb : -1;
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) r.tg[k] = __shfl_sync(FULL, (k & 1) ? r.mineB : r.mineA, k >> 1, G);
                if (!MULTI || t.j == 0) {
                    const uint64_t p = row_addr(base0, (uint32_t)t.last, pitch);
#pragma unroll
                    for (int v = 0; v < VPL; v++) ldcg4_into(r.v0[v], p + v * 64, t.act && live[v]);
                }
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) {
                    const uint64_t p = row_addr(base1, (uint32_t)r.tg[k], pitch);
#pragma unroll
                    for (int v = 0; v < VPL; v++) ldcg4_into(r.row[k][v], p + v * 64, r.tg[k] >= 0 && live[v]);
                }
            };
            auto compute = [&](const stage_r &r) {
                if (!__any_sync(FULL, r.act)) return;
                const bool first = !MULTI || r.j == 0;
                if (first) {
                    npairs += r.act;
#pragma unroll
                    for (int v = 0; v < VPL; v++) neu[v] = zero4;
                }
                if (MULTI && first) {
#pragma unroll
                    for (int v = 0; v < VPL; v++) v0p[v] = r.v0[v];
                }
                float4 v0[VPL];
#pragma unroll
                for (int v = 0; v < VPL; v++) v0[v] = MULTI ? v0p[v] : r.v0[v];
                float d[8];
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) {
                    d[k] = dot4(v0[0], r.row[k][0]);
                    if (VPL == 2) d[k] += dot4(v0[1], r.row[k][1]);
                }
                d[5] = 0.f;
                if (first) {
                    d[5] = dot4(v0[0], cur[0]);
                    if (VPL == 2) d[5] += dot4(v0[1], cur[1]);
                }
                d[6] = d[7] = 0.f;
                // transposed reduction, 8 values over 4 lanes: offset 2 (bit 1 clear keeps values 0..3), then offset 1
                float e[4];
#pragma unroll
                for (int j = 0; j < 4; j++) e[j] = (up2 ? d[j + 4] : d[j]) + __shfl_xor_sync(FULL, up2 ? d[j] : d[j + 4], 2);
                const float totA = (up1 ? e[2] : e[0]) + __shfl_xor_sync(FULL, up1 ? e[0] : e[2], 1); // value 2*lane
                const float totB = (up1 ? e[3] : e[1]) + __shfl_xor_sync(FULL, up1 ? e[1] : e[3], 1); // value 2*lane + 1
                float gA = sgns_g_lane(totA, 0.f, alpha, gA_hi, 0.f, s_exp, E, idx_scale);
                float gB = sgns_g_lane(totB, labelB, alpha, gB_hi, gB_lo, s_exp, E, idx_scale);
                if (r.mineA < 0) gA = 0.f;
                const bool okB = kB < SGNS_CH ? r.mineB >= 0 : (kB == SGNS_CH && r.act && first);
                if (!okB) gB = 0.f;
                float gk[SGNS_CH + 1];
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) gk[k] = __shfl_sync(FULL, (k & 1) ? gB : gA, k >> 1, G);
                gk[SGNS_CH] = first ? __shfl_sync(FULL, gB, SGNS_CH >> 1, G) : 0.f;
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) {
                    const uint64_t p = row_addr(base1, (uint32_t)r.tg[k], pitch);
#pragma unroll
                    for (int v = 0; v < VPL; v++) {
                        axpy4(neu[v], gk[k], r.row[k][v]);
                        red_add4_if(p + v * 64, scale4(gk[k], v0[v]), gk[k] != 0.f && live[v] && !(a.dbg & 1));
                    }
                }
                if (first) {
#pragma unroll
                    for (int v = 0; v < VPL; v++) {
                        axpy4(neu[v], gk[SGNS_CH], cur[v]);
                        axpy4(d1[v], gk[SGNS_CH], v0[v]);
                        axpy4(cur[v], gk[SGNS_CH], v0[v]);
                    }
                }
                if (!MULTI || r.j == NCH - 1) { // the pair is complete: syn0[last] += neu
                    const uint64_t p0 = row_addr(base0, (uint32_t)r.last, pitch);
#pragma unroll
                    for (int v = 0; v < VPL; v++) {
                        if (same_s) { // one row for the whole warp (inactive groups carry neu = 0)
                            float4 ns = neu[v];
#pragma unroll
                            for (int o = G; o < 32; o <<= 1) {
                                ns.x += __shfl_xor_sync(FULL, ns.x, o); ns.y += __shfl_xor_sync(FULL, ns.y, o);
                                ns.z += __shfl_xor_sync(FULL, ns.z, o); ns.w += __shfl_xor_sync(FULL, ns.w, o);
                            }
                            red_add4_if(p0 + v * 64, ns, gw == 0 && live[v] && !(a.dbg & 1));
                        } else {
                            red_add4_if(p0 + v * 64, neu[v], r.act && live[v] && !(a.dbg & 1));
                        }
                    }
                }
            };

            const int U = (c_max - c_min + 1) * NCH;
            stage_r rA; // rows not (re)loaded keep stale finite values, cancelled by g = 0; start from zeros
#pragma unroll
            for (int v = 0; v < VPL; v++) {
                rA.v0[v] = zero4;
#pragma unroll
                for (int k = 0; k < SGNS_CH; k++) rA.row[k][v] = zero4;
            }
            stage_t t1 = stageT();
            for (int u = 0; u < U; u++) {
                stageR(t1, rA);
                t1 = stageT();
                compute(rA);
            }
            const uint64_t pw = row_addr(base1, (uint32_t)w1, pitch);
#pragma unroll
            for (int v = 0; v < VPL; v++) red_add4_if(pw + v * 64, d1[v], valid && live[v] && !(a.dbg & 1));
            pairs += (unsigned)npairs;
        }
    }
    if (lane == 0 && pairs) atomicAdd(a.pairs, pairs);
}

// ---------------------------------------------------------------------------------------------------------
// Kernel E: the item kernel for SMALL VOCABULARIES with narrow rows (the reference's own community-area run:
// V = 1 848, D = 8).  There the staleness bound (SGNS_STALE_BOUND * V / (K + 1) pairs in flight) leaves ~4 warps per
// SM, every warp scheduler holds one warp, and the epoch time is  pairs / in-flight pairs x (latency of one pair) --
// kernels C / D spend ~480 dependent-issue slots per pair step (2 800 cycles measured, profiles/r1s15_bench_ca.json).
// This kernel shortens that chain instead of widening the machine: the K + 1 targets of a pair are handled by
// DIFFERENT lanes (target slot ts = 0: the positive target, 1..K: the negatives; NL lanes per target row, one
// 128-bit slot each), so a pair step is ONE row load, ONE dot product, ONE sigmoid lookup and ONE reduction deep,
// and the rows of the next pair step are requested before the current one is computed (software pipeline: table
// lookups two steps ahead, rows one step ahead).  Same draws and arithmetic per target as kernels B-D.
template <int NL>
__global__ void __launch_bounds__(128)
k_sgns_items_tp(const sgns_args a) {
    static_assert(NL == 1 || NL == 2 || NL == 4, "1, 2 or 4 lanes (128-bit slots) per target row");
    constexpr int GP = 8 * NL;   // lanes per item: 8 target slots (1 positive + up to 7 negatives) x NL
    constexpr int GPW = 32 / GP; // items per warp, in lockstep
    extern __shared__ int32_t smem[];
    float *s_exp = reinterpret_cast<float *>(smem);
    constexpr unsigned FULL = 0xffffffffu;
    const int gpw_eff = (a.dbg & 8) ? 1 : GPW; // test mode: one item at a time (n_groups = 1)
    const int lane = threadIdx.x % GP;
    const int gw = (threadIdx.x & 31) / GP;
    const int ts = lane / NL, q = lane % NL;
    int32_t *mytok = smem + a.exp_table_size + (threadIdx.x / GP) * a.Lmax;
    for (int i = threadIdx.x; i < a.exp_table_size; i += blockDim.x) s_exp[i] = a.exp_table[i];
    __syncthreads();
    const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int E = a.exp_table_size;
    const float idx_scale = (float)E / SGNS_MAX_EXP / 2.0f;
    const int win = a.window;
    const int64_t N = a.n_sent;
    const int Lmax = a.Lmax;
    const int64_t item_lo = a.s_lo * Lmax, n_items = a.s_hi * Lmax;
    const double inv_total = 1.0 / (double)((int64_t)a.epochs * a.n_global);
    const uint32_t tsize = (uint32_t)a.neg_table_size, vm1 = (uint32_t)(a.V > 1 ? a.V - 1 : 1);
    const double inv_tsize = 1.0 / (double)tsize, inv_vm1 = 1.0 / (double)vm1;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int K = a.V >= 2 ? a.negative : 0; // <= 7 (host)
    const bool is_pos = ts == 0, is_neg = ts >= 1 && ts <= K;
    const bool live = q < a.n4;
    const float label = is_pos ? 1.f : 0.f;
    const uint32_t pitch = (uint32_t)a.stride * 4u;
    const char *base0 = reinterpret_cast<const char *>(a.syn0) + (live ? q : 0) * 16;
    const char *base1 = reinterpret_cast<const char *>(a.syn1neg) + (live ? q : 0) * 16;
    const uint64_t my_a = a.lcg_a[is_neg ? ts - 1 : 0], my_c = a.lcg_c[is_neg ? ts - 1 : 0]; // negative ts-1 of the pair
    const bool reds_on = !(a.dbg & 1);
    unsigned long long pairs = 0;

    struct stage_t { int32_t last; bool act; uint64_t nsk; int32_t traw; };
    struct stage_r { int32_t last; bool act; int32_t mine; float4 row, v0; };

    for (int ep = a.ep_lo; ep < a.ep_hi; ep++) {
        for (int64_t base = item_lo + warp_id * gpw_eff; base < n_items; base += a.n_groups) { // warp-uniform trip count
            const int64_t item = base + gw;
            bool valid = item < n_items && gw < gpw_eff;
            const int64_t s = valid ? item / Lmax : 0;
            const int i = valid ? (int)(item - s * Lmax) : 0;
            __syncwarp();
            int n_tok = 0;
            for (int j = lane; j < Lmax; j += GP) { const int32_t tk = a.wtok[(int64_t)j * N + s]; mytok[j] = tk; n_tok += tk >= 0; }
            __syncwarp();
#pragma unroll
            for (int o = GP >> 1; o > 0; o >>= 1) n_tok += __shfl_xor_sync(FULL, n_tok, o);
            const int32_t w1 = mytok[i];
            valid = valid && w1 >= 0;
            if (!__any_sync(FULL, valid)) continue;
            float alpha = a.lr * (float)(1.0 - (double)((int64_t)ep * a.n_global + a.s_off + s) * inv_total);
            if (alpha < a.min_lr) alpha = a.min_lr;
            const float g_hi = (label - 1.f) * alpha, g_lo = label * alpha; // saturated sigmoid: dot > 6, dot < -6
            const uint64_t S = sgns_sentence_rng(a.seed, ep, s + a.s_off);
            const int b = (int32_t)(uint32_t)sgns_position_rng(S, i) % win;
            const int lo = valid ? i - win + b : 1, hi = valid ? i + win - b : 0; // inclusive context range; empty if invalid
            const int c_min = __reduce_min_sync(FULL, valid ? max(lo, 0) : Lmax);
            const int c_max = __reduce_max_sync(FULL, valid ? min(hi, n_tok - 1) : -1);
            if (c_max < c_min) continue;
            float4 cur = zero4, d1 = zero4; // positive-slot lanes: private copy of syn1neg[w1] and its accumulated delta
            ldcg4_into(cur, row_addr(base1, (uint32_t)w1, pitch), valid && live && is_pos);
            int npairs = 0;
            int cT = c_min;
            uint64_t hc = 0; // pair hash of context position hcb * GP + lane
            int hcb = -1;

            auto stageT = [&]() { // next pair step: which context, and this lane's negative-table entry
                stage_t t;
                t.last = cT < Lmax ? mytok[cT] : -1;
                t.act = cT >= lo && cT <= hi && cT != i && t.last >= 0 && t.last != w1;
                if (cT / GP != hcb) { hcb = cT / GP; hc = sgns_pair_rng(S, i, hcb * GP + lane); } // warp-uniform condition
                const uint64_t ns0 = shfl64(hc, cT & (GP - 1), GP);
                t.nsk = my_a * ns0 + my_c; // the LCG is affine: state after ts steps
                t.traw = -2;               // "draws nothing"
                if (is_neg && t.act) t.traw = a.neg_table[mod48(t.nsk >> 16, tsize, inv_tsize)];
                cT++;
                return t;
            };
            auto stageR = [&](const stage_t &t, stage_r &r) { // resolve the negative, request this lane's rows
                r.last = t.last; r.act = t.act;
                int32_t tt = t.traw;
                const bool redraw = tt != -2 && (tt <= 0 || tt >= a.V); // DL4J: target = r % (V-1) + 1
                if (__any_sync(FULL, redraw)) {
                    if (redraw) tt = (int32_t)mod64(t.nsk, vm1, inv_vm1) + 1;
                }
                r.mine = (tt != -2 && tt != w1) ? tt : -1;
                ldcg4_into(r.v0, row_addr(base0, (uint32_t)t.last, pitch), t.act && live);     // same row in all 8 slots: one sector
                ldcg4_into(r.row, row_addr(base1, (uint32_t)r.mine, pitch), r.mine >= 0 && live);
            };
            auto compute = [&](const stage_r &r) {
                if (!__any_sync(FULL, r.act)) return;
                npairs += r.act;
                const float4 rowv = is_pos ? cur : r.row;
                float dot = live ? dot4(r.v0, rowv) : 0.f;
#pragma unroll
                for (int o = NL >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);
                float g = sgns_g_lane(dot, label, alpha, g_hi, g_lo, s_exp, E, idx_scale);
                if (!(is_pos ? r.act : r.mine >= 0)) g = 0.f; // idle slots, skipped negatives, inactive items
                const float4 upd = scale4(g, r.v0);            // target row += g * syn0[last]
                red_add4_if(row_addr(base1, (uint32_t)r.mine, pitch), upd, g != 0.f && live && !is_pos && reds_on);
                float4 ns = scale4(g, rowv);                   // this target's share of neu1e
                if (is_pos) { axpy4(d1, 1.f, upd); axpy4(cur, 1.f, upd); }
#pragma unroll
                for (int o = NL; o < GP; o <<= 1) {
                    ns.x += __shfl_xor_sync(FULL, ns.x, o); ns.y += __shfl_xor_sync(FULL, ns.y, o);
                    ns.z += __shfl_xor_sync(FULL, ns.z, o); ns.w += __shfl_xor_sync(FULL, ns.w, o);
                }
                red_add4_if(row_addr(base0, (uint32_t)r.last, pitch), ns, is_pos && r.act && live && reds_on); // syn0[last] += neu1e
            };

            const int U = c_max - c_min + 1;
            stage_r rA, rB; // rows not (re)loaded keep stale finite values, cancelled by g = 0; start from zeros
            rA.v0 = rA.row = rB.v0 = rB.row = zero4;
            // table entries one pair step ahead, rows one pair step ahead of their use.  (Requesting the table entries
            // two steps ahead measured the same 2.6 G pairs/s on the CA workload, profiles/logs/gpurun_out_session17.log:
            // the pair step is bound by its own dependent instruction chain, ~250 issue slots at ~7 cycles each.)
            stage_t t1 = stageT();
            stageR(t1, rA); // rows of step 0
            t1 = stageT();  // table entry of step 1
            for (int u = 0; u < U; u += 2) {
                stageR(t1, rB); // rows of step u + 1 (no-ops past the end: act is false there)
                t1 = stageT();
                compute(rA);
                if (u + 1 < U) {
                    stageR(t1, rA);
                    t1 = stageT();
                    compute(rB);
                }
            }
            red_add4_if(row_addr(base1, (uint32_t)w1, pitch), d1, valid && live && is_pos && reds_on);
            pairs += (unsigned)npairs;
        }
    }
    if (lane == 0 && pairs) atomicAdd(a.pairs, pairs);
}

// ---------------------------------------------------------------------------------------------------------
// Kernel J: kernel G's wavefront with the block split into CRITICAL and HELPER warps.  What bounds kernel G
// (profiles/r2s13_sgns_block_tract24.json): parity caps the sentences in flight at two blocks per SM, a sentence is a
// chain of 2 n - 3 barrier-separated rounds, and a round costs ~1 200 cycles because the warp that owns a centre runs ~330
// instructions in order between two barriers although only a third of them lie on the dependency path
// (pending delta -> K + 1 dot products -> sigmoid -> neu1e -> pending delta); the rest stages the next round (addresses,
// L2 loads) and sends the reductions.  Here that rest is done by a second set of warps:
//   * helper warp h serves the four centres of critical warp h.  In round u it sends the negative-row reductions of round
//     u - 1 (the critical warp leaves each pair's K gradient scales and the context row it used in shared memory), the
//     context-row delta that became final, and requests the rows of round u + ST - 1 with cp.async (LDGSTS, L2 only) into a
//     ring of ST stages in shared memory -- no register staging, L2 latency hidden over ST - 1 rounds;
//   * the critical warp reads its pair's K + 1 rows from the ring, runs the dependency path and nothing else.
// One block barrier per round, as before; same pair / negative enumeration, same wavefront order, same flush points as
// kernel G (a negative-row reduction leaves one round later).  Rows of up to 8 slots, K <= 5, sentences of up to 32 tokens.
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 2)
k_sgns_duo(const sgns_args a) {
    constexpr int G = 8;
    constexpr int KM = SGNS_CH;
    extern __shared__ __align__(16) int32_t smem_j[];
    constexpr unsigned FULL = 0xffffffffu;
    const int n4 = a.n4, Lmax = a.Lmax, ST = a.stages;
    const int nwords = (a.neg_table_size + 31) >> 5;
    const int K = a.V >= 2 ? a.negative : 0;
    const bool smem_neg = a.neg_bits != nullptr;
    const int ROWS = KM + 1;                                  // rows of a pair in the ring: the context row, then the negatives
    float4 *stage = reinterpret_cast<float4 *>(smem_j);       // [ST][Lmax][ROWS][n4]
    float4 *delta = stage + (size_t)ST * Lmax * ROWS * n4;    // [Lmax][n4] pending syn0 updates of the sentence's context rows
    float4 *xv = delta + Lmax * n4;                           // [2][Lmax][n4] the context row a pair used (for the helper's reductions)
    float *xg = reinterpret_cast<float *>(xv + 2 * Lmax * n4); // [2][Lmax][8] its K gradient scales
    float *s_exp = xg + 2 * Lmax * 8;
    int32_t *tok = reinterpret_cast<int32_t *>(s_exp + a.exp_table_size);
    int32_t *s_lo = tok + Lmax, *s_hi = s_lo + Lmax;
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(s_hi + Lmax); // [2 Lmax] centres with a pair in round u
    uint32_t *s_bits = s_mask + 2 * Lmax;
    uint32_t *s_pref = s_bits + nwords;
    int32_t *s_tg = reinterpret_cast<int32_t *>(s_bits + (smem_neg ? 2 * nwords : 0)); // [Lmax][Lmax][K] negatives of every pair
    const int tid = threadIdx.x;
    const int NW = (blockDim.x >> 5) >> 1;                    // critical warps = helper warps
    const bool helper = (tid >> 5) >= NW;
    const int rt = helper ? tid - NW * 32 : tid;              // thread index within the role
    const int lane = rt % G, i = rt / G;                      // slot of the row; centre position served
    for (int q = tid; q < a.exp_table_size; q += blockDim.x) s_exp[q] = a.exp_table[q];
    if (smem_neg)
        for (int q = tid; q < 2 * nwords; q += blockDim.x) s_bits[q] = a.neg_bits[q];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = tid; q < ST * Lmax * ROWS * n4; q += blockDim.x) stage[q] = zero4; // the ring only ever holds table rows afterwards
    const int E = a.exp_table_size;
    const float idx_scale = (float)E / SGNS_MAX_EXP / 2.0f;
    const int win = a.window;
    const int64_t N = a.n_sent;
    const double inv_total = 1.0 / (double)((int64_t)a.epochs * a.n_global);
    const uint32_t tsize = (uint32_t)a.neg_table_size, vm1 = (uint32_t)(a.V > 1 ? a.V - 1 : 1);
    const double inv_tsize = 1.0 / (double)tsize, inv_vm1 = 1.0 / (double)vm1;
    const bool live = lane < n4;
    const uint32_t pitch = (uint32_t)a.stride * 4u;
    const char *base0 = reinterpret_cast<const char *>(a.syn0) + (live ? lane : 0) * 16;
    const char *base1 = reinterpret_cast<const char *>(a.syn1neg) + (live ? lane : 0) * 16;
    const int L8 = lane & 7;
    const bool up4 = (lane & 4) != 0, up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
    const float my_label = L8 == KM ? 1.f : 0.f;
    const bool reds_on = !(a.dbg & 1);
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    unsigned long long pairs = 0;

    for (int ep = a.ep_lo; ep < a.ep_hi; ep++) {
        int32_t tk_next = -1;
        if (tid < Lmax && a.s_lo + blockIdx.x < a.s_hi) tk_next = a.wtok[(int64_t)tid * N + a.s_lo + blockIdx.x];
        for (int64_t s = a.s_lo + blockIdx.x; s < a.s_hi; s += a.n_groups) { // n_groups = blocks = sentences in flight
            __syncthreads(); // the previous sentence's final flush has read the caches
            const int32_t tk = tk_next;
            if (tid < Lmax) {
                tok[tid] = tk;
                if (s + a.n_groups < a.s_hi) tk_next = a.wtok[(int64_t)tid * N + s + a.n_groups]; // lands during this sentence's rounds
            }
            for (int e = tid; e < Lmax * n4; e += blockDim.x) delta[e] = zero4;
            const int n_tok = __syncthreads_count(tid < Lmax && tk >= 0); // the compacted sentence: tokens first, then padding
            if (n_tok < 2) continue;
            float alpha = a.lr * (float)(1.0 - (double)((int64_t)ep * a.n_global + a.s_off + s) * inv_total);
            if (alpha < a.min_lr) alpha = a.min_lr;
            const float g_hi = (my_label - 1.f) * alpha, g_lo = my_label * alpha;
            const uint64_t S = sgns_sentence_rng(a.seed, ep, s + a.s_off);
            const int R = 2 * n_tok - 3;
            const bool valid = i < n_tok;
            const int32_t w1 = valid ? tok[i] : 0;
            float4 cur = zero4, d1 = zero4;
            if (!helper) ldcg4_into(cur, row_addr(base1, (uint32_t)w1, pitch), valid && live); // private copy of syn1neg[w1]
            if (tid < n_tok) { // the centre's window (word2vec's random shrink), clamped to the sentence
                const int b = (int32_t)(uint32_t)sgns_position_rng(S, tid) % win;
                s_lo[tid] = max(tid - win + b, 0);
                s_hi[tid] = min(tid + win - b, n_tok - 1);
            }
            __syncthreads();
            // ---- draw phase: the K negatives of every pair inside a window; the centres of every round
            for (int e = tid; e < n_tok * n_tok * K; e += blockDim.x) {
                const int kq = e % K, ic = e / K;
                const int cc = ic % n_tok, ii = ic / n_tok;
                if (cc == ii || cc < s_lo[ii] || cc > s_hi[ii]) continue;
                const uint64_t nsk = a.lcg_a[kq] * sgns_pair_rng(S, ii, cc) + a.lcg_c[kq]; // the LCG is affine: state after kq + 1 steps
                const uint32_t idx = mod48(nsk >> 16, tsize, inv_tsize);
                int32_t tg = smem_neg ? neg_lookup(s_bits, s_pref, idx) : a.neg_table[idx];
                if (tg <= 0 || tg >= a.V) tg = (int32_t)mod64(nsk, vm1, inv_vm1) + 1;   // DL4J: target = r % (V - 1) + 1
                s_tg[(ii * Lmax + cc) * K + kq] = tg == tok[ii] ? -1 : tg;
            }
            if (tid >= 1 && tid <= R + 1) { // round u = tid: centre ii meets context u - ii
                uint32_t m = 0;
                if (tid <= R)
                    for (int ii = max(0, tid - (n_tok - 1)); ii <= min(n_tok - 1, tid); ii++) {
                        const int c = tid - ii;
                        if (c != ii && c >= s_lo[ii] && c <= s_hi[ii] && tok[c] != tok[ii]) m |= 1u << ii;
                    }
                s_mask[tid] = m; // round R + 1 is empty
            }
            __syncthreads();
            const int fr = valid ? i + s_hi[i] + 1 : 0; // the round after the centre's last context: its output-row delta is sent then

            if (!helper) {
                // ================= critical warps: the dependency path of the rounds =================
                int npairs = 0;
                uint32_t m = s_mask[1];
                int su = 1 % ST; // u % ST
                for (int u = 1; u <= R; u++, su = (su + 1 == ST ? 0 : su + 1)) {
                    // what does not depend on round u - 1 is read before the barrier
                    const bool act = valid && ((m >> i) & 1u);
                    const int c = act ? u - i : 0;
                    const int32_t mine = (act && L8 < K) ? s_tg[(i * Lmax + c) * K + L8] : -1;
                    const float4 *st = stage + ((size_t)(su * Lmax + (valid ? i : 0)) * ROWS) * n4 + (live ? lane : 0);
                    m = s_mask[u + 1];
                    asm volatile("bar.sync 0;" ::: "memory");
                    if (u == fr) red_add4_if(row_addr(base1, (uint32_t)w1, pitch), d1, live && reds_on);
                    if (!__any_sync(FULL, act)) continue;
                    float4 v0 = zero4, row[KM];
#pragma unroll
                    for (int k = 0; k < KM; k++) row[k] = zero4;
                    float4 dl = zero4;
                    if (live) {
                        v0 = st[0];
#pragma unroll
                        for (int k = 0; k < KM; k++) row[k] = st[(k + 1) * n4];
                        dl = delta[c * n4 + lane];
                    }
                    npairs += act;
                    const float4 v0p = add4(v0, dl); // L2's value + what this sentence has added to the row so far
                    float d0 = dot4(v0p, row[0]), d1v = dot4(v0p, row[1]), d2 = dot4(v0p, row[2]), d3 = dot4(v0p, row[3]);
                    float d4 = dot4(v0p, row[4]), d5 = dot4(v0p, cur);
                    float e0 = (up4 ? d4 : d0) + __shfl_xor_sync(FULL, up4 ? d0 : d4, 4);
                    float e1 = (up4 ? d5 : d1v) + __shfl_xor_sync(FULL, up4 ? d1v : d5, 4);
                    float e2 = (up4 ? 0.f : d2) + __shfl_xor_sync(FULL, up4 ? d2 : 0.f, 4);
                    float e3 = (up4 ? 0.f : d3) + __shfl_xor_sync(FULL, up4 ? d3 : 0.f, 4);
                    float f0 = (up2 ? e2 : e0) + __shfl_xor_sync(FULL, up2 ? e0 : e2, 2);
                    float f1 = (up2 ? e3 : e1) + __shfl_xor_sync(FULL, up2 ? e1 : e3, 2);
                    float tot = (up1 ? f1 : f0) + __shfl_xor_sync(FULL, up1 ? f0 : f1, 1);
                    float g = sgns_g_lane(tot, my_label, alpha, g_hi, g_lo, s_exp, E, idx_scale);
                    {
                        const bool mine_ok = L8 < KM ? mine >= 0 : (L8 == KM && act);
                        if (!mine_ok) g = 0.f;
                    }
                    float gk[KM + 1];
#pragma unroll
                    for (int k = 0; k <= KM; k++) gk[k] = __shfl_sync(FULL, g, k, G);
                    float4 neu = scale4(gk[KM], cur);
#pragma unroll
                    for (int k = 0; k < KM; k++) axpy4(neu, gk[k], row[k]);
                    if (act && live) delta[c * n4 + lane] = add4(dl, neu); // syn0[last] += neu1e, pending in the block's cache
                    axpy4(d1, gk[KM], v0p);
                    axpy4(cur, gk[KM], v0p);
                    // for the helper: the scales of the K negative rows and the context row they multiply
                    if (valid && L8 < KM) xg[((u & 1) * Lmax + i) * 8 + L8] = g;
                    if (valid && live) xv[((u & 1) * Lmax + i) * n4 + lane] = v0p;
                }
                asm volatile("bar.sync 0;" ::: "memory");
                if (fr > R) red_add4_if(row_addr(base1, (uint32_t)w1, pitch), d1, valid && live && reds_on);
                pairs += (unsigned)npairs;
            } else {
                // ================= helper warps: staging and reductions, off the dependency path =================
                int sq = 1 % ST; // ring slot of the next round to request
                auto stage_round = [&](int u) { // request the rows of the pairs of round u (called for u = 1, 2, ... in order)
                    const bool act = valid && u <= R && ((s_mask[u] >> i) & 1u);
                    const int c = act ? u - i : 0;
                    const uint32_t dst = stage_s + (uint32_t)(((sq * Lmax + (valid ? i : 0)) * ROWS * n4 + (live ? lane : 0)) * 16);
                    sq = sq + 1 == ST ? 0 : sq + 1;
                    cp_async16_if(dst, row_addr(base0, (uint32_t)tok[c], pitch), act && live);
                    const int32_t *tgp = s_tg + (i * Lmax + c) * K;
#pragma unroll
                    for (int k = 0; k < KM; k++) {
                        const int32_t tg = (act && k < K) ? tgp[k] : -1;
                        cp_async16_if(dst + (uint32_t)((k + 1) * n4 * 16), row_addr(base1, (uint32_t)max(tg, 0), pitch), tg >= 0 && live);
                    }
                    cp_async_commit();
                };
                auto send_round = [&](int u) { // the negative-row reductions of round u, from what the critical warp left
                    const bool act = valid && ((s_mask[u] >> i) & 1u);
                    if (!__any_sync(FULL, act)) return;
                    const int c = act ? u - i : 0;
                    const int32_t *tgp = s_tg + (i * Lmax + c) * K;
                    const float *gp = xg + ((u & 1) * Lmax + (valid ? i : 0)) * 8;
                    float4 v = zero4;
                    if (valid && live) v = xv[((u & 1) * Lmax + i) * n4 + lane];
#pragma unroll
                    for (int k = 0; k < KM; k++) {
                        const int32_t tg = (act && k < K) ? tgp[k] : -1;
                        const float gv = act ? gp[k] : 0.f;
                        red_add4_if(row_addr(base1, (uint32_t)max(tg, 0), pitch), scale4(gv, v), tg >= 0 && gv != 0.f && live && reds_on);
                    }
                };
                for (int u = 1; u < ST; u++) stage_round(u);
                for (int u = 1; u <= R; u++) {
                    if (ST == 4) cp_async_wait<2>(); else if (ST == 3) cp_async_wait<1>(); else cp_async_wait<0>(); // round u has landed
                    asm volatile("bar.sync 0;" ::: "memory");
                    if (u > 1) send_round(u - 1);
                    const int cf = u - n_tok; // context row cf saw its last centre in round cf + n_tok - 1 at the latest
                    if (rt < n4 && cf >= 0) {
                        const float4 dl = delta[cf * n4 + rt];
                        if (reds_on && (dl.x != 0.f || dl.y != 0.f || dl.z != 0.f || dl.w != 0.f))
                            red_add4(reinterpret_cast<float4 *>(a.syn0 + (int64_t)tok[cf] * a.stride) + rt, dl);
                    }
                    stage_round(u + ST - 1);
                }
                cp_async_wait<0>();
                asm volatile("bar.sync 0;" ::: "memory");
                send_round(R);
                // the context rows whose last centre came in the final rounds
                for (int e = rt; e < n_tok * 8; e += NW * 32) {
                    const int row = e >> 3, slot = e & 7;
                    if (slot >= n4 || !reds_on || row + n_tok <= R) continue;
                    const float4 dl = delta[row * n4 + slot];
                    if (dl.x != 0.f || dl.y != 0.f || dl.z != 0.f || dl.w != 0.f)
                        red_add4(reinterpret_cast<float4 *>(a.syn0 + (int64_t)tok[row] * a.stride) + slot, dl);
                }
            }
        }
    }
    if (lane == 0 && pairs) atomicAdd(a.pairs, pairs);
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
    dge_free(m->ctx, m->syn0); dge_free(m->ctx, m->syn1neg);
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
    bool ok = dge_malloc(ctx, &m->syn0, nel) == cudaSuccess && dge_malloc(ctx, &m->syn1neg, nel) == cudaSuccess;
    if (!multi) { t0 = m->syn0; t1 = m->syn1neg; }
    ok = ok && dge_malloc(ctx, &m->id_of_word, (size_t)V) == cudaSuccess && tmp.get(&d_word_of_id, (size_t)n_ids) == cudaSuccess &&
         tmp.get(&d_table, (size_t)p->neg_table_size) == cudaSuccess && tmp.get(&d_exp, (size_t)p->exp_table_size) == cudaSuccess &&
         tmp.get(&d_pairs, 2) == cudaSuccess && (!train || tmp.get(&d_wtok, (size_t)n_sent * (size_t)Lmax) == cudaSuccess);
    local = ok ? DGE_OK : dge_fail(ctx, DGE_E_CUDA, std::string("dge_sgns_train: cudaMalloc failed (") + cudaGetErrorString(cudaGetLastError()) + ")");
    if (multi) local = dge_comm_agree(ctx, local, "dge_sgns_train (tables)");
    if (local != DGE_OK) { model_release(m); return local; }
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
