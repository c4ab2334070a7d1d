// Skip-gram (stage 2): kernel arguments, the draw definitions shared with oracle/sgns_oracle.c, the vocabulary / corpus
// preparation kernels and the small device helpers every training kernel uses.  Included by sgns.cu only.
#pragma once
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

