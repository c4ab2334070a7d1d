// Skip-gram training kernels A-E and C': the exact-order kernel and the round-1 ITEM kernels (work item = (sentence, centre)).
// Kernel A is the parity reference on the device (concurrency = 1); B-E stay for A/B runs (DGE_SGNS_F_ITEM_KERNELS) and for rows
// beyond 32 slots.  Included by sgns.cu after sgns_common.cuh.
#pragma once
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

