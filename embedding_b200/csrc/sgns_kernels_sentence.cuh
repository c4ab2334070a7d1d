// Skip-gram training kernels F-J: the SENTENCE-RESIDENT kernels (a warp or a block owns a sentence; intra-sentence updates in
// the reference's order).  Kernel F is what the automatic schedule runs; G is its fallback; H, I, J are measured experiments
// behind flags.  Included by sgns.cu after sgns_kernels_items.cuh.
#pragma once
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
