/*
 * sgns_oracle.c -- CPU ORACLE for stage 2.  TEST INFRASTRUCTURE ONLY (see sgns_oracle.h).
 * PARITY UNPINNED: restates the published word2vec skip-gram algorithm with the DL4J 0.7.2
 * parameterisation of SURVEY.md 8(a) A14; the DL4J sources are not under /root/reference.
 * Call sites it stands in for: DeepWalk.java:73-76 (builder), :79 (fit).
 */
#include "sgns_oracle.h"
#include "dge_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

#define MAX_EXP 6.0f
#define LCG_MUL 25214903917ULL
#define LCG_ADD 11ULL

struct ora_vocab {
    int32_t n_ids, V;
    int64_t total;
    int32_t *word_of_id; /* n_ids */
    int32_t *id_of_word; /* V */
    int64_t *count;      /* V */
};

typedef struct { int64_t c; int32_t id; } cnt_id;
static int cmp_cnt_desc(const void *a, const void *b) {
    const cnt_id *x = (const cnt_id *)a, *y = (const cnt_id *)b;
    if (x->c != y->c) return x->c > y->c ? -1 : 1;
    return x->id < y->id ? -1 : (x->id > y->id);
}

ora_vocab *ora_vocab_build(const int32_t *tokens, int64_t n_tokens, int32_t n_ids, int32_t min_count) {
    ora_vocab *v = (ora_vocab *)calloc(1, sizeof(*v));
    v->n_ids = n_ids;
    int64_t *cnt = (int64_t *)calloc((size_t)(n_ids ? n_ids : 1), sizeof(int64_t));
    for (int64_t i = 0; i < n_tokens; i++)
        if (tokens[i] >= 0) cnt[tokens[i]]++;
    cnt_id *arr = (cnt_id *)malloc(sizeof(cnt_id) * (size_t)(n_ids ? n_ids : 1));
    int32_t V = 0;
    for (int32_t i = 0; i < n_ids; i++)
        if (cnt[i] >= min_count && cnt[i] > 0) { arr[V].c = cnt[i]; arr[V].id = i; V++; }
    qsort(arr, (size_t)V, sizeof(cnt_id), cmp_cnt_desc);
    v->V = V;
    v->word_of_id = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_ids ? n_ids : 1));
    v->id_of_word = (int32_t *)malloc(sizeof(int32_t) * (size_t)(V ? V : 1));
    v->count = (int64_t *)malloc(sizeof(int64_t) * (size_t)(V ? V : 1));
    for (int32_t i = 0; i < n_ids; i++) v->word_of_id[i] = -1;
    for (int32_t w = 0; w < V; w++) {
        v->word_of_id[arr[w].id] = w;
        v->id_of_word[w] = arr[w].id;
        v->count[w] = arr[w].c;
        v->total += arr[w].c;
    }
    free(arr); free(cnt);
    return v;
}
void ora_vocab_free(ora_vocab *v) {
    if (!v) return;
    free(v->word_of_id); free(v->id_of_word); free(v->count); free(v);
}
int32_t ora_vocab_size(const ora_vocab *v) { return v->V; }
int64_t ora_vocab_total_words(const ora_vocab *v) { return v->total; }
void ora_vocab_tables(const ora_vocab *v, int32_t *word_of_id, int32_t *id_of_word, int64_t *count) {
    if (word_of_id) memcpy(word_of_id, v->word_of_id, sizeof(int32_t) * (size_t)v->n_ids);
    if (id_of_word) memcpy(id_of_word, v->id_of_word, sizeof(int32_t) * (size_t)v->V);
    if (count) memcpy(count, v->count, sizeof(int64_t) * (size_t)v->V);
}

/* unigram^0.75 cumulative table (word2vec.c InitUnigramTable; DL4J makeTable). */
void ora_neg_table(const ora_vocab *v, int32_t table_size, int32_t *table) {
    int32_t V = v->V;
    if (V == 0) { for (int32_t i = 0; i < table_size; i++) table[i] = 0; return; }
    double pow_sum = 0;
    for (int32_t w = 0; w < V; w++) pow_sum += pow((double)v->count[w], 0.75);
    int32_t wi = 0;
    double d1 = pow((double)v->count[0], 0.75) / pow_sum;
    for (int32_t i = 0; i < table_size; i++) {
        table[i] = wi;
        if ((double)i / (double)table_size > d1) {
            if (wi < V - 1) wi++;
            d1 += pow((double)v->count[wi], 0.75) / pow_sum;
        }
    }
}

void ora_init_syn0(int32_t V, int32_t dim, uint64_t seed, float *syn0) {
    /* (U[0,1) - 0.5) / dim ; element e = row*dim + col is draw e of a Philox stream. */
    for (int64_t e = 0; e < (int64_t)V * dim; e++) {
        uint32_t ctr[4] = {(uint32_t)(e >> 2), (uint32_t)((uint64_t)(e >> 2) >> 32), 0x5347u, 0u};
        uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        uint32_t r[4];
        ora_philox4x32_10(ctr, key, r);
        float u = (float)(r[e & 3] >> 8) * 0x1.0p-24f;
        syn0[e] = (u - 0.5f) / (float)dim;
    }
}

uint64_t ora_sentence_rng(uint64_t seed, int32_t epoch, int64_t sentence) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(sentence + 1) + 0xD1B54A32D192ED03ULL * (uint64_t)epoch;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return z & 0x7FFFFFFFFFFFFFFFULL;
}

float ora_alpha(const ora_sgns_params *p, int32_t epoch, int64_t sentence, int64_t n_sent) {
    double progress = (double)((int64_t)epoch * n_sent + sentence) / (double)((int64_t)p->epochs * n_sent);
    float a = p->lr * (float)(1.0 - progress);
    return a < p->min_lr ? p->min_lr : a;
}

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* DL4J draws the window shrink b and the negative-sampling seed of each pair from a per-thread LCG
 * chain (nextRandom), whose values depend on thread scheduling and are not reproducible.  The
 * restatement keeps the same distributions but makes every draw a pure function of
 * (seed, epoch, sentence, position[, context]) so that any schedule enumerates the same pairs:
 *   position i : r = 63-bit uniform -> b = ((int) r) % window            (SkipGram.learnSequence)
 *   pair (i,c) : 64-bit initial state of the aggregate's negative-sampling LCG (AggregateSkipGram) */
uint64_t ora_position_rng(uint64_t S, int32_t i) {
    return mix64(S + 0x9E3779B97F4A7C15ULL * (uint64_t)(i + 1)) & 0x7FFFFFFFFFFFFFFFULL;
}
uint64_t ora_pair_rng(uint64_t S, int32_t i, int32_t c) {
    return mix64(S ^ (0xD6E8FEB86659FD93ULL * (uint64_t)((int64_t)i * 65536 + c + 1)));
}

/* Huffman tree (word2vec.c CreateBinaryTree): codes/points per word, for the optional HS rounds. */
typedef struct { int32_t *points; int8_t *codes; int32_t *len; int32_t max_len; } huff;
static void huff_build(const int64_t *count, int32_t V, huff *h) {
    h->max_len = 40;
    h->points = (int32_t *)calloc((size_t)V * 40, sizeof(int32_t));
    h->codes = (int8_t *)calloc((size_t)V * 40, 1);
    h->len = (int32_t *)calloc((size_t)V, sizeof(int32_t));
    if (V < 2) return;
    int64_t *cnt = (int64_t *)malloc(sizeof(int64_t) * (size_t)(2 * V + 1));
    int8_t *bin = (int8_t *)calloc((size_t)(2 * V + 1), 1);
    int32_t *parent = (int32_t *)calloc((size_t)(2 * V + 1), sizeof(int32_t));
    for (int32_t a = 0; a < V; a++) cnt[a] = count[a];
    for (int32_t a = V; a < 2 * V; a++) cnt[a] = (int64_t)1e15;
    int32_t pos1 = V - 1, pos2 = V;
    for (int32_t a = 0; a < V - 1; a++) {
        int32_t m1, m2;
        if (pos1 >= 0 && cnt[pos1] < cnt[pos2]) m1 = pos1--; else m1 = pos2++;
        if (pos1 >= 0 && cnt[pos1] < cnt[pos2]) m2 = pos1--; else m2 = pos2++;
        cnt[V + a] = cnt[m1] + cnt[m2];
        parent[m1] = V + a; parent[m2] = V + a;
        bin[m2] = 1;
    }
    for (int32_t a = 0; a < V; a++) {
        int8_t code[40]; int32_t point[40];
        int32_t b = a, i = 0;
        while (1) {
            code[i] = bin[b]; point[i] = b; i++;
            b = parent[b];
            if (b == V * 2 - 2) break;
        }
        h->len[a] = i;
        h->points[(int64_t)a * 40] = V - 2;
        for (b = 0; b < i; b++) {
            h->codes[(int64_t)a * 40 + i - b - 1] = code[b];
            h->points[(int64_t)a * 40 + i - b] = point[b] - V;
        }
    }
    free(cnt); free(bin); free(parent);
}
static void huff_free(huff *h) { free(h->points); free(h->codes); free(h->len); }

struct ora_model {
    int32_t V, dim;
    float *syn0, *syn1neg, *syn1;
    int32_t *id_of_word;
};

typedef struct {
    const int32_t *tokens; int64_t n_sent; int32_t L;
    const ora_sgns_params *p;
    const ora_vocab *vocab;
    const int32_t *table; const float *exp_table;
    const huff *hf;
    ora_model *m;
    int tid, nthreads;
    int count_only;
    int64_t pairs;
    int64_t s_lo, s_hi; /* sentences [s_lo, s_hi) of this call (s_hi == 0: all) */
    int32_t only_epoch;  /* >= 0: this epoch only (data-parallel emulation runs an epoch slice by slice) */
    int64_t sent_offset; /* data-parallel shard: global index of local sentence 0 (RNG keys and the learning-rate */
    int64_t n_total;     /* schedule use GLOBAL sentence indices, so the shards enumerate exactly the pairs and  */
                         /* negatives of the single-process run over the whole corpus); 0 = n_sent              */
} worker;

static void *worker_run(void *arg) {
    worker *w = (worker *)arg;
    const ora_sgns_params *p = w->p;
    const int32_t D = p->dim, V = w->vocab->V, E = p->exp_table_size, T = p->neg_table_size;
    const int32_t win = p->window;
    float *neu1e = (float *)malloc(sizeof(float) * (size_t)(D ? D : 1));
    int32_t *sent = (int32_t *)malloc(sizeof(int32_t) * (size_t)(w->L ? w->L : 1));
    float *syn0 = w->m ? w->m->syn0 : NULL, *syn1neg = w->m ? w->m->syn1neg : NULL, *syn1 = w->m ? w->m->syn1 : NULL;
    const float idx_scale = (float)E / MAX_EXP / 2.0f;
    int64_t pairs = 0;
    const int64_t s_lo = w->s_hi > 0 ? w->s_lo : 0, s_hi = w->s_hi > 0 ? w->s_hi : w->n_sent;
    for (int32_t ep = 0; ep < p->epochs; ep++) {
        if (w->only_epoch >= 0 && ep != w->only_epoch) continue;
        for (int64_t s = s_lo + w->tid; s < s_hi; s += w->nthreads) {
            int32_t n = 0;
            for (int32_t j = 0; j < w->L; j++) {
                int32_t id = w->tokens[s * w->L + j];
                if (id < 0) continue;
                int32_t wd = w->vocab->word_of_id[id];
                if (wd >= 0) sent[n++] = wd;
            }
            float alpha = ora_alpha(p, ep, s + w->sent_offset, w->n_total > 0 ? w->n_total : w->n_sent);
            const uint64_t S = ora_sentence_rng(p->seed, ep, s + w->sent_offset);
            for (int32_t i = 0; i < n; i++) {
                uint64_t r = ora_position_rng(S, i);
                int32_t b = (int32_t)(uint32_t)r % win; /* ((int) nextRandom) % window, may be negative */
                int32_t w1 = sent[i];
                int32_t end = win * 2 + 1 - b;
                for (int32_t a = b; a < end; a++) {
                    if (a == win) continue;
                    int32_t c = i - win + a;
                    if (c < 0 || c >= n) continue;
                    int32_t last = sent[c];
                    if (last == w1) continue;
                    uint64_t ns = ora_pair_rng(S, i, c);
                    pairs++;
                    if (w->count_only) continue;
                    float *v0 = syn0 + (int64_t)last * D;
                    for (int32_t d = 0; d < D; d++) neu1e[d] = 0;
                    if (p->use_hs) {
                        for (int32_t q = 0; q < w->hf->len[w1]; q++) {
                            int32_t pt = w->hf->points[(int64_t)w1 * 40 + q];
                            int32_t code = w->hf->codes[(int64_t)w1 * 40 + q];
                            if (pt < 0 || pt >= V) continue;
                            float *v1 = syn1 + (int64_t)pt * D;
                            float dot = 0;
                            for (int32_t d = 0; d < D; d++) dot += v0[d] * v1[d];
                            if (dot < -MAX_EXP || dot >= MAX_EXP) continue;
                            int32_t idx = (int32_t)((dot + MAX_EXP) * idx_scale);
                            if (idx < 0 || idx >= E) continue;
                            float g = (1.0f - (float)code - w->exp_table[idx]) * alpha;
                            for (int32_t d = 0; d < D; d++) neu1e[d] += g * v1[d];
                            for (int32_t d = 0; d < D; d++) v1[d] += g * v0[d];
                        }
                    }
                    for (int32_t k = 0; k < p->negative + 1; k++) {
                        int32_t target, label;
                        if (k == 0) { target = w1; label = 1; }
                        else {
                            if (V < 2) break;
                            ns = ns * LCG_MUL + LCG_ADD;
                            target = w->table[(ns >> 16) % (uint64_t)T];
                            if (target <= 0 || target >= V) target = (int32_t)(ns % (uint64_t)(V - 1)) + 1;
                            if (target == w1) continue;
                            label = 0;
                        }
                        float *v1 = syn1neg + (int64_t)target * D;
                        float dot = 0;
                        for (int32_t d = 0; d < D; d++) dot += v0[d] * v1[d];
                        float g;
                        if (dot > MAX_EXP) g = ((float)label - 1.0f) * alpha;
                        else if (dot < -MAX_EXP) g = ((float)label - 0.0f) * alpha;
                        else {
                            int32_t idx = (int32_t)((dot + MAX_EXP) * idx_scale);
                            if (idx < 0 || idx >= E) continue;
                            g = ((float)label - w->exp_table[idx]) * alpha;
                        }
                        for (int32_t d = 0; d < D; d++) neu1e[d] += g * v1[d];
                        for (int32_t d = 0; d < D; d++) v1[d] += g * v0[d];
                    }
                    for (int32_t d = 0; d < D; d++) v0[d] += neu1e[d];
                }
            }
        }
    }
    w->pairs = pairs;
    free(neu1e); free(sent);
    return NULL;
}

static int64_t run_workers(const int32_t *tokens, int64_t n_sent, int32_t L, const ora_sgns_params *p,
                           const ora_vocab *vocab, const int32_t *table, const float *exp_table,
                           const huff *hf, ora_model *m, int count_only) {
    int nt = p->threads > 0 ? p->threads : 1;
    if (nt > 256) nt = 256;
    worker *ws = (worker *)calloc((size_t)nt, sizeof(worker));
    pthread_t *th = (pthread_t *)calloc((size_t)nt, sizeof(pthread_t));
    for (int t = 0; t < nt; t++) {
        ws[t].tokens = tokens; ws[t].n_sent = n_sent; ws[t].L = L; ws[t].p = p; ws[t].vocab = vocab;
        ws[t].table = table; ws[t].exp_table = exp_table; ws[t].hf = hf; ws[t].m = m;
        ws[t].tid = t; ws[t].nthreads = nt; ws[t].count_only = count_only;
        ws[t].s_lo = 0; ws[t].s_hi = 0; ws[t].only_epoch = -1;
    }
    if (nt == 1) worker_run(&ws[0]);
    else {
        for (int t = 0; t < nt; t++) pthread_create(&th[t], NULL, worker_run, &ws[t]);
        for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
    }
    int64_t pairs = 0;
    for (int t = 0; t < nt; t++) pairs += ws[t].pairs;
    free(ws); free(th);
    return pairs;
}

ora_model *ora_sgns_train(const int32_t *tokens, int64_t n_sent, int32_t L, int32_t n_ids,
                          const ora_sgns_params *p, int64_t *pairs_out) {
    ora_vocab *vocab = ora_vocab_build(tokens, n_sent * L, n_ids, p->min_count);
    int32_t V = vocab->V, D = p->dim;
    ora_model *m = (ora_model *)calloc(1, sizeof(*m));
    m->V = V; m->dim = D;
    size_t n = (size_t)(V ? V : 1) * (size_t)D;
    m->syn0 = (float *)malloc(sizeof(float) * n);
    m->syn1neg = (float *)calloc(n, sizeof(float));
    m->syn1 = p->use_hs ? (float *)calloc(n, sizeof(float)) : NULL;
    m->id_of_word = (int32_t *)malloc(sizeof(int32_t) * (size_t)(V ? V : 1));
    memcpy(m->id_of_word, vocab->id_of_word, sizeof(int32_t) * (size_t)V);
    ora_init_syn0(V, D, p->seed, m->syn0);
    int32_t *table = (int32_t *)malloc(sizeof(int32_t) * (size_t)p->neg_table_size);
    ora_neg_table(vocab, p->neg_table_size, table);
    float *exp_table = (float *)malloc(sizeof(float) * (size_t)p->exp_table_size);
    for (int32_t i = 0; i < p->exp_table_size; i++) {
        double e = exp(((double)i / (double)p->exp_table_size * 2.0 - 1.0) * (double)MAX_EXP);
        exp_table[i] = (float)(e / (e + 1.0));
    }
    huff hf; memset(&hf, 0, sizeof(hf));
    if (p->use_hs) huff_build(vocab->count, V, &hf);
    int64_t pairs = run_workers(tokens, n_sent, L, p, vocab, table, exp_table, &hf, m, 0);
    if (pairs_out) *pairs_out = pairs;
    if (p->use_hs) huff_free(&hf);
    free(table); free(exp_table);
    ora_vocab_free(vocab);
    return m;
}

/* Data-parallel emulation (checker of the multi-GPU scheme of libdge's stage 2; the reference itself is single
 * process, its only hint being "model averaging", DeepWalk.java:43).  `world` ranks own contiguous shards of the
 * sentences (embedding_b200/parallel.py walk_shard), share one vocabulary built from the whole corpus, and keep
 * replicas of syn0 / syn1neg.  Every epoch is cut into `rounds` slices; in a slice each rank trains sequentially on
 * its own sentences (GLOBAL sentence indices for the RNG and the learning-rate schedule, as on the GPUs: the shards
 * together enumerate exactly the pairs and negatives of the sequential run), then the replicas are recombined from the
 * per-rank deltas d_r = cur_r - base, per ROW:
 *   combine 0: base += sum_r d_r                      (all updates applied)
 *   combine 1: base += mean_r d_r                     (parameter averaging)
 *   combine 2: base += sum_r d_r / max(1, #ranks whose delta of that row is non-zero)   (average over contributors)
 *   combine 3: base += sum_r d_r / sqrt(max(1, #contributors))
 *   combine 4: base += sum_r d_r / max(1, |sum_r d_r|^2 / sum_r |d_r|^2)                (alignment-weighted)
 *              deltas that point the same way (every rank pushed the row towards the same optimum: adding them
 *              overshoots) are averaged, deltas that are orthogonal (independent information) are summed; the
 *              divisor lies in [1, #contributors].
 * combine + 16 (DELAYED): the combined delta of slice t reaches the replicas one slice late, while they already
 * trained slice t + 1 from their own state (the exchange overlapped with compute): replica_r += combined_t - d_r,t.
 */
static float combine_div(int combine, int world, int touched, double sq_of_sum, double sum_of_sq) {
    const float c = (float)(touched > 1 ? touched : 1);
    switch (combine) {
        case 1: return (float)world;
        case 2: return c;
        case 3: return sqrtf(c);
        case 4: { if (!(sum_of_sq > 0.0)) return 1.0f; float a = (float)(sq_of_sum / sum_of_sq); return a > 1.0f ? a : 1.0f; }
        default: return 1.0f;
    }
}

ora_model *ora_sgns_train_dp(const int32_t *tokens, int64_t n_sent, int32_t L, int32_t n_ids,
                             const ora_sgns_params *p, int32_t world, int32_t rounds, int32_t combine,
                             int64_t *pairs_out) {
    return ora_sgns_train_dp_tiered(tokens, n_sent, L, n_ids, p, world, rounds, combine, 1, 0, pairs_out);
}

/* Tiered exchange: the vocabulary is sorted by descending count, so the rows that several ranks touch within a slice
 * (the ones whose deltas interfere) are a contiguous PREFIX of both tables.  `rounds` exchanges per epoch; every
 * exchange covers rows [0, hot_rows); every `full_every`-th one (and the last) covers all rows.  full_every = 1 is
 * the plain scheme. */
ora_model *ora_sgns_train_dp_tiered(const int32_t *tokens, int64_t n_sent, int32_t L, int32_t n_ids,
                                    const ora_sgns_params *p, int32_t world, int32_t rounds, int32_t combine,
                                    int32_t full_every, int32_t hot_rows, int64_t *pairs_out) {
    const int delayed = (combine & 16) != 0;
    combine &= 15;
    ora_vocab *vocab = ora_vocab_build(tokens, n_sent * L, n_ids, p->min_count);
    int32_t V = vocab->V, D = p->dim;
    size_t n = (size_t)(V ? V : 1) * (size_t)D;
    ora_model *m = (ora_model *)calloc(1, sizeof(*m));
    m->V = V; m->dim = D;
    m->syn0 = (float *)malloc(sizeof(float) * n);
    m->syn1neg = (float *)calloc(n, sizeof(float));
    m->syn1 = NULL;
    m->id_of_word = (int32_t *)malloc(sizeof(int32_t) * (size_t)(V ? V : 1));
    memcpy(m->id_of_word, vocab->id_of_word, sizeof(int32_t) * (size_t)V);
    ora_init_syn0(V, D, p->seed, m->syn0);
    int32_t *table = (int32_t *)malloc(sizeof(int32_t) * (size_t)p->neg_table_size);
    ora_neg_table(vocab, p->neg_table_size, table);
    float *exp_table = (float *)malloc(sizeof(float) * (size_t)p->exp_table_size);
    for (int32_t i = 0; i < p->exp_table_size; i++) {
        double e = exp(((double)i / (double)p->exp_table_size * 2.0 - 1.0) * (double)MAX_EXP);
        exp_table[i] = (float)(e / (e + 1.0));
    }
    if (world < 1) world = 1;
    if (rounds < 1) rounds = 1;
    /* per rank: the replica (cur) and its base = the replica minus the local progress not yet handed to an exchange */
    ora_model *rep = (ora_model *)calloc((size_t)world, sizeof(ora_model));
    float **bas[2], **own[2], **pown[2];
    for (int t = 0; t < 2; t++) {
        bas[t] = (float **)calloc((size_t)world, sizeof(float *));
        own[t] = (float **)calloc((size_t)world, sizeof(float *));
        pown[t] = (float **)calloc((size_t)world, sizeof(float *));
    }
    for (int r = 0; r < world; r++) {
        rep[r].V = V; rep[r].dim = D;
        rep[r].syn0 = (float *)malloc(sizeof(float) * n);
        rep[r].syn1neg = (float *)malloc(sizeof(float) * n);
        memcpy(rep[r].syn0, m->syn0, sizeof(float) * n);
        memcpy(rep[r].syn1neg, m->syn1neg, sizeof(float) * n);
        for (int t = 0; t < 2; t++) {
            bas[t][r] = (float *)malloc(sizeof(float) * n);
            memcpy(bas[t][r], t == 0 ? m->syn0 : m->syn1neg, sizeof(float) * n);
            own[t][r] = (float *)malloc(sizeof(float) * n);
            pown[t][r] = (float *)calloc(n, sizeof(float));
        }
    }
    float *acc = (float *)malloc(sizeof(float) * n);      /* combined delta of the newest exchange */
    float *pend[2] = {(float *)calloc(n, sizeof(float)), (float *)calloc(n, sizeof(float))}; /* delayed: in flight */
    int have_pending = 0;
    int32_t *touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)(V ? V : 1));
    double *sumsq = (double *)malloc(sizeof(double) * (size_t)(V ? V : 1));
    int64_t pairs = 0;
    ora_sgns_params q = *p;
    q.use_hs = 0;
    for (int32_t ep = 0; ep < p->epochs; ep++) {
        for (int32_t rd = 0; rd < rounds; rd++) {
            for (int r = 0; r < world; r++) {
                int64_t base = n_sent / world, rem = n_sent % world;
                int64_t first = r * base + (r < rem ? r : rem), count = base + (r < rem ? 1 : 0);
                worker w; memset(&w, 0, sizeof(w));
                w.tokens = tokens + first * L; w.n_sent = count; w.L = L; w.p = &q; w.vocab = vocab;
                w.table = table; w.exp_table = exp_table; w.hf = NULL; w.m = &rep[r];
                w.tid = 0; w.nthreads = 1; w.count_only = 0; w.only_epoch = ep;
                w.sent_offset = first; w.n_total = n_sent;
                w.s_lo = count * rd / rounds; w.s_hi = count * (rd + 1) / rounds;
                if (w.s_hi > w.s_lo) { worker_run(&w); pairs += w.pairs; }
            }
            const int last = ep == p->epochs - 1 && rd == rounds - 1;
            const int full = full_every <= 1 || last || (rd + 1) % full_every == 0;
            const int32_t Vx = full ? V : (hot_rows < V ? hot_rows : V);   /* rows [0, Vx) take part in this exchange */
            for (int t = 0; t < 2; t++) {
                /* a delayed exchange lands now, one slice late: replica and base += combined - own */
                if (have_pending) {
                    for (int r = 0; r < world; r++) {
                        float *cur = t == 0 ? rep[r].syn0 : rep[r].syn1neg;
                        for (size_t i = 0; i < n; i++) { const float c = pend[t][i] - pown[t][r][i]; cur[i] += c; bas[t][r][i] += c; }
                    }
                }
                /* snapshot: own = cur - base (the local progress of this slice); base = cur */
                memset(acc, 0, sizeof(float) * n);
                memset(touched, 0, sizeof(int32_t) * (size_t)(V ? V : 1));
                memset(sumsq, 0, sizeof(double) * (size_t)(V ? V : 1));
                for (int r = 0; r < world; r++) {
                    const float *cur = t == 0 ? rep[r].syn0 : rep[r].syn1neg;
                    for (int32_t v = 0; v < Vx; v++) {
                        int nz = 0;
                        double sq = 0.0;
                        for (int32_t d = 0; d < D; d++) {
                            const size_t i = (size_t)v * D + d;
                            const float dl = cur[i] - bas[t][r][i];
                            own[t][r][i] = dl;
                            bas[t][r][i] = cur[i];
                            acc[i] += dl;
                            sq += (double)dl * (double)dl;
                            nz |= dl != 0.0f;
                        }
                        touched[v] += nz;
                        sumsq[v] += sq;
                    }
                }
                for (int32_t v = 0; v < Vx; v++) {
                    double sq_of_sum = 0.0;
                    for (int32_t d = 0; d < D; d++) sq_of_sum += (double)acc[(size_t)v * D + d] * (double)acc[(size_t)v * D + d];
                    const float div = combine_div(combine, world, touched[v], sq_of_sum, sumsq[v]);
                    for (int32_t d = 0; d < D; d++) acc[(size_t)v * D + d] /= div;
                }
                const size_t nx = (size_t)Vx * (size_t)D;
                if (delayed && !last) { /* in flight during the next slice */
                    memset(pend[t], 0, sizeof(float) * n);
                    memcpy(pend[t], acc, sizeof(float) * nx);
                    for (int r = 0; r < world; r++) { memset(pown[t][r], 0, sizeof(float) * n); memcpy(pown[t][r], own[t][r], sizeof(float) * nx); }
                } else {
                    for (int r = 0; r < world; r++) {
                        float *cur = t == 0 ? rep[r].syn0 : rep[r].syn1neg;
                        for (size_t i = 0; i < nx; i++) { const float c = acc[i] - own[t][r][i]; cur[i] += c; bas[t][r][i] += c; }
                    }
                }
            }
            have_pending = delayed && !last;
        }
    }
    /* the replicas agree up to fp32 rounding of (cur += combined - own); rank 0's copy is the result (the GPUs broadcast it) */
    memcpy(m->syn0, rep[0].syn0, sizeof(float) * n);
    memcpy(m->syn1neg, rep[0].syn1neg, sizeof(float) * n);
    if (pairs_out) *pairs_out = pairs;
    for (int r = 0; r < world; r++) {
        free(rep[r].syn0); free(rep[r].syn1neg);
        for (int t = 0; t < 2; t++) { free(bas[t][r]); free(own[t][r]); free(pown[t][r]); }
    }
    for (int t = 0; t < 2; t++) { free(bas[t]); free(own[t]); free(pown[t]); }
    free(rep); free(pend[0]); free(pend[1]);
    free(acc); free(touched); free(sumsq); free(table); free(exp_table);
    ora_vocab_free(vocab);
    return m;
}

int64_t ora_sgns_count_pairs(const int32_t *tokens, int64_t n_sent, int32_t L, int32_t n_ids,
                             const ora_sgns_params *p) {
    ora_vocab *vocab = ora_vocab_build(tokens, n_sent * L, n_ids, p->min_count);
    int64_t pairs = run_workers(tokens, n_sent, L, p, vocab, NULL, NULL, NULL, NULL, 1);
    ora_vocab_free(vocab);
    return pairs;
}

void ora_model_free(ora_model *m) {
    if (!m) return;
    free(m->syn0); free(m->syn1neg); free(m->syn1); free(m->id_of_word); free(m);
}
int32_t ora_model_vocab_size(const ora_model *m) { return m->V; }
void ora_model_get(const ora_model *m, float *syn0, float *syn1neg, int32_t *id_of_word) {
    size_t n = (size_t)m->V * (size_t)m->dim;
    if (syn0) memcpy(syn0, m->syn0, sizeof(float) * n);
    if (syn1neg) memcpy(syn1neg, m->syn1neg, sizeof(float) * n);
    if (id_of_word) memcpy(id_of_word, m->id_of_word, sizeof(int32_t) * (size_t)m->V);
}
