/*
 * dge_oracle.c -- CPU ORACLE for stage 1 (graph core, alias tables, walks).
 * TEST INFRASTRUCTURE ONLY: see dge_oracle.h.  Never linked into the product.
 *
 * Restates (does not copy) the algorithms of the reference's
 *   LayeredGraph.java, CrossTimeGraph.java, SpatialGraph.java,
 *   CommunityAreas.java:240-245, Tracts.java:477-482
 * with integer vertex ids instead of String names (the host owns naming).
 * All floating point is IEEE-754 double, one rounding per Java operator:
 * compile with -ffp-contract=off and without -ffast-math.
 */
#include "dge_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>

/* ------------------------------------------------------------------ alias */

/* LayeredGraph.java:59-62: probTable[i] = k * w / outDegree  (int*double, then divide). */
static void alias_init(int32_t k, const double *w, double out_degree, double *prob, int32_t *alias) {
    for (int32_t i = 0; i < k; i++) {
        double kw = (double)k * w[i];
        prob[i] = kw / out_degree;
        alias[i] = -1;
    }
}

/* LayeredGraph.java:65-81 (same loop again at :209-225 for the source table). */
static void alias_pair_literal(int32_t k, double *prob, int32_t *alias) {
    for (int32_t l1 = 0; l1 < k; l1++) {
        if (prob[l1] != 1.0 && alias[l1] == -1) {
            for (int32_t l2 = 0; l2 < k; l2++) {
                if (l2 != l1 && alias[l2] == -1) {
                    if (prob[l1] > 1.0 && prob[l2] < 1.0) {
                        alias[l2] = l1;
                        double d = 1 - prob[l2];
                        prob[l1] -= d;
                    } else if (prob[l1] < 1.0 && prob[l2] > 1.0) {
                        alias[l1] = l2;
                        double d = 1 - prob[l1];
                        prob[l2] -= d;
                        break; /* l1 is exactly full */
                    }
                }
            }
        }
    }
}

void ora_alias_literal(int32_t k, const double *w, double out_degree, double *prob, int32_t *alias) {
    alias_init(k, w, out_degree, prob, alias);
    alias_pair_literal(k, prob, alias);
}

/* Hierarchical bitmap: ordered set of ints in [0,n) with successor queries. */
typedef struct {
    int levels;
    uint64_t *lv[8];
    int64_t nw[8];
} hset;

static void hset_init(hset *h, int64_t n) {
    h->levels = 0;
    int64_t bits = n > 0 ? n : 1;
    do {
        int64_t words = (bits + 63) >> 6;
        h->lv[h->levels] = (uint64_t *)calloc((size_t)words, sizeof(uint64_t));
        h->nw[h->levels] = words;
        h->levels++;
        bits = words;
    } while (bits > 1);
}
static void hset_free(hset *h) {
    for (int l = 0; l < h->levels; l++) free(h->lv[l]);
}
static void hset_add(hset *h, int64_t i) {
    for (int l = 0; l < h->levels; l++) {
        h->lv[l][i >> 6] |= 1ULL << (i & 63);
        i >>= 6;
    }
}
static void hset_del(hset *h, int64_t i) {
    for (int l = 0; l < h->levels; l++) {
        uint64_t *wp = &h->lv[l][i >> 6];
        *wp &= ~(1ULL << (i & 63));
        if (*wp) break;
        i >>= 6;
    }
}
/* smallest member >= p, or -1 */
static int64_t hset_succ(const hset *h, int64_t p) {
    int64_t pos = p;
    for (int l = 0; l < h->levels; l++) {
        int64_t wi = pos >> 6;
        if (wi >= h->nw[l]) return -1;
        uint64_t m = h->lv[l][wi] & (~0ULL << (pos & 63));
        if (m) {
            int64_t idx = (wi << 6) + __builtin_ctzll(m);
            for (int d = l - 1; d >= 0; d--) idx = (idx << 6) + __builtin_ctzll(h->lv[d][idx]);
            return idx;
        }
        pos = wi + 1;
    }
    return -1;
}

/* Ordered-set form of LayeredGraph.java:65-81.  Invariants that make it exact:
 *  - an entry is "small" (prob<1), "large" (prob>1) or inert (==1.0, NaN, or alias set);
 *    small entries never become large; large entries only shrink; alias is written once.
 *  - a large l1 visits l2 ascending from 0 and absorbs every small unassigned l2 while
 *    prob[l1] > 1.0 (:69-71); no entry joins the small set during that scan, so "next l2"
 *    is "minimum of the small set".
 *  - once prob[l1] < 1.0 at scan position q, the scan continues from q+1 and stops at the
 *    first large unassigned l2 (:72-76).
 *  - a small l1 scans from 0 to the first large l2 (:72-76).
 * The sequence of double subtractions is identical, hence bit-identical tables. */
static void alias_pair_fast(int32_t k, double *prob, int32_t *alias) {
    hset S, G;
    hset_init(&S, k);
    hset_init(&G, k);
    for (int32_t i = 0; i < k; i++) {
        if (prob[i] < 1.0) hset_add(&S, i);
        else if (prob[i] > 1.0) hset_add(&G, i);
    }
    for (int32_t l1 = 0; l1 < k; l1++) {
        if (!(prob[l1] != 1.0 && alias[l1] == -1)) continue;
        if (prob[l1] > 1.0) {
            double p1 = prob[l1];
            int64_t pos = -1;
            int exhausted = 0;
            while (p1 > 1.0) {
                int64_t l2 = hset_succ(&S, 0);
                if (l2 < 0) { exhausted = 1; break; }
                alias[l2] = l1;
                double d = 1 - prob[l2];
                p1 -= d;
                hset_del(&S, l2);
                pos = l2;
            }
            prob[l1] = p1;
            if (exhausted) continue; /* still large; inner scan ran to the end */
            hset_del(&G, l1);
            if (p1 < 1.0) {
                int64_t l2 = hset_succ(&G, pos + 1);
                if (l2 >= 0) {
                    alias[l1] = (int32_t)l2;
                    double d = 1 - p1;
                    prob[l2] -= d;
                    if (!(prob[l2] > 1.0)) {
                        hset_del(&G, l2);
                        if (prob[l2] < 1.0) hset_add(&S, l2);
                    }
                } else {
                    hset_add(&S, l1); /* dangling small; a later large may take it */
                }
            }
        } else if (prob[l1] < 1.0) {
            int64_t l2 = hset_succ(&G, 0);
            if (l2 >= 0) {
                alias[l1] = (int32_t)l2;
                double d = 1 - prob[l1];
                prob[l2] -= d;
                hset_del(&S, l1);
                if (!(prob[l2] > 1.0)) {
                    hset_del(&G, l2);
                    if (prob[l2] < 1.0) hset_add(&S, l2);
                }
            }
        }
        /* NaN: passes the != 1.0 test but no comparison fires */
    }
    hset_free(&S);
    hset_free(&G);
}

void ora_alias_fast(int32_t k, const double *w, double out_degree, double *prob, int32_t *alias) {
    alias_init(k, w, out_degree, prob, alias);
    alias_pair_fast(k, prob, alias);
}

/* ------------------------------------------------------------------ graph */

struct ora_graph {
    int32_t nv, ns;
    int64_t ne;
    int64_t *row_ptr;
    int32_t *col;
    double *w, *prob;
    int32_t *alias;
    double *out_degree;
    int32_t *sources;
    double *src_prob;
    int32_t *src_alias;
    double source_weight_sum;
};

ora_graph *ora_graph_build(int32_t nv, int64_t ne, const int32_t *src, const int32_t *dst, const double *w,
                           int32_t ns, const int32_t *sources, const double *out_degree_override,
                           const double *source_weight_sum_override, int alias_mode) {
    ora_graph *g = (ora_graph *)calloc(1, sizeof(*g));
    g->nv = nv; g->ne = ne; g->ns = ns;
    g->row_ptr = (int64_t *)calloc((size_t)nv + 1, sizeof(int64_t));
    g->col = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ne ? ne : 1));
    g->w = (double *)malloc(sizeof(double) * (size_t)(ne ? ne : 1));
    g->prob = (double *)malloc(sizeof(double) * (size_t)(ne ? ne : 1));
    g->alias = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ne ? ne : 1));
    g->out_degree = (double *)calloc((size_t)(nv ? nv : 1), sizeof(double));
    g->sources = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ns ? ns : 1));
    g->src_prob = (double *)malloc(sizeof(double) * (size_t)(ns ? ns : 1));
    g->src_alias = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ns ? ns : 1));

    /* addEdge :157-174 -> Vertex.addOutEdge :46-49: per-vertex list in insertion order,
     * outDegree += weight in that order.  Stable counting sort by source == the lists. */
    for (int64_t e = 0; e < ne; e++) g->row_ptr[src[e] + 1]++;
    for (int32_t v = 0; v < nv; v++) g->row_ptr[v + 1] += g->row_ptr[v];
    int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nv ? nv : 1));
    memcpy(cur, g->row_ptr, sizeof(int64_t) * (size_t)nv);
    for (int64_t e = 0; e < ne; e++) {
        int64_t p = cur[src[e]]++;
        g->col[p] = dst[e];
        g->w[p] = w[e];
        g->out_degree[src[e]] += w[e];
    }
    free(cur);
    if (out_degree_override) memcpy(g->out_degree, out_degree_override, sizeof(double) * (size_t)nv);

    /* initiateAliasTables :197 */
    for (int32_t v = 0; v < nv; v++) {
        int64_t b = g->row_ptr[v];
        int32_t k = (int32_t)(g->row_ptr[v + 1] - b);
        if (alias_mode == ORA_ALIAS_FAST) ora_alias_fast(k, g->w + b, g->out_degree[v], g->prob + b, g->alias + b);
        else ora_alias_literal(k, g->w + b, g->out_degree[v], g->prob + b, g->alias + b);
    }
    /* addSourceVertex :180-189 then the source table :199-225 */
    double sws = 0;
    double *sw = (double *)malloc(sizeof(double) * (size_t)(ns ? ns : 1));
    for (int32_t i = 0; i < ns; i++) {
        g->sources[i] = sources[i];
        sw[i] = g->out_degree[sources[i]];
        sws += sw[i];
    }
    if (source_weight_sum_override) sws = *source_weight_sum_override;
    g->source_weight_sum = sws;
    if (alias_mode == ORA_ALIAS_FAST) ora_alias_fast(ns, sw, sws, g->src_prob, g->src_alias);
    else ora_alias_literal(ns, sw, sws, g->src_prob, g->src_alias);
    free(sw);
    return g;
}

void ora_graph_free(ora_graph *g) {
    if (!g) return;
    free(g->row_ptr); free(g->col); free(g->w); free(g->prob); free(g->alias);
    free(g->out_degree); free(g->sources); free(g->src_prob); free(g->src_alias);
    free(g);
}
int64_t ora_graph_num_edges(const ora_graph *g) { return g->ne; }
int32_t ora_graph_num_vertices(const ora_graph *g) { return g->nv; }
int32_t ora_graph_num_sources(const ora_graph *g) { return g->ns; }

void ora_graph_tables(const ora_graph *g, int64_t *row_ptr, int32_t *col, double *w, double *prob,
                      int32_t *alias, double *out_degree, double *src_prob, int32_t *src_alias,
                      double *source_weight_sum) {
    if (row_ptr) memcpy(row_ptr, g->row_ptr, sizeof(int64_t) * ((size_t)g->nv + 1));
    if (col) memcpy(col, g->col, sizeof(int32_t) * (size_t)g->ne);
    if (w) memcpy(w, g->w, sizeof(double) * (size_t)g->ne);
    if (prob) memcpy(prob, g->prob, sizeof(double) * (size_t)g->ne);
    if (alias) memcpy(alias, g->alias, sizeof(int32_t) * (size_t)g->ne);
    if (out_degree) memcpy(out_degree, g->out_degree, sizeof(double) * (size_t)g->nv);
    if (src_prob) memcpy(src_prob, g->src_prob, sizeof(double) * (size_t)g->ns);
    if (src_alias) memcpy(src_alias, g->src_alias, sizeof(int32_t) * (size_t)g->ns);
    if (source_weight_sum) *source_weight_sum = g->source_weight_sum;
}

/* ---------------------------------------------------------------- sampling */

/* LayeredGraph.java:107-115 / :124-131: one uniform drives column and coin. */
static inline int32_t alias_draw(int32_t k, const double *prob, const int32_t *alias, double x) {
    double xk = x * (double)k;
    int32_t i = (int32_t)xk;
    double y = xk - (double)i;
    if (y < prob[i]) return i;
    int32_t a = alias[i];
    return a < 0 ? i : a; /* Java: ArrayList.get(-1) throws; we define "no alias" = self */
}

int32_t ora_sample_next(const ora_graph *g, int32_t v, double x) {
    int64_t b = g->row_ptr[v];
    int32_t k = (int32_t)(g->row_ptr[v + 1] - b);
    if (k == 0) return -1;
    return g->col[b + alias_draw(k, g->prob + b, g->alias + b, x)];
}

/* LayeredGraph.java:89-98 */
int32_t ora_sample_next_ov(const ora_graph *g, int32_t v, double x) {
    int64_t b = g->row_ptr[v], e = g->row_ptr[v + 1];
    double s = x * g->out_degree[v];
    double cnt = 0;
    for (int64_t j = b; j < e; j++) {
        cnt += g->w[j];
        if (cnt >= s) return g->col[j];
    }
    return -1;
}

int32_t ora_sample_source(const ora_graph *g, double x, int sampler) {
    if (g->ns == 0) return -1;
    if (sampler == ORA_SAMPLER_ALIAS) /* :234-242 */
        return g->sources[alias_draw(g->ns, g->src_prob, g->src_alias, x)];
    /* :261-270 */
    double s = x * g->source_weight_sum;
    double cnt = 0;
    for (int32_t i = 0; i < g->ns; i++) {
        cnt += g->out_degree[g->sources[i]];
        if (cnt >= s) return g->sources[i];
    }
    return -1;
}

/* -------------------------------------------------------------------- RNGs */

static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

/* Philox4x32-10 (Salmon et al., SC'11), the published counter-based generator. */
void ora_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; r++) {
        if (r) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
        philox_round(c, k);
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* Draw number `draw` of walk `walk_id`: block = draw/2 yields two 53-bit uniforms, the same
 * (hi<<32|lo)>>11 * 2^-53 grid that java.util.Random.nextDouble() lands on. */
double ora_philox_uniform(uint64_t seed, uint64_t walk_id, uint32_t draw) {
    uint32_t ctr[4] = {(uint32_t)walk_id, (uint32_t)(walk_id >> 32), draw >> 1, 0u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t r[4];
    ora_philox4x32_10(ctr, key, r);
    uint64_t bits = (draw & 1) ? (((uint64_t)r[3] << 32) | r[2]) : (((uint64_t)r[1] << 32) | r[0]);
    return (double)(bits >> 11) * 0x1.0p-53;
}

/* java.util.Random (JDK): 48-bit LCG; nextDouble = ((next(26)<<27)+next(27)) * 2^-53. */
typedef struct { uint64_t s; } jrandom;
static void jr_seed(jrandom *r, int64_t seed) { r->s = ((uint64_t)seed ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); }
static inline uint32_t jr_next(jrandom *r, int bits) {
    r->s = (r->s * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (uint32_t)(r->s >> (48 - bits));
}
static inline double jr_double(jrandom *r) {
    uint64_t a = jr_next(r, 26);
    uint64_t b = jr_next(r, 27);
    return (double)((a << 27) + b) * 0x1.0p-53;
}
void ora_java_random_doubles(int64_t seed, int32_t n, double *out) {
    jrandom r;
    jr_seed(&r, seed);
    for (int32_t i = 0; i < n; i++) out[i] = jr_double(&r);
}

/* -------------------------------------------------------------------- walks */

void ora_walk(const ora_graph *g, int64_t n_walks, int64_t first_walk_id, int32_t L, uint64_t seed,
              int sampler, int rng_kind, int32_t *tokens) {
    jrandom jr;
    jr_seed(&jr, (int64_t)seed);
    for (int64_t i = 0; i < n_walks; i++) {
        int32_t *seq = tokens + i * (int64_t)L;
        uint64_t wid = (uint64_t)(first_walk_id + i);
        uint32_t draw = 0;
        int32_t len = 0;
        for (int32_t j = 0; j < L; j++) seq[j] = -1;
        if (L <= 0) continue;
        double x = rng_kind == ORA_RNG_PHILOX ? ora_philox_uniform(seed, wid, draw++) : jr_double(&jr);
        int32_t v = ora_sample_source(g, x, sampler);
        if (v < 0) continue;
        seq[len++] = v;
        while (len < L) { /* :244-250 */
            int64_t b = g->row_ptr[v];
            int32_t k = (int32_t)(g->row_ptr[v + 1] - b);
            if (k == 0) break; /* sampleNextVertex returns null before drawing :105-107 */
            x = rng_kind == ORA_RNG_PHILOX ? ora_philox_uniform(seed, wid, draw++) : jr_double(&jr);
            int32_t nn = sampler == ORA_SAMPLER_ALIAS ? ora_sample_next(g, v, x) : ora_sample_next_ov(g, v, x);
            if (nn < 0) break;
            seq[len++] = nn;
            v = nn;
        }
    }
}

/* ------------------------------------------------------------ spatial graph */

/* JDK 8 DoubleStream.sum(): Collectors.sumWithCompensation + computeFinalSum (which in
 * JDK 8 ADDS the compensation term); SpatialGraph.java:34,57 go through it. */
double ora_java8_stream_sum(const double *v, int64_t n) {
    double sum = 0, comp = 0, simple = 0;
    for (int64_t i = 0; i < n; i++) {
        double tmp = v[i] - comp;
        double velvel = sum + tmp;
        comp = (velvel - sum) - tmp;
        sum = velvel;
        simple += v[i];
    }
    double tmp = sum + comp;
    if (isnan(tmp) && isinf(simple)) return simple;
    return tmp;
}

typedef struct { double w; int32_t c; } wcol;
static void stable_sort_desc(wcol *a, wcol *tmp, int32_t n) {
    /* bottom-up merge sort: stable, like List.sort (TimSort) with -Double.compare */
    for (int32_t width = 1; width < n; width *= 2) {
        for (int32_t lo = 0; lo < n; lo += 2 * width) {
            int32_t mid = lo + width < n ? lo + width : n;
            int32_t hi = lo + 2 * width < n ? lo + 2 * width : n;
            int32_t i = lo, j = mid, o = lo;
            while (i < mid && j < hi) {
                /* take right only if strictly greater (Double.compare order on non-NaN, no -0.0 here) */
                if (a[j].w > a[i].w) tmp[o++] = a[j++]; else tmp[o++] = a[i++];
            }
            while (i < mid) tmp[o++] = a[i++];
            while (j < hi) tmp[o++] = a[j++];
        }
        memcpy(a, tmp, sizeof(wcol) * (size_t)n);
    }
}

void ora_keep_nearest_k(int32_t n, const double *w, int32_t k, int32_t *col, double *wk, double *out_degree) {
    wcol *a = (wcol *)malloc(sizeof(wcol) * (size_t)n);
    wcol *t = (wcol *)malloc(sizeof(wcol) * (size_t)n);
    for (int32_t r = 0; r < n; r++) {
        for (int32_t c = 0; c < n; c++) { a[c].w = w[(int64_t)r * n + c]; a[c].c = c; }
        stable_sort_desc(a, t, n);
        for (int32_t j = 0; j < k; j++) { col[(int64_t)r * k + j] = a[j].c; wk[(int64_t)r * k + j] = a[j].w; }
        out_degree[r] = ora_java8_stream_sum(wk + (int64_t)r * k, k);
    }
    free(a); free(t);
}

/* --------------------------------------------------------- cross-time graph */

int32_t ora_flow_ca(const int32_t *F, int32_t n, int32_t src, int32_t dst, int32_t lo, int32_t hi) {
    int32_t cnt = 0;
    for (int32_t h = lo; h != hi; h = (h + 1) % 24) cnt += F[((int64_t)src * 24 + h) * n + dst];
    return cnt;
}
int32_t ora_flow_tract(const int32_t *F, int32_t n, int32_t src, int32_t dst, int32_t lo, int32_t hi) {
    int32_t cnt = 0;
    for (int32_t h = lo; h <= hi; h++) cnt += F[((int64_t)src * 24 + h) * n + dst];
    return cnt;
}

int64_t ora_crosstime_edges(const int32_t *F, int32_t n, const int32_t *order, int32_t L, int mode,
                            const int32_t *intervals, int32_t *src, int32_t *dst, double *w,
                            int32_t *v_layer, int32_t *v_region, int32_t *n_vertices_out,
                            int32_t *sources, int32_t *n_sources_out) {
    /* name -> id map of LayeredGraph.addEdge :157-174, keyed by (layer, region index) */
    int32_t *vid = (int32_t *)malloc(sizeof(int32_t) * (size_t)L * (size_t)n);
    for (int64_t i = 0; i < (int64_t)L * n; i++) vid[i] = -1;
    int32_t nv = 0;
    int64_t ne = 0;
    int32_t time_step = 24 / L; /* CrossTimeGraph.java:31 */
    for (int32_t h = 0; h < L; h++) {
        for (int32_t a = 0; a < n; a++) {
            int32_t s = order[a];
            for (int32_t b = 0; b < n; b++) {
                int32_t d = order[b];
                int32_t f = mode == 0 ? ora_flow_ca(F, n, s, d, intervals[h], intervals[h + 1])   /* :79 */
                                      : ora_flow_tract(F, n, s, d, h, h + time_step - 1);        /* :36 */
                if (f > 0) {
                    int32_t h2 = (h + 1) % L;
                    int64_t ks = (int64_t)h * n + s, kd = (int64_t)h2 * n + d;
                    if (vid[ks] < 0) { vid[ks] = nv; v_layer[nv] = h; v_region[nv] = s; nv++; }
                    if (vid[kd] < 0) { vid[kd] = nv; v_layer[nv] = h2; v_region[nv] = d; nv++; }
                    src[ne] = vid[ks]; dst[ne] = vid[kd]; w[ne] = (double)f;
                    ne++;
                }
            }
        }
    }
    int32_t nsrc = 0;
    for (int32_t a = 0; a < n; a++) { /* :43-47 / :86-90 */
        int32_t s = order[a];
        if (vid[s] >= 0) sources[nsrc++] = vid[s];
    }
    free(vid);
    *n_vertices_out = nv;
    *n_sources_out = nsrc;
    return ne;
}

/* CrossTimeGraph.sampleSequenceHelper :134-141: fout.write(String.join(" ", seq) + "\n") per walk through a
 * BufferedWriter -- one thread, tokens "<layer>-<region>".  The CPU arm of the reference's published timing experiment
 * (CrossTimeGraph.java:152-159 includes this write).  Returns bytes written, -1 on an I/O error. */
int64_t ora_write_seq(const int32_t *tokens, int64_t n_walks, int32_t L, const int32_t *layer, const int32_t *region,
                      const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    static char buf[1 << 16];
    setvbuf(f, buf, _IOFBF, sizeof(buf));
    int64_t bytes = 0;
    char line[64 * 26 + 2];
    for (int64_t i = 0; i < n_walks; i++) {
        char *p = line;
        for (int32_t j = 0; j < L && j < 64; j++) {
            int32_t t = tokens[i * L + j];
            if (t < 0) break;
            if (j) *p++ = ' ';
            p += sprintf(p, "%d-%d", layer[t], region[t]);
        }
        *p++ = '\n';
        if (fwrite(line, 1, (size_t)(p - line), f) != (size_t)(p - line)) { fclose(f); return -1; }
        bytes += p - line;
    }
    return fclose(f) == 0 ? bytes : -1;
}
