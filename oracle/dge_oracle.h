/*
 * dge_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's stage-1 hot path (graph core, alias
 * tables, weighted random walks) of thekingofkings/embedding.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (embedding_b200/libdge.so) never links
 * or calls it.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/embedding/src/main/java/embedding/).
 *
 * Parity status: stage 1 is pinned by the reference's only golden vector
 * (LayeredGraphTest.java:12-44, checked in tests/test_oracle_golden.py).
 * The reference itself (Java) cannot run in this image (no JDK), so there is
 * no oracle/_ref build.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no fast-math).
 */
#ifndef DGE_ORACLE_H
#define DGE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_graph ora_graph;

enum { ORA_SAMPLER_ALIAS = 0, ORA_SAMPLER_CDF = 1 };
enum { ORA_RNG_PHILOX = 0, ORA_RNG_JAVA_LCG = 1 };
enum { ORA_ALIAS_LITERAL = 0, ORA_ALIAS_FAST = 1 };

/* Vertex.initiateAliasTable(), LayeredGraph.java:54-82 -- literal O(k^2). */
void ora_alias_literal(int32_t k, const double *w, double out_degree, double *prob, int32_t *alias);
/* Same result bit for bit, ordered-set form (hierarchical bitmaps), O(k log k). */
void ora_alias_fast(int32_t k, const double *w, double out_degree, double *prob, int32_t *alias);

/* LayeredGraph.addEdge :157-174, addSourceVertex :180-189, initiateAliasTables :195-226.
 * Vertex ids are the host's first-appearance indices; edges arrive in insertion order.
 * out_degree_override / source_weight_sum_override: NULL => accumulate left to right as
 * Vertex.addOutEdge :46-49 and addSourceVertex :188 do; non-NULL => host-owned values
 * (SpatialGraph.java:34,57 recompute them with DoubleStream.sum()). */
ora_graph *ora_graph_build(int32_t n_vertices, int64_t n_edges, const int32_t *src, const int32_t *dst,
                           const double *w, int32_t n_sources, const int32_t *sources,
                           const double *out_degree_override, const double *source_weight_sum_override,
                           int alias_mode);
void ora_graph_free(ora_graph *g);
int64_t ora_graph_num_edges(const ora_graph *g);
int32_t ora_graph_num_vertices(const ora_graph *g);
int32_t ora_graph_num_sources(const ora_graph *g);
/* Any output pointer may be NULL. Sizes: row_ptr[nv+1], col/w/prob/alias[ne], out_degree[nv],
 * src_prob/src_alias[ns], source_weight_sum[1]. */
void ora_graph_tables(const ora_graph *g, int64_t *row_ptr, int32_t *col, double *w, double *prob,
                      int32_t *alias, double *out_degree, double *src_prob, int32_t *src_alias,
                      double *source_weight_sum);

/* Vertex.sampleNextVertex(double x), LayeredGraph.java:123-132 (and :104-116). Returns the
 * destination vertex id, or -1 when the vertex has no out-edges (Java: null).
 * alias == -1 is resolved to the column itself (Java would throw). */
int32_t ora_sample_next(const ora_graph *g, int32_t v, double x);
/* Vertex.sampleNextVertex_OV(), LayeredGraph.java:89-98, with x supplied. */
int32_t ora_sample_next_ov(const ora_graph *g, int32_t v, double x);
/* Source draw of sampleVertexSequence :234-242 / _OV :261-270; returns a vertex id or -1. */
int32_t ora_sample_source(const ora_graph *g, double x, int sampler);

/* LayeredGraph.sampleVertexSequence() :232-252 (sampler=ALIAS) / _OV :260-279 (sampler=CDF),
 * called n_walks times as CrossTimeGraph.sampleSequenceHelper :134-140 does.
 * tokens[n_walks*L] int32 vertex ids, -1 padded after a dead end.
 * rng_kind PHILOX: draw t of walk id (first_walk_id+i) is Philox4x32-10 keyed by seed (the
 * counter-based replacement of the unseeded java.util.Random, LayeredGraph.java:14).
 * rng_kind JAVA_LCG: one java.util.Random(seed) stream shared by all walks, as the reference. */
void ora_walk(const ora_graph *g, int64_t n_walks, int64_t first_walk_id, int32_t L, uint64_t seed,
              int sampler, int rng_kind, int32_t *tokens);

/* RNG building blocks (exposed for known-answer tests). */
void ora_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double ora_philox_uniform(uint64_t seed, uint64_t walk_id, uint32_t draw);
void ora_java_random_doubles(int64_t seed, int32_t n, double *out);

/* SpatialGraph.keepNearestKVertices(k), SpatialGraph.java:29-35: per row a stable sort by
 * descending weight, truncation to k, out_degree = DoubleStream.sum() (JDK 8 compensated sum).
 * In: dense row-major weights w[n*n] (row = source, insertion order = column order).
 * Out: col[n*k], wk[n*k], out_degree[n]. */
void ora_keep_nearest_k(int32_t n, const double *w, int32_t k, int32_t *col, double *wk, double *out_degree);
double ora_java8_stream_sum(const double *v, int64_t n);

/* CommunityArea.getFlowTo(dst,lo,hi) CommunityAreas.java:240-245 (circular half-open) and
 * Tract.getFlowTo(dst,lo,hi) Tracts.java:477-482 (inclusive). F is dense [n][24][n] int32. */
int32_t ora_flow_ca(const int32_t *F, int32_t n, int32_t src, int32_t dst, int32_t lo, int32_t hi);
int32_t ora_flow_tract(const int32_t *F, int32_t n, int32_t src, int32_t dst, int32_t lo, int32_t hi);

/* CrossTimeGraph.constructGraph_CA(int[]) :68-95 (mode 0) / constructGraph_tract() :25-52 (mode 1).
 * order[n]: region indices in the host's HashMap iteration order.
 * intervals[L+1] used by mode 0 only.  Outputs (caller-sized: edges <= L*n*n, vertices <= L*n):
 * COO in insertion order, vertex labels (layer, region index) by first appearance, sources.
 * Returns the number of edges; *n_vertices_out, *n_sources_out are set. */
int64_t ora_crosstime_edges(const int32_t *F, int32_t n, const int32_t *order, int32_t L, int mode,
                            const int32_t *intervals, int32_t *src, int32_t *dst, double *w,
                            int32_t *v_layer, int32_t *v_region, int32_t *n_vertices_out,
                            int32_t *sources, int32_t *n_sources_out);

/* String.join(" ", seq) + "\n" per walk into a file, one thread (CrossTimeGraph.java:134-141); bytes written or -1. */
int64_t ora_write_seq(const int32_t *tokens, int64_t n_walks, int32_t L, const int32_t *layer, const int32_t *region,
                      const char *path);

#ifdef __cplusplus
}
#endif
#endif
