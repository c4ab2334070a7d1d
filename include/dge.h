/*
 * dge.h -- C ABI of libdge.so: the B200-native (sm_100a) replacement of the region-embedding
 * hot path of thekingofkings/embedding:
 *     stage 1  weighted random walks over the time-sliced mobility-flow graph
 *     stage 2  skip-gram negative-sampling SGD over the walk corpus
 *
 * The reference is pure Java and has no FFI of its own.  Each entry point below names the Java
 * symbol it stands under (paths relative to embedding/src/main/java/embedding/ of the reference);
 * INTEGRATION.md shows the JNI / Panama stubs a maintainer adds to the unchanged Java classes.
 *
 * Conventions
 *   - plain C types only; every buffer in a signature is HOST memory owned by the caller
 *     (device memory never crosses the ABI);
 *   - opaque handles are library-owned and released only by the matching *_free; every handle keeps its ctx alive, so
 *     a *_free that comes after dge_destroy is safe (the ctx is torn down when its last handle goes);
 *   - device memory comes from a stream-ordered pool PRIVATE to the ctx (the process-wide default pool is not touched);
 *   - every call is blocking; a ctx (one per GPU / process) is used by one thread at a time;
 *   - return 0 on success, a negative dge_status otherwise; dge_last_error() has the text;
 *   - there is NO CPU fallback: without a usable sm_100 device dge_create fails with DGE_E_NO_DEVICE.
 */
#ifndef DGE_H
#define DGE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dge_ctx dge_ctx;
typedef struct dge_graph dge_graph;
typedef struct dge_corpus dge_corpus;
typedef struct dge_model dge_model;
typedef struct dge_flows dge_flows;

typedef enum {
    DGE_OK = 0,
    DGE_E_INVALID = -1,   /* bad argument (null pointer, out-of-range id, size) */
    DGE_E_NO_DEVICE = -2, /* no CUDA device / not sm_100 */
    DGE_E_CUDA = -3,      /* CUDA runtime error */
    DGE_E_IO = -4,        /* file error */
    DGE_E_LIMIT = -5,     /* documented size limit exceeded */
    DGE_E_COMM = -6       /* NCCL / multi-GPU error */
} dge_status;

enum { DGE_SAMPLER_ALIAS = 0, /* LayeredGraph.sampleVertexSequence()    :232-252 */
       DGE_SAMPLER_CDF = 1 }; /* LayeredGraph.sampleVertexSequence_OV() :260-279 */

/* ------------------------------------------------------------------ context */
int dge_version(void);
/* device: CUDA ordinal (one process per GPU uses LOCAL_RANK). */
int dge_create(int device, dge_ctx **out);
void dge_destroy(dge_ctx *ctx);
/* ctx may be NULL: returns the calling thread's last error (e.g. from a failed dge_create). */
const char *dge_last_error(const dge_ctx *ctx);
/* Pinned host buffers (optional; any host pointer is accepted everywhere, pinned ones copy faster). */
void *dge_host_alloc(size_t bytes);
void dge_host_free(void *p);
/* Device time in ms of the most recent run of a named internal phase, measured with CUDA events on
 * the launching stream ("csr", "alias", "pack", "walk", "vocab", "sgns", "tokens_d2h", ...).
 * Returns DGE_E_INVALID for an unknown name. */
int dge_phase_ms(const dge_ctx *ctx, const char *phase, float *ms);
/* Number of library kernels launched since dge_create (for bench.py's gpu_launches). */
int64_t dge_kernel_launches(const dge_ctx *ctx);
/* Device-side stopwatch: CUDA events recorded on the ctx stream (the stream every library kernel is launched on)
 * around whatever calls the host makes in between.  Replaces the System.currentTimeMillis() prints of
 * CrossTimeGraph.java:128,146-147 / SpatialGraph.java:98,119-120. */
int dge_timer_start(dge_ctx *ctx);
int dge_timer_stop(dge_ctx *ctx, float *ms);

/* ------------------------------------------------------------------ multi-GPU (one process per GPU)
 * The reference is single-device (only a commented-out hint, DeepWalk.java:43).  Walks shard by walk id with no
 * collective (first_walk_id of dge_walk).  Stage 2 becomes data-parallel when the ctx carries a communicator:
 * every rank trains on its own corpora (global sentence order = rank 0's sentences, then rank 1's, ...: the ranks
 * together enumerate exactly the pairs and negatives of a single-GPU run over the whole corpus) and dge_sgns_train
 * exchanges the embedding deltas sync_rounds times per epoch -- one kernel over NVLink peer memory (the replicas of all
 * ranks are mapped with cudaIpc), or NCCL all-reduces where peer mapping is unavailable; the per-row combine rule is
 * dge_sgns_params.combine.  dge_sgns_train is then COLLECTIVE: every rank must call it with the same parameters
 * (checked; a failure on one rank makes every rank return an error).  The host moves the opaque id from rank 0 to
 * the other ranks by any means it has (a file, a socket, MPI, torch.distributed). */
#define DGE_COMM_ID_BYTES 128
int dge_comm_unique_id(void *id, size_t bytes);
int dge_comm_init(dge_ctx *ctx, int rank, int world, const void *id, size_t bytes);
int dge_comm_shape(const dge_ctx *ctx, int *rank, int *world);
void dge_comm_destroy(dge_ctx *ctx); /* also done by dge_destroy */
int dge_comm_nccl_version(void);     /* 0 when NCCL cannot be loaded */

/* ------------------------------------------------------------------ stage 0: flow records -> CrossTimeGraph
 * The hourly flow counts of every region: the dense tensor F[src][hour 0..23][dst] int32 stands for the reference's
 * per-region `taxiFlows: List<Map<dst, count>>` (CommunityAreas.java:220, Tracts.java:459).  Region INDICES
 * (0..n_regions-1) are positions in the host's region-id array; the host keeps the ids and the geometry.
 * F != NULL uploads counts (the deserialised CA-serialize-<year>.seq / tracts-serialize-<year>.seq maps,
 * CommunityAreas.java:115-124, Tracts.java:115-125); F == NULL starts from zero.  Limit: n_regions <= 8192. */
int dge_flows_create(dge_ctx *ctx, int32_t n_regions, const int32_t *F, dge_flows **out);
/* Counting of CommunityAreas.mapTripsIntoCommunities :55-103 / Tracts.mapTripsIntoTracts :71-102 after the host's
 * point-in-polygon lookup: per trip F[src][start hour][dst] += 1.  A trip with src or dst == -1 (outside every
 * region) is ignored as in the reference. */
int dge_flows_add_trips(dge_flows *f, int64_t n_trips, const int32_t *src_region, const int32_t *dst_region,
                        const int32_t *start_hour);
int dge_flows_tensor(const dge_flows *f, int32_t *F /* [n][24][n] */);
void dge_flows_free(dge_flows *f);
/* Slot weights of every (src, dst) pair: W[src * n + dst] = getFlowTo(dst, lo, hi) of region src, by region index.
 * mode 0 = CommunityArea.getFlowTo CommunityAreas.java:240-245 (hours lo, lo+1, ... circular, stopping BEFORE hi: the
 * call getFlowTo(j, 0, 23) of outputStaticFlowGraph :133 therefore leaves hour 23 out, as the reference does);
 * mode 1 = Tract.getFlowTo Tracts.java:477-482 (hours lo..hi inclusive; lo == hi is the single-hour
 * getFlowTo(dst, hour) :236 / Tracts.java:474). */
int dge_flows_slot_weights(const dge_flows *f, int mode, int32_t lo, int32_t hi, int32_t *W);
/* The static-graph exports that feed the LINE / matrix-factorisation baselines.  Rows / sources
 * are visited in rows[0..n_rows), columns / destinations in cols[0..n_cols) (region indices; the host passes ids
 * 1..77, the sorted tract ids or its HashMap iteration order as the Java loops do).
 *   dge_flows_write_matrix: one line per row, the weights joined by `sep` -- CommunityAreas.outputStaticFlowGraph
 *     :129-144 (','), outputAdjacencyMatrix :147-164 (' '), Tracts.outputAdjacencyMatrix :265-301 (',').
 *   dge_flows_write_od: one line "<src id> <dst id> <w>" per pair -- CommunityAreas.outputStaticFlowGraph :130,136-137
 *     (w > 0 only), outputEdgeGraph_LINE :169-184 (keep_zero != 0, and the Java loop's `j < size` drops the last
 *     column: the host passes n_cols = 76), Tracts.outputEdgeFile :236-260.  presence_hour >= 0 visits only the
 *     destinations with a trip in that single hour (Tracts.java:243 iterates taxiFlows.get(h).keySet()).
 *     region_ids[n] are the printed ids. */
int dge_flows_write_matrix(const dge_flows *f, int mode, int32_t lo, int32_t hi, const int32_t *rows, int32_t n_rows,
                           const int32_t *cols, int32_t n_cols, char sep, const char *path);
int dge_flows_write_od(const dge_flows *f, int mode, int32_t lo, int32_t hi, const int32_t *rows, int32_t n_rows,
                       const int32_t *cols, int32_t n_cols, const int32_t *region_ids, int keep_zero,
                       int32_t presence_hour, const char *path);
/* CrossTimeGraph.constructGraph_CA(int[]) :68-95 (mode 0; intervals[num_layer+1], slot h = circular half-open
 * [intervals[h], intervals[h+1]) of CommunityArea.getFlowTo :240-245) and constructGraph_tract() :25-52 (mode 1;
 * slot h = hours [h, h + 24/num_layer - 1] inclusive of Tract.getFlowTo Tracts.java:477-482), followed by
 * initiateAliasTables :195-226, entirely on device: slot sums, the (h, src, dst) edge enumeration with w > 0, vertex
 * ids by first appearance (LayeredGraph.addEdge :159-170), the layer-0 source list, CSR and alias tables.
 * order[n_regions] = region indices in the host's HashMap iteration order (it defines edge and id order).
 * The result is bit-identical to dge_graph_build on the COO the Java loops would produce. */
int dge_crosstime_graph_build(const dge_flows *f, const int32_t *order, int32_t num_layer, int mode,
                              const int32_t *intervals, dge_graph **out);
/* Labels by vertex id of a graph built by dge_crosstime_graph_build: layer h and region index of "<h>-<region>";
 * sources[n_sources] works for every graph.  Any pointer may be NULL. */
int dge_graph_labels(const dge_graph *g, int32_t *v_layer, int32_t *v_region_index, int32_t *sources);

/* ------------------------------------------------------------------ stage 1a: graph + alias tables
 * Stands under LayeredGraph.addEdge :157-174, addSourceVertex :180-189, initiateAliasTables :195-226
 * (row tables: Vertex.initiateAliasTable :54-82).  The Java host owns names and order: it passes
 * vertex ids in first-appearance order, COO edges in insertion order, the source list in order.
 * CSR rows keep insertion order; out_degree is the left-to-right double sum of the row
 * (Vertex.addOutEdge :46-49) unless out_degree != NULL (host-owned values, e.g. after
 * SpatialGraph.keepNearestKVertices :29-35); source_weight_sum likewise (:188 / SpatialGraph.java:57).
 * prob/alias tables are bit-identical to the Java ones (IEEE double, same operation order).
 * Limits: n_edges < 2^31, a single row (or the source list) <= 2^25 entries. */
int dge_graph_build(dge_ctx *ctx, int32_t n_vertices, int64_t n_edges, const int32_t *src, const int32_t *dst,
                    const double *w, int32_t n_sources, const int32_t *sources, const double *out_degree,
                    const double *source_weight_sum, dge_graph **out);
int dge_graph_sizes(const dge_graph *g, int32_t *n_vertices, int64_t *n_edges, int32_t *n_sources);
/* Read-back (Vertex.probTable / aliasTable / outDegree, LayeredGraph.probTable / aliasTable); any
 * pointer may be NULL.  row_ptr[nv+1], col/w/prob/alias[ne] in CSR order, out_degree[nv],
 * src_prob/src_alias[ns], source_weight_sum[1].  alias == -1 means "no alias" as in Java. */
int dge_graph_tables(const dge_graph *g, int64_t *row_ptr, int32_t *col, double *w, double *prob,
                     int32_t *alias, double *out_degree, double *src_prob, int32_t *src_alias,
                     double *source_weight_sum);
/* Batched Vertex.sampleNextVertex(double x) :123-132 (sampler ALIAS) / sampleNextVertex_OV :89-98 with the
 * uniform supplied (sampler CDF): out[i] = next vertex of v[i] for uniform x[i], -1 if v[i] has no
 * out-edge.  v[i] == -1 draws a source vertex instead (sampleVertexSequence :233-242). */
int dge_graph_sample_next(const dge_graph *g, int64_t n, const int32_t *v, const double *x, int sampler,
                          int32_t *out);
void dge_graph_free(dge_graph *g);

/* ------------------------------------------------------------------ stage 1b: walks
 * Stands under the loop of CrossTimeGraph.sampleSequenceHelper :134-140 / SpatialGraph.outputSampleSequence
 * :103-110 calling LayeredGraph.sampleVertexSequence() :232-252.  Walk i uses the counter-based stream
 * Philox4x32-10(key = seed, counter = (first_walk_id + i, draw/2)): draw 0 picks the source, draw t the
 * t-th step; results do not depend on how walks are split over calls or GPUs.  (The reference's
 * java.util.Random is unseeded, LayeredGraph.java:14, so its walks are not reproducible.) */
int dge_walk(const dge_graph *g, int64_t n_walks, int64_t first_walk_id, int32_t num_layer, uint64_t seed,
             int sampler, dge_corpus **out);
/* Corpus from host tokens [n_walks, L] (int32 ids in [0,n_ids), -1 = padding). */
int dge_corpus_from_tokens(dge_ctx *ctx, const int32_t *tokens, int64_t n_walks, int32_t L, int32_t n_ids,
                           dge_corpus **out);
int dge_corpus_shape(const dge_corpus *c, int64_t *n_walks, int32_t *L, int32_t *n_ids);
/* tokens[n_walks * L] walk-major, -1 padded after a dead end (Java: sampleNextVertex() == null). */
int dge_corpus_tokens(const dge_corpus *c, int32_t *tokens);
/* Same, as 16-bit tokens (0xFFFF = padding) for id spaces below 65 535 -- every reference-scale graph (CA: 1 848
 * vertices, tract x 24: 19 224).  The download of the walk corpus is bound by the PCIe link, so half the
 * bytes is half the time; a JNI host receives a short[] / ShortBuffer.  DGE_E_LIMIT when n_ids > 65 535. */
int dge_corpus_tokens_u16(const dge_corpus *c, uint16_t *tokens);
/* In-place change of id space: token t at walk position j becomes id_map[t] + j * position_stride; the
 * corpus then reports new_n_ids.  This is how the host puts both corpora of the "usespatial" run into ONE
 * vocabulary, exactly as the token strings do in the reference: a cross-time token "<h>-<region>"
 * (CrossTimeGraph.java:80-82) and a spatial token "<j>-<region>" (SpatialGraph.java:105-108) are the same
 * word when h == j.  id_map has the corpus' current n_ids entries, all results must lie in [0, new_n_ids). */
int dge_corpus_relabel(dge_corpus *c, const int32_t *id_map, int32_t new_n_ids, int32_t position_stride);
/* Number of tokens != -1 (walk steps taken, counting the source draw). */
int dge_corpus_count_tokens(const dge_corpus *c, int64_t *n_tokens);
/* `.seq` text (CrossTimeGraph.java:136-137, SpatialGraph.java:105-110): one walk per line, tokens
 * "<layer>-<region>" joined by one space, '\n' terminated, no header.  label_layer / label_region are
 * per vertex id; position_prefix != 0 writes "<j>-<region>" with j the position in the walk (spatial
 * mode) and ignores label_layer.  append != 0 appends to an existing file. */
int dge_corpus_write_seq(const dge_corpus *c, const int32_t *label_layer, const int32_t *label_region,
                         int position_prefix, const char *path, int append);
/* Reads a `.seq` file back into a device corpus: what FileSentenceIterator / LineSentenceIterator +
 * DefaultTokenizerFactory feed to Word2Vec when the corpus files already exist (DeepWalk.java:47-59,70; generation
 * is skipped by checkInputFile :86-87,99-100).  Tokens "<layer>-<region>" are mapped to the id whose labels match;
 * with position_prefix != 0 the first number is the walk position (spatial corpus) and only the region is matched.
 * L = the longest line; shorter lines are -1 padded.  An unknown or malformed token is DGE_E_INVALID. */
int dge_corpus_read_seq(dge_ctx *ctx, const char *path, const int32_t *label_layer, const int32_t *label_region,
                        int32_t n_ids, int position_prefix, dge_corpus **out);
void dge_corpus_free(dge_corpus *c);

/* ------------------------------------------------------------------ stage 2: skip-gram
 * Stands under DeepWalk.learnEmbedding :32-83: Word2Vec.Builder()...build() :73-76 and w2v.fit() :79
 * (DL4J 0.7.2, external) and WordVectorSerializer.writeWordVectors :82. */
enum { DGE_SCHEDULE_ITEMS = 0,     /* work item = (sentence, centre); row updates are 128-bit L2 reductions */
       DGE_SCHEDULE_SENTENCE = 1 }; /* work item = sentence; plain atomic-free row stores (classic Hogwild) */
/* How the deltas d_r = (replica of rank r) - (common base) of one embedding row are combined at an exchange:
 * base += sum_r d_r / div.  SUM applies every update but overshoots when several ranks saturate the same row (it
 * diverges at world = 8); MEAN is parameter averaging (a row only one rank saw learns world times slower);
 * CONTRIBUTORS divides by the number of ranks whose delta is non-zero; SQRT by its square root; ALIGNED (default) by
 * max(1, |sum_r d_r|^2 / sum_r |d_r|^2), which is 1 for orthogonal deltas (independent information: summed) and the
 * number of contributors for parallel ones (every rank made the same move: averaged).  DESIGN.md 3.4. */
enum { DGE_COMBINE_DEFAULT = 0, DGE_COMBINE_MEAN = 1, DGE_COMBINE_CONTRIBUTORS = 2, DGE_COMBINE_SQRT = 3,
       DGE_COMBINE_ALIGNED = 4, DGE_COMBINE_SUM = 5 };
enum { DGE_TRANSPORT_AUTO = 0, DGE_TRANSPORT_PEER = 1, DGE_TRANSPORT_NCCL = 2 };
/* dge_sgns_params.flags (tests, A/B measurements; every one of them changes speed or schedule, none the arithmetic of
 * an update, except NO_UPDATES which trains nothing and PLAIN_STORES which may lose updates) */
enum { DGE_SGNS_F_NO_UPDATES = 1,     /* timing experiment: compute everything, send no row update */
       DGE_SGNS_F_NO_NARROW = 2,      /* never use the 4-lane-group kernel for D <= 16 */
       DGE_SGNS_F_ONE_WARP = 8,       /* item kernel on ONE warp, one item at a time, corpus order (arithmetic check) */
       DGE_SGNS_F_NARROW = 32,        /* always use the 4-lane-group kernel for D <= 16 */
       DGE_SGNS_F_TARGET_PARALLEL = 64,    /* always use the target-parallel kernel for D <= 16, K <= 7 */
       DGE_SGNS_F_NO_TARGET_PARALLEL = 128,
       DGE_SGNS_F_STAGED_ROWS = 256,  /* item kernel with the rows of the next unit staged in shared memory (cp.async) */
       DGE_SGNS_F_PLAIN_STORES = 512, /* atomic-free item kernel: plain 128-bit row stores instead of L2 reductions */
       DGE_SGNS_F_BLOCK_PER_SENTENCE = 4, /* kernel G (a block owns a sentence) even with concurrency = 1 (else the default for narrow rows) */
       DGE_SGNS_F_PAIR_WARPS = 262144,    /* kernel I: kernel G's wavefront with a warp per pair and the round's pairs handed to the block's warps dynamically;
                                             bits 12-15 (4 .. 15) = warps per block, default 8 */
       DGE_SGNS_F_HELPER_WARPS = 524288,  /* kernel J: kernel G's wavefront with helper warps that stage the rows (cp.async ring in shared memory)
                                             and send the reductions; bits 12-15 (2 .. 4) = stages of the ring */
       DGE_SGNS_F_HOT_SHIFT = 20,         /* not a flag: bits 20-23 = v > 0 makes the 2^(v-1) most frequent words write-through in kernel F */
       DGE_SGNS_F_ROW_PREFETCH = 1 << 24, /* kernel F, rows of up to 8 slots, K <= 5: the rows of the next unit are requested before the current one is computed */
       DGE_SGNS_F_DYNAMIC = 1 << 25,      /* kernel F: sentences handed out in corpus order from a counter instead of strided by warp */
       DGE_SGNS_F_ROW_PREFETCH_SMEM = 1 << 26, /* the same with the requested rows landing in shared memory (cp.async) instead of registers */
       DGE_SGNS_F_PIPELINED = 131072,     /* kernel H: kernel G with the sentences of a block pipelined through the wavefront */
       DGE_SGNS_F_ITEM_KERNELS = 65536,   /* the round-1 item kernels B-E with their automatic choice (centres of a sentence in flight at once) */
       DGE_SGNS_F_SMALL_BLOCKS = 16,  /* sentence-resident kernel: 128-thread blocks instead of one 640-thread block per SM */
       DGE_SGNS_F_SENTENCE_RESIDENT = 2048, /* kernel F (a warp owns a sentence for all its centres) where kernel G would be chosen */
       DGE_SGNS_F_SMEM_NEG_TABLE = 1024, /* rows of up to 8 slots, V < 65536: one 640-thread block per SM, negative table in shared memory */
       DGE_SGNS_F_BLOCKS_6 = 6 << 12, /* with STAGED_ROWS, rows of up to 8 slots: register allocation for 6 / 7 resident */
       DGE_SGNS_F_BLOCKS_7 = 7 << 12  /* blocks per SM instead of 5 */ };
typedef struct {
    int32_t dim;            /* layerSize(...)         DeepWalk.java:62-66,74 */
    int32_t window;         /* windowSize(numLayer)   :74 */
    int32_t negative;       /* negativeSample(5)      :75 */
    int32_t min_count;      /* minWordFrequency(2)    :73 */
    int32_t epochs;         /* iterations(1) x epochs(1) */
    int32_t neg_table_size; /* DL4J: 100000 */
    int32_t exp_table_size; /* sigmoid lookup table entries (1000) */
    int32_t concurrency;    /* sentences in flight, the analogue of workers(8) :75.  0 = automatic (bounded by
                               the vocabulary size); 1 = one sentence at a time in the oracle's
                               exact order (parity tests); N = N sentences in flight */
    int32_t schedule;       /* DGE_SCHEDULE_ITEMS (default) or DGE_SCHEDULE_SENTENCE */
    int32_t sync_rounds;    /* multi-GPU only: exchanges of the embedding deltas per epoch; 0 = automatic (one per
                               ~2^19 local sentences, at least 8).  Without a communicator of world > 1 a value > 0
                               still cuts the epoch into that many launches (same arithmetic, exchange = identity). */
    int32_t combine;        /* multi-GPU only: DGE_COMBINE_* rule for the per-rank deltas of a row; 0 = default
                               (DGE_COMBINE_ALIGNED).  Must agree on all ranks (checked). */
    int32_t transport;      /* multi-GPU only: DGE_TRANSPORT_AUTO (peer-memory kernel over NVLink when the replicas can
                               be mapped, else NCCL), _PEER (fail if they cannot), _NCCL */
    uint32_t flags;         /* DGE_SGNS_F_*: kernel-selection / measurement hooks for tests and A/B runs; 0 in production */
    float lr;               /* 0.025 */
    float min_lr;           /* 1e-4 */
    uint64_t seed;
} dge_sgns_params;
void dge_sgns_default_params(dge_sgns_params *p);
/* Trains on the concatenation of n corpora (FileSentenceIterator over both .seq files, :48-50). All
 * corpora must share one id space (n_ids).  Vocabulary: ids with count >= min_count, indexed by
 * descending count (ties: ascending id). */
int dge_sgns_train(dge_ctx *ctx, const dge_corpus *const *corpora, int32_t n_corpora,
                   const dge_sgns_params *params, dge_model **out);
int dge_model_shape(const dge_model *m, int32_t *vocab_size, int32_t *dim, int64_t *pairs_trained);
/* syn0 / syn1neg [V*dim] row-major by vocabulary index; id_of_word[V] maps back to corpus ids. NULLs ok. */
int dge_model_vectors(const dge_model *m, float *syn0, float *syn1neg, int32_t *id_of_word);
/* Health of the trained tables without downloading them (no reference counterpart; DL4J has none): mean Euclidean
 * norm of the syn0 rows, largest |element| of syn0 and syn1neg, number of non-finite elements.  A diverging
 * multi-GPU exchange shows up here as exploding norms.  NULLs ok. */
int dge_model_stats(const dge_model *m, double *mean_row_norm, double *max_abs, int64_t *n_nonfinite);
/* `.vec` text (writeWordVectors, consumers python/embeddingEvaluation_tract.py:139-166): no header, one line
 * per vocabulary word "<layer>-<region> v1 ... vD".  label arrays are indexed by corpus id. */
int dge_model_write_vec(const dge_model *m, const int32_t *label_layer, const int32_t *label_region,
                        const char *path);
void dge_model_free(dge_model *m);

/* ------------------------------------------------------------------ downstream metric
 * The reference's pairwise-similarity evaluation of an embedding layer on device, fp64 as numpy / scipy:
 * python/embeddingEvaluation_tract.py:169-196 (pairwiseEstimator: cosine distance of every pair of rows, NaN -> 2,
 * the topk nearest OTHER rows; ties by ascending row index like numpy's stable argsort) and :249-260 (dcg_atK /
 * ndcg_atK with relevance 1 - ground-truth distance, discount 1 / log2(rank + 2), normalised by the DCG of the
 * ground truth's own ordering, generatePairWiseGT :63-103).
 * dge_eval_knn: X[m*dim] row-major -> nbr[m*topk] (row indices, -1 padded when m <= topk), dist[m*topk] or NULL.
 * dge_eval_ndcg: gt_index[m] = row of every X row in the ground-truth distance matrix gt_dist[n*n]; ndcg[m] (or NULL)
 * per row and their mean.  With m <= topk the reference reports nothing: *mean = NaN. */
int dge_eval_knn(dge_ctx *ctx, const float *X, int32_t m, int32_t dim, int32_t topk, int32_t *nbr, double *dist);
int dge_eval_ndcg(dge_ctx *ctx, const float *X, int32_t m, int32_t dim, const int32_t *gt_index, const double *gt_dist,
                  int32_t n, int32_t topk, double *ndcg, double *mean);

#ifdef __cplusplus
}
#endif
#endif
