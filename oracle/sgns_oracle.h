/*
 * sgns_oracle.h -- CPU ORACLE for stage 2 (skip-gram, negative sampling [+ optional
 * hierarchical softmax]).  TEST INFRASTRUCTURE ONLY, never linked into the product.
 *
 * PARITY UNPINNED.  The reference contributes only a builder call (DeepWalk.java:73-76)
 * and `w2v.fit()` (:79); the arithmetic lives in un-vendored Maven dependencies that are
 * absent from /root/reference and unreachable here:
 *     org.deeplearning4j:deeplearning4j-nlp:0.7.2, org.nd4j:nd4j-native:0.7.2 (pom.xml:14-16,42-51)
 * No reference test pins any Word2Vec output.  This file restates the published word2vec
 * skip-gram algorithm (Mikolov et al. 2013; word2vec.c) with the DL4J 0.7.2 parameterisation
 * listed in SURVEY.md section 8(a) row A14 (from knowledge of that release, unverifiable here).
 * Parity for stage 2 is therefore defined through (i) exact pair / negative enumeration and
 * fp32-tolerance agreement with this oracle on seeded inputs in a sequential schedule, and
 * (ii) the reference's own downstream metrics within the oracle's seed-to-seed spread.
 */
#ifndef SGNS_ORACLE_H
#define SGNS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t dim;            /* layerSize            DeepWalk.java:62-66,74 */
    int32_t window;         /* windowSize           DeepWalk.java:74 */
    int32_t negative;       /* negativeSample(5)    DeepWalk.java:75 */
    int32_t min_count;      /* minWordFrequency(2)  DeepWalk.java:73 */
    int32_t epochs;         /* iterations(1) x epochs(1) */
    int32_t threads;        /* workers(8)           DeepWalk.java:75 */
    int32_t use_hs;         /* DL4J default true; north_star asks for pure SGNS => 0 */
    int32_t neg_table_size; /* DL4J 100000 */
    int32_t exp_table_size; /* 1000 */
    float lr;               /* 0.025 */
    float min_lr;           /* 1e-4 */
    uint64_t seed;
} ora_sgns_params;

typedef struct ora_vocab ora_vocab;
typedef struct ora_model ora_model;

/* Vocabulary: count tokens (ids in [0,n_ids), -1 = padding), keep count >= min_count, index by
 * descending count, ties by ascending raw id (the host owns tie order). */
ora_vocab *ora_vocab_build(const int32_t *tokens, int64_t n_tokens, int32_t n_ids, int32_t min_count);
void ora_vocab_free(ora_vocab *v);
int32_t ora_vocab_size(const ora_vocab *v);
int64_t ora_vocab_total_words(const ora_vocab *v);
/* word_of_id[n_ids] (-1 = dropped); id_of_word[V]; count[V]. NULLs allowed. */
void ora_vocab_tables(const ora_vocab *v, int32_t *word_of_id, int32_t *id_of_word, int64_t *count);
/* unigram^0.75 table, word2vec.c InitUnigramTable / DL4J InMemoryLookupTable.makeTable */
void ora_neg_table(const ora_vocab *v, int32_t table_size, int32_t *table);

/* syn0 init (U[0,1)-0.5)/dim from Philox(seed); syn1neg = 0 */
void ora_init_syn0(int32_t V, int32_t dim, uint64_t seed, float *syn0);

/* Train.  tokens[n_sent*L] raw ids (-1 padded).  Returns the model (syn0, syn1neg [V*dim]).
 * pairs_out: number of (centre, context) updates executed. */
ora_model *ora_sgns_train(const int32_t *tokens, int64_t n_sent, int32_t L, int32_t n_ids,
                          const ora_sgns_params *p, int64_t *pairs_out);
void ora_model_free(ora_model *m);
int32_t ora_model_vocab_size(const ora_model *m);
void ora_model_get(const ora_model *m, float *syn0, float *syn1neg, int32_t *id_of_word);

/* Data-parallel emulation: `world` ranks on contiguous sentence shards, `rounds` delta exchanges per epoch,
 * combine 0 = sum of deltas, 1 = mean, 2 = per-row average over the ranks that touched the row (see the .c file). */
ora_model *ora_sgns_train_dp(const int32_t *tokens, int64_t n_sent, int32_t L, int32_t n_ids,
                             const ora_sgns_params *p, int32_t world, int32_t rounds, int32_t combine,
                             int64_t *pairs_out);

/* Tiered exchange: every exchange covers the hot prefix rows [0, hot_rows) of the frequency-sorted tables, every
 * full_every-th one (and the last) all rows. */
ora_model *ora_sgns_train_dp_tiered(const int32_t *tokens, int64_t n_sent, int32_t L, int32_t n_ids,
                                    const ora_sgns_params *p, int32_t world, int32_t rounds, int32_t combine,
                                    int32_t full_every, int32_t hot_rows, int64_t *pairs_out);

/* Count pairs only (same enumeration, no arithmetic). */
int64_t ora_sgns_count_pairs(const int32_t *tokens, int64_t n_sent, int32_t L, int32_t n_ids,
                             const ora_sgns_params *p);

/* building blocks exposed for tests */
uint64_t ora_sentence_rng(uint64_t seed, int32_t epoch, int64_t sentence);
uint64_t ora_position_rng(uint64_t sentence_key, int32_t position);
uint64_t ora_pair_rng(uint64_t sentence_key, int32_t position, int32_t context);
float ora_alpha(const ora_sgns_params *p, int32_t epoch, int64_t sentence, int64_t n_sent);

#ifdef __cplusplus
}
#endif
#endif
