/*
 * dge_jni.c -- JNI glue between java/embedding/DgeNative.java and libdge.so (include/dge.h): one
 * Java_embedding_DgeNative_<name> body for EVERY native method the class declares (tests/test_jni_glue.py checks the
 * two lists against each other and syntax-checks this file with gcc on the CPU box).
 *
 * The build image has no JDK: the maintainer compiles the shim with the JDK's headers,
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude java/jni/dge_jni.c \
 *       -Lembedding_b200 -ldge -o libdge_jni.so
 * (java/jni/include/jni.h is a declaration subset for `gcc -fsyntax-only` only -- never link against it).
 *
 * Arrays are taken with Get<Type>ArrayElements, not GetPrimitiveArrayCritical: every libdge call blocks on CUDA work,
 * and a critical region must not block (it can stall the garbage collector of every other Java thread).  Input
 * arrays are released with JNI_ABORT (nothing to copy back), output arrays with mode 0.  No Java object is kept
 * across calls; handles travel as long.  A non-zero status becomes a RuntimeException carrying dge_last_error().
 */
#include <jni.h>
#include <stddef.h>
#include <stdint.h>
#include "dge.h"

#define H(type, j) ((type *)(intptr_t)(j))
#define NAT(ret, name) JNIEXPORT ret JNICALL Java_embedding_DgeNative_##name

static void throw_dge(JNIEnv *env, const dge_ctx *ctx) {
    jclass cls = (*env)->FindClass(env, "java/lang/RuntimeException");
    if (cls) (*env)->ThrowNew(env, cls, dge_last_error(ctx));
}
/* input arrays: pointer or NULL for a null reference */
static jint *in_i(JNIEnv *env, jintArray a) { return a ? (*env)->GetIntArrayElements(env, a, NULL) : NULL; }
static void done_i(JNIEnv *env, jintArray a, jint *p, jint mode) { if (a && p) (*env)->ReleaseIntArrayElements(env, a, p, mode); }
static jdouble *in_d(JNIEnv *env, jdoubleArray a) { return a ? (*env)->GetDoubleArrayElements(env, a, NULL) : NULL; }
static void done_d(JNIEnv *env, jdoubleArray a, jdouble *p, jint mode) { if (a && p) (*env)->ReleaseDoubleArrayElements(env, a, p, mode); }
static jfloat *in_f(JNIEnv *env, jfloatArray a) { return a ? (*env)->GetFloatArrayElements(env, a, NULL) : NULL; }
static void done_f(JNIEnv *env, jfloatArray a, jfloat *p, jint mode) { if (a && p) (*env)->ReleaseFloatArrayElements(env, a, p, mode); }
static jlong *in_l(JNIEnv *env, jlongArray a) { return a ? (*env)->GetLongArrayElements(env, a, NULL) : NULL; }
static void done_l(JNIEnv *env, jlongArray a, jlong *p, jint mode) { if (a && p) (*env)->ReleaseLongArrayElements(env, a, p, mode); }

/* ------------------------------------------------------------------ context */
NAT(jlong, create)(JNIEnv *env, jclass c, jint device) {
    (void)c;
    dge_ctx *ctx = NULL;
    if (dge_create(device, &ctx) != DGE_OK) { throw_dge(env, NULL); return 0; }
    return (jlong)(intptr_t)ctx;
}
NAT(void, destroy)(JNIEnv *env, jclass c, jlong ctx) { (void)env; (void)c; dge_destroy(H(dge_ctx, ctx)); }

/* ------------------------------------------------------------------ multi-GPU (one JVM per GPU) */
NAT(jbyteArray, commUniqueId)(JNIEnv *env, jclass c) {
    (void)c;
    jbyte id[DGE_COMM_ID_BYTES];
    if (dge_comm_unique_id(id, sizeof(id)) != DGE_OK) { throw_dge(env, NULL); return NULL; }
    jbyteArray out = (*env)->NewByteArray(env, DGE_COMM_ID_BYTES);
    if (out) (*env)->SetByteArrayRegion(env, out, 0, DGE_COMM_ID_BYTES, id);
    return out;
}
NAT(void, commInit)(JNIEnv *env, jclass c, jlong ctx, jint rank, jint world, jbyteArray jid) {
    (void)c;
    jsize n = (*env)->GetArrayLength(env, jid);
    jbyte *id = (*env)->GetByteArrayElements(env, jid, NULL);
    int rc = dge_comm_init(H(dge_ctx, ctx), rank, world, id, (size_t)n);
    (*env)->ReleaseByteArrayElements(env, jid, id, JNI_ABORT);
    if (rc != DGE_OK) throw_dge(env, H(dge_ctx, ctx));
}

/* ------------------------------------------------------------------ stage 1a: graph + alias tables */
NAT(jlong, graphBuild)(JNIEnv *env, jclass c, jlong jctx, jint nv, jintArray jsrc, jintArray jdst, jdoubleArray jw,
                       jintArray jsources, jdoubleArray jod, jdoubleArray jsws) {
    (void)c;
    dge_ctx *ctx = H(dge_ctx, jctx);
    jsize ne = (*env)->GetArrayLength(env, jsrc), ns = (*env)->GetArrayLength(env, jsources);
    jint *src = in_i(env, jsrc), *dst = in_i(env, jdst), *sources = in_i(env, jsources);
    jdouble *w = in_d(env, jw), *od = in_d(env, jod), *sws = in_d(env, jsws);
    dge_graph *g = NULL;
    int rc = dge_graph_build(ctx, nv, ne, (const int32_t *)src, (const int32_t *)dst, w, ns, (const int32_t *)sources, od, sws, &g);
    done_d(env, jsws, sws, JNI_ABORT); done_d(env, jod, od, JNI_ABORT); done_d(env, jw, w, JNI_ABORT);
    done_i(env, jsources, sources, JNI_ABORT); done_i(env, jdst, dst, JNI_ABORT); done_i(env, jsrc, src, JNI_ABORT);
    if (rc != DGE_OK) { throw_dge(env, ctx); return 0; }
    return (jlong)(intptr_t)g;
}
NAT(void, graphTables)(JNIEnv *env, jclass c, jlong jg, jlongArray jrow, jintArray jcol, jdoubleArray jw, jdoubleArray jprob,
                       jintArray jalias, jdoubleArray jod, jdoubleArray jsp, jintArray jsa, jdoubleArray jsws) {
    (void)c;
    jlong *row = in_l(env, jrow);
    jint *col = in_i(env, jcol), *alias = in_i(env, jalias), *sa = in_i(env, jsa);
    jdouble *w = in_d(env, jw), *prob = in_d(env, jprob), *od = in_d(env, jod), *sp = in_d(env, jsp), *sws = in_d(env, jsws);
    int rc = dge_graph_tables(H(const dge_graph, jg), (int64_t *)row, (int32_t *)col, w, prob, (int32_t *)alias, od, sp, (int32_t *)sa, sws);
    done_d(env, jsws, sws, 0); done_d(env, jsp, sp, 0); done_d(env, jod, od, 0); done_d(env, jprob, prob, 0); done_d(env, jw, w, 0);
    done_i(env, jsa, sa, 0); done_i(env, jalias, alias, 0); done_i(env, jcol, col, 0); done_l(env, jrow, row, 0);
    if (rc != DGE_OK) throw_dge(env, NULL);
}
NAT(void, graphSampleNext)(JNIEnv *env, jclass c, jlong jg, jintArray jv, jdoubleArray jx, jint sampler, jintArray jout) {
    (void)c;
    jsize n = (*env)->GetArrayLength(env, jv);
    jint *v = in_i(env, jv), *out = in_i(env, jout);
    jdouble *x = in_d(env, jx);
    int rc = dge_graph_sample_next(H(const dge_graph, jg), n, (const int32_t *)v, x, sampler, (int32_t *)out);
    done_i(env, jout, out, 0); done_d(env, jx, x, JNI_ABORT); done_i(env, jv, v, JNI_ABORT);
    if (rc != DGE_OK) throw_dge(env, NULL);
}
NAT(void, graphLabels)(JNIEnv *env, jclass c, jlong jg, jintArray jlayer, jintArray jregion, jintArray jsources) {
    (void)c;
    jint *layer = in_i(env, jlayer), *region = in_i(env, jregion), *sources = in_i(env, jsources);
    int rc = dge_graph_labels(H(const dge_graph, jg), (int32_t *)layer, (int32_t *)region, (int32_t *)sources);
    done_i(env, jsources, sources, 0); done_i(env, jregion, region, 0); done_i(env, jlayer, layer, 0);
    if (rc != DGE_OK) throw_dge(env, NULL);
}
NAT(void, graphFree)(JNIEnv *env, jclass c, jlong g) { (void)env; (void)c; dge_graph_free(H(dge_graph, g)); }

/* ------------------------------------------------------------------ stage 1b: walks and the corpus */
NAT(jlong, walk)(JNIEnv *env, jclass c, jlong jg, jlong n, jlong first, jint L, jlong seed, jint sampler) {
    (void)c;
    dge_corpus *corpus = NULL;
    if (dge_walk(H(const dge_graph, jg), n, first, L, (uint64_t)seed, sampler, &corpus) != DGE_OK) { throw_dge(env, NULL); return 0; }
    return (jlong)(intptr_t)corpus;
}
NAT(void, corpusTokens)(JNIEnv *env, jclass c, jlong jc, jintArray jout) {
    (void)c;
    jint *out = in_i(env, jout);
    int rc = dge_corpus_tokens(H(const dge_corpus, jc), (int32_t *)out);
    done_i(env, jout, out, 0);
    if (rc != DGE_OK) throw_dge(env, NULL);
}
NAT(void, corpusTokensU16)(JNIEnv *env, jclass c, jlong jc, jshortArray jout) {
    (void)c;
    jshort *out = jout ? (*env)->GetShortArrayElements(env, jout, NULL) : NULL;
    int rc = dge_corpus_tokens_u16(H(const dge_corpus, jc), (uint16_t *)out);
    if (out) (*env)->ReleaseShortArrayElements(env, jout, out, 0);
    if (rc != DGE_OK) throw_dge(env, NULL);
}
NAT(jlong, corpusCountTokens)(JNIEnv *env, jclass c, jlong jc) {
    (void)c;
    int64_t n = 0;
    if (dge_corpus_count_tokens(H(const dge_corpus, jc), &n) != DGE_OK) { throw_dge(env, NULL); return 0; }
    return (jlong)n;
}
NAT(void, corpusRelabel)(JNIEnv *env, jclass c, jlong jc, jintArray jmap, jint newNIds, jint positionStride) {
    (void)c;
    jint *map = in_i(env, jmap);
    int rc = dge_corpus_relabel(H(dge_corpus, jc), (const int32_t *)map, newNIds, positionStride);
    done_i(env, jmap, map, JNI_ABORT);
    if (rc != DGE_OK) throw_dge(env, NULL);
}
NAT(void, corpusWriteSeq)(JNIEnv *env, jclass c, jlong jc, jintArray jlayer, jintArray jregion, jboolean positionPrefix,
                          jstring jpath, jboolean append) {
    (void)c;
    const char *path = (*env)->GetStringUTFChars(env, jpath, NULL);
    jint *layer = in_i(env, jlayer), *region = in_i(env, jregion);
    int rc = dge_corpus_write_seq(H(const dge_corpus, jc), (const int32_t *)layer, (const int32_t *)region, positionPrefix ? 1 : 0, path, append ? 1 : 0);
    done_i(env, jregion, region, JNI_ABORT); done_i(env, jlayer, layer, JNI_ABORT);
    (*env)->ReleaseStringUTFChars(env, jpath, path);
    if (rc != DGE_OK) throw_dge(env, NULL);
}
NAT(jlong, corpusReadSeq)(JNIEnv *env, jclass c, jlong jctx, jstring jpath, jintArray jlayer, jintArray jregion, jint nIds,
                          jboolean positionPrefix) {
    (void)c;
    const char *path = (*env)->GetStringUTFChars(env, jpath, NULL);
    jint *layer = in_i(env, jlayer), *region = in_i(env, jregion);
    dge_corpus *corpus = NULL;
    int rc = dge_corpus_read_seq(H(dge_ctx, jctx), path, (const int32_t *)layer, (const int32_t *)region, nIds, positionPrefix ? 1 : 0, &corpus);
    done_i(env, jregion, region, JNI_ABORT); done_i(env, jlayer, layer, JNI_ABORT);
    (*env)->ReleaseStringUTFChars(env, jpath, path);
    if (rc != DGE_OK) { throw_dge(env, H(dge_ctx, jctx)); return 0; }
    return (jlong)(intptr_t)corpus;
}
NAT(void, corpusFree)(JNIEnv *env, jclass c, jlong corpus) { (void)env; (void)c; dge_corpus_free(H(dge_corpus, corpus)); }

/* ------------------------------------------------------------------ stage 2: skip-gram */
static jlong train(JNIEnv *env, jlong jctx, jlongArray jcorp, const dge_sgns_params *p) {
    dge_ctx *ctx = H(dge_ctx, jctx);
    jsize n = (*env)->GetArrayLength(env, jcorp);
    jlong *h = in_l(env, jcorp);
    const dge_corpus *corp[4] = {NULL, NULL, NULL, NULL};
    for (jsize i = 0; i < n && i < 4; i++) corp[i] = H(const dge_corpus, h[i]);
    done_l(env, jcorp, h, JNI_ABORT);
    dge_model *m = NULL;
    if (dge_sgns_train(ctx, corp, n, p, &m) != DGE_OK) { throw_dge(env, ctx); return 0; }
    return (jlong)(intptr_t)m;
}
NAT(jlong, sgnsTrain)(JNIEnv *env, jclass c, jlong jctx, jlongArray jcorp, jint dim, jint window, jint negative, jint minCount,
                      jint epochs, jfloat lr, jfloat minLr, jlong seed) {
    (void)c;
    dge_sgns_params p;
    dge_sgns_default_params(&p);
    p.dim = dim; p.window = window; p.negative = negative; p.min_count = minCount; p.epochs = epochs;
    p.lr = lr; p.min_lr = minLr; p.seed = (uint64_t)seed;
    return train(env, jctx, jcorp, &p);
}
NAT(jlong, sgnsTrainDataParallel)(JNIEnv *env, jclass c, jlong jctx, jlongArray jcorp, jint dim, jint window, jint negative,
                                  jint minCount, jint epochs, jfloat lr, jfloat minLr, jlong seed, jint syncRounds, jint combine,
                                  jint transport) {
    (void)c;
    dge_sgns_params p;
    dge_sgns_default_params(&p);
    p.dim = dim; p.window = window; p.negative = negative; p.min_count = minCount; p.epochs = epochs;
    p.lr = lr; p.min_lr = minLr; p.seed = (uint64_t)seed;
    p.sync_rounds = syncRounds; p.combine = combine; p.transport = transport;
    return train(env, jctx, jcorp, &p);
}
NAT(void, modelWriteVec)(JNIEnv *env, jclass c, jlong jm, jintArray jlayer, jintArray jregion, jstring jpath) {
    (void)c;
    const char *path = (*env)->GetStringUTFChars(env, jpath, NULL);
    jint *layer = in_i(env, jlayer), *region = in_i(env, jregion);
    int rc = dge_model_write_vec(H(const dge_model, jm), (const int32_t *)layer, (const int32_t *)region, path);
    done_i(env, jregion, region, JNI_ABORT); done_i(env, jlayer, layer, JNI_ABORT);
    (*env)->ReleaseStringUTFChars(env, jpath, path);
    if (rc != DGE_OK) throw_dge(env, NULL);
}
NAT(void, modelVectors)(JNIEnv *env, jclass c, jlong jm, jfloatArray j0, jfloatArray j1, jintArray jids) {
    (void)c;
    jfloat *s0 = in_f(env, j0), *s1 = in_f(env, j1);
    jint *ids = in_i(env, jids);
    int rc = dge_model_vectors(H(const dge_model, jm), s0, s1, (int32_t *)ids);
    done_i(env, jids, ids, 0); done_f(env, j1, s1, 0); done_f(env, j0, s0, 0);
    if (rc != DGE_OK) throw_dge(env, NULL);
}
NAT(jlong, modelVocabSize)(JNIEnv *env, jclass c, jlong jm) {
    (void)c;
    int32_t V = 0;
    if (dge_model_shape(H(const dge_model, jm), &V, NULL, NULL) != DGE_OK) { throw_dge(env, NULL); return 0; }
    return (jlong)V;
}
NAT(jdoubleArray, modelStats)(JNIEnv *env, jclass c, jlong jm) {
    (void)c;
    double v[3] = {0, 0, 0};
    int64_t bad = 0;
    if (dge_model_stats(H(const dge_model, jm), &v[0], &v[1], &bad) != DGE_OK) { throw_dge(env, NULL); return NULL; }
    v[2] = (double)bad;
    jdoubleArray out = (*env)->NewDoubleArray(env, 3);
    if (out) (*env)->SetDoubleArrayRegion(env, out, 0, 3, v);
    return out;
}
NAT(void, modelFree)(JNIEnv *env, jclass c, jlong m) { (void)env; (void)c; dge_model_free(H(dge_model, m)); }

/* ------------------------------------------------------------------ stage 0: flow records, static exports */
NAT(jlong, flowsCreate)(JNIEnv *env, jclass c, jlong jctx, jint n, jintArray jF) {
    (void)c;
    jint *F = in_i(env, jF);
    dge_flows *f = NULL;
    int rc = dge_flows_create(H(dge_ctx, jctx), n, (const int32_t *)F, &f);
    done_i(env, jF, F, JNI_ABORT);
    if (rc != DGE_OK) { throw_dge(env, H(dge_ctx, jctx)); return 0; }
    return (jlong)(intptr_t)f;
}
NAT(void, flowsAddTrips)(JNIEnv *env, jclass c, jlong jf, jintArray js, jintArray jd, jintArray jh) {
    (void)c;
    jsize n = (*env)->GetArrayLength(env, js);
    jint *s = in_i(env, js), *d = in_i(env, jd), *h = in_i(env, jh);
    int rc = dge_flows_add_trips(H(dge_flows, jf), n, (const int32_t *)s, (const int32_t *)d, (const int32_t *)h);
    done_i(env, jh, h, JNI_ABORT); done_i(env, jd, d, JNI_ABORT); done_i(env, js, s, JNI_ABORT);
    if (rc != DGE_OK) throw_dge(env, NULL);
}
NAT(void, flowsTensor)(JNIEnv *env, jclass c, jlong jf, jintArray jF) {
    (void)c;
    jint *F = in_i(env, jF);
    int rc = dge_flows_tensor(H(const dge_flows, jf), (int32_t *)F);
    done_i(env, jF, F, 0);
    if (rc != DGE_OK) throw_dge(env, NULL);
}
NAT(void, flowsFree)(JNIEnv *env, jclass c, jlong f) { (void)env; (void)c; dge_flows_free(H(dge_flows, f)); }
NAT(jlong, crosstimeGraphBuild)(JNIEnv *env, jclass c, jlong jf, jintArray jorder, jint numLayer, jint mode, jintArray jiv) {
    (void)c;
    jint *order = in_i(env, jorder), *iv = in_i(env, jiv);
    dge_graph *g = NULL;
    int rc = dge_crosstime_graph_build(H(const dge_flows, jf), (const int32_t *)order, numLayer, mode, (const int32_t *)iv, &g);
    done_i(env, jiv, iv, JNI_ABORT); done_i(env, jorder, order, JNI_ABORT);
    if (rc != DGE_OK) { throw_dge(env, NULL); return 0; }
    return (jlong)(intptr_t)g;
}
NAT(void, flowsWriteMatrix)(JNIEnv *env, jclass c, jlong jf, jint mode, jint lo, jint hi, jintArray jrows, jintArray jcols, jchar sep,
                            jstring jpath) {
    (void)c;
    const char *path = (*env)->GetStringUTFChars(env, jpath, NULL);
    jsize nr = (*env)->GetArrayLength(env, jrows), nc = (*env)->GetArrayLength(env, jcols);
    jint *rows = in_i(env, jrows), *cols = in_i(env, jcols);
    int rc = dge_flows_write_matrix(H(const dge_flows, jf), mode, lo, hi, (const int32_t *)rows, nr, (const int32_t *)cols, nc, (char)sep, path);
    done_i(env, jcols, cols, JNI_ABORT); done_i(env, jrows, rows, JNI_ABORT);
    (*env)->ReleaseStringUTFChars(env, jpath, path);
    if (rc != DGE_OK) throw_dge(env, NULL);
}
NAT(void, flowsWriteOd)(JNIEnv *env, jclass c, jlong jf, jint mode, jint lo, jint hi, jintArray jrows, jintArray jcols, jintArray jids,
                        jboolean keepZero, jint presenceHour, jstring jpath) {
    (void)c;
    const char *path = (*env)->GetStringUTFChars(env, jpath, NULL);
    jsize nr = (*env)->GetArrayLength(env, jrows), nc = (*env)->GetArrayLength(env, jcols);
    jint *rows = in_i(env, jrows), *cols = in_i(env, jcols), *ids = in_i(env, jids);
    int rc = dge_flows_write_od(H(const dge_flows, jf), mode, lo, hi, (const int32_t *)rows, nr, (const int32_t *)cols, nc, (const int32_t *)ids,
                                keepZero ? 1 : 0, presenceHour, path);
    done_i(env, jids, ids, JNI_ABORT); done_i(env, jcols, cols, JNI_ABORT); done_i(env, jrows, rows, JNI_ABORT);
    (*env)->ReleaseStringUTFChars(env, jpath, path);
    if (rc != DGE_OK) throw_dge(env, NULL);
}

/* ------------------------------------------------------------------ downstream metric */
NAT(jdouble, evalNdcg)(JNIEnv *env, jclass c, jlong jctx, jfloatArray jx, jint m, jint dim, jintArray jgi, jdoubleArray jgd, jint n, jint topk,
                       jdoubleArray jper) {
    (void)c;
    jfloat *x = in_f(env, jx);
    jint *gi = in_i(env, jgi);
    jdouble *gd = in_d(env, jgd), *per = in_d(env, jper);
    double mean = 0.0;
    int rc = dge_eval_ndcg(H(dge_ctx, jctx), x, m, dim, (const int32_t *)gi, gd, n, topk, per, &mean);
    done_d(env, jper, per, 0); done_d(env, jgd, gd, JNI_ABORT); done_i(env, jgi, gi, JNI_ABORT); done_f(env, jx, x, JNI_ABORT);
    if (rc != DGE_OK) { throw_dge(env, H(dge_ctx, jctx)); return 0.0; }
    return mean;
}
