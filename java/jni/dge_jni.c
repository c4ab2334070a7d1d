/*
 * dge_jni.c -- JNI glue between java/embedding/DgeNative.java and libdge.so (include/dge.h).
 * Source only: jni.h does not exist in the build image; compiled by the maintainer with
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -Iinclude java/jni/dge_jni.c \
 *       -Lembedding_b200 -ldge -o libdge_jni.so
 * The graph / walk / train / flows entry points are shown in full; the remaining ones (tables, labels, exports,
 * evaluation) pin their arrays and forward in exactly the same way.
 */
#if defined(__has_include)
#if __has_include(<jni.h>)
#include <jni.h>
#include "dge.h"

static void throw_dge(JNIEnv *env, dge_ctx *ctx) {
    jclass cls = (*env)->FindClass(env, "java/lang/RuntimeException");
    (*env)->ThrowNew(env, cls, dge_last_error(ctx));
}

JNIEXPORT jlong JNICALL Java_embedding_DgeNative_create(JNIEnv *env, jclass c, jint device) {
    dge_ctx *ctx = NULL;
    if (dge_create(device, &ctx) != DGE_OK) { throw_dge(env, NULL); return 0; }
    return (jlong)(intptr_t)ctx;
}

JNIEXPORT jlong JNICALL Java_embedding_DgeNative_graphBuild(JNIEnv *env, jclass c, jlong jctx, jint nv, jintArray jsrc,
        jintArray jdst, jdoubleArray jw, jintArray jsources, jdoubleArray jod, jdoubleArray jsws) {
    dge_ctx *ctx = (dge_ctx *)(intptr_t)jctx;
    jsize ne = (*env)->GetArrayLength(env, jsrc), ns = (*env)->GetArrayLength(env, jsources);
    jint *src = (*env)->GetPrimitiveArrayCritical(env, jsrc, NULL);
    jint *dst = (*env)->GetPrimitiveArrayCritical(env, jdst, NULL);
    jdouble *w = (*env)->GetPrimitiveArrayCritical(env, jw, NULL);
    jint *sources = (*env)->GetPrimitiveArrayCritical(env, jsources, NULL);
    jdouble *od = jod ? (*env)->GetPrimitiveArrayCritical(env, jod, NULL) : NULL;
    jdouble *sws = jsws ? (*env)->GetPrimitiveArrayCritical(env, jsws, NULL) : NULL;
    dge_graph *g = NULL;
    int rc = dge_graph_build(ctx, nv, ne, (const int32_t *)src, (const int32_t *)dst, w, ns, (const int32_t *)sources, od, sws, &g);
    if (sws) (*env)->ReleasePrimitiveArrayCritical(env, jsws, sws, JNI_ABORT);
    if (od) (*env)->ReleasePrimitiveArrayCritical(env, jod, od, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jsources, sources, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jw, w, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jdst, dst, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jsrc, src, JNI_ABORT);
    if (rc != DGE_OK) { throw_dge(env, ctx); return 0; }
    return (jlong)(intptr_t)g;
}

JNIEXPORT jlong JNICALL Java_embedding_DgeNative_walk(JNIEnv *env, jclass c, jlong jg, jlong n, jlong first, jint L,
        jlong seed, jint sampler) {
    dge_corpus *corpus = NULL;
    if (dge_walk((dge_graph *)(intptr_t)jg, n, first, L, (uint64_t)seed, sampler, &corpus) != DGE_OK) { throw_dge(env, NULL); return 0; }
    return (jlong)(intptr_t)corpus;
}

JNIEXPORT void JNICALL Java_embedding_DgeNative_corpusTokensU16(JNIEnv *env, jclass c, jlong jc, jshortArray jout) {
    jshort *out = (*env)->GetPrimitiveArrayCritical(env, jout, NULL);
    int rc = dge_corpus_tokens_u16((const dge_corpus *)(intptr_t)jc, (uint16_t *)out);
    (*env)->ReleasePrimitiveArrayCritical(env, jout, out, 0);
    if (rc != DGE_OK) throw_dge(env, NULL);
}

JNIEXPORT jlong JNICALL Java_embedding_DgeNative_sgnsTrain(JNIEnv *env, jclass c, jlong jctx, jlongArray jcorp, jint dim,
        jint window, jint negative, jint minCount, jint epochs, jfloat lr, jfloat minLr, jlong seed) {
    dge_ctx *ctx = (dge_ctx *)(intptr_t)jctx;
    jsize n = (*env)->GetArrayLength(env, jcorp);
    jlong *h = (*env)->GetLongArrayElements(env, jcorp, NULL);
    const dge_corpus *corp[4];
    for (jsize i = 0; i < n && i < 4; i++) corp[i] = (const dge_corpus *)(intptr_t)h[i];
    (*env)->ReleaseLongArrayElements(env, jcorp, h, JNI_ABORT);
    dge_sgns_params p;
    dge_sgns_default_params(&p);
    p.dim = dim; p.window = window; p.negative = negative; p.min_count = minCount; p.epochs = epochs;
    p.lr = lr; p.min_lr = minLr; p.seed = (uint64_t)seed;
    dge_model *m = NULL;
    if (dge_sgns_train(ctx, corp, n, &p, &m) != DGE_OK) { throw_dge(env, ctx); return 0; }
    return (jlong)(intptr_t)m;
}

JNIEXPORT void JNICALL Java_embedding_DgeNative_modelVectors(JNIEnv *env, jclass c, jlong jm, jfloatArray j0, jfloatArray j1,
        jintArray jids) {
    jfloat *s0 = (*env)->GetPrimitiveArrayCritical(env, j0, NULL);
    jfloat *s1 = j1 ? (*env)->GetPrimitiveArrayCritical(env, j1, NULL) : NULL;
    jint *ids = jids ? (*env)->GetPrimitiveArrayCritical(env, jids, NULL) : NULL;
    int rc = dge_model_vectors((const dge_model *)(intptr_t)jm, s0, s1, (int32_t *)ids);
    if (ids) (*env)->ReleasePrimitiveArrayCritical(env, jids, ids, 0);
    if (s1) (*env)->ReleasePrimitiveArrayCritical(env, j1, s1, 0);
    (*env)->ReleasePrimitiveArrayCritical(env, j0, s0, 0);
    if (rc != DGE_OK) throw_dge(env, NULL);
}

JNIEXPORT jdoubleArray JNICALL Java_embedding_DgeNative_modelStats(JNIEnv *env, jclass c, jlong jm) {
    double v[3] = {0, 0, 0};
    int64_t bad = 0;
    if (dge_model_stats((const dge_model *)(intptr_t)jm, &v[0], &v[1], &bad) != DGE_OK) { throw_dge(env, NULL); return NULL; }
    v[2] = (double)bad;
    jdoubleArray out = (*env)->NewDoubleArray(env, 3);
    if (out) (*env)->SetDoubleArrayRegion(env, out, 0, 3, v);
    return out;
}

JNIEXPORT jlong JNICALL Java_embedding_DgeNative_flowsCreate(JNIEnv *env, jclass c, jlong jctx, jint n, jintArray jF) {
    dge_ctx *ctx = (dge_ctx *)(intptr_t)jctx;
    jint *F = jF ? (*env)->GetPrimitiveArrayCritical(env, jF, NULL) : NULL;
    dge_flows *f = NULL;
    int rc = dge_flows_create(ctx, n, (const int32_t *)F, &f);
    if (F) (*env)->ReleasePrimitiveArrayCritical(env, jF, F, JNI_ABORT);
    if (rc != DGE_OK) { throw_dge(env, ctx); return 0; }
    return (jlong)(intptr_t)f;
}

JNIEXPORT void JNICALL Java_embedding_DgeNative_flowsAddTrips(JNIEnv *env, jclass c, jlong jf, jintArray js, jintArray jd,
        jintArray jh) {
    jsize n = (*env)->GetArrayLength(env, js);
    jint *s = (*env)->GetPrimitiveArrayCritical(env, js, NULL);
    jint *d = (*env)->GetPrimitiveArrayCritical(env, jd, NULL);
    jint *h = (*env)->GetPrimitiveArrayCritical(env, jh, NULL);
    int rc = dge_flows_add_trips((dge_flows *)(intptr_t)jf, n, (const int32_t *)s, (const int32_t *)d, (const int32_t *)h);
    (*env)->ReleasePrimitiveArrayCritical(env, jh, h, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jd, d, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, js, s, JNI_ABORT);
    if (rc != DGE_OK) throw_dge(env, NULL);
}

JNIEXPORT jlong JNICALL Java_embedding_DgeNative_crosstimeGraphBuild(JNIEnv *env, jclass c, jlong jf, jintArray jorder,
        jint numLayer, jint mode, jintArray jiv) {
    jint *order = (*env)->GetPrimitiveArrayCritical(env, jorder, NULL);
    jint *iv = jiv ? (*env)->GetPrimitiveArrayCritical(env, jiv, NULL) : NULL;
    dge_graph *g = NULL;
    int rc = dge_crosstime_graph_build((const dge_flows *)(intptr_t)jf, (const int32_t *)order, numLayer, mode, (const int32_t *)iv, &g);
    if (iv) (*env)->ReleasePrimitiveArrayCritical(env, jiv, iv, JNI_ABORT);
    (*env)->ReleasePrimitiveArrayCritical(env, jorder, order, JNI_ABORT);
    if (rc != DGE_OK) { throw_dge(env, NULL); return 0; }
    return (jlong)(intptr_t)g;
}
#endif
#endif
