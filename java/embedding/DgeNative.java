package embedding;

/**
 * JNI declarations over libdge.so (include/dge.h).  Source only: this image has no JDK (SURVEY F6), so the
 * class is not compiled or tested here; tests drive the same C ABI through ctypes (embedding_b200/abi.py).
 *
 * Handles are opaque native pointers carried as long.  Every array argument is a plain Java array; the glue
 * (java/jni/dge_jni.c, one body per native below -- tests/test_jni_glue.py checks the two lists against each other)
 * takes it with Get<Type>ArrayElements for the duration of the call only, so no Java object is retained across calls
 * (SURVEY 8(b) "Ownership").  A non-zero status becomes a RuntimeException carrying dge_last_error().
 */
public final class DgeNative {
    static { System.loadLibrary("dge_jni"); }   // libdge_jni.so links against libdge.so

    private DgeNative() {}

    public static native long create(int device);
    public static native void destroy(long ctx);
    /** multi-GPU, one JVM per GPU: rank 0 makes the id, the host ships it to the other ranks -> dge_comm_unique_id / _init. */
    public static native byte[] commUniqueId();
    public static native void commInit(long ctx, int rank, int world, byte[] id);

    /** LayeredGraph.addEdge / addSourceVertex / initiateAliasTables -> dge_graph_build. */
    public static native long graphBuild(long ctx, int nVertices, int[] src, int[] dst, double[] w, int[] sources,
                                         double[] outDegreeOrNull, double[] sourceWeightSumOrNull);
    /** Vertex.probTable / aliasTable / outDegree and LayeredGraph.probTable / aliasTable -> dge_graph_tables. */
    public static native void graphTables(long graph, long[] rowPtr, int[] col, double[] w, double[] prob, int[] alias,
                                          double[] outDegree, double[] srcProb, int[] srcAlias, double[] sourceWeightSum);
    /** Vertex.sampleNextVertex(double x) -> dge_graph_sample_next. */
    public static native void graphSampleNext(long graph, int[] v, double[] x, int sampler, int[] out);
    public static native void graphFree(long graph);

    /** the loop of CrossTimeGraph.sampleSequenceHelper / SpatialGraph.outputSampleSequence -> dge_walk. */
    public static native long walk(long graph, long nWalks, long firstWalkId, int numLayer, long seed, int sampler);
    public static native void corpusTokens(long corpus, int[] tokensOut);
    /** 16-bit tokens (0xFFFF = padding) for id spaces below 65 535: half the PCIe bytes -> dge_corpus_tokens_u16. */
    public static native void corpusTokensU16(long corpus, short[] tokensOut);
    public static native long corpusCountTokens(long corpus);
    public static native void corpusRelabel(long corpus, int[] idMap, int newNIds, int positionStride);
    public static native void corpusWriteSeq(long corpus, int[] labelLayer, int[] labelRegion, boolean positionPrefix,
                                             String path, boolean append);
    /** an existing `.seq` file back into a device corpus (DeepWalk.checkInputFile skips generation) -> dge_corpus_read_seq. */
    public static native long corpusReadSeq(long ctx, String path, int[] labelLayer, int[] labelRegion, int nIds, boolean positionPrefix);
    public static native void corpusFree(long corpus);

    /** Word2Vec.Builder()...build().fit() -> dge_sgns_train; writeWordVectors -> dge_model_write_vec. */
    public static native long sgnsTrain(long ctx, long[] corpora, int dim, int window, int negative, int minCount,
                                        int epochs, float lr, float minLr, long seed);
    /** the same on a ctx with a communicator: collective, embedding deltas exchanged syncRounds times per epoch. */
    public static native long sgnsTrainDataParallel(long ctx, long[] corpora, int dim, int window, int negative, int minCount,
                                                    int epochs, float lr, float minLr, long seed, int syncRounds, int combine, int transport);
    public static native void modelWriteVec(long model, int[] labelLayer, int[] labelRegion, String path);
    /** in-memory access to the trained tables (syn1negOrNull / idOfWordOrNull may be null) -> dge_model_vectors. */
    public static native void modelVectors(long model, float[] syn0, float[] syn1negOrNull, int[] idOfWordOrNull);
    /** {mean |syn0 row|, max |element|, non-finite elements} computed on the device -> dge_model_stats. */
    public static native double[] modelStats(long model);
    public static native long modelVocabSize(long model);
    public static native void modelFree(long model);

    /** CommunityAreas.mapTripsIntoCommunities / Tracts.mapTripsIntoTracts (counting) -> dge_flows_create / _add_trips. */
    public static native long flowsCreate(long ctx, int nRegions, int[] flowTensorOrNull);
    public static native void flowsAddTrips(long flows, int[] srcRegion, int[] dstRegion, int[] startHour);
    public static native void flowsTensor(long flows, int[] flowTensorOut);
    public static native void flowsFree(long flows);
    /** CrossTimeGraph.constructGraph_CA(int[]) (mode 0) / constructGraph_tract() (mode 1) + initiateAliasTables. */
    public static native long crosstimeGraphBuild(long flows, int[] order, int numLayer, int mode, int[] intervalsOrNull);
    public static native void graphLabels(long graph, int[] vLayer, int[] vRegionIndex, int[] sources);
    /** outputStaticFlowGraph / outputAdjacencyMatrix / outputEdgeGraph_LINE / outputEdgeFile -> dge_flows_write_*. */
    public static native void flowsWriteMatrix(long flows, int mode, int lo, int hi, int[] rows, int[] cols, char sep, String path);
    public static native void flowsWriteOd(long flows, int mode, int lo, int hi, int[] rows, int[] cols, int[] regionIds,
                                           boolean keepZero, int presenceHour, String path);
    /** python/embeddingEvaluation_tract.py pairwiseEstimator + ndcg_atK -> dge_eval_ndcg (returns the mean). */
    public static native double evalNdcg(long ctx, float[] x, int m, int dim, int[] gtIndex, double[] gtDist, int n, int topk,
                                         double[] ndcgPerRowOrNull);
}
