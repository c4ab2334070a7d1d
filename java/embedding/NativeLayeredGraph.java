package embedding;

import java.util.ArrayList;
import java.util.HashMap;
import java.util.List;
import java.util.Map;

/**
 * Drop-in core for the reference's LayeredGraph: same public methods (addEdge, addSourceVertex,
 * initiateAliasTables, sampleVertexSequence), same String vertex names, but the adjacency lists are kept as
 * primitive COO arrays and all heavy work (CSR, alias tables, walks) runs in libdge.so on the GPU.
 * The Java host keeps what it owns in the reference: names, first-appearance ids, insertion order.
 *
 * Source only (no JDK in the build image).  Written from the reference's public interface
 * (LayeredGraph.java:157-252); it shares no code with it.
 */
public class NativeLayeredGraph {
    public static int numLayer = 8;
    public static long seed = 2013L;          // replaces the unseeded static Random

    protected final long ctx;
    protected final Map<String, Integer> ids = new HashMap<>();
    protected final List<String> names = new ArrayList<>();
    protected int[] src = new int[1024], dst = new int[1024];
    protected double[] w = new double[1024];
    protected int nEdges = 0;
    protected final List<Integer> sources = new ArrayList<>();
    protected double[] outDegreeOverride = null;      // SpatialGraph.keepNearestKVertices recomputes it
    protected double[] sourceWeightSumOverride = null;
    protected long graph = 0L;
    protected long walksDrawn = 0L;

    public NativeLayeredGraph(long ctx) { this.ctx = ctx; }

    private int id(String name) {
        Integer i = ids.get(name);
        if (i == null) { i = names.size(); ids.put(name, i); names.add(name); }
        return i;
    }

    public void addEdge(String fn, String tn, double weight) {
        int f = id(fn), t = id(tn);
        if (nEdges == src.length) {
            src = java.util.Arrays.copyOf(src, 2 * nEdges);
            dst = java.util.Arrays.copyOf(dst, 2 * nEdges);
            w = java.util.Arrays.copyOf(w, 2 * nEdges);
        }
        src[nEdges] = f; dst[nEdges] = t; w[nEdges] = weight; nEdges++;
    }

    public void addSourceVertex(String vn) { sources.add(id(vn)); }

    public void initiateAliasTables() {
        int[] s = sources.stream().mapToInt(Integer::intValue).toArray();
        if (graph != 0L) DgeNative.graphFree(graph);
        graph = DgeNative.graphBuild(ctx, names.size(), java.util.Arrays.copyOf(src, nEdges),
                java.util.Arrays.copyOf(dst, nEdges), java.util.Arrays.copyOf(w, nEdges), s,
                outDegreeOverride, sourceWeightSumOverride);
    }

    /** numSamples x sampleVertexSequence() in one launch; the corpus stays on the device. */
    public long sample(long numSamples) {
        long c = DgeNative.walk(graph, numSamples, walksDrawn, numLayer, seed, 0);
        walksDrawn += numSamples;
        return c;
    }

    public List<String> sampleVertexSequence() {
        long c = sample(1);
        int[] tok = new int[numLayer];
        DgeNative.corpusTokens(c, tok);
        DgeNative.corpusFree(c);
        List<String> seq = new ArrayList<>();
        for (int t : tok) { if (t < 0) break; seq.add(names.get(t)); }
        return seq;
    }
}
