"""Host-side mirror of the reference's Java entry points over the C ABI (libdge.so).

The reference's host language is Java; no JDK exists in this image, so the host side the
tests and bench drive is this Python mirror: same class / method / static-knob names, same argument
meaning and error behaviour as
    LayeredGraph.java, CrossTimeGraph.java, SpatialGraph.java, DeepWalk.java
(embedding/src/main/java/embedding/ of the reference).  java/ holds the JNI shim a maintainer drops
into the unchanged Java classes.  All heavy work happens in libdge.so; nothing here
computes tables, walks or SGD on the CPU.

Differences forced by the missing ingest layer (out of scope): the deserialised
`CommunityAreas` / `Tracts` objects are replaced by a `Flows` value (region ids in the host's HashMap
iteration order + the dense F[src, hour, dst] count tensor).
"""
import os

import numpy as np

from . import abi

_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = abi.Context()
    return _default_ctx


def java_hashmap_order(keys):
    """Iteration order of a java.util.HashMap<Integer, V> filled with `keys` in this insertion order
    (JDK 8: bucket index (h ^ h>>>16) & (cap-1), insertion order inside a bin, resize keeps relative
    order; bins here never reach the treeify threshold).  SURVEY Q5: this order defines edge order."""
    keys = [int(k) for k in keys]
    cap = 16
    while len(keys) > 0.75 * cap:
        cap *= 2
    buckets = {}
    for k in keys:
        h = k & 0xFFFFFFFF
        b = (h ^ (h >> 16)) & (cap - 1)
        buckets.setdefault(b, []).append(k)
    return [k for b in sorted(buckets) for k in buckets[b]]


class Flows:
    """Region ids (iteration order) + F[src, hour, dst] int32; stands in for CommunityAreas / Tracts."""

    def __init__(self, region_ids, F, order=None):
        self.region_ids = np.asarray(region_ids, np.int32)
        self.F = np.ascontiguousarray(F, np.int32)
        n = len(self.region_ids)
        assert self.F.shape == (n, 24, n)
        if order is None:  # indices into region_ids in java.util.HashMap iteration order
            pos = {int(r): i for i, r in enumerate(self.region_ids)}
            order = [pos[k] for k in java_hashmap_order(self.region_ids)]
        self.order = np.asarray(order, np.int32)

    @classmethod
    def from_trips(cls, region_ids, src_index, dst_index, start_hour, ctx=None, order=None):
        """CommunityAreas.mapTripsIntoCommunities :55-103 / Tracts.mapTripsIntoTracts :71-102 after the host's
        point-in-polygon lookup: the trips are counted into F on the device (dge_flows_add_trips)."""
        ctx = ctx or default_context()
        n = len(region_ids)
        dev = abi.Flows(ctx, n)
        dev.add_trips(src_index, dst_index, start_hour)
        fl = cls(region_ids, dev.tensor(), order)
        fl._dev = (ctx, dev)
        return fl

    def device(self, ctx):
        """The same counts as a device tensor (uploaded once per ctx)."""
        cached = getattr(self, "_dev", None)
        if cached is None or cached[0] is not ctx or cached[1]._h is None:
            self._dev = (ctx, abi.Flows(ctx, len(self.region_ids), self.F))
        return self._dev[1]

    def slot_weights_ca(self, lo, hi):
        """CommunityArea.getFlowTo(dst, lo, hi) for all (src, dst): circular half-open [lo, hi)."""
        W = np.zeros(self.F.shape[::2], np.int64)
        h = lo
        while h != hi:
            W += self.F[:, h, :]
            h = (h + 1) % 24
        return W

    def slot_weights_tract(self, lo, hi):
        """Tract.getFlowTo(dst, lo, hi) for all (src, dst): inclusive [lo, hi]."""
        return self.F[:, lo:hi + 1, :].sum(axis=1, dtype=np.int64)


    # ---- static flow-graph exports for the LINE / matrix baselines.  `out_dir` stands for the
    # reference's "../miscs/<year>"; file names, separators, loop bounds and the hour ranges are the Java ones.
    # The slot sums run on the device (dge_flows_slot_weights); ctx defaults to default_context().

    def _by_id(self, ids):
        pos = {int(r): i for i, r in enumerate(self.region_ids)}
        return np.array([pos[int(i)] for i in ids], np.int32)

    def outputStaticFlowGraph(self, out_dir, ctx=None):
        """CommunityAreas.outputStaticFlowGraph :127-145: taxi-CA-static.matrix (',' joined) and taxi-CA-static.od
        (w > 0) over ids 1..n with getFlowTo(j, 0, 23) -- circular half-open, i.e. hours 0..22 as in the reference."""
        dev = self.device(ctx or default_context())
        idx = self._by_id(range(1, len(self.region_ids) + 1))
        dev.write_matrix(0, 0, 23, idx, idx, ",", os.path.join(out_dir, "taxi-CA-static.matrix"))
        dev.write_od(0, 0, 23, idx, idx, self.region_ids, os.path.join(out_dir, "taxi-CA-static.od"))

    def outputAdjacencyMatrix_CA(self, out_dir, ctx=None):
        """CommunityAreas.outputAdjacencyMatrix :147-164: taxi-CA-h<hour>.matrix, ' ' joined, single-hour flows."""
        dev = self.device(ctx or default_context())
        idx = self._by_id(range(1, len(self.region_ids) + 1))
        for hour in range(24):
            dev.write_matrix(1, hour, hour, idx, idx, " ", os.path.join(out_dir, "taxi-CA-h%d.matrix" % hour))

    def outputEdgeGraph_LINE(self, out_dir, ctx=None):
        """CommunityAreas.outputEdgeGraph_LINE :169-184: taxi-CA-h<hour>.od with every (i, j) pair, zeros included;
        the Java inner loop runs `j < communities.size()`, so the last destination id is never written."""
        dev = self.device(ctx or default_context())
        n = len(self.region_ids)
        idx = self._by_id(range(1, n + 1))
        for hour in range(24):
            dev.write_od(1, hour, hour, idx, idx[:n - 1], self.region_ids, os.path.join(out_dir, "taxi-CA-h%d.od" % hour),
                         keep_zero=True)

    def outputEdgeFile(self, out_dir, numTimeSlot=8, ctx=None):
        """Tracts.outputEdgeFile :236-260: taxi-h<h>.od per time slot (sources in HashMap iteration order, destinations
        = those with a trip in hour h itself, weight = getFlowTo(dst, h, h+timeStep-1) > 0) and taxi-all.od (hours
        0..23).  Destinations of one source are written in the host's region iteration order; the reference uses the
        per-hour HashMap<Integer,Integer> key order, which depends on trip arrival order (consumers are order-free)."""
        dev = self.device(ctx or default_context())
        timeStep = 24 // numTimeSlot
        for h in range(numTimeSlot):
            dev.write_od(1, h, h + timeStep - 1, self.order, self.order, self.region_ids,
                         os.path.join(out_dir, "taxi-h%d.od" % h), presence_hour=h)
        dev.write_od(1, 0, 23, self.order, self.order, self.region_ids, os.path.join(out_dir, "taxi-all.od"))

    def outputAdjacencyMatrix_tract(self, out_dir, numTimeSlot=8, ctx=None):
        """Tracts.outputAdjacencyMatrix :265-301: taxi-h<h>.matrix per slot and taxi-all.matrix, ids sorted, ',' joined."""
        dev = self.device(ctx or default_context())
        idx = self._by_id(sorted(int(r) for r in self.region_ids))
        timeStep = 24 // numTimeSlot
        for h in range(numTimeSlot):
            dev.write_matrix(1, h, h + timeStep - 1, idx, idx, ",", os.path.join(out_dir, "taxi-h%d.matrix" % h))
        dev.write_matrix(1, 0, 23, idx, idx, ",", os.path.join(out_dir, "taxi-all.matrix"))


class Vertex:
    """View of LayeredGraph.Vertex (LayeredGraph.java:29-133) after initiateAliasTables()."""

    def __init__(self, graph, name, vid):
        self._g, self.name, self.id = graph, name, vid

    @property
    def outDegree(self):
        return float(self._g._tables()["out_degree"][self.id])

    def _row(self):
        t = self._g._tables()
        return int(t["row_ptr"][self.id]), int(t["row_ptr"][self.id + 1])

    @property
    def probTable(self):
        b, e = self._row()
        return self._g._tables()["prob"][b:e]

    @property
    def aliasTable(self):
        b, e = self._row()
        return self._g._tables()["alias"][b:e]

    @property
    def edgesOut(self):
        b, e = self._row()
        t = self._g._tables()
        return [(self._g._names[int(c)], float(w)) for c, w in zip(t["col"][b:e], t["w"][b:e])]

    def sampleNextVertex(self, x=None):
        """Vertex.sampleNextVertex(double x) :123-132; x=None draws the uniform from the graph's stream."""
        if x is None:
            x = self._g._next_uniform()
        nxt = int(self._g._graph.sample_next([self.id], [x])[0])
        return None if nxt < 0 else self._g.vertex(self._g._names[nxt])

    def sampleNextVertex_OV(self, x=None):
        if x is None:
            x = self._g._next_uniform()
        nxt = int(self._g._graph.sample_next([self.id], [x], abi.SAMPLER_CDF)[0])
        return None if nxt < 0 else self._g.vertex(self._g._names[nxt])


class LayeredGraph:
    """LayeredGraph.java:142-281.  Names are Strings as in the reference; `add_edges_bulk` is the
    integer-id fast path for graphs too large for per-edge Python calls."""

    numLayer = 8        # LayeredGraph.java:15
    seed = 2013         # replaces the unseeded `rnd` (LayeredGraph.java:14) with a Philox key

    def __init__(self, ctx=None):
        self._ctx = ctx  # created lazily: building the edge list needs no GPU
        self._ids = {}
        self._names = []
        self._src, self._dst, self._w = [], [], []
        self._bulk = None
        self.sourceVertices = []
        self._out_degree_override = None
        self._sws_override = None
        self._graph = None
        self._tab = None
        self._walks_drawn = 0
        self._uniforms_drawn = 0
        self.v_layer = None
        self.v_region = None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    # ---- construction
    def _vid(self, name):
        i = self._ids.get(name)
        if i is None:
            i = len(self._names)
            self._ids[name] = i
            self._names.append(name)
        return i

    def addEdge(self, fn, tn, weight):
        """LayeredGraph.addEdge :157-174."""
        f = self._vid(fn)
        t = self._vid(tn)
        self._src.append(f)
        self._dst.append(t)
        self._w.append(float(weight))
        self._graph = None

    def add_edges_bulk(self, n_vertices, src, dst, w, names=None):
        self._bulk = (int(n_vertices), np.asarray(src, np.int32), np.asarray(dst, np.int32),
                      np.asarray(w, np.float64))
        if names is not None:
            self._names = list(names)
            self._ids = {n: i for i, n in enumerate(self._names)}
        self._graph = None

    def addSourceVertex(self, vn):
        """LayeredGraph.addSourceVertex :180-189 (call after all edges are added)."""
        self.sourceVertices.append(self._vid(vn) if not isinstance(vn, (int, np.integer)) else int(vn))
        self._graph = None

    @property
    def allVertices(self):
        return {n: self.vertex(n) for n in self._names}

    def vertex(self, name):
        return Vertex(self, name, self._ids[name])

    @property
    def n_vertices(self):
        if getattr(self, "_device_built", False):
            return self._graph.nv
        return self._bulk[0] if self._bulk is not None else len(self._names)

    def initiateAliasTables(self):
        """LayeredGraph.initiateAliasTables :195-226 -> dge_graph_build (CSR + all alias tables on the GPU)."""
        if getattr(self, "_device_built", False) and self._graph is not None:
            return self                      # dge_crosstime_graph_build already built the tables
        if self._bulk is not None:
            nv, src, dst, w = self._bulk
        else:
            nv, src, dst, w = len(self._names), self._src, self._dst, self._w
        if self._graph is not None:
            self._graph.free()
        self._graph = abi.Graph(self.ctx, nv, src, dst, w, self.sourceVertices,
                                out_degree=self._out_degree_override, source_weight_sum=self._sws_override)
        self._tab = None
        return self

    def _tables(self):
        if self._graph is None:
            raise RuntimeError("initiateAliasTables() has not been called")  # Java: NullPointerException
        if self._tab is None:
            self._tab = self._graph.tables()
        return self._tab

    @property
    def probTable(self):
        return self._tables()["src_prob"]

    @property
    def aliasTable(self):
        return self._tables()["src_alias"]

    @property
    def sourceWeightSum(self):
        return self._tables()["source_weight_sum"]

    # ---- sampling
    def _next_uniform(self):
        # single-draw test hook: a dedicated Philox stream (walk id 2^62) so it never collides with walks
        from . import philox
        x = philox.uniform(self.seed, (1 << 62), self._uniforms_drawn)
        self._uniforms_drawn += 1
        return x

    def sample(self, n_walks, sampler=abi.SAMPLER_ALIAS, num_layer=None):
        """n_walks x sampleVertexSequence() in one launch -> device corpus."""
        if self._graph is None:
            raise RuntimeError("initiateAliasTables() has not been called")
        L = type(self).numLayer if num_layer is None else num_layer
        c = self._graph.walk(n_walks, L, self.seed, sampler, first_walk_id=self._walks_drawn)
        self._walks_drawn += n_walks
        return c

    def _seq_names(self, tokens):
        return [self._names[t] if self._names else int(t) for t in tokens if t >= 0]

    def sampleVertexSequence(self):
        """LayeredGraph.sampleVertexSequence :232-252 -> List<String>."""
        c = self.sample(1, abi.SAMPLER_ALIAS, LayeredGraph.numLayer)
        return self._seq_names(c.tokens()[0])

    def sampleVertexSequence_OV(self):
        """LayeredGraph.sampleVertexSequence_OV :260-279 (the reference's slow CDF baseline)."""
        c = self.sample(1, abi.SAMPLER_CDF, LayeredGraph.numLayer)
        return self._seq_names(c.tokens()[0])


def _first_appearance_ids(src_keys, dst_keys):
    """Vertex ids by first appearance in the interleaved mention order s0,d0,s1,d1,... (addEdge :159-170)."""
    inter = np.empty(2 * len(src_keys), np.int64)
    inter[0::2] = src_keys
    inter[1::2] = dst_keys
    uniq, first = np.unique(inter, return_index=True)
    rank = np.empty(len(uniq), np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
    ids = rank[np.searchsorted(uniq, inter)]
    keys_by_id = uniq[np.argsort(first, kind="stable")]
    return ids[0::2].astype(np.int32), ids[1::2].astype(np.int32), keys_by_id


class CrossTimeGraph(LayeredGraph):
    """CrossTimeGraph.java:16-161."""

    numSamples = 10_000_000          # :18
    numLayer = LayeredGraph.numLayer  # :19 (a separate static that shadows the parent's)

    @classmethod
    def _construct(cls, flows, slot_weights, L, ctx=None):
        n = len(flows.region_ids)
        o = flows.order.astype(np.int64)
        srcs, dsts, ws = [], [], []
        for h in range(L):
            W = slot_weights(h)[np.ix_(o, o)]        # rows/cols in the host's iteration order
            a, b = np.nonzero(W > 0)                 # row-major == (src loop, dst loop) order
            srcs.append(h * n + o[a])
            dsts.append(((h + 1) % L) * n + o[b])
            ws.append(W[a, b].astype(np.float64))
        src_k = np.concatenate(srcs) if srcs else np.empty(0, np.int64)
        dst_k = np.concatenate(dsts) if dsts else np.empty(0, np.int64)
        w = np.concatenate(ws) if ws else np.empty(0, np.float64)
        src, dst, keys = _first_appearance_ids(src_k, dst_k)
        g = cls(ctx)
        g.v_layer = (keys // n).astype(np.int32)
        g.v_region = flows.region_ids[keys % n].astype(np.int32)
        names = ["%d-%d" % (l, r) for l, r in zip(g.v_layer, g.v_region)]
        g.add_edges_bulk(len(keys), src, dst, w, names)
        # sources: layer-0 vertices present in the vertex map, in region iteration order (:43-47, :86-90)
        for r in o:
            name = "0-%d" % flows.region_ids[r]
            if name in g._ids:
                g.addSourceVertex(name)
        return g

    @classmethod
    def _construct_device(cls, flows, L, mode, intervals, ctx):
        """Same graph, enumerated on the GPU from the device-resident flow tensor (dge_crosstime_graph_build):
        slot sums, edges with w > 0 in (h, src, dst) order, first-appearance vertex ids, sources, CSR and alias
        tables in one call -- nothing but names is produced on the host."""
        ctx = ctx or default_context()
        g = cls(ctx)
        g._graph = flows.device(ctx).crosstime_graph(flows.order, L, mode, intervals)
        vl, vr, so = g._graph.labels()
        g.v_layer, g.v_region = vl, flows.region_ids[vr].astype(np.int32)
        g._names = ["%d-%d" % (l, r) for l, r in zip(g.v_layer, g.v_region)]
        g._ids = {nm: i for i, nm in enumerate(g._names)}
        g.sourceVertices = [int(x) for x in so]
        g._device_built = True
        return g

    @classmethod
    def constructGraph_tract(cls, flows, ctx=None, on_device=False):
        """CrossTimeGraph.constructGraph_tract :25-52 (window [h, h+timeStep-1] inclusive).
        on_device=False enumerates the COO on the host (what a Java host would pass to dge_graph_build);
        on_device=True leaves the whole construction to the GPU."""
        L = cls.numLayer
        time_step = 24 // L
        if on_device:
            return cls._construct_device(flows, L, 1, None, ctx)
        return cls._construct(flows, lambda h: flows.slot_weights_tract(h, h + time_step - 1), L, ctx)

    @classmethod
    def constructGraph_CA(cls, flows, timeIntervals=None, ctx=None, on_device=False):
        """CrossTimeGraph.constructGraph_CA() :54-61 and constructGraph_CA(int[]) :68-95."""
        if timeIntervals is None:
            L = cls.numLayer
            time_step = 24 // L
            timeIntervals = [0] * (L + 1)
            for i in range(0, L + 1, time_step):       # :57-58, including its quirk for timeStep != 1 (Q1)
                timeIntervals[i] = (i * time_step) % L
        cls.numLayer = len(timeIntervals) - 1           # :71
        ti = list(timeIntervals)
        if on_device:
            return cls._construct_device(flows, cls.numLayer, 0, ti, ctx)
        return cls._construct(flows, lambda h: flows.slot_weights_ca(ti[h], ti[h + 1]), cls.numLayer, ctx)

    @classmethod
    def outputSampleSequence(cls, regionLevel, flows, path, timeIntervals=None, ctx=None):
        """CrossTimeGraph.outputSampleSequence :103-124 + sampleSequenceHelper :127-148: writes the .seq file
        and returns (graph, device corpus) for the in-memory hand-off to DeepWalk."""
        LayeredGraph.numLayer = cls.numLayer             # :104 / :116
        if regionLevel == "tract":
            g = cls.constructGraph_tract(flows, ctx, on_device=True)
        else:
            g = cls.constructGraph_CA(flows, timeIntervals, ctx, on_device=True)
        g.initiateAliasTables()
        corpus = g.sample(cls.numSamples, abi.SAMPLER_ALIAS, LayeredGraph.numLayer)
        if path is not None:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            corpus.write_seq(path, g.v_region, g.v_layer)
        return g, corpus


class SpatialGraph(LayeredGraph):
    """SpatialGraph.java:13-128.  Geometry (centroid distances) is host input: `weights[src, dst]`."""

    numSamples = 5_000_000            # :16
    numLayer = LayeredGraph.numLayer  # :17

    @staticmethod
    def keepNearestKVertices(weights, k):
        """SpatialGraph.keepNearestKVertices :29-35 on a dense weight matrix: per row stable sort by
        descending weight, first k kept, outDegree = DoubleStream.sum() (JDK-8 compensated sum)."""
        n = weights.shape[0]
        idx = np.argsort(-weights, axis=1, kind="stable")[:, :k]
        wk = np.take_along_axis(weights, idx, axis=1)
        out_degree = np.empty(n, np.float64)
        for r in range(n):
            s = c = 0.0
            for v in wk[r]:                      # Collectors.sumWithCompensation
                tmp = v - c
                vel = s + tmp
                c = (vel - s) - tmp
                s = vel
            out_degree[r] = s + c                # JDK 8 computeFinalSum adds the compensation
        return idx.astype(np.int32), wk, out_degree

    @classmethod
    def constructGraph(cls, region_ids, weights, order=None, ctx=None):
        """constructGraph_tract :37-62 / constructGraph_CA :65-88 with w = exp(-100 d) supplied by the host."""
        region_ids = np.asarray(region_ids, np.int32)
        n = len(region_ids)
        if order is None:
            pos = {int(r): i for i, r in enumerate(region_ids)}
            order = [pos[k] for k in java_hashmap_order(region_ids)]
        o = np.asarray(order, np.int64)
        W = np.asarray(weights, np.float64)[np.ix_(o, o)]  # iteration order on both loops
        k = min(10, n)
        idx, wk, out_degree = cls.keepNearestKVertices(W, k)
        # vertex ids: first appearance over the FULL all-pairs addEdge loop: src o[0] first, then every dst in order
        g = cls(ctx)
        g.v_region = region_ids[o].astype(np.int32)
        g.v_layer = np.zeros(n, np.int32)
        names = [str(int(r)) for r in g.v_region]
        src = np.repeat(np.arange(n, dtype=np.int32), k)
        g.add_edges_bulk(n, src, idx.reshape(-1), wk.reshape(-1), names)
        g._out_degree_override = out_degree
        # sourceVertices = new LinkedList<>(allVertices.values()) :56 -- HashMap<String,Vertex> order; the
        # host owns it.  We use vertex-id order; sourceWeightSum = stream sum over that order :57.
        g.sourceVertices = list(range(n))
        s = c = 0.0
        for v in out_degree:
            tmp = v - c
            vel = s + tmp
            c = (vel - s) - tmp
            s = vel
        g._sws_override = s + c
        return g

    @classmethod
    def outputSampleSequence(cls, regionLevel, region_ids, weights, path, ctx=None):
        """SpatialGraph.outputSampleSequence :91-121: token j is renamed "<j>-<region>" (:105-108)."""
        LayeredGraph.numLayer = cls.numLayer             # :92
        g = cls.constructGraph(region_ids, weights, ctx=ctx)
        g.initiateAliasTables()
        corpus = g.sample(cls.numSamples, abi.SAMPLER_ALIAS, LayeredGraph.numLayer)
        if path is not None:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            corpus.write_seq(path, g.v_region, position_prefix=True)
        return g, corpus


class DeepWalk:
    """DeepWalk.java:24-141."""

    Year = 2013                        # :25
    base_dir = ".."                    # the reference resolves ../miscs/<Year>/... from embedding/
    negativeSample = 5                 # :75
    minWordFrequency = 2               # :73
    seed = 2013

    @classmethod
    def _seq_path(cls, regionLevel, which):
        return os.path.join(cls.base_dir, "miscs", str(cls.Year), "deepwalkseq-%s" % regionLevel,
                            "taxi-%s.seq" % which)

    @classmethod
    def checkInputFile(cls, regionLevel, spatialGF, flows, spatial_weights, ctx=None):
        """DeepWalk.checkInputFile :85-112: generates the corpora with the per-level sizes.  Returns the
        device corpora in FileSentenceIterator order (crosstime, spatial) with position-tagged spatial ids."""
        out = {}
        if spatialGF in ("usespatial", "onlyspatial"):
            if regionLevel == "tract":
                SpatialGraph.numSamples, SpatialGraph.numLayer = 600_000, 8        # :89-91
            else:
                SpatialGraph.numSamples, SpatialGraph.numLayer = 80_000, 24        # :93-94
            out["spatial"] = SpatialGraph.outputSampleSequence(regionLevel, flows.region_ids, spatial_weights,
                                                               cls._seq_path(regionLevel, "spatial"), ctx)
        if spatialGF in ("nospatial", "usespatial"):
            if regionLevel == "tract":
                CrossTimeGraph.numSamples, CrossTimeGraph.numLayer = 15_000_000, 8  # :102-104
            else:
                CrossTimeGraph.numSamples, CrossTimeGraph.numLayer = 8_000_000, 24  # :106-107
            out["crosstime"] = CrossTimeGraph.outputSampleSequence(regionLevel, flows,
                                                                   cls._seq_path(regionLevel, "crosstime"), ctx=ctx)
        return out

    @classmethod
    def learnEmbedding(cls, regionLevel, spatialGF, corpora, labels, layerSize=None, ctx=None, out=None, **kw):
        """DeepWalk.learnEmbedding :32-83.  `corpora`: device corpora sharing one id space; `labels`:
        (layer, region) per id.  Writes the .vec file and returns the model."""
        ctx = ctx or default_context()
        if layerSize is None:
            layerSize = 2 if regionLevel == "CA" else 20                            # :62-66
        if out is None:
            out = os.path.join(cls.base_dir, "miscs", str(cls.Year),
                               "taxi-deepwalk-%s-%s-2D.vec" % (regionLevel, spatialGF))  # :61
        p = abi.sgns_params(dim=layerSize, window=LayeredGraph.numLayer, negative=cls.negativeSample,
                            min_count=cls.minWordFrequency, epochs=1, seed=cls.seed, **kw)  # :73-76
        model = abi.Model.train(ctx, corpora, p)                                     # :79
        os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
        model.write_vec(out, labels[0], labels[1])                                   # :82
        return model

    _sample_scale = 1.0                # test hook (not in the reference): scales the per-level numSamples of checkInputFile

    @classmethod
    def main(cls, argv, flows, spatial_weights, ctx=None):
        """DeepWalk.main :120-140: `[regionLevel] [spatialGF] [Year]` with the reference's defaults, then
        checkInputFile + learnEmbedding.  The reference deserialises the flow maps and reads the shapefiles itself
        (absent here), so the host hands in `flows` and the spatial weight matrix.  Exceptions are reported
        and swallowed like the reference's printStackTrace() (:137-139); returns the model or None."""
        import traceback
        regionLevel, spatialGF = "tract", "usespatial"                                 # :116-118
        if len(argv) > 0:
            regionLevel = argv[0]
        if len(argv) > 1:
            spatialGF = argv[1]
        if len(argv) > 2:
            cls.Year = int(argv[2])
        try:
            if regionLevel not in ("tract", "CA") or spatialGF not in ("usespatial", "nospatial", "onlyspatial"):
                raise ValueError("usage: DeepWalk [tract|CA] [usespatial|nospatial|onlyspatial] [Year]")
            ctx = ctx or default_context()
            made = cls._check_input_scaled(regionLevel, spatialGF, flows, spatial_weights, ctx)
            ids = np.asarray(flows.region_ids)
            n = len(ids)
            pos = {int(r): i for i, r in enumerate(ids)}
            L = LayeredGraph.numLayer
            corpora = []
            if "crosstime" in made:                                                    # FileSentenceIterator order
                g, c = made["crosstime"]
                c.relabel((g.v_layer.astype(np.int64) * n + np.array([pos[int(r)] for r in g.v_region])).astype(np.int32), L * n)
                corpora.append(c)
            if "spatial" in made:
                g, c = made["spatial"]
                c.relabel(np.array([pos[int(r)] for r in g.v_region], np.int32), L * n, position_stride=n)
                corpora.append(c)
            labels = ((np.arange(L * n) // n).astype(np.int32), ids[np.arange(L * n) % n].astype(np.int32))
            return cls.learnEmbedding(regionLevel, spatialGF, corpora, labels, ctx=ctx)
        except Exception:                                                              # :137-139
            traceback.print_exc()
            return None

    @classmethod
    def _check_input_scaled(cls, regionLevel, spatialGF, flows, spatial_weights, ctx):
        """checkInputFile with the per-level sizes multiplied by _sample_scale (1.0 = the reference's sizes)."""
        if cls._sample_scale == 1.0:
            return cls.checkInputFile(regionLevel, spatialGF, flows, spatial_weights, ctx)
        out = {}
        k = cls._sample_scale
        if spatialGF in ("usespatial", "onlyspatial"):
            SpatialGraph.numSamples, SpatialGraph.numLayer = (int(600_000 * k), 8) if regionLevel == "tract" else (int(80_000 * k), 24)
            out["spatial"] = SpatialGraph.outputSampleSequence(regionLevel, flows.region_ids, spatial_weights,
                                                               cls._seq_path(regionLevel, "spatial"), ctx)
        if spatialGF in ("nospatial", "usespatial"):
            CrossTimeGraph.numSamples, CrossTimeGraph.numLayer = (int(15_000_000 * k), 8) if regionLevel == "tract" else (int(8_000_000 * k), 24)
            out["crosstime"] = CrossTimeGraph.outputSampleSequence(regionLevel, flows, cls._seq_path(regionLevel, "crosstime"), ctx=ctx)
        return out

