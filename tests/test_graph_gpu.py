"""GPU parity, stage 1a: CSR + alias tables from libdge.so (through the C ABI) must be BIT-EXACT against the
CPU oracle, which is pinned to the reference's golden vector (tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def bits(a):
    return np.asarray(a, np.float64).view(np.int64)


def assert_tables_equal(gpu_t, ora_t):
    assert np.array_equal(gpu_t["row_ptr"], ora_t["row_ptr"])
    assert np.array_equal(gpu_t["col"], ora_t["col"])
    assert np.array_equal(bits(gpu_t["w"]), bits(ora_t["w"]))
    assert np.array_equal(bits(gpu_t["out_degree"]), bits(ora_t["out_degree"]))
    assert np.array_equal(gpu_t["alias"], ora_t["alias"])
    assert np.array_equal(bits(gpu_t["prob"]), bits(ora_t["prob"]))
    assert np.array_equal(gpu_t["src_alias"], ora_t["src_alias"])
    assert np.array_equal(bits(gpu_t["src_prob"]), bits(ora_t["src_prob"]))
    assert bits([gpu_t["source_weight_sum"]])[0] == bits([ora_t["source_weight_sum"]])[0]


def both(dge_lib, oracle, ctx, nv, src, dst, w, sources, **kw):
    g = dge_lib.Graph(ctx, nv, src, dst, w, sources, **kw)
    o = oracle.Graph(nv, src, dst, w, sources, out_degree=kw.get("out_degree"),
                     source_weight_sum=kw.get("source_weight_sum"), alias_mode=oracle.ALIAS_FAST)
    return g, o


def test_reference_golden_vector_through_the_abi(dge_lib, ctx):
    """LayeredGraphTest.java:12-44 replayed against the CUDA library, exact double equality."""
    gv = load_golden("layered_graph_test.json")
    g = dge_lib.Graph(ctx, 4, [0, 0, 0], [1, 2, 3], gv["weights"], [0])
    t = g.tables()
    assert t["prob"].tolist() == gv["prob"]
    assert t["alias"].tolist() == gv["alias"]
    assert t["out_degree"][0] == gv["out_degree"]
    xs = [x for x, _ in gv["samples"]]
    assert g.sample_next([0] * len(xs), xs).tolist() == [i for _, i in gv["samples"]]
    # dead-end vertices return "null"
    assert g.sample_next([1, 2, 3], [0.5] * 3).tolist() == [-1, -1, -1]


def test_host_mirror_reads_like_the_java_test(dge_lib, ctx):
    from embedding_b200.host import LayeredGraph
    gv = load_golden("layered_graph_test.json")
    g = LayeredGraph(ctx)
    g.addEdge("start", "d1", 2)
    g.addEdge("start", "d2", 10)
    g.addEdge("start", "d3", 8)
    g.addSourceVertex("start")
    g.initiateAliasTables()
    org = g.vertex("start")
    assert org.aliasTable.tolist() == gv["alias"] and org.probTable.tolist() == gv["prob"]
    assert org.outDegree == 20.0
    assert [org.sampleNextVertex(x).id for x, _ in gv["samples"]] == [i for _, i in gv["samples"]]
    assert g.vertex("d1").sampleNextVertex(0.3) is None
    LayeredGraph.numLayer = 8
    seq = g.sampleVertexSequence()
    assert seq[0] == "start" and len(seq) == 2 and seq[1] in ("d1", "d2", "d3")


@pytest.mark.parametrize("level,L", [("CA", 24), ("tract", 8), ("tract", 24)])
def test_flow_graph_tables_bit_exact(dge_lib, oracle, ctx, level, L):
    """Synthetic flows of the CA 77x24 and tract 801x{8,24} shapes (BASELINE configs 1-2)."""
    from embedding_b200 import host, synth
    if level == "CA":
        ids, dens = synth.ca_ids(), 0.6
    else:
        ids, dens = synth.tract_ids(), 0.03
    fl = host.Flows(ids, synth.flow_tensor(len(ids), seed=2013, density=dens))
    host.CrossTimeGraph.numLayer = L
    g = host.CrossTimeGraph.constructGraph_CA(fl, ctx=ctx) if level == "CA" else host.CrossTimeGraph.constructGraph_tract(fl, ctx=ctx)
    host.CrossTimeGraph.numLayer = 8
    nv, src, dst, w = g._bulk
    G, O = both(dge_lib, oracle, ctx, nv, src, dst, w, g.sourceVertices)
    assert ctx.phase_ms("grouped_input") == 1.0
    assert_tables_equal(G.tables(), O.tables())


def test_ungrouped_input_keeps_insertion_order(dge_lib, oracle, ctx):
    """addEdge in arbitrary order: rows must keep insertion order (stable), exercised via the sort path."""
    rng = np.random.default_rng(0)
    nv, ne = 300, 20000
    src = rng.integers(0, nv, ne).astype(np.int32)
    dst = rng.integers(0, nv, ne).astype(np.int32)
    w = (np.floor(rng.pareto(1.2, ne) * 3) + 1).astype(np.float64)
    sources = rng.choice(nv, 50, replace=False).astype(np.int32)
    G, O = both(dge_lib, oracle, ctx, nv, src, dst, w, sources)
    assert ctx.phase_ms("grouped_input") == 0.0
    assert_tables_equal(G.tables(), O.tables())


@pytest.mark.parametrize("weights", ["int", "real", "ties"])
def test_random_rows_bit_exact(dge_lib, oracle, ctx, weights):
    rng = np.random.default_rng({"int": 1, "real": 2, "ties": 3}[weights])
    nv = 4000
    deg = np.minimum(rng.zipf(1.5, nv), 900)
    deg[rng.random(nv) < 0.1] = 0  # destination-only vertices
    src = np.repeat(np.arange(nv, dtype=np.int32), deg)
    ne = len(src)
    dst = rng.integers(0, nv, ne).astype(np.int32)
    if weights == "int":
        w = np.floor(rng.pareto(1.2, ne) * 3) + 1
    elif weights == "real":
        w = np.exp(-rng.random(ne) * 5)  # like exp(-100 d)
    else:
        w = rng.integers(1, 3, ne).astype(np.float64)
    sources = np.flatnonzero(deg > 0)[:700].astype(np.int32)
    G, O = both(dge_lib, oracle, ctx, nv, src, dst, w, sources)
    assert_tables_equal(G.tables(), O.tables())


def test_hub_rows_and_long_source_list(dge_lib, oracle, ctx):
    """Rows > 1024 entries and a source list of 50K go through the hierarchical-bitmap kernel."""
    rng = np.random.default_rng(7)
    nv = 60000
    deg = np.ones(nv, np.int64)
    deg[[5, 77, 4000]] = [1025, 5000, 70000]
    src = np.repeat(np.arange(nv, dtype=np.int32), deg)
    ne = len(src)
    dst = rng.integers(0, nv, ne).astype(np.int32)
    w = np.floor(rng.pareto(1.1, ne) * 2) + 1
    sources = np.arange(50000, dtype=np.int32)
    G, O = both(dge_lib, oracle, ctx, nv, src, dst, w, sources)
    assert_tables_equal(G.tables(), O.tables())


def test_host_owned_out_degree_spatial(dge_lib, oracle, ctx):
    """SpatialGraph: outDegree / sourceWeightSum come from the host (DoubleStream.sum), top-10 rows."""
    from embedding_b200 import host, synth
    W = synth.spatial_weights(77, seed=1)
    g = host.SpatialGraph.constructGraph(synth.ca_ids(), W, ctx=ctx)
    nv, src, dst, w = g._bulk
    G, O = both(dge_lib, oracle, ctx, nv, src, dst, w, g.sourceVertices, out_degree=g._out_degree_override,
                source_weight_sum=g._sws_override)
    tg = G.tables()
    assert_tables_equal(tg, O.tables())
    assert np.all(np.diff(tg["row_ptr"]) == 10)  # SpatialGraphTest: 10 out-edges per vertex


def test_edge_cases(dge_lib, oracle, ctx):
    # empty graph
    g = dge_lib.Graph(ctx, 0, [], [], [], [])
    t = g.tables()
    assert t["row_ptr"].tolist() == [0] and g.walk(5, 4, 1).tokens().tolist() == [[-1] * 4] * 5
    # vertices but no edges, sources present: every walk is the source token only
    G, O = both(dge_lib, oracle, ctx, 3, [], [], [], [2, 0])
    assert_tables_equal(G.tables(), O.tables())
    # single edge, single source
    G, O = both(dge_lib, oracle, ctx, 2, [0], [1], [3.5], [0])
    assert_tables_equal(G.tables(), O.tables())
    # no sources
    G, O = both(dge_lib, oracle, ctx, 2, [0], [1], [1.0], [])
    assert_tables_equal(G.tables(), O.tables())
    assert G.walk(3, 5, 9).tokens().tolist() == [[-1] * 5] * 3


def test_error_behaviour(dge_lib, ctx):
    with pytest.raises(dge_lib.DgeError) as e:
        dge_lib.Graph(ctx, 3, [0, 5], [1, 2], [1.0, 1.0], [0])
    assert e.value.code == -1 and "out of" in str(e.value)
    with pytest.raises(dge_lib.DgeError):
        dge_lib.Graph(ctx, 3, [0], [1], [1.0], [7])
    g = dge_lib.Graph(ctx, 3, [0], [1], [1.0], [0])
    with pytest.raises(dge_lib.DgeError):
        g.sample_next([9], [0.5])
    with pytest.raises(dge_lib.DgeError):
        g.walk(-1, 4, 0)


def test_sample_next_matches_oracle_everywhere(dge_lib, oracle, ctx):
    """The packed 32 B walk records (what the walk kernel reads) agree with the oracle's prob/alias lookup for
    uniforms on and around every acceptance threshold."""
    rng = np.random.default_rng(11)
    nv = 500
    deg = rng.integers(0, 40, nv)
    src = np.repeat(np.arange(nv, dtype=np.int32), deg)
    ne = len(src)
    dst = rng.integers(0, nv, ne).astype(np.int32)
    w = np.floor(rng.pareto(1.2, ne) * 3) + 1
    sources = np.flatnonzero(deg > 0)[:100].astype(np.int32)
    G, O = both(dge_lib, oracle, ctx, nv, src, dst, w, sources)
    t = O.tables()
    vs, xs = [], []
    for v in range(nv):
        b, e = t["row_ptr"][v], t["row_ptr"][v + 1]
        k = e - b
        for i in range(k):
            for y in (t["prob"][b + i], np.nextafter(t["prob"][b + i], 0), 0.0, 0.999999):
                x = (i + min(max(y, 0.0), 0.9999999)) / k
                vs.append(v)
                xs.append(x)
    vs += [-1] * 2000
    xs += rng.random(2000).tolist()
    vs = np.array(vs, np.int32)
    xs = np.array(xs)
    for sampler, fn in ((dge_lib.SAMPLER_ALIAS, O.sample_next), (dge_lib.SAMPLER_CDF, O.sample_next_ov)):
        got = G.sample_next(vs, xs, sampler)
        want = np.array([fn(int(v), float(x)) if v >= 0 else O.sample_source(float(x), sampler) for v, x in zip(vs, xs)])
        assert np.array_equal(got, want)
