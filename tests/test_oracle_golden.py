"""Pins the CPU oracle (oracle/) against every golden vector the reference's own tests hold for the hot
path (exactly one, LayeredGraphTest.java:12-44) and against published known answers of the
generators it uses.  CPU only."""
import numpy as np
import pytest

from conftest import load_golden


def test_layered_graph_test_golden_vector(oracle):
    """LayeredGraphTest.testAliasTable: weights 2,10,8 -> prob [0.3,0.8,1.0], alias [1,2,-1], outDegree 20,
    five sampleNextVertex(x) lookups; exact double equality as JUnit's assertEquals(Object, Object)."""
    g = load_golden("layered_graph_test.json")
    for mode in (oracle.ALIAS_LITERAL, oracle.ALIAS_FAST):
        # vertices: 0 = org ("start"), 1..3 = d1..d3 as in the Java test
        gr = oracle.Graph(4, [0, 0, 0], [1, 2, 3], g["weights"], [0], alias_mode=mode)
        t = gr.tables()
        assert t["prob"].tolist() == g["prob"]
        assert t["alias"].tolist() == g["alias"]
        assert t["out_degree"][0] == g["out_degree"]
        for x, expect in g["samples"]:
            assert gr.sample_next(0, x) == expect


def test_philox_known_answers(oracle):
    """Random123 kat_vectors for Philox4x32-10."""
    assert oracle.philox4x32_10([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_java_util_random_known_answer(oracle):
    """new java.util.Random(42).nextDouble() twice (documented JDK LCG)."""
    d = oracle.java_random_doubles(42, 2)
    assert d[0] == 0.7275636800328681
    assert d[1] == 0.6832234717598454


def test_host_philox_matches_oracle(oracle):
    from embedding_b200 import philox
    for seed, wid, draw in [(0, 0, 0), (7, 12345678901, 3), (2 ** 63 + 5, 2 ** 40, 17)]:
        assert philox.uniform(seed, wid, draw) == oracle.philox_uniform(seed, wid, draw)


@pytest.mark.parametrize("kind", ["int", "pareto", "uniform", "ties", "exact_one"])
def test_fast_alias_equals_literal(oracle, kind):
    """The ordered-set form must reproduce LayeredGraph.java:65-81 bit for bit, including rows with
    entries exactly 1.0, ties, dangling alias -1 and role switches mid-scan."""
    rng = np.random.default_rng(hash(kind) % 2 ** 32)
    for trial in range(400):
        k = int(rng.integers(1, 300))
        if kind == "int":
            w = rng.integers(1, 50, k).astype(float)
        elif kind == "pareto":
            w = np.floor(rng.pareto(1.2, k) * 3) + 1
        elif kind == "uniform":
            w = rng.random(k) + 1e-3
        elif kind == "ties":
            w = rng.integers(1, 4, k).astype(float)
        else:  # many entries with prob exactly 1.0: w == mean
            w = np.full(k, 4.0)
            m = k // 3
            if m:
                idx = rng.choice(k, 2 * m, replace=False) if 2 * m <= k else np.arange(0)
                w[idx[:m]] += 2.0
                w[idx[m:2 * m]] -= 2.0
        p1, a1 = oracle.alias_table(w, mode=oracle.ALIAS_LITERAL)
        p2, a2 = oracle.alias_table(w, mode=oracle.ALIAS_FAST)
        assert np.array_equal(p1.view(np.int64), p2.view(np.int64))
        assert np.array_equal(a1, a2)


def test_alias_table_is_a_distribution(oracle):
    """Property: the table reproduces w/outDegree (up to the dangling -1 entries' ~1e-15 deficit)."""
    rng = np.random.default_rng(5)
    for _ in range(50):
        k = int(rng.integers(1, 120))
        w = np.floor(rng.pareto(1.2, k) * 3) + 1
        prob, alias = oracle.alias_table(w)
        mass = prob.copy()
        for i in range(k):
            j = alias[i] if alias[i] >= 0 else i
            mass[j] += 1.0 - prob[i]
        assert np.allclose(mass / k, w / w.sum(), rtol=0, atol=1e-12)


def test_crosstime_edges_literal_vs_host_mirror(oracle):
    """CrossTimeGraph.constructGraph_CA / _tract restated literally (oracle) vs the vectorised host mirror."""
    from embedding_b200 import host, synth
    ids = np.array([17, 3, 99, 40, 8, 1, 64, 23, 5], np.int32)
    F = synth.flow_tensor(len(ids), seed=3, density=0.4)
    fl = host.Flows(ids, F)
    for level, L in (("CA", 24), ("tract", 8), ("tract", 24), ("CA", 8)):
        host.CrossTimeGraph.numLayer = L
        if level == "CA":
            g = host.CrossTimeGraph.constructGraph_CA(fl)
            step = 24 // L
            iv = [0] * (L + 1)
            for i in range(0, L + 1, step):
                iv[i] = (i * step) % L
            ref = oracle.crosstime_edges(F, fl.order, L, 0, iv)
        else:
            g = host.CrossTimeGraph.constructGraph_tract(fl)
            ref = oracle.crosstime_edges(F, fl.order, L, 1)
        nv, src, dst, w = g._bulk
        assert nv == ref["n_vertices"]
        assert np.array_equal(src, ref["src"]) and np.array_equal(dst, ref["dst"]) and np.array_equal(w, ref["w"])
        assert np.array_equal(g.v_layer, ref["v_layer"])
        assert np.array_equal(g.v_region, ids[ref["v_region"]])
        assert g.sourceVertices == ref["sources"].tolist()
    host.CrossTimeGraph.numLayer = 8


def test_flow_slot_semantics(oracle):
    """CommunityArea.getFlowTo is circular half-open (lo==hi -> 0, (0,23) drops hour 23); Tract.getFlowTo is
    inclusive."""
    import ctypes as C
    n = 3
    F = np.arange(n * 24 * n, dtype=np.int32).reshape(n, 24, n)
    p = F.ctypes.data_as(C.POINTER(C.c_int32))
    L = oracle.lib()
    assert L.ora_flow_ca(p, n, 1, 2, 5, 5) == 0
    assert L.ora_flow_ca(p, n, 1, 2, 0, 23) == int(F[1, 0:23, 2].sum())
    assert L.ora_flow_ca(p, n, 1, 2, 22, 2) == int(F[1, 22, 2] + F[1, 23, 2] + F[1, 0, 2] + F[1, 1, 2])
    assert L.ora_flow_tract(p, n, 1, 2, 3, 5) == int(F[1, 3:6, 2].sum())


def test_keep_nearest_k(oracle):
    """SpatialGraphTest.java:18-21 pins: 10 out-edges, weights non-increasing; plus host mirror equality."""
    from embedding_b200 import host, synth
    W = synth.spatial_weights(40, seed=9)
    col, wk, od = oracle.keep_nearest_k(W, 10)
    assert col.shape == (40, 10)
    assert np.all(np.diff(wk, axis=1) <= 0)
    assert np.all(wk[:, 0] == 1.0) and np.all(col[:, 0] == np.arange(40))  # self edge, exp(0) = 1
    idx, wk2, od2 = host.SpatialGraph.keepNearestKVertices(W, 10)
    assert np.array_equal(idx, col) and np.array_equal(wk2, wk)
    assert np.array_equal(od2.view(np.int64), od.view(np.int64))


def test_java_hashmap_order():
    """CA ids 1..77 iterate ascending; the 801 tract ids iterate in bucket order
    (id ^ id>>>16) & 2047 starting 51200, 831500, 823304."""
    from embedding_b200 import host, synth
    assert host.java_hashmap_order(range(1, 78)) == list(range(1, 78))
    order = host.java_hashmap_order(synth.tract_ids())
    assert len(order) == 801 and order[:3] == [51200, 831500, 823304]


def test_oracle_walks_are_valid_and_deterministic(oracle):
    from embedding_b200 import synth
    g = synth.powerlaw_flow_graph(50, L=6, seed=4, mean_degree=5, cap=20)
    gr = oracle.Graph(g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"])
    a = gr.walk(2000, 6, seed=11)
    b = np.concatenate([gr.walk(700, 6, seed=11), gr.walk(1300, 6, seed=11, first_walk_id=700)])
    assert np.array_equal(a, b)  # counter-based: independent of how walks are split
    edges = set(zip(g["src"].tolist(), g["dst"].tolist()))
    assert np.all(np.isin(a[:, 0], g["sources"]))
    for row in a[:300]:
        for u, v in zip(row[:-1], row[1:]):
            assert (int(u), int(v)) in edges
    lcg = gr.walk(100, 6, seed=11, rng=oracle.RNG_JAVA_LCG)
    assert lcg.shape == (100, 6) and np.all(np.isin(lcg[:, 0], g["sources"]))
