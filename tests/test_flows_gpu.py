"""GPU parity of the step in front of stage 1: trip counting, slot sums and the
CrossTimeGraph edge enumeration on device, against the literal C restatement of the Java loops (oracle) and against
the host-enumerated COO fed through dge_graph_build.  Everything here is integer / table work: bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ca_intervals(L):
    step = 24 // L
    iv = [0] * (L + 1)
    for i in range(0, L + 1, step):            # CrossTimeGraph.java:55-58 incl. its quirk for timeStep != 1 (Q1)
        iv[i] = (i * step) % L
    return iv


def check_against_oracle(dge_lib, oracle, ctx, ids, F, level, L):
    from embedding_b200 import host
    fl = host.Flows(ids, F)
    mode = 0 if level == "CA" else 1
    iv = ca_intervals(L) if level == "CA" else None
    ref = oracle.crosstime_edges(F, fl.order, L, mode, iv)
    dev = dge_lib.Flows(ctx, len(ids), F)
    G = dev.crosstime_graph(fl.order, L, mode, iv)
    assert (G.nv, G.ne, G.ns) == (ref["n_vertices"], len(ref["src"]), len(ref["sources"]))
    vl, vr, so = G.labels()
    assert np.array_equal(vl, ref["v_layer"]) and np.array_equal(vr, ref["v_region"])
    assert np.array_equal(so, ref["sources"])
    # CSR + alias tables: identical to building from the COO the Java loops produce
    Og = oracle.Graph(ref["n_vertices"], ref["src"], ref["dst"], ref["w"], ref["sources"])
    tg, to = G.tables(), Og.tables()
    for k in ("row_ptr", "col", "alias", "src_alias"):
        assert np.array_equal(tg[k], to[k]), k
    for k in ("w", "prob", "out_degree", "src_prob"):
        assert np.array_equal(tg[k].view(np.int64), to[k].view(np.int64)), k
    assert tg["source_weight_sum"] == to["source_weight_sum"]
    return G, ref


@pytest.mark.parametrize("level,L", [("CA", 24), ("tract", 8), ("tract", 24), ("CA", 8), ("tract", 1), ("CA", 12)])
def test_crosstime_graph_from_device_flows_bit_exact(dge_lib, oracle, ctx, level, L):
    from embedding_b200 import synth
    ids = np.array([17, 3, 99, 40, 8, 1, 64, 23, 5, 77, 12], np.int32)
    F = synth.flow_tensor(len(ids), seed=3 + L, density=0.35)
    F[4] = 0                                      # a region that never emits: appears only as a destination
    F[:, :, 7] = 0                                # a region nobody reaches
    check_against_oracle(dge_lib, oracle, ctx, ids, F, level, L)


def test_real_shapes_ca77_and_tract801(dge_lib, oracle, ctx):
    """The two region sets of the reference at full size (77 community areas x 24, 801 tracts x 8)."""
    from embedding_b200 import synth
    F = synth.planted_flow_tensor(synth.ca_latents(), mean_trips_per_pair_hour=2.0)
    check_against_oracle(dge_lib, oracle, ctx, synth.ca_ids(), F, "CA", 24)
    F = synth.planted_flow_tensor(synth.poi_latents())
    G, ref = check_against_oracle(dge_lib, oracle, ctx, synth.tract_ids(), F, "tract", 8)
    assert G.ne > 100_000


def test_trip_counting_matches_numpy_and_reference_rules(dge_lib, ctx):
    rng = np.random.default_rng(5)
    n, trips = 37, 200_000
    s = rng.integers(-1, n, trips).astype(np.int32)          # -1: outside every region (never counted)
    d = rng.integers(-1, n, trips).astype(np.int32)
    h = rng.integers(0, 24, trips).astype(np.int32)
    fl = dge_lib.Flows(ctx, n)
    fl.add_trips(s[:120_000], d[:120_000], h[:120_000])      # two batches accumulate
    fl.add_trips(s[120_000:], d[120_000:], h[120_000:])
    ok = (s >= 0) & (d >= 0)
    ref = np.zeros((n, 24, n), np.int32)
    np.add.at(ref, (s[ok], h[ok], d[ok]), 1)
    assert np.array_equal(fl.tensor(), ref)
    with pytest.raises(dge_lib.DgeError):
        fl.add_trips([0], [1], [24])                          # hour of day out of range
    with pytest.raises(dge_lib.DgeError):
        fl.add_trips([n], [1], [3])


def test_empty_and_degenerate_flows(dge_lib, ctx):
    fl = dge_lib.Flows(ctx, 5)                                # no trips at all: no vertices, no edges, no sources
    G = fl.crosstime_graph(np.arange(5), 8, 1)
    assert (G.nv, G.ne, G.ns) == (0, 0, 0)
    c = G.walk(10, 8, seed=1)
    assert (c.tokens() == -1).all()
    fl0 = dge_lib.Flows(ctx, 0)
    G0 = fl0.crosstime_graph(np.zeros(0, np.int32), 24, 0, [0] * 25)
    assert (G0.nv, G0.ne) == (0, 0)
    with pytest.raises(dge_lib.DgeError):
        fl.crosstime_graph([0, 1, 2, 3, 3], 8, 1)             # not a permutation
    with pytest.raises(dge_lib.DgeError):
        fl.crosstime_graph(np.arange(5), 8, 0, [0, 1, 2, 3, 4, 5, 6, 7, 24])   # Java's loop would never end


def test_host_mirror_device_path_equals_host_enumeration(dge_lib, ctx):
    """CrossTimeGraph.constructGraph_* with on_device=True yields the same names, sources, tables and walks as
    the host-enumerated COO through dge_graph_build."""
    from embedding_b200 import host, synth
    ids = synth.tract_ids()[:120]
    F = synth.flow_tensor(len(ids), seed=9, density=0.05)
    fl = host.Flows(ids, F)
    host.CrossTimeGraph.numLayer = 8
    a = host.CrossTimeGraph.constructGraph_tract(fl, ctx).initiateAliasTables()
    b = host.CrossTimeGraph.constructGraph_tract(fl, ctx, on_device=True).initiateAliasTables()
    assert a._names == b._names and a.sourceVertices == b.sourceVertices
    ta, tb = a._graph.tables(), b._graph.tables()
    for k in ta:
        assert np.array_equal(np.asarray(ta[k]), np.asarray(tb[k])), k
    assert np.array_equal(a._graph.walk(5000, 8, 3).tokens(), b._graph.walk(5000, 8, 3).tokens())
    # trips -> flows on device -> same graph
    s, h, d = np.nonzero(F)
    reps = F[s, h, d]
    fl2 = host.Flows.from_trips(ids, np.repeat(s, reps), np.repeat(d, reps), np.repeat(h, reps), ctx=ctx)
    assert np.array_equal(fl2.F, F)
    host.CrossTimeGraph.numLayer = 8


def test_static_exports_match_the_java_loops(dge_lib, oracle, ctx, tmp_path):
    """the .matrix / .od exports (CommunityAreas.java:127-184, Tracts.java:236-301) written from the
    device slot sums equal a literal restatement of the Java loops over the oracle's getFlowTo, byte for byte."""
    from embedding_b200 import host, synth
    n = 13
    ids = np.arange(1, n + 1, dtype=np.int32)                    # community-area ids are 1..n
    F = synth.flow_tensor(n, seed=21, density=0.4)
    fl = host.Flows(ids[::-1].copy(), F)                          # ids stored in another order than 1..n
    pos = {int(r): i for i, r in enumerate(fl.region_ids)}
    ca = lambda i, j, lo, hi: oracle.flow_ca(F, pos[i], pos[j], lo, hi)
    hour = lambda i, j, h: int(F[pos[i], h, pos[j]])
    d = str(tmp_path)
    fl.outputStaticFlowGraph(d, ctx)
    rows, od = [], []
    for i in range(1, n + 1):
        w = [ca(i, j, 0, 23) for j in range(1, n + 1)]
        rows.append(",".join(map(str, w)) + "\n")
        od += ["%d %d %d\n" % (i, j, w[j - 1]) for j in range(1, n + 1) if w[j - 1] > 0]
    assert open(d + "/taxi-CA-static.matrix").read() == "".join(rows)
    assert open(d + "/taxi-CA-static.od").read() == "".join(od)
    assert sum(ca(i, j, 0, 23) for i in range(1, n + 1) for j in range(1, n + 1)) == int(F[:, :23, :].sum())  # hour 23 left out
    fl.outputAdjacencyMatrix_CA(d, ctx)
    fl.outputEdgeGraph_LINE(d, ctx)
    for h in (0, 7, 23):
        ref = "".join(" ".join(str(hour(i, j, h)) for j in range(1, n + 1)) + "\n" for i in range(1, n + 1))
        assert open(d + "/taxi-CA-h%d.matrix" % h).read() == ref
        ref = "".join("%d %d %d\n" % (i, j, hour(i, j, h)) for i in range(1, n + 1) for j in range(1, n))
        assert open(d + "/taxi-CA-h%d.od" % h).read() == ref
    # tract level
    tids = synth.tract_ids()[:40]
    Ft = synth.flow_tensor(len(tids), seed=22, density=0.2)
    ft = host.Flows(tids, Ft)
    tp = {int(r): i for i, r in enumerate(tids)}
    tr = lambda a, b, lo, hi: oracle.flow_tract(Ft, a, b, lo, hi)
    ft.outputEdgeFile(d, 8, ctx)
    ft.outputAdjacencyMatrix_tract(d, 8, ctx)
    for h in (0, 5, 7):
        ref = []
        for a in ft.order:                                       # tracts.values(): HashMap iteration order
            for b in ft.order:                                   # keySet of hour h: destinations with a trip in hour h
                if Ft[a, h, b] > 0 and tr(a, b, h, h + 2) > 0:
                    ref.append("%d %d %d\n" % (tids[a], tids[b], tr(a, b, h, h + 2)))
        assert open(d + "/taxi-h%d.od" % h).read() == "".join(ref)
        srt = sorted(int(r) for r in tids)
        ref = "".join(",".join(str(tr(tp[a], tp[b], h, h + 2)) for b in srt) + "\n" for a in srt)
        assert open(d + "/taxi-h%d.matrix" % h).read() == ref
    ref = "".join("%d %d %d\n" % (tids[a], tids[b], tr(a, b, 0, 23)) for a in ft.order for b in ft.order if tr(a, b, 0, 23) > 0)
    assert open(d + "/taxi-all.od").read() == ref
    assert np.array_equal(ft.device(ctx).slot_weights(1, 0, 23), Ft.sum(axis=1))
    with pytest.raises(dge_lib.DgeError):
        ft.device(ctx).slot_weights(1, 0, 24)
