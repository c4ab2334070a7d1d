"""GPU parity, stage 1b: walks.  With the same counter-based Philox stream the CUDA walks are identical to the
oracle's edge by edge; statistically they follow the reference's transition distribution (chi-square)."""
import os

import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu


def small_graph(seed=4, n_regions=60, L=8):
    from embedding_b200 import synth
    return synth.powerlaw_flow_graph(n_regions, L=L, seed=seed, mean_degree=6, cap=40)


@pytest.mark.parametrize("sampler", [0, 1])
def test_walks_equal_oracle_token_for_token(dge_lib, oracle, ctx, sampler):
    g = small_graph()
    G = dge_lib.Graph(ctx, g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"])
    O = oracle.Graph(g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"])
    for L in (1, 5, 8, 24):
        got = G.walk(20000, L, seed=99, sampler=sampler).tokens()
        want = O.walk(20000, L, seed=99, sampler=sampler)
        assert np.array_equal(got, want)


def test_walks_with_dead_ends_are_padded(dge_lib, oracle, ctx):
    rng = np.random.default_rng(3)
    nv = 200
    deg = rng.integers(0, 6, nv)
    deg[rng.random(nv) < 0.3] = 0
    src = np.repeat(np.arange(nv, dtype=np.int32), deg)
    dst = rng.integers(0, nv, len(src)).astype(np.int32)
    w = rng.integers(1, 9, len(src)).astype(np.float64)
    sources = np.flatnonzero(deg > 0)[:40].astype(np.int32)
    G = dge_lib.Graph(ctx, nv, src, dst, w, sources)
    O = oracle.Graph(nv, src, dst, w, sources)
    c = G.walk(50000, 12, seed=5)
    got = c.tokens()
    assert np.array_equal(got, O.walk(50000, 12, seed=5))
    assert (got == -1).any()
    # padding is a suffix, and count_tokens agrees
    valid = got >= 0
    assert np.all(valid[:, :-1] >= valid[:, 1:])
    assert c.count_tokens() == int(valid.sum())


def test_sharding_by_walk_id_is_invisible(dge_lib, ctx):
    """Multi-GPU partitioning: walk ids [a,b) on any device give the same tokens."""
    g = small_graph(seed=8)
    G = dge_lib.Graph(ctx, g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"])
    whole = G.walk(10000, 8, seed=1).tokens()
    parts = [G.walk(n, 8, seed=1, first_walk_id=f).tokens() for f, n in ((0, 2500), (2500, 2500), (5000, 5000))]
    assert np.array_equal(whole, np.concatenate(parts))


def test_transition_frequencies_chi_square(dge_lib, ctx):
    """Empirical transition counts out of the busiest vertices vs w/outDegree, and the source distribution vs
    outDegree/sourceWeightSum (LayeredGraph.java:199-225): chi-square must not reject at 1e-4."""
    g = small_graph(seed=21, n_regions=30, L=4)
    G = dge_lib.Graph(ctx, g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"])
    t = G.tables()
    tok = G.walk(2_000_000, 4, seed=77).tokens()
    # sources
    cnt = np.bincount(tok[:, 0], minlength=g["n_vertices"])[g["sources"]]
    p = t["out_degree"][g["sources"]] / t["source_weight_sum"]
    keep = p > 0
    chi = stats.chisquare(cnt[keep], p[keep] / p[keep].sum() * cnt[keep].sum())
    assert chi.pvalue > 1e-4, chi
    # transitions out of the 20 most visited layer-0 vertices
    first = tok[:, 0]
    for v in np.argsort(-np.bincount(first, minlength=g["n_vertices"]))[:20]:
        b, e = t["row_ptr"][v], t["row_ptr"][v + 1]
        nxt = tok[first == v, 1]
        # merge multi-edges to the same destination
        dests, inv = np.unique(t["col"][b:e], return_inverse=True)
        pw = np.bincount(inv, weights=t["w"][b:e]) / t["out_degree"][v]
        obs = np.array([(nxt == d).sum() for d in dests])
        assert obs.sum() == len(nxt)  # every step lands on a real edge
        ok = pw * len(nxt) >= 5
        if ok.sum() >= 2:
            chi = stats.chisquare(np.append(obs[ok], obs[~ok].sum()) if (~ok).any() else obs[ok],
                                  np.append(pw[ok], pw[~ok].sum()) * len(nxt) if (~ok).any() else pw[ok] * len(nxt))
            assert chi.pvalue > 1e-4, (v, chi)


def test_every_step_is_an_edge_and_layers_advance(dge_lib, ctx):
    """Size-independent property on a larger graph (100K vertices, 5M walks): consecutive tokens are CSR edges
    and the layer advances by exactly one (mod L)."""
    from embedding_b200 import synth
    L = 8
    g = synth.powerlaw_flow_graph(12500, L=L, seed=5, mean_degree=20, cap=2000)
    G = dge_lib.Graph(ctx, g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"])
    tok = G.walk(5_000_000, L, seed=3).tokens()
    assert (tok >= 0).all()  # this generator has no dead ends
    lay = g["v_layer"][tok]
    assert np.array_equal(lay, np.tile(np.arange(L, dtype=np.int32), (len(tok), 1)))
    key = g["src"].astype(np.int64) * g["n_vertices"] + g["dst"]
    key = np.unique(key)
    sub = tok[:: 50]
    pair = sub[:, :-1].astype(np.int64) * g["n_vertices"] + sub[:, 1:]
    assert np.isin(pair.ravel(), key).all()


def test_tokens_u16_equal_tokens(dge_lib, ctx):
    g = small_graph(seed=12)
    G = dge_lib.Graph(ctx, g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"])
    for n, L in ((1, 1), (33, 7), (50_000, 24)):
        c = G.walk(n, L, seed=4)
        a, b = c.tokens(), c.tokens_u16()
        assert np.array_equal(np.where(a < 0, 0xFFFF, a).astype(np.uint16), b)
    big = dge_lib.Corpus.from_tokens(ctx, np.array([[70_000, -1]], np.int32), 70_001)
    with pytest.raises(dge_lib.DgeError):
        big.tokens_u16()                      # id space does not fit 16 bits


def test_corpus_roundtrip_relabel_and_seq_format(dge_lib, ctx, tmp_path):
    rng = np.random.default_rng(2)
    tok = rng.integers(0, 50, size=(1000, 7)).astype(np.int32)
    tok[rng.random(tok.shape) < 0.1] = -1
    tok = np.where(np.cumsum(tok < 0, axis=1) > 0, -1, tok).astype(np.int32)  # padding is a suffix
    c = dge_lib.Corpus.from_tokens(ctx, tok, 50)
    assert np.array_equal(c.tokens(), tok)
    layer = (np.arange(50) % 7).astype(np.int32)
    region = (1000 + np.arange(50) * 3).astype(np.int32)
    p = tmp_path / "taxi-crosstime.seq"
    c.write_seq(str(p), region, layer)
    lines = p.read_text().split("\n")
    assert lines[-1] == "" and len(lines) == 1001
    for row, line in zip(tok[:200], lines):
        want = " ".join("%d-%d" % (layer[t], region[t]) for t in row if t >= 0)
        assert line == want  # String.join(" ", seq), CrossTimeGraph.java:136
    p2 = tmp_path / "taxi-spatial.seq"
    c.write_seq(str(p2), region, position_prefix=True)
    l2 = p2.read_text().split("\n")
    for row, line in zip(tok[:200], l2):
        assert line == " ".join("%d-%d" % (j, region[t]) for j, t in enumerate(row) if t >= 0)  # SpatialGraph.java:105-108
    # read both files back (DeepWalk.learnEmbedding on existing .seq files, DeepWalk.java:47-59)
    back = dge_lib.Corpus.read_seq(ctx, str(p), region, layer)
    assert (back.n_walks, back.L, back.n_ids) == (1000, 7, 50) and np.array_equal(back.tokens(), tok)
    back2 = dge_lib.Corpus.read_seq(ctx, str(p2), region, position_prefix=True)
    assert np.array_equal(back2.tokens(), tok)
    bad = tmp_path / "bad.seq"
    bad.write_text("0-1000 3-9999\n")
    with pytest.raises(dge_lib.DgeError):
        dge_lib.Corpus.read_seq(ctx, str(bad), region, layer)          # unknown label
    with pytest.raises(dge_lib.DgeError):
        dge_lib.Corpus.read_seq(ctx, str(tmp_path / "missing.seq"), region, layer)
    # relabel into a (position, region-index) space
    c.relabel(np.arange(50, dtype=np.int32), 50 * 7, position_stride=50)
    got = c.tokens()
    want = np.where(tok >= 0, tok + np.arange(7)[None, :] * 50, -1)
    assert np.array_equal(got, want)
    with pytest.raises(dge_lib.DgeError):
        dge_lib.Corpus.from_tokens(ctx, np.array([[99]], np.int32), 50)


def test_seq_text_of_a_corpus_larger_than_one_chunk(dge_lib, ctx, tmp_path):
    """dge_corpus_write_seq formats the text on the device in chunks of ~64 MB that travel through two pinned buffers:
    the file must be byte-identical to String.join(" ", seq) + "\\n" per walk (CrossTimeGraph.java:136-137) across chunk
    and 128-line block boundaries, with ragged lines (dead ends), one-token lines, negative and 10-digit labels."""
    rng = np.random.default_rng(5)
    n, L, n_ids = 700_000, 24, 4000
    tok = rng.integers(0, n_ids, size=(n, L)).astype(np.int32)
    cut = rng.integers(1, L + 1, size=n)
    tok[np.arange(L)[None, :] >= cut[:, None]] = -1
    layer = rng.integers(0, 24, size=n_ids).astype(np.int32)
    region = rng.integers(10100, 980100, size=n_ids).astype(np.int32)
    region[:3] = (-7, 0, 2147483647)
    c = dge_lib.Corpus.from_tokens(ctx, tok, n_ids)
    p = tmp_path / "big.seq"
    c.write_seq(str(p), region, layer)
    lab = np.array(["%d-%d" % (a, b) for a, b in zip(layer, region)])
    data = p.read_bytes()
    lines = data.split(b"\n")
    assert lines[-1] == b"" and len(lines) == n + 1
    for i in list(range(300)) + list(range(n - 300, n)) + list(rng.integers(0, n, 2000)):
        assert lines[i].decode() == " ".join(lab[t] for t in tok[i] if t >= 0), i
    lab_len = np.array([len(x) for x in lab])
    want_bytes = int(lab_len[tok[tok >= 0]].sum()) + int((tok >= 0).sum() - n) + n      # labels + separators + newlines
    assert len(data) == want_bytes
    back = dge_lib.Corpus.read_seq(ctx, str(p), region, layer)
    assert np.array_equal(back.tokens(), tok)
    c.free()
    back.free()
