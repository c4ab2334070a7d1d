"""GPU parity, stage 2: skip-gram negative sampling.  Parity is UNPINNED against DL4J (sources absent, see
oracle/sgns_oracle.h); what is checked: (i) identical vocabulary / pair / negative enumeration and fp32-
tolerance agreement with the CPU oracle in a sequential schedule, (ii) the parallel Hogwild schedule reaches
the same embedding quality."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# fp32, ~1e4 dependent updates, GPU uses FMA + tree-reduced dots, the oracle sequential non-fused adds
SEQ_ATOL = 2e-4


def community_corpus(n=3000, L=8, seed=1):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 10, size=(n, L))
    b = rng.integers(10, 20, size=(n, L))
    t = np.concatenate([a, b]).astype(np.int32)
    rng.shuffle(t)
    return t


@pytest.mark.parametrize("dim", [2, 8, 20, 64, 128])
def test_sequential_schedule_matches_oracle(dge_lib, oracle, ctx, dim):
    rng = np.random.default_rng(dim)
    tok = rng.integers(0, 40, size=(300, 8)).astype(np.int32)
    tok[rng.random(tok.shape) < 0.08] = -1
    tok[:, 0] = np.maximum(tok[:, 0], 0)
    kw = dict(dim=dim, window=5, negative=5, min_count=2, seed=17)
    ref = oracle.sgns_train(tok, 41, oracle.sgns_params(threads=1, **kw))
    c = dge_lib.Corpus.from_tokens(ctx, tok, 41)
    m = dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(concurrency=1, **kw))
    syn0, syn1, ids = m.vectors(want_syn1neg=True)
    assert m.V == len(ref["id_of_word"]) and np.array_equal(ids, ref["id_of_word"])
    assert m.pairs == ref["pairs"]
    assert np.allclose(syn0, ref["syn0"], rtol=0, atol=SEQ_ATOL)
    assert np.allclose(syn1, ref["syn1neg"], rtol=0, atol=SEQ_ATOL)


def test_init_is_bit_identical_to_oracle(dge_lib, oracle, ctx):
    """One sentence of one token trains nothing: syn0 is the (U-0.5)/dim initialisation."""
    tok = np.zeros((5, 1), np.int32)
    tok[:, 0] = [0, 1, 2, 1, 0]
    for dim in (2, 20, 128):
        c = dge_lib.Corpus.from_tokens(ctx, tok, 3)
        m = dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(dim=dim, min_count=1, seed=5))
        syn0, ids = m.vectors()
        assert m.pairs == 0
        assert np.array_equal(syn0, oracle.init_syn0(3, dim, 5))


def test_two_corpora_share_one_vocabulary(dge_lib, oracle, ctx):
    """usespatial: FileSentenceIterator over both .seq files (DeepWalk.java:48-50)."""
    rng = np.random.default_rng(4)
    a = rng.integers(0, 30, size=(200, 8)).astype(np.int32)
    b = rng.integers(0, 30, size=(100, 5)).astype(np.int32)
    kw = dict(dim=8, window=8, min_count=1, seed=2)
    bp = np.full((100, 8), -1, np.int32)
    bp[:, :5] = b
    ref = oracle.sgns_train(np.concatenate([a, bp]), 30, oracle.sgns_params(**kw))
    ca, cb = dge_lib.Corpus.from_tokens(ctx, a, 30), dge_lib.Corpus.from_tokens(ctx, b, 30)
    m = dge_lib.Model.train(ctx, [ca, cb], dge_lib.sgns_params(concurrency=1, **kw))
    syn0, ids = m.vectors()
    assert m.pairs == ref["pairs"] and np.array_equal(ids, ref["id_of_word"])
    assert np.allclose(syn0, ref["syn0"], rtol=0, atol=SEQ_ATOL)


@pytest.mark.parametrize("dim", [8, 20, 128])
def test_parallel_schedule_learns_the_same_structure(dge_lib, oracle, ctx, dim):
    tok = community_corpus()
    kw = dict(dim=dim, window=5, min_count=1, seed=3)
    ref = oracle.sgns_train(tok, 20, oracle.sgns_params(threads=8, **kw))
    c = dge_lib.Corpus.from_tokens(ctx, tok, 20)
    m = dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(**kw))
    assert m.pairs == ref["pairs"]  # the enumeration does not depend on the schedule
    syn0, ids = m.vectors()
    assert np.isfinite(syn0).all()

    def gap(emb, ids):
        e = np.zeros((20, emb.shape[1]), np.float32)
        e[ids] = emb
        e /= np.linalg.norm(e, axis=1, keepdims=True)
        s = e @ e.T
        within = ((s[:10, :10].sum() - 10) + (s[10:, 10:].sum() - 10)) / 180
        return within - s[:10, 10:].mean()

    g_ref, g_gpu = gap(ref["syn0"], ref["id_of_word"]), gap(syn0, ids)
    assert g_ref > 0.3
    assert g_gpu > 0.5 * g_ref, (g_gpu, g_ref)


@pytest.mark.parametrize("dim,negative", [(2, 5), (4, 7), (8, 5), (8, 10), (12, 3), (16, 5), (20, 5), (20, 12), (32, 5), (64, 5), (100, 7), (128, 5), (128, 20)])
def test_item_kernel_arithmetic_matches_oracle(dge_lib, oracle, ctx, dim, negative):
    """The throughput kernel (work item = (sentence, centre), 128-bit L2 reductions, software pipeline) enumerates
    the oracle's pairs and negatives and applies the same update arithmetic; only the interleaving differs.
    Run on ONE warp, one item at a time (flags = DGE_SGNS_F_ONE_WARP: items strictly in corpus order) the interleaving is
    the oracle's up to the prefetch of a pair's rows, so the learned change of both tables must agree with the
    sequential oracle to ~1 %: any slip in the dot products, the sigmoid table, the negative draws or the row
    addressing is an O(1) error here.  With sentences in flight the deviation grows like sqrt(stale fraction)."""
    rng = np.random.default_rng(100 + dim + negative)
    n_ids = 3000
    tok = rng.integers(0, n_ids, size=(500, 12)).astype(np.int32)
    tok[rng.random(tok.shape) < 0.1] = -1
    tok[:, 0] = np.maximum(tok[:, 0], 0)
    kw = dict(dim=dim, window=6, negative=negative, min_count=1, seed=23)
    ref = oracle.sgns_train(tok, n_ids, oracle.sgns_params(threads=1, **kw))
    init = oracle.init_syn0(len(ref["id_of_word"]), dim, 23)
    learned0 = np.linalg.norm(ref["syn0"] - init)
    learned1 = np.linalg.norm(ref["syn1neg"])
    assert learned0 > 0 and learned1 > 0
    c = dge_lib.Corpus.from_tokens(ctx, tok, n_ids)

    def rel_err(**p):
        m = dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(**p, **kw))
        syn0, syn1, ids = m.vectors(want_syn1neg=True)
        assert m.pairs == ref["pairs"] and np.array_equal(ids, ref["id_of_word"])
        return np.linalg.norm(syn0 - ref["syn0"]) / learned0, np.linalg.norm(syn1 - ref["syn1neg"]) / learned1

    L = dge_lib
    flag_sets = [L.F_ONE_WARP, L.F_ONE_WARP | L.F_ITEM_KERNELS]           # kernel F / kernel C on one lane group
    if dim <= 16:
        flag_sets.append(L.F_ONE_WARP | L.F_NARROW)                  # the 4-lane-group kernel for D <= 16
        if negative <= 7:
            flag_sets.append(L.F_ONE_WARP | L.F_TARGET_PARALLEL)     # the target-parallel kernel (D <= 16, K <= 7)
    if dim <= 128:
        flag_sets.append(L.F_ONE_WARP | L.F_SENTENCE_RESIDENT)       # kernel F on one lane group: the oracle's exact pair order
        flag_sets.append(L.F_ONE_WARP | L.F_SENTENCE_RESIDENT | L.F_SMALL_BLOCKS)
        flag_sets.append(L.F_ONE_WARP | L.F_STAGED_ROWS)             # kernel C': rows staged in shared memory by cp.async
        flag_sets.append(L.F_ONE_WARP | L.F_PLAIN_STORES)            # atomic-free build: one item at a time loses no update
    if dim <= 32 and negative <= 5:
        # kernel F with write-through words (bits 20-23: the 2^(v-1) most frequent words are re-read for every pair and
        # updated at once), without and with the row prefetch (cached last update + unit tag): still the oracle's order
        flag_sets.append(L.F_ONE_WARP | L.F_SENTENCE_RESIDENT | (8 << 20))
        flag_sets.append(L.F_ONE_WARP | L.F_SENTENCE_RESIDENT | L.F_ROW_PREFETCH)
        flag_sets.append(L.F_ONE_WARP | L.F_SENTENCE_RESIDENT | L.F_ROW_PREFETCH | (8 << 20))
        flag_sets.append(L.F_ONE_WARP | L.F_SENTENCE_RESIDENT | L.F_ROW_PREFETCH | (12 << 20))
        flag_sets.append(L.F_ONE_WARP | L.F_SENTENCE_RESIDENT | L.F_ROW_PREFETCH_SMEM)             # ... requested into shared memory (cp.async)
        flag_sets.append(L.F_ONE_WARP | L.F_SENTENCE_RESIDENT | L.F_ROW_PREFETCH_SMEM | (8 << 20))
    for flags in flag_sets:
        e0, e1 = rel_err(flags=flags)
        # plain stores lose an update whenever one row is drawn twice within a unit (both copies start from the same load)
        tol = 0.06 if flags & L.F_PLAIN_STORES else 0.02
        assert e0 < tol and e1 < tol, (flags, e0, e1)
    if dim <= 128:
        # kernel G (a block per sentence, one lane group per centre, wavefront over the contexts), ONE sentence in flight: a
        # schedule that is conflict-equivalent to the oracle's centre-major loop (up to a negative that happens to be a
        # word of the same sentence)
        e0, e1 = rel_err(concurrency=1, schedule=dge_lib.SCHEDULE_ITEMS, flags=L.F_SENTENCE_RESIDENT | L.F_BLOCK_PER_SENTENCE | L.F_NO_TARGET_PARALLEL)
        assert e0 < 0.05 and e1 < 0.05, ("kernel G", e0, e1)
        if dim <= 32 and negative <= 7:
            # kernel I: the same wavefront, a warp per pair, the round's pairs handed to the warps dynamically (8 and 12 warps)
            for warps in (0, 12 << 12):
                e0, e1 = rel_err(concurrency=1, schedule=dge_lib.SCHEDULE_ITEMS,
                                 flags=L.F_SENTENCE_RESIDENT | L.F_BLOCK_PER_SENTENCE | L.F_NO_TARGET_PARALLEL | L.F_PAIR_WARPS | warps)
                assert e0 < 0.05 and e1 < 0.05, ("kernel I", warps, e0, e1)
        if dim <= 32 and negative <= 5:
            # kernel J: the same wavefront, helper warps stage the rows (cp.async ring) and send the reductions (4 and 2 stages)
            for stages in (0, 2 << 12):
                e0, e1 = rel_err(concurrency=1, schedule=dge_lib.SCHEDULE_ITEMS,
                                 flags=L.F_SENTENCE_RESIDENT | L.F_BLOCK_PER_SENTENCE | L.F_NO_TARGET_PARALLEL | L.F_HELPER_WARPS | stages)
                assert e0 < 0.05 and e1 < 0.05, ("kernel J", stages, e0, e1)
        if dim <= 32 and negative <= 5:
            # kernel H: the same wavefront with the block's sentences pipelined -- one block: two sentences overlap at most
            e0, e1 = rel_err(concurrency=1, schedule=dge_lib.SCHEDULE_ITEMS, flags=L.F_SENTENCE_RESIDENT | L.F_BLOCK_PER_SENTENCE | L.F_NO_TARGET_PARALLEL | L.F_PIPELINED)
            assert e0 < 0.1 and e1 < 0.1, ("kernel H", e0, e1)
    if dim <= 32 and negative <= 5:
        # one warp, its four lane groups staggered over the contexts (a legitimate sequential order of the sentence's pairs),
        # write-through words and row prefetch: the cached-update bookkeeping between the groups must be exact
        b0, b1 = rel_err(concurrency=1, schedule=dge_lib.SCHEDULE_ITEMS, flags=L.F_SENTENCE_RESIDENT)
        for extra in (8 << 20, L.F_ROW_PREFETCH | (8 << 20), L.F_ROW_PREFETCH | (12 << 20), L.F_DYNAMIC | (8 << 20), L.F_ROW_PREFETCH_SMEM | (8 << 20)):
            e0, e1 = rel_err(concurrency=1, schedule=dge_lib.SCHEDULE_ITEMS, flags=L.F_SENTENCE_RESIDENT | extra)
            assert e0 < 1.25 * b0 + 0.02 and e1 < 1.25 * b1 + 0.02, ("kernel F, one sentence in flight", extra, e0, e1, b0, b1)
    e0, e1 = rel_err(concurrency=2)                 # two sentences in flight
    assert e0 < 0.35 and e1 < 0.35, (e0, e1)
    e0, e1 = rel_err()                              # automatic full-GPU schedule
    # everything in flight at once on a corpus this small: syn1neg (first-order updates) still agrees; syn0 learns
    # through syn1neg rows that are still zero when read, so only its scale is checked
    assert e1 < 0.5 and e0 < 1.5, (e0, e1)


def test_automatic_schedule_on_a_mid_size_vocabulary(dge_lib, oracle, ctx):
    """flags = 0, concurrency = 0 on narrow rows and a vocabulary of thousands of words: kernel F (a warp per sentence), the
    sentences handed out from the device counter, the most frequent words written through, a whole number of warps per SM in
    flight and never more sentences than vocabulary words.  The enumeration does not depend on the schedule:
    exactly the oracle's pairs.  The same corpus with write-through forced off / on by flags trains the same pairs too."""
    rng = np.random.default_rng(77)
    n_ids, n_sent, L = 12000, 60000, 12
    p = 1.0 / (np.arange(n_ids) + 20.0)                      # a Zipf-like word distribution: a few hub words
    tok = rng.choice(n_ids, size=(n_sent, L), p=p / p.sum()).astype(np.int32)
    kw = dict(dim=20, window=6, negative=5, min_count=2, seed=9)
    ref = oracle.sgns_train(tok, n_ids, oracle.sgns_params(threads=8, **kw))
    c = dge_lib.Corpus.from_tokens(ctx, tok, n_ids)
    m = dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(**kw))
    assert m.pairs == ref["pairs"] and m.V == len(ref["id_of_word"])
    assert int(ctx.phase_ms("sgns_kernel")) == 8                           # k_sgns_sent
    in_flight, hot = int(ctx.phase_ms("sgns_groups")), int(ctx.phase_ms("sgns_write_through"))
    assert 0 < hot <= m.V and in_flight <= m.V and in_flight <= n_sent
    syn0, _ = m.vectors()
    assert np.isfinite(syn0).all()
    norm_auto = float(np.linalg.norm(syn0, axis=1).mean())
    norm_ref = float(np.linalg.norm(ref["syn0"], axis=1).mean())
    assert 0.5 * norm_ref < norm_auto < 2.0 * norm_ref, (norm_auto, norm_ref)   # neither collapsed nor blown up (5 % of this corpus is in flight at once)
    L_ = dge_lib
    for flags in (L_.F_SENTENCE_RESIDENT | L_.F_DYNAMIC, L_.F_SENTENCE_RESIDENT | L_.F_DYNAMIC | (9 << 20)):
        m2 = dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(concurrency=in_flight, flags=flags, **kw))
        assert m2.pairs == ref["pairs"]
        assert int(ctx.phase_ms("sgns_write_through")) == (256 if (flags >> 20) & 15 else 0)


def test_vec_file_format(dge_lib, ctx, tmp_path):
    """WordVectorSerializer.writeWordVectors as its consumers read it (embeddingEvaluation_tract.py:139-166,
    skipheader=0): "<layer>-<region> v1 ... vD", one line per vocabulary word."""
    tok = community_corpus(200)
    c = dge_lib.Corpus.from_tokens(ctx, tok, 20)
    m = dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(dim=8, window=5, min_count=1))
    layer = (np.arange(20) % 4).astype(np.int32)
    region = (17000 + np.arange(20)).astype(np.int32)
    p = tmp_path / "taxi-deepwalk-tract-usespatial.vec"
    m.write_vec(str(p), layer, region)
    syn0, ids = m.vectors()
    lines = p.read_text().strip().split("\n")
    assert len(lines) == m.V
    for wd, line in enumerate(lines):
        parts = line.split(" ")
        k1, k2 = parts[0].split("-")
        assert int(k1) == layer[ids[wd]] and int(k2) == region[ids[wd]]
        assert np.array_equal(np.array(parts[1:], np.float32), syn0[wd])


def test_model_stats_match_numpy(dge_lib, ctx):
    tok = community_corpus(400)
    c = dge_lib.Corpus.from_tokens(ctx, tok, 20)
    for dim in (8, 20, 100):
        m = dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(dim=dim, window=5, min_count=1))
        syn0, syn1, _ = m.vectors(want_syn1neg=True)
        st = m.stats()
        assert st["nonfinite"] == 0
        assert np.isclose(st["mean_row_norm"], np.linalg.norm(syn0.astype(np.float64), axis=1).mean(), rtol=1e-5)
        assert np.isclose(st["max_abs"], max(np.abs(syn0).max(), np.abs(syn1).max()), rtol=1e-6)
        m.free()


def test_sgns_error_behaviour(dge_lib, ctx):
    c = dge_lib.Corpus.from_tokens(ctx, np.zeros((2, 2), np.int32), 1)
    with pytest.raises(dge_lib.DgeError):
        dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(dim=0))
    with pytest.raises(dge_lib.DgeError):
        dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(window=0))
    with pytest.raises(dge_lib.DgeError):
        dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(dim=1024))
    # everything below min_count: empty vocabulary, not an error
    m = dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(min_count=100))
    assert m.V == 0 and m.pairs == 0


def test_large_vocabulary_is_ranked_on_the_device(dge_lib, oracle, ctx):
    """Id spaces of 65 536 and more are ranked by a 64-bit radix sort on the device (descending count, ties by ascending
    id) and the unigram^0.75 table takes pow() once per run of equal counts: vocabulary order, negative table (through
    the pair count and the sequential schedule's result) and minWordFrequency filtering must still be the oracle's."""
    rng = np.random.default_rng(9)
    n_ids = 200_000
    tok = np.minimum((rng.pareto(0.9, size=(60_000, 6)) * 40).astype(np.int64), n_ids - 1).astype(np.int32)   # heavy ties in the tail
    tok[rng.random(tok.shape) < 0.05] = -1
    tok[:, 0] = np.maximum(tok[:, 0], 0)
    kw = dict(dim=8, window=3, negative=5, min_count=2, seed=4)
    _, want = oracle.vocab(tok, n_ids, 2)
    c = dge_lib.Corpus.from_tokens(ctx, tok, n_ids)
    m = dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(concurrency=1, **kw))
    syn0, ids = m.vectors()
    assert np.array_equal(ids, want["id_of_word"]) and m.V == want["V"]
    ref = oracle.sgns_train(tok, n_ids, oracle.sgns_params(threads=1, **kw))
    assert m.pairs == ref["pairs"]
    assert np.allclose(syn0, ref["syn0"], atol=2e-4)       # same negatives => the sequential schedule reproduces the oracle
    m.free()
    c.free()
