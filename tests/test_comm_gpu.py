"""GPU tests of the multi-GPU row: the NCCL communicator of a ctx and the data-parallel skip-gram
(embedding deltas combined per row by the peer-memory kernel over NVLink, or by NCCL all-reduces).  The 1-GPU cases run the same code path with a world of one; the 2-rank case needs two
GPUs on the box (gpurun --gpus 2) and is skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def corpus_tokens(n=4000, L=8, seed=2):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 30, size=(n // 2, L))
    b = rng.integers(30, 60, size=(n - n // 2, L))
    t = np.concatenate([a, b]).astype(np.int32)
    rng.shuffle(t)
    return t


def test_rounds_are_invisible_in_the_sequential_schedule(dge_lib, ctx):
    """Cutting the epoch into slices with a delta exchange in between (identity in a world of one) must not change
    what a sequential run computes beyond the rounding of (cur - base) + base."""
    tok = corpus_tokens()
    c = dge_lib.Corpus.from_tokens(ctx, tok, 60)
    kw = dict(dim=20, window=5, negative=5, min_count=2, seed=5, concurrency=1)
    m1 = dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(**kw))
    m2 = dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(sync_rounds=7, **kw))
    assert ctx.phase_ms("sgns_rounds") == 7
    a0, a1, ia = m1.vectors(want_syn1neg=True)
    b0, b1, ib = m2.vectors(want_syn1neg=True)
    assert m1.pairs == m2.pairs and np.array_equal(ia, ib)
    # fp32 rounding of the extra subtract / add is amplified by the ~1e5 dependent updates that follow
    assert np.allclose(a0, b0, rtol=0, atol=2e-4) and np.allclose(a1, b1, rtol=0, atol=2e-4)


def test_communicator_of_one(dge_lib):
    """dge_comm_unique_id / dge_comm_init / dge_comm_shape / dge_comm_destroy on a single GPU."""
    assert dge_lib.lib().dge_comm_nccl_version() >= 22000
    c = dge_lib.Context(0)
    assert c.comm_shape() == (0, 1)
    uid = dge_lib.Context.comm_unique_id()
    assert len(uid) == dge_lib.COMM_ID_BYTES and any(uid)
    c.comm_init(0, 1, uid)
    assert c.comm_shape() == (0, 1)
    with pytest.raises(dge_lib.DgeError):
        c.comm_init(0, 1, uid)                       # a ctx carries one communicator
    tok = corpus_tokens(1000)
    corp = dge_lib.Corpus.from_tokens(c, tok, 60)
    m = dge_lib.Model.train(c, [corp], dge_lib.sgns_params(dim=8, window=5, seed=1, sync_rounds=3))
    assert np.isfinite(m.vectors()[0]).all() and m.pairs > 0
    m.free()
    corp.free()
    c.close()


def test_comm_error_behaviour(dge_lib, ctx):
    with pytest.raises(dge_lib.DgeError):
        ctx.comm_init(2, 2, b"\0" * dge_lib.COMM_ID_BYTES)      # rank out of range
    with pytest.raises(ValueError):
        ctx.comm_init(0, 1, b"short")


def _device_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.skipif(_device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("transport", [1, 2])          # 1: peer-memory kernel over NVLink (cudaIpc), 2: NCCL all-reduce
def test_two_rank_data_parallel_skipgram(dge_lib, tmp_path, transport):
    """Two processes, one GPU each: walk ids sharded, vocabulary from the all-reduced counts, embedding deltas combined
    per row (alignment-weighted rule) by the peer-memory kernel / by NCCL.  Both ranks must end with bit-identical
    tables, the vocabulary must be the one of the whole corpus, the two ranks together must train EXACTLY the pairs of a
    single-GPU run over the whole corpus (global sentence indices key the RNG), and the embeddings must carry the same
    neighbourhood structure as that run."""
    from embedding_b200 import evaluation as ev
    worker = os.path.join(ROOT, "tests", "helpers", "dp_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(tmp_path), "24", str(transport)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert r0["transport"] == transport and r1["transport"] == transport
    assert np.array_equal(r0["ids"], r1["ids"])
    assert np.array_equal(r0["syn0"], r1["syn0"]) and np.array_equal(r0["syn1"], r1["syn1"])
    assert r0["rounds"] == 24 and np.isfinite(r0["syn0"]).all()
    assert np.linalg.norm(r0["syn0"], axis=1).mean() < 5       # a diverging combine rule shows up as exploding rows
    whole = np.concatenate([r0["tok"], r1["tok"]])
    nv = 300 * 8
    cnt = np.bincount(whole[whole >= 0], minlength=nv)
    order = sorted([i for i in range(nv) if cnt[i] >= 2], key=lambda i: (-cnt[i], i))
    assert np.array_equal(r0["ids"], np.array(order, np.int32))
    # single-GPU run over the whole corpus, same hyper-parameters
    ctx = dge_lib.Context(0)
    c = dge_lib.Corpus.from_tokens(ctx, whole, nv)
    kw = dict(dim=32, window=5, negative=5, min_count=2)
    m = dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(seed=3, **kw))
    s0, ids = m.vectors()
    assert np.array_equal(ids, r0["ids"])
    assert int(r0["pairs"]) + int(r1["pairs"]) == m.pairs       # same pairs, same negatives: only the interleaving differs
    m2 = dge_lib.Model.train(ctx, [c], dge_lib.sgns_params(seed=4, **kw))
    zeros, ar = np.zeros(nv, np.int32), np.arange(nv, dtype=np.int32)
    la = ev.layers_from_model(s0, ids, zeros, ar)
    lb = ev.layers_from_model(r0["syn0"], r0["ids"], zeros, ar)
    ls = ev.layers_from_model(m2.vectors()[0], ids, zeros, ar)
    lc = {0: (np.random.default_rng(0).standard_normal(s0.shape), la[0][1])}
    ov, ov_seed, ov_rand = ev.knn_overlap(la, lb, 10), ev.knn_overlap(la, ls, 10), ev.knn_overlap(la, lc, 10)
    print("kNN agreement with the single-GPU run: data-parallel %.3f, another seed %.3f, random %.3f" % (ov, ov_seed, ov_rand))
    # Calibration: the oracle's emulation of this exchange (24 rounds, two ranks, alignment-weighted rule) agrees with the
    # sequential run to 0.65, two sequential runs with different seeds to 0.50 (scripts/dp_emulation_sweep.py, DESIGN.md
    # 3.4); the plain sum of the deltas of round 1 fell to 0.10 on the GPUs.  Bar: at least 0.35, and not worse than
    # 0.8 x what two single-GPU seeds agree to.
    assert ov > 20 * ov_rand and ov >= 0.35 and ov >= 0.8 * ov_seed, (ov, ov_seed, ov_rand)
    ctx.close()
