"""GPU parity of the downstream pairwise-similarity metric: dge_eval_knn / dge_eval_ndcg against
the Python 3 port of python/embeddingEvaluation_tract.py (embedding_b200/evaluation.py), on the real POI ground truth
fixture.  Index work (neighbour lists) must be identical wherever the fp64 distances are not within rounding of a tie;
nDCG within 1e-9."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_knn_matches_numpy(dge_lib, ctx):
    from embedding_b200 import evaluation as ev
    rng = np.random.default_rng(5)
    X = rng.normal(size=(700, 20)).astype(np.float32)
    X[13] = 0                                            # zero vector: distance NaN -> 2 in the reference
    X[200] = X[100]                                      # exact tie: both rows are at the same distance from everyone
    nbr, dist = ctx.eval_knn(X, 30, want_dist=True)
    D = ev.cosine_distance_matrix(X)
    D[np.arange(700), np.arange(700)] = np.inf
    ref = np.argsort(D, axis=1, kind="stable")[:, :30]
    same = (nbr == ref).all(axis=1)
    assert same.mean() > 0.995                           # the rest: distances equal to within an ulp in another order
    assert np.allclose(dist, np.take_along_axis(D, ref, axis=1), rtol=0, atol=1e-12)
    assert (np.diff(dist, axis=1) >= 0).all() and (nbr != np.arange(700)[:, None]).all()
    assert dist[13, 0] == 2.0 and (dist[13] == 2.0).all() and np.array_equal(nbr[13], np.r_[0:13, 14:31])   # ties by index
    small = ctx.eval_knn(X[:5], 8)                       # fewer than topk other rows: -1 padded
    assert (small[:, 4:] == -1).all() and (small[:, :4] >= 0).all()


def test_ndcg_matches_the_python_port(dge_lib, ctx):
    from embedding_b200 import evaluation as ev, synth
    ids, z = synth.tract_ids(), synth.poi_latents()
    gt = ev.PairwiseGroundTruth(ids, z)
    rng = np.random.default_rng(9)
    layers = {}
    for h in range(3):
        keep = np.sort(rng.choice(len(ids), size=600 + 50 * h, replace=False))
        f = (z[keep] @ rng.normal(size=(z.shape[1], 20)) + rng.normal(size=(len(keep), 20))).astype(np.float32)
        layers[h] = (f, ids[keep])
    cpu = ev.pairwise_ndcg(gt, layers, ks=(5, 20, 70))
    gpu = ev.pairwise_ndcg_device(ctx, gt, layers, ks=(5, 20, 70))
    for k in cpu:
        assert abs(cpu[k] - gpu[k]) < 1e-9, (k, cpu[k], gpu[k])
    f, r = layers[0]
    gi = np.array([gt.index[int(x)] for x in r], np.int32)
    per_row, mean = ctx.eval_ndcg(f, gi, gt.D, 10)
    assert per_row.shape == (len(r),) and abs(per_row.mean() - mean) < 1e-12 and 0 < mean <= 1
    assert np.isnan(ctx.eval_ndcg(f[:8], gi[:8], gt.D, 10)[1])       # m <= topk: the reference reports nothing
    with pytest.raises(dge_lib.DgeError):
        ctx.eval_ndcg(f, gi + 10_000, gt.D, 10)
