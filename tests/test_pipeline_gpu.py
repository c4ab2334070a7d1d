"""GPU end-to-end: the reference's entry points (host mirror) over the C ABI, file formats, and downstream-metric
parity of stage 2 against the CPU oracle (the yardstick of SURVEY 8(c): pairwise nDCG@k vs POI ground truth)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# |nDCG@k(GPU) - nDCG@k(oracle)| on identical walks; the oracle's own seed-to-seed / thread-count spread on this
# input is ~0.003 (profiles/quality_tract_r1.json)
NDCG_TOL = 0.006


@pytest.fixture(scope="module")
def tract_setup(ctx):
    from embedding_b200 import evaluation as ev, host, synth
    ids, z = synth.tract_ids(), synth.poi_latents()
    fl = host.Flows(ids, synth.planted_flow_tensor(z))
    host.CrossTimeGraph.numLayer = 8
    g = host.CrossTimeGraph.constructGraph_tract(fl, ctx=ctx)
    g.initiateAliasTables()
    return dict(ids=ids, z=z, g=g, gt=ev.PairwiseGroundTruth(ids, z))


def test_downstream_metric_parity_with_oracle(dge_lib, oracle, ctx, tract_setup):
    from embedding_b200 import evaluation as ev
    g, gt = tract_setup["g"], tract_setup["gt"]
    corpus = g._graph.walk(600_000, 8, seed=2013)
    tok = corpus.tokens()
    kw = dict(dim=20, window=8, negative=5, min_count=2)
    ref = oracle.sgns_train(tok, g.n_vertices, oracle.sgns_params(threads=8, seed=1, **kw))
    m = dge_lib.Model.train(ctx, [corpus], dge_lib.sgns_params(seed=1, **kw))   # automatic schedule
    assert m.pairs == ref["pairs"]
    syn0, idw = m.vectors()
    a = ev.pairwise_ndcg(gt, ev.layers_from_model(ref["syn0"], ref["id_of_word"], g.v_layer, g.v_region), ks=(5, 20, 50))
    b = ev.pairwise_ndcg(gt, ev.layers_from_model(syn0, idw, g.v_layer, g.v_region), ks=(5, 20, 50))
    rng = np.random.default_rng(0)
    chance = ev.pairwise_ndcg(gt, {0: (rng.normal(size=(801, 20)), tract_setup["ids"])}, ks=(5,))[5]
    assert a[5] > chance + 0.03          # the planted structure is learnt at all
    for k in (5, 20, 50):
        assert abs(a[k] - b[k]) < NDCG_TOL, (k, a, b)


def test_deepwalk_entry_points_write_reference_formats(dge_lib, ctx, tmp_path, tract_setup):
    """CrossTimeGraph.outputSampleSequence + SpatialGraph.outputSampleSequence + DeepWalk.learnEmbedding through
    the host mirror; files are parsed with the reference consumers' logic (embeddingEvaluation_tract.py:139-166)."""
    from embedding_b200 import evaluation as ev, host, synth
    ids = tract_setup["ids"]
    fl = host.Flows(ids, synth.planted_flow_tensor(tract_setup["z"]))
    host.DeepWalk.base_dir = str(tmp_path)
    host.CrossTimeGraph.numSamples, host.CrossTimeGraph.numLayer = 50_000, 8
    host.SpatialGraph.numSamples, host.SpatialGraph.numLayer = 5_000, 8
    seq1 = host.DeepWalk._seq_path("tract", "crosstime")
    seq2 = host.DeepWalk._seq_path("tract", "spatial")
    g1, c1 = host.CrossTimeGraph.outputSampleSequence("tract", fl, seq1, ctx=ctx)
    g2, c2 = host.SpatialGraph.outputSampleSequence("tract", ids, synth.spatial_weights(len(ids)), seq2, ctx=ctx)
    l1 = open(seq1).read().split("\n")
    l2 = open(seq2).read().split("\n")
    assert len(l1) == 50_001 and len(l2) == 5_001
    tok = c1.tokens()
    for row, line in zip(tok[:50], l1):
        assert line == " ".join(g1._names[t] for t in row if t >= 0)
    assert all(p.split("-")[0] == str(j) for j, p in enumerate(l2[0].split(" ")))   # "<j>-<region>"
    # joint vocabulary: (layer | position, region) -> id
    n = len(ids)
    pos = {int(r): i for i, r in enumerate(ids)}
    c1.relabel((g1.v_layer.astype(np.int64) * n + np.array([pos[int(r)] for r in g1.v_region])).astype(np.int32), 8 * n)
    c2.relabel(np.array([pos[int(r)] for r in g2.v_region], np.int32), 8 * n, position_stride=n)
    lab_layer = (np.arange(8 * n) // n).astype(np.int32)
    lab_region = ids[np.arange(8 * n) % n].astype(np.int32)
    model = host.DeepWalk.learnEmbedding("tract", "usespatial", [c1, c2], (lab_layer, lab_region), ctx=ctx)
    out = os.path.join(str(tmp_path), "miscs", "2013", "taxi-deepwalk-tract-usespatial-2D.vec")
    layers = ev.read_vec(out)
    assert sorted(layers) == list(range(8))
    assert sum(len(r) for _, r in layers.values()) == model.V
    assert all(f.shape[1] == 20 for f, _ in layers.values())            # flowFeatureGeneration_tract.py:82 asserts 20
    syn0, idw = model.vectors()
    f0, r0 = layers[int(lab_layer[idw[0]])]
    assert np.allclose(f0[0], syn0[0], rtol=1e-6) and r0[0] == lab_region[idw[0]]


def test_deepwalk_main_runs_the_whole_path(dge_lib, ctx, tmp_path, tract_setup):
    """DeepWalk.main [regionLevel] [spatialGF] [Year] (DeepWalk.java:120-140) through the host mirror at 0.2 % of the
    reference's corpus sizes: both .seq files and the .vec file appear under ../miscs/<Year>/ with the reference's names."""
    from embedding_b200 import evaluation as ev, host, synth
    ids = tract_setup["ids"]
    fl = host.Flows(ids, synth.planted_flow_tensor(tract_setup["z"]))
    host.DeepWalk.base_dir = str(tmp_path)
    host.DeepWalk._sample_scale = 0.002
    try:
        model = host.DeepWalk.main(["tract", "usespatial", "2014"], fl, synth.spatial_weights(len(ids)), ctx=ctx)
    finally:
        host.DeepWalk._sample_scale = 1.0
    assert model is not None and host.DeepWalk.Year == 2014
    d = os.path.join(str(tmp_path), "miscs", "2014")
    assert len(open(os.path.join(d, "deepwalkseq-tract", "taxi-crosstime.seq")).read().split("\n")) == 30_001
    assert len(open(os.path.join(d, "deepwalkseq-tract", "taxi-spatial.seq")).read().split("\n")) == 1_201
    layers = ev.read_vec(os.path.join(d, "taxi-deepwalk-tract-usespatial-2D.vec"))
    assert sorted(layers) == list(range(8)) and all(f.shape[1] == 20 for f, _ in layers.values())
    assert host.DeepWalk.main(["county"], fl, None, ctx=ctx) is None       # bad argument: reported, swallowed (:137-139)
    host.DeepWalk.Year = 2013

