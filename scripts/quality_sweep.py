"""Stage-2 parity sweep (run on the GPU box): trains the same walk corpus with the CPU oracle (1 and 8 Hogwild
threads, two seeds) and with libdge under several schedules, and reports the reference's downstream metric
(pairwise nDCG@k vs POI ground truth for tracts; 10-fold CV accuracy for CAs) plus kNN agreement with the
single-thread oracle.  Output: gpurun_out/quality_<level>.json

    python scripts/quality_sweep.py tract 2000000
    python scripts/quality_sweep.py CA 300000
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embedding_b200 import abi, evaluation as ev, host, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)


def main():
    level = sys.argv[1] if len(sys.argv) > 1 else "tract"
    n_walks = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    ctx = abi.Context(0)
    if level == "tract":
        ids, z, L, dim = synth.tract_ids(), synth.poi_latents(), 8, 20
        fl = host.Flows(ids, synth.planted_flow_tensor(z))
        host.CrossTimeGraph.numLayer = L
        g = host.CrossTimeGraph.constructGraph_tract(fl, ctx=ctx)
        gt = ev.PairwiseGroundTruth(ids, z)
    else:
        ids, z, L, dim = synth.ca_ids(), synth.ca_latents(), 24, 8
        fl = host.Flows(ids, synth.planted_flow_tensor(z, mean_trips_per_pair_hour=2.0))
        host.CrossTimeGraph.numLayer = L
        g = host.CrossTimeGraph.constructGraph_CA(fl, ctx=ctx)
        with open(os.path.join(ROOT, "tests", "golden", "ca_labels.json")) as f:
            d = json.load(f)
        labels = {"crime": d["crime-label"], "lehd": d["lehd-label"]}
        labels.update(d["demo-label"])
        labels.update(d["poi-label"])
    g.initiateAliasTables()
    nv = g.n_vertices
    corpus = g._graph.walk(n_walks, L, seed=2013)
    tok = corpus.tokens()
    kw = dict(dim=dim, window=L, negative=5, min_count=2)
    results = []

    def metric(layers):
        if level == "tract":
            return ev.pairwise_ndcg(gt, layers, ks=(5, 20, 50))
        return ev.ca_classification_accuracy(layers, labels, ids)

    base_layers = None

    def record(name, syn0, idw, seconds, pairs, extra=None):
        nonlocal base_layers
        layers = ev.layers_from_model(syn0, idw, g.v_layer, g.v_region)
        if base_layers is None:
            base_layers = layers
        r = dict(name=name, seconds=seconds, pairs=int(pairs), mpairs_per_s=pairs / seconds / 1e6,
                 metric=metric(layers), knn_overlap_vs_first=ev.knn_overlap(base_layers, layers, 10),
                 finite=bool(np.isfinite(syn0).all()), mean_norm=float(np.linalg.norm(syn0, axis=1).mean()))
        if extra:
            r.update(extra)
        results.append(r)
        print(json.dumps(r), flush=True)

    for threads, seed in ((1, 1), (8, 1), (8, 2), (1, 2)):
        t = time.time()
        m = O.sgns_train(tok, nv, O.sgns_params(threads=threads, seed=seed, **kw))
        record("oracle_t%d_seed%d" % (threads, seed), m["syn0"], m["id_of_word"], time.time() - t, m["pairs"])

    def gpu(name, **extra):
        t = time.time()
        m = abi.Model.train(ctx, [corpus], abi.sgns_params(seed=1, **kw, **extra))
        wall = time.time() - t
        syn0, idw = m.vectors()
        record(name, syn0, idw, ctx.phase_ms("sgns") / 1e3, m.pairs,
               dict(wall_s=wall, groups=ctx.phase_ms("sgns_groups")))

    gpu("gpu_items_auto")
    for c in (64, 256, 512, 1024, 2048, 4096, 16384):
        gpu("gpu_items_c%d" % c, concurrency=c)
    for c in (1024, 0):
        gpu("gpu_sentence_c%d" % c, concurrency=c, schedule=abi.SCHEDULE_SENTENCE)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "quality_%s.json" % level), "w") as f:
        json.dump(dict(level=level, n_walks=n_walks, L=L, dim=dim, nv=nv, ne=g._graph.ne, results=results), f, indent=1)


if __name__ == "__main__":
    main()
