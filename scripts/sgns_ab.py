"""A/B of skip-gram kernel variants / schedules on ONE corpus (run on the GPU box): throughput, and -- with --quality --
the reference's downstream metric and the neighbourhood agreement with the 8-thread CPU oracle on the same corpus.

    python scripts/sgns_ab.py tract24 4000000 --variants v2:0,staged:256,plain:512,noupd:1 [--quality]
    python scripts/sgns_ab.py ca 2000000 --conc 0,206,412,824,1648 --quality
    python scripts/sgns_ab.py synth 100000 2000000 --dim 128 --variants v2:0,staged:256

variants = name:flags (dge.h DGE_SGNS_F_*); --conc = sentences in flight (0 = the library's automatic bound).
Output: one JSON line per run, and gpurun_out/sgns_ab_<tag>.json.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embedding_b200 import abi, evaluation as ev, host, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("level")
    ap.add_argument("n", type=int, nargs="+")
    ap.add_argument("--variants", default="v2:0")
    ap.add_argument("--conc", default="0")
    ap.add_argument("--dim", type=int, default=0)
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--negative", type=int, default=5)
    ap.add_argument("--quality", action="store_true")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    ctx = abi.Context(0)
    level = args.level
    labels = gt = None
    if level == "synth":
        n_regions, n_walks = args.n[0], args.n[1]
        g = synth.powerlaw_flow_graph(n_regions, L=24, seed=100000)
        G = abi.Graph(ctx, g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"])
        L, dim, window, nv, v_layer, v_region, ids = 24, args.dim or 128, args.window or 10, g["n_vertices"], g["v_layer"], g["v_region"], None
    else:
        n_walks = args.n[0]
        if level == "ca":
            ids, z, L, dim = synth.ca_ids(), synth.ca_latents(), 24, args.dim or 8
            fl = host.Flows(ids, synth.planted_flow_tensor(z, mean_trips_per_pair_hour=2.0))
            host.CrossTimeGraph.numLayer = L
            gh = host.CrossTimeGraph.constructGraph_CA(fl, ctx=ctx)
            with open(os.path.join(ROOT, "tests", "golden", "ca_labels.json")) as f:
                d = json.load(f)
            labels = {"crime": d["crime-label"], "lehd": d["lehd-label"]}
            labels.update(d["demo-label"])
            labels.update(d["poi-label"])
        else:
            ids, z, L, dim = synth.tract_ids(), synth.poi_latents(), (24 if level == "tract24" else 8), args.dim or 20
            fl = host.Flows(ids, synth.planted_flow_tensor(z))
            host.CrossTimeGraph.numLayer = L
            gh = host.CrossTimeGraph.constructGraph_tract(fl, ctx=ctx)
            gt = ev.PairwiseGroundTruth(ids, z)
        gh.initiateAliasTables()
        G, nv, v_layer, v_region, window = gh._graph, gh.n_vertices, gh.v_layer, gh.v_region, args.window or L
    corpus = G.walk(n_walks, L, seed=2013)
    kw = dict(dim=dim, window=window, negative=args.negative, min_count=2)
    results = []
    base_table = None

    def quality(syn0, idw):
        nonlocal base_table
        if not args.quality or ids is None:
            return {}
        layers = ev.layers_from_model(syn0, idw, v_layer, v_region)
        table = ev.knn_table(layers, ids, L, 10)
        if base_table is None:
            base_table = table
        q = dict(knn_overlap_vs_oracle=round(ev.knn_table_overlap(base_table, table), 4),
                 mean_row_norm=float(np.linalg.norm(syn0, axis=1).mean()), finite=bool(np.isfinite(syn0).all()))
        if gt is not None:
            q["ndcg"] = {str(k): round(v, 4) for k, v in ev.pairwise_ndcg(gt, layers, ks=(5, 20, 50)).items()}
        if labels is not None:
            q["cv_accuracy"] = ev.ca_classification_accuracy(layers, labels, ids)
        return q

    if args.quality and ids is not None:   # the checker: 8-thread CPU oracle on the same corpus, two seeds (noise floor)
        from oracle import oracle as O
        tok = corpus.tokens()
        for seed in (1, 2):
            t = time.time()
            m = O.sgns_train(tok, nv, O.sgns_params(threads=min(8, os.cpu_count() or 1), seed=seed, **kw))
            r = dict(name="oracle_t8_seed%d" % seed, seconds=round(time.time() - t, 1), pairs=int(m["pairs"]), **quality(m["syn0"], m["id_of_word"]))
            results.append(r)
            print(json.dumps(r), flush=True)
    for spec in args.variants.split(","):
        name, flags = spec.split(":")
        for conc in [int(c) for c in args.conc.split(",")]:
            best = None
            for rep in range(args.reps):
                m = abi.Model.train(ctx, [corpus], abi.sgns_params(seed=1, flags=int(flags), concurrency=conc, **kw))
                ms = ctx.phase_ms("sgns")
                if best is None or ms < best[0]:
                    best = (ms, m)
            ms, m = best
            r = dict(name=name, flags=int(flags), concurrency=conc, sgns_ms=round(ms, 3), pairs=int(m.pairs), gpairs_per_s=round(m.pairs / ms / 1e6, 4),
                     groups=ctx.phase_ms("sgns_groups"), kernel=int(ctx.phase_ms("sgns_kernel")))
            if args.quality and not (int(flags) & 1):
                syn0, idw = m.vectors()
                r.update(quality(syn0, idw))
            results.append(r)
            print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = args.tag or "%s_%d" % (level, n_walks)
    json.dump(dict(level=level, n_walks=n_walks, L=L, params=kw, results=results), open(os.path.join(ROOT, "gpurun_out", "sgns_ab_%s.json" % tag), "w"), indent=1)


if __name__ == "__main__":
    main()
