"""Data-parallel skip-gram, emulated on the CPU (oracle/sgns_oracle.c ora_sgns_train_dp): which rule should combine
the per-rank embedding deltas, and what does an exchange that lands one slice late (overlapped with compute) cost?
Reports kNN agreement (k = 10) with the sequential run on the whole corpus -- two sequential runs with different
seeds give the noise floor -- and, on the tract workload, the reference's nDCG@k.

    python scripts/dp_emulation_sweep.py [small|tract8] [out.json]
"""
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embedding_b200 import evaluation as ev, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

RULES = {0: "sum", 1: "mean", 2: "contributors", 3: "sqrt", 4: "aligned"}


def setup(kind):
    if kind == "small":       # the corpus of tests/test_comm_gpu.py, 10 x the walks
        g = synth.powerlaw_flow_graph(3000, L=8, seed=5, mean_degree=8, cap=64)
        G = O.Graph(g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"], alias_mode=O.ALIAS_FAST)
        tok = G.walk(400_000, 8, seed=11)
        nv = g["n_vertices"]
        return dict(tok=tok, nv=nv, v_layer=np.zeros(nv, np.int32), v_region=np.arange(nv, dtype=np.int32), gt=None,
                    kw=dict(dim=32, window=5, negative=5, min_count=2, seed=3, threads=1))
    from embedding_b200 import host
    ids, z = synth.tract_ids(), synth.poi_latents()
    fl = host.Flows(ids, synth.planted_flow_tensor(z))
    host.CrossTimeGraph.numLayer = 8
    g = host.CrossTimeGraph.constructGraph_tract(fl)
    nv, src, dst, wt = g._bulk
    G = O.Graph(nv, src, dst, wt, np.array(g.sourceVertices, np.int32), alias_mode=O.ALIAS_FAST)
    tok = G.walk(600_000, 8, seed=2013)
    return dict(tok=tok, nv=nv, v_layer=g.v_layer, v_region=g.v_region, gt=(ids, z),
                kw=dict(dim=20, window=8, negative=5, min_count=2, seed=1, threads=1))


_S = None


def _run(job):
    global _S
    kind, name, world, rounds, combine, seed = job
    if _S is None:
        _S = setup(kind)
    S = _S
    kw = dict(S["kw"])
    if seed is not None:
        kw["seed"] = seed
    t = time.time()
    if world == 0:
        m = O.sgns_train(S["tok"], S["nv"], O.sgns_params(**kw))
    else:
        m = O.sgns_train_dp(S["tok"], S["nv"], O.sgns_params(**kw), world, rounds, combine)
    return name, m["syn0"], m["id_of_word"], time.time() - t


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "small"
    out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "dp_emulation_%s.json" % kind)
    S = setup(kind)
    jobs = [(kind, "sequential", 0, 0, 0, None), (kind, "sequential_seed2", 0, 0, 0, S["kw"]["seed"] + 1)]
    for world in (2, 8):
        for rounds in (8, 32, 128):
            for combine in (0, 2, 3, 4, 16 + 2, 16 + 4):
                if combine == 0 and world == 8 and rounds > 8:
                    continue
                jobs.append((kind, "w%d_r%d_%s%s" % (world, rounds, RULES[combine & 15], "_delayed" if combine & 16 else ""),
                             world, rounds, combine, None))
    res = []
    base = None
    gt = ev.PairwiseGroundTruth(*S["gt"]) if S["gt"] else None
    with ProcessPoolExecutor(max_workers=int(os.environ.get("DP_SWEEP_PROCS", "4"))) as ex:
        for name, syn0, idw, sec in ex.map(_run, jobs):
            layers = ev.layers_from_model(syn0, idw, S["v_layer"], S["v_region"])
            if base is None:
                base = layers
            r = dict(name=name, knn_overlap_vs_sequential=round(ev.knn_overlap(base, layers, 10), 4),
                     mean_row_norm=float(np.linalg.norm(syn0, axis=1).mean()), finite=bool(np.isfinite(syn0).all()), seconds=round(sec, 1))
            if gt is not None:
                r["ndcg"] = {str(k): round(v, 4) for k, v in ev.pairwise_ndcg(gt, layers, ks=(5, 20, 50)).items()}
            res.append(r)
            print(json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump(dict(kind=kind, n_sentences=int(S["tok"].shape[0]), params=S["kw"], results=res), open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
