"""Generates tests/golden/fullsize_<workload>_oracle.json: the reference's downstream metric of the CPU ORACLE's
skip-gram at the FULL bench size (BASELINE.json configs[1]: tract x 24, 15,000,000 flow walks + 600,000 spatial walks,
D=20, window=24, K=5 -- DeepWalk.java:89-110,73-76), on exactly the corpus the GPU produces for the same seeds (the
oracle's Philox walks are token-for-token the GPU's, tests/test_walk_gpu.py).

    python scripts/make_fullsize_fixture.py [tract24|ca] [n_runs] [threads]

Runs the 8-thread Hogwild oracle (workers(8), DeepWalk.java:75) `n_runs` times with seeds 1..n_runs (thread
interleaving differs run to run as well), then once with hierarchical softmax switched on (DL4J's default objective,
SURVEY F9), and records pairwise nDCG@k (tract; python/embeddingEvaluation_tract.py:285-367) or the 10-fold CV
accuracy (CA; python/binaryClassification_CA.py:33-58) of every run.  CPU only; ~5 min per run on 8 cores at tract24.
The GPU parity tests (tests/test_full_size_gpu.py) compare against the mean with tolerance 2 x (max - min).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from embedding_b200 import evaluation as ev, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

WALK_SEED = 2013
KS = (5, 10, 20, 50)


def oracle_corpus(w, seed=WALK_SEED, n_flow=None, n_spatial=None):
    """Relabelled host corpus [n, L] of the workload, from the oracle's walks (== libdge's for the same seed)."""
    f, sp, L = w["flow"], w["spatial"], w["L"]
    g = O.Graph(f["nv"], f["src"], f["dst"], f["w"], f["sources"], alias_mode=O.ALIAS_FAST)
    tok = g.walk(n_flow or f["n_walks"], L, seed)
    if f["id_map"] is not None:
        tok = np.where(tok >= 0, f["id_map"][np.maximum(tok, 0)], -1).astype(np.int32)
    parts = [tok]
    if sp is not None:
        s = O.Graph(sp["nv"], sp["src"], sp["dst"], sp["w"], sp["sources"], out_degree=sp["out_degree"],
                    source_weight_sum=sp["sws"], alias_mode=O.ALIAS_FAST)
        ts = s.walk(n_spatial or sp["n_walks"], L, seed + 1)
        pos = np.arange(L, dtype=np.int64)[None, :] * w["n_regions"]
        ts = np.where(ts >= 0, sp["id_map"][np.maximum(ts, 0)] + pos, -1).astype(np.int32)
        parts.append(ts)
    return np.concatenate(parts)


def layers_of(w, syn0, ids):
    n = w["n_regions"]
    idx = np.arange(w["n_ids"])
    return ev.layers_from_model(syn0, ids, (idx // n).astype(np.int32), np.asarray(w["region_ids"])[idx % n])


def metric_of(w, layers):
    if w["name"] == "ca":
        with open(os.path.join(ROOT, "tests", "golden", "ca_labels.json")) as f:
            d = json.load(f)
        labels = {"crime": d["crime-label"], "lehd": d["lehd-label"]}
        labels.update(d["demo-label"])
        labels.update(d["poi-label"])
        return ev.ca_classification_accuracy(layers, labels, w["region_ids"])
    gt = ev.PairwiseGroundTruth(synth.tract_ids(), synth.poi_latents())
    return {str(k): v for k, v in ev.pairwise_ndcg(gt, layers, ks=KS).items()}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "tract24"
    n_runs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    w = bench.make_workload(name)
    t = time.time()
    dry = int(os.environ.get("FIXTURE_DRY_RUN_WALKS", "0"))      # smoke test of this script only
    tok = oracle_corpus(w, n_flow=dry or None, n_spatial=(dry // 20) or None)
    print("corpus %s in %.1f s" % (tok.shape, time.time() - t), flush=True)
    kw = dict(dim=w["dim"], window=w["window"], negative=w["negative"], min_count=2, threads=threads)
    runs = []
    base = None
    for i in range(n_runs + 1):
        hs = i == n_runs
        seed = 1 if hs else i + 1
        t = time.time()
        m = O.sgns_train(tok, w["n_ids"], O.sgns_params(seed=seed, use_hs=int(hs), **kw))
        sec = time.time() - t
        layers = layers_of(w, m["syn0"], m["id_of_word"])
        table = ev.knn_table(layers, w["region_ids"], w["L"], 10)
        if base is None:
            base = table
            if not dry:   # run 0's neighbourhood fingerprint: the GPU embedding must agree with it as well as runs 1.. do
                np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fullsize_%s_oracle_knn.npz" % name), knn=table)
        r = dict(objective="hs+ns" if hs else "ns", seed=seed, threads=threads, pairs=int(m["pairs"]), seconds=sec,
                 metric=metric_of(w, layers), knn_overlap_vs_run0=ev.knn_table_overlap(base, table),
                 mean_row_norm=float(np.linalg.norm(m["syn0"], axis=1).mean()))
        runs.append(r)
        print(json.dumps(r), flush=True)
        save(w, name, tok, runs, dry)          # after every run: a partial fixture is usable
    print("wrote", save(w, name, tok, runs, dry))


def save(w, name, tok, runs, dry):
    ns = [r for r in runs if r["objective"] == "ns"]
    keys = list(ns[0]["metric"].keys())
    summary = {k: dict(mean=float(np.mean([r["metric"][k] for r in ns])), min=min(r["metric"][k] for r in ns),
                       max=max(r["metric"][k] for r in ns)) for k in keys}
    out = dict(workload=w["desc"], walk_seed=WALK_SEED, n_sentences=int(tok.shape[0]), made_by="scripts/make_fullsize_fixture.py",
               note="oracle = oracle/sgns_oracle.c (word2vec skip-gram, DL4J 0.7.2 parameterisation; parity unpinned against DL4J itself)",
               runs=runs, summary=summary)
    path = os.path.join(ROOT, "tests", "golden", "fullsize_%s_oracle.json" % name)
    if dry:
        path = "/tmp/fixture_dry_%s.json" % name
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    return path


if __name__ == "__main__":
    main()
