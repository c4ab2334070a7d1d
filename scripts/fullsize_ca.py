"""BASELINE configs[0] at its full size (77 community areas x 24: 8,000,000 flow + 80,000 spatial walks, D=8, window=24, K=5)
for a given skip-gram schedule: throughput, the reference's CA-level metric (10-fold CV accuracy, python/binaryClassification_CA.py)
and the kNN agreement with oracle run 0 of tests/golden/fullsize_ca_oracle.json (the same checks as
tests/test_full_size_gpu.py::test_full_size_ca_metric_matches_the_oracle).

    python scripts/fullsize_ca.py [conc,conc,...] [flags[,flags...]] [tag]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from embedding_b200 import abi, evaluation as ev  # noqa: E402


def main():
    concs = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0").split(",")]
    flag_sets = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")]
    tag = sys.argv[3] if len(sys.argv) > 3 else "f%d" % flag_sets[0]
    G = os.path.join(ROOT, "tests", "golden")
    fx = json.load(open(os.path.join(G, "fullsize_ca_oracle.json")))
    ref = np.load(os.path.join(G, "fullsize_ca_oracle_knn.npz"))["knn"]
    d = json.load(open(os.path.join(G, "ca_labels.json")))
    labels = {"crime": d["crime-label"], "lehd": d["lehd-label"]}
    labels.update(d["demo-label"])
    labels.update(d["poi-label"])
    w = bench.make_workload("ca")
    f, sp, L = w["flow"], w["spatial"], w["L"]
    ctx = abi.Context(0)
    Gf = abi.Graph(ctx, f["nv"], f["src"], f["dst"], f["w"], f["sources"])
    S = abi.Graph(ctx, sp["nv"], sp["src"], sp["dst"], sp["w"], sp["sources"], out_degree=sp["out_degree"], source_weight_sum=sp["sws"])
    c1, c2 = Gf.walk(f["n_walks"], L, seed=2013), S.walk(sp["n_walks"], L, seed=2014)
    c1.relabel(f["id_map"], w["n_ids"], 0)
    c2.relabel(sp["id_map"], w["n_ids"], w["n_regions"])
    n = w["n_regions"]
    idx = np.arange(w["n_ids"])
    vl, vr = (idx // n).astype(np.int32), np.asarray(w["region_ids"])[idx % n]
    ns = [r for r in fx["runs"] if r["objective"] == "ns"]
    print("oracle:", {k: {a: round(b, 4) for a, b in v.items()} for k, v in fx["summary"].items()},
          "kNN agreement of oracle runs with run 0:", [round(r["knn_overlap_vs_run0"], 4) for r in ns[1:]], flush=True)
    out = []
    for flags, conc in [(f_, c_) for f_ in flag_sets for c_ in concs]:
        m = abi.Model.train(ctx, [c1, c2], abi.sgns_params(dim=w["dim"], window=w["window"], negative=5, min_count=2, seed=1, concurrency=conc, flags=flags))
        ms = ctx.phase_ms("sgns")
        syn0, ids = m.vectors()
        layers = ev.layers_from_model(syn0, ids, vl, vr)
        acc = ev.ca_classification_accuracy(layers, labels, w["region_ids"])
        ov = ev.knn_table_overlap(ref, ev.knn_table(layers, w["region_ids"], L, 10))
        r = dict(concurrency=conc, flags=flags, groups=ctx.phase_ms("sgns_groups"), kernel=int(ctx.phase_ms("sgns_kernel")),
                 write_through_words=int(ctx.phase_ms("sgns_write_through")), sgns_ms=round(ms, 1), gpairs_per_s=round(m.pairs / ms / 1e6, 3),
                 accuracy={k: round(v, 4) for k, v in acc.items()}, knn_agreement_with_oracle_run0=round(ov, 4),
                 mean_row_norm=round(float(np.linalg.norm(syn0, axis=1).mean()), 3))
        out.append(r)
        print(json.dumps(r), flush=True)
        m.free()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fullsize_ca_%s.json" % tag), "w"), indent=1)


if __name__ == "__main__":
    main()
