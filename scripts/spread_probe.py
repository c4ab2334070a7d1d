"""Run-to-run spread of the skip-gram kernel time on tract x 24 (985 ms on most calls, 1 140 ms on some): does it follow where
the two tables land?  Trains the same corpus repeatedly, with and without a dummy allocation in between that shifts the pool.
    python scripts/spread_probe.py [reps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from embedding_b200 import abi  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    w = bench.make_workload("tract24")
    f, sp, L = w["flow"], w["spatial"], w["L"]
    ctx = abi.Context(0)
    Gf = abi.Graph(ctx, f["nv"], f["src"], f["dst"], f["w"], f["sources"])
    S = abi.Graph(ctx, sp["nv"], sp["src"], sp["dst"], sp["w"], sp["sources"], out_degree=sp["out_degree"], source_weight_sum=sp["sws"])
    c1, c2 = Gf.walk(f["n_walks"], L, seed=2013), S.walk(sp["n_walks"], L, seed=2014)
    c1.relabel(f["id_map"], w["n_ids"], 0)
    c2.relabel(sp["id_map"], w["n_ids"], w["n_regions"])
    out, keep = [], []
    for rep in range(reps):
        if rep >= reps // 2:   # second half: a small corpus stays allocated between the calls, so the pool hands out other blocks
            keep.append(abi.Corpus.from_tokens(ctx, np.zeros((1000 + 777 * rep, 3), np.int32), 10))
        m = abi.Model.train(ctx, [c1, c2], abi.sgns_params(dim=w["dim"], window=w["window"], negative=5, min_count=2, seed=1))
        r = dict(rep=rep, sgns_ms=round(ctx.phase_ms("sgns"), 1),
                 syn0=(int(ctx.phase_ms("sgns_syn0_addr_hi")), int(ctx.phase_ms("sgns_syn0_addr_lo"))),
                 syn1=(int(ctx.phase_ms("sgns_syn1_addr_hi")), int(ctx.phase_ms("sgns_syn1_addr_lo"))))
        r["placement"] = dict(k=int(ctx.phase_ms("sgns_placement")), best_us=round(ctx.phase_ms("sgns_placement_best_us"), 1), worst_us=round(ctx.phase_ms("sgns_placement_worst_us"), 1))
        r["delta_256B"] = ((r["syn1"][0] << 12) + r["syn1"][1]) - ((r["syn0"][0] << 12) + r["syn0"][1])
        out.append(r)
        print(json.dumps(r), flush=True)
        m.free()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "spread_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
