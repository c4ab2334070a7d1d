"""Where the end-to-end (host buffers through the C ABI) time of one bench step goes: wall clock per ABI call next to
the device phases each call reports.   python scripts/e2e_breakdown.py [tract24|tract8|ca|synth100k] [reps]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from embedding_b200 import abi  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "tract24"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    w = bench.make_workload(name)
    f, L = w["flow"], w["L"]
    ctx = abi.Context(0)
    pin = abi.PinnedArray((f["n_walks"], L), np.int32)
    params = abi.sgns_params(dim=w["dim"], window=w["window"], negative=w["negative"], min_count=2, seed=1)
    rows = []
    for rep in range(reps):
        r = {}
        t0 = time.perf_counter()
        G = abi.Graph(ctx, f["nv"], f["src"], f["dst"], f["w"], f["sources"])
        t1 = time.perf_counter()
        r["graph_build_wall_ms"] = (t1 - t0) * 1e3
        for ph in ("coo_h2d", "csr", "alias", "pack"):
            r["graph_" + ph + "_dev_ms"] = ctx.phase_ms(ph)
        c = G.walk(f["n_walks"], L, seed=10 + rep)
        n_steps = c.count_tokens()
        t2 = time.perf_counter()
        r["walk_wall_ms"] = (t2 - t1) * 1e3
        r["walk_dev_ms"] = ctx.phase_ms("walk")
        c.tokens(pin.array)
        t3 = time.perf_counter()
        r["tokens_d2h_wall_ms"] = (t3 - t2) * 1e3
        r["tokens_d2h_GBps"] = pin.array.nbytes / (t3 - t2) / 1e9
        c.free()
        t4 = time.perf_counter()
        h = abi.Corpus.from_tokens(ctx, pin.array, f["nv"])
        t5 = time.perf_counter()
        r["tokens_h2d_wall_ms"] = (t5 - t4) * 1e3
        r["tokens_h2d_GBps"] = pin.array.nbytes / (t5 - t4) / 1e9
        m = abi.Model.train(ctx, [h], params)
        t6 = time.perf_counter()
        r["sgns_train_wall_ms"] = (t6 - t5) * 1e3
        for ph in ("vocab", "compact", "sgns"):
            r["sgns_" + ph + "_dev_ms"] = ctx.phase_ms(ph)
        syn0, ids = m.vectors()
        t7 = time.perf_counter()
        r["vectors_d2h_wall_ms"] = (t7 - t6) * 1e3
        r["steps"], r["pairs"] = n_steps, m.pairs
        r["walk_e2e_steps_per_s"] = n_steps / (t3 - t0)
        r["sgns_e2e_pairs_per_s"] = m.pairs / (t7 - t4)
        m.free()
        h.free()
        G.free()
        rows.append(r)
    out = dict(workload=w["desc"], reps=rows[1:], note="first repetition (allocator warm-up) dropped")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
