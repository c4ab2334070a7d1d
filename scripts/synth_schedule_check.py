"""Does the automatic skip-gram schedule on the synthetic HBM-resident workload (100K regions x 24, D = 128: kernel F on wide
rows, sentence counter, write-through words, a full GPU of sentences in flight) learn the embedding a near-sequential run
learns?  No CPU oracle finishes at this size, so the yardstick is the library itself with FEW sentences in flight (strided
hand-out, nothing written through -- the schedule whose parity tests/test_sgns_gpu.py pins at small sizes), two seeds of it
giving the noise floor; metric = overlap of the 10 nearest cosine neighbours of 2 000 words among the 50 000 most frequent.

    python scripts/synth_schedule_check.py [n_walks] [reference sentences in flight]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from embedding_b200 import abi  # noqa: E402


def main():
    n_walks = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    ref_conc = int(sys.argv[2]) if len(sys.argv) > 2 else 148
    w = bench.make_workload("synth100k", 0)
    f, L, dim, neg = w["flow"], w["L"], w["dim"], w["negative"]
    ctx = abi.Context(0)
    G = abi.Graph(ctx, f["nv"], f["src"], f["dst"], f["w"], f["sources"])
    c = G.walk(n_walks, L, 777)
    kw = dict(dim=dim, window=w["window"], negative=neg, min_count=2)
    out = []

    def run(name, seed, **extra):
        m = abi.Model.train(ctx, [c], abi.sgns_params(seed=seed, **kw, **extra))
        ms = ctx.phase_ms("sgns")
        syn0, ids = m.vectors()
        st = m.stats()
        r = dict(name=name, seed=seed, sgns_ms=round(ms, 1), gpairs_per_s=round(m.pairs / ms / 1e6, 3), sentences_in_flight=ctx.phase_ms("sgns_groups"),
                 kernel=int(ctx.phase_ms("sgns_kernel")), write_through_words=int(ctx.phase_ms("sgns_write_through")), hub_bound=ctx.phase_ms("sgns_hub_bound"),
                 mean_row_norm=round(st["mean_row_norm"], 4) if isinstance(st, dict) else None, **extra)
        m.free()
        return r, syn0, ids

    r_ref, s_ref, ids_ref = run("reference: few sentences in flight, strided", 1, concurrency=ref_conc)
    r_ref2, s_ref2, ids2 = run("reference, another seed", 2, concurrency=ref_conc)
    assert np.array_equal(ids_ref, ids2)
    r_ref2["knn_overlap_vs_reference"] = bench.knn_overlap_sample(s_ref, s_ref2)
    out += [r_ref, r_ref2]
    print(json.dumps(r_ref), flush=True)
    print(json.dumps(r_ref2), flush=True)
    hub = int(r_ref["hub_bound"])
    for name, extra in [("automatic schedule", dict()), ("hub-bounded, strided (the schedule before this change)", dict(concurrency=max(1, min(hub, 2368)))),
                        ("full GPU, strided, nothing written through", dict(concurrency=2368))]:
        r, s0, ids = run(name, 1, **extra)
        assert np.array_equal(ids, ids_ref)
        r["knn_overlap_vs_reference"] = bench.knn_overlap_sample(s_ref, s0)
        out.append(r)
        print(json.dumps(r), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "synth_schedule_check.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
