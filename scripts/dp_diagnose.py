"""Where does the data-parallel skip-gram lose neighbourhood agreement?  (gpurun --gpus 2)
kNN overlap (k = 10) between: the single-thread oracle, single-GPU runs under several schedules, and 2-rank
data-parallel runs with 6 / 24 / 96 delta exchanges, all on the corpus of tests/test_comm_gpu.py."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embedding_b200 import abi, evaluation as ev, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only)


def main():
    nv = 300 * 8
    zeros, ar = np.zeros(nv, np.int32), np.arange(nv, dtype=np.int32)
    runs = {}
    worker = os.path.join(ROOT, "tests", "helpers", "dp_worker.py")
    whole = None
    for rounds in (6, 24, 96):
        d = tempfile.mkdtemp()
        procs = [subprocess.Popen([sys.executable, worker, str(r), "2", d, str(rounds)]) for r in range(2)]
        assert all(p.wait(timeout=600) == 0 for p in procs)
        r0, r1 = np.load(os.path.join(d, "rank0.npz")), np.load(os.path.join(d, "rank1.npz"))
        whole = np.concatenate([r0["tok"], r1["tok"]])
        runs["dp2_rounds%d" % rounds] = ev.layers_from_model(r0["syn0"], r0["ids"], zeros, ar)
    kw = dict(dim=32, window=5, negative=5, min_count=2)
    for name, th, seed in (("oracle_t1_s3", 1, 3), ("oracle_t1_s4", 1, 4), ("oracle_t8_s3", 8, 3)):
        m = O.sgns_train(whole, nv, O.sgns_params(seed=seed, threads=th, **kw))
        runs[name] = ev.layers_from_model(m["syn0"], m["id_of_word"], zeros, ar)
    ctx = abi.Context(0)
    c = abi.Corpus.from_tokens(ctx, whole, nv)
    for name, extra in (("gpu1_auto_s3", dict(seed=3)), ("gpu1_auto_s4", dict(seed=4)), ("gpu1_c1_s3", dict(seed=3, concurrency=1)),
                        ("gpu1_c16_s3", dict(seed=3, concurrency=16)), ("gpu1_rounds6_s3", dict(seed=3, sync_rounds=6))):
        m = abi.Model.train(ctx, [c], abi.sgns_params(**kw, **extra))
        s0, ids = m.vectors()
        runs[name] = ev.layers_from_model(s0, ids, zeros, ar)
        print(name, "groups", ctx.phase_ms("sgns_groups"), "kernel", ctx.phase_ms("sgns_kernel"), flush=True)
    names = list(runs)
    out = {}
    for i, a in enumerate(names):
        for b in names[i + 1:]:
            out["%s|%s" % (a, b)] = round(float(ev.knn_overlap(runs[a], runs[b], 10)), 3)
    for a in names:
        print("%-18s" % a, " ".join("%5.2f" % (1.0 if a == b else out.get("%s|%s" % (a, b), out.get("%s|%s" % (b, a)))) for b in names))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(names=names, knn_overlap=out), open(os.path.join(ROOT, "gpurun_out", "dp_diagnose.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
