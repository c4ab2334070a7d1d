"""The reference's ONLY published performance experiment, reproduced: wall time of sampling {0.5, 1, 2, 5, 10} M walk
sequences INCLUDING writing them to the `.seq` file (CrossTimeGraph.main, CrossTimeGraph.java:152-159; numbers hard-coded in
python/running_time.py:16-20: CA / tract graph, alias method vs "random interval" = the CDF scan of
LayeredGraph.sampleVertexSequence_OV :260-279).

    python bench.py --workload running_time            (or: python scripts/running_time.py)

For every (graph, sampler, n): libdge on the GPU (dge_walk + dge_corpus_write_seq to a file on the box's local disk, graph
already built -- as in the Java experiment, which samples from a constructed graph) beside the CPU port (oracle/, one
thread, java.util.Random LCG, the text formatted and written by a single-threaded C loop standing in for String.join +
BufferedWriter; measured on a bounded sample and scaled linearly) and the authors' published seconds (different, unnamed 2017 hardware).  Prints one JSON line; writes gpurun_out/running_time.json.
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SIZES = [500_000, 1_000_000, 2_000_000, 5_000_000, 10_000_000]            # CrossTimeGraph.java:153
PUBLISHED = {                                                                # python/running_time.py:17-20 (seconds)
    ("ca", "alias"): [1.038, 2.001, 3.959, 9.739, 19.462],
    ("ca", "cdf"): [1.291, 2.423, 4.635, 11.496, 23.44],
    ("tract", "alias"): [1.815, 3.617, 7.161, 18.003, 35.766],
    ("tract", "cdf"): [3.467, 6.685, 13.76, 33.652, 65.752],
}


def cpu_seconds_per_walk(O, w, sampler, n_sample, path):
    """CPU port: one thread samples n_sample walks (java.util.Random LCG) and writes them as text; seconds per walk."""
    f = w["flow"]
    g = O.Graph(f["nv"], f["src"], f["dst"], f["w"], f["sources"], alias_mode=O.ALIAS_FAST)
    t0 = time.perf_counter()
    tok = g.walk(n_sample, w["L"], seed=7, sampler=O.SAMPLER_ALIAS if sampler == "alias" else O.SAMPLER_CDF, rng=O.RNG_JAVA_LCG)
    t1 = time.perf_counter()
    O.write_seq(tok, f["v_layer"], f["v_region"], path)                     # String.join(" ", seq) + "\n" per walk, one thread, in C
    t2 = time.perf_counter()
    return (t1 - t0) / n_sample, (t2 - t1) / n_sample


def main(args=None):
    import bench
    from embedding_b200 import abi
    from oracle import oracle as O
    ctx = abi.Context(int(os.environ.get("LOCAL_RANK", "0")))
    rows = []
    tmpdir = tempfile.mkdtemp(prefix="dge_running_time_")
    for level, wname in (("ca", "ca"), ("tract", "tract24")):
        w = bench.make_workload(wname)
        f, L = w["flow"], w["L"]                                             # numLayer = 24 (CrossTimeGraph.java:19 default used by main)
        G = abi.Graph(ctx, f["nv"], f["src"], f["dst"], f["w"], f["sources"])
        for sampler in ("alias", "cdf"):
            smp = abi.SAMPLER_ALIAS if sampler == "alias" else abi.SAMPLER_CDF
            cpu_walk, cpu_text = cpu_seconds_per_walk(O, w, sampler, 200_000, os.path.join(tmpdir, "cpu.seq"))
            for n, pub in zip(SIZES, PUBLISHED[(level, sampler)]):
                path = os.path.join(tmpdir, "gpu.seq")
                best = None
                for rep in range(2):
                    t0 = time.perf_counter()
                    c = G.walk(n, L, seed=11 + rep, sampler=smp)
                    t1 = time.perf_counter()
                    c.write_seq(path, f["v_region"], f["v_layer"])
                    t2 = time.perf_counter()
                    kms, fmt_ms = ctx.phase_ms("walk"), ctx.phase_ms("seq_format")
                    size = os.path.getsize(path)
                    c.free()
                    if best is None or t2 - t0 < best["gpu_total_s"]:
                        best = dict(gpu_total_s=t2 - t0, gpu_walk_s=t1 - t0, gpu_walk_kernel_ms=kms, gpu_seq_write_s=t2 - t1, gpu_seq_format_and_copy_ms=fmt_ms, seq_bytes=size)
                r = dict(graph=level, sampler=sampler, n_walks=n, published_java_s=pub, cpu_port_s=(cpu_walk + cpu_text) * n, cpu_port_walk_only_s=cpu_walk * n, **best)
                r["speedup_vs_published"] = pub / r["gpu_total_s"]
                r["speedup_vs_cpu_port"] = r["cpu_port_s"] / r["gpu_total_s"]
                rows.append(r)
                print(json.dumps(r), file=sys.stderr, flush=True)
        G.free()
    out = dict(metric="seconds to sample n walk sequences and write the .seq file", unit="s", higher_is_better=False,
               source="CrossTimeGraph.java:152-159; published numbers python/running_time.py:16-20 (authors' 2017 workstation, one JVM thread)",
               note="GPU = dge_walk + dge_corpus_write_seq (text formatted on the device) to local disk, best of 2; CPU port = oracle walk (1 thread, java.util.Random LCG) + String.join-style text write, measured on 200 000 walks and scaled linearly, same host",
               rows=rows)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "running_time.json"), "w"), indent=1)
    print(json.dumps(out), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
