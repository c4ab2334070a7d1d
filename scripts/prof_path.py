"""One short pass of the hot path for profiling under ncu (graph build -> walks -> SGNS).
    python scripts/prof_path.py tract 500000 [concurrency]
    python scripts/prof_path.py synth 20000 2000000 [dim]
    python scripts/prof_path.py ca 2000000 [concurrency] [dim]      (77 community areas x 24, D = 8)
A trailing argument flags=N passes dge_sgns_params.flags (DGE_SGNS_F_*: kernel selection for A/B captures).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embedding_b200 import abi, host, synth  # noqa: E402


def main():
    flags = 0
    if sys.argv[-1].startswith("flags="):
        flags = int(sys.argv.pop()[6:])
    level = sys.argv[1]
    ctx = abi.Context(0)
    if level == "synth":
        n_regions, n_walks = int(sys.argv[2]), int(sys.argv[3])
        dim = int(sys.argv[4]) if len(sys.argv) > 4 else 128
        g = synth.powerlaw_flow_graph(n_regions, L=24, seed=100000)
        G = abi.Graph(ctx, g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"])
        L, window, conc = 24, 10, 0
    else:
        n_walks = int(sys.argv[2])
        conc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
        if level == "ca":
            ids, z, L, dim = synth.ca_ids(), synth.ca_latents(), 24, (int(sys.argv[4]) if len(sys.argv) > 4 else 8)
            fl = host.Flows(ids, synth.planted_flow_tensor(z, mean_trips_per_pair_hour=2.0))
            host.CrossTimeGraph.numLayer = L
            gh = host.CrossTimeGraph.constructGraph_CA(fl, ctx=ctx)
        else:
            ids, z, L, dim = synth.tract_ids(), synth.poi_latents(), (24 if level == "tract24" else 8), 20
            fl = host.Flows(ids, synth.planted_flow_tensor(z))
            host.CrossTimeGraph.numLayer = L
            gh = host.CrossTimeGraph.constructGraph_tract(fl, ctx=ctx)
        gh.initiateAliasTables()
        G = gh._graph
        window = L
    for rep in range(2):
        corpus = G.walk(n_walks, L, seed=1 + rep)
        m = abi.Model.train(ctx, [corpus], abi.sgns_params(dim=dim, window=window, seed=1, concurrency=conc, flags=flags))
        print("walk ms", ctx.phase_ms("walk"), "steps", corpus.count_tokens(), "sgns ms", ctx.phase_ms("sgns"),
              "pairs", m.pairs, "groups", ctx.phase_ms("sgns_groups"), "kernel", ctx.phase_ms("sgns_kernel"),
              "Mpairs/s", m.pairs / ctx.phase_ms("sgns") / 1e3, "Msteps/s", corpus.count_tokens() / ctx.phase_ms("walk") / 1e3)


if __name__ == "__main__":
    main()
