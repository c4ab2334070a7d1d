"""BASELINE.json configs[3] on ONE GPU: synthetic 1M regions x 24 slices (~1.0e9 weighted edges).

    python scripts/config4_1m.py [--regions 1000000] [--walks 32000000] [--sgns-walks 8000000] [--dim 128]

The graph (16 GB of COO; 24 GB of CSR tables + 32 GB of packed walk records on the device) fits one 180 GB B200, so
the single-GPU numbers of the config exist before the 8-GPU run: graph build (CSR + bit-exact alias tables incl. the
1M-entry source table), walks (truly HBM-resident: the record array is 250x the L2), skip-gram with 24M x 128 tables
(12.3 GB each).  Multi-GPU: torchrun the same script; walk ids shard by rank with no collective, stage 2 exchanges
embedding deltas over NCCL (dge_sgns_train with a communicator).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embedding_b200 import abi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--regions", type=int, default=1_000_000)
    ap.add_argument("--walks", type=int, default=32_000_000)
    ap.add_argument("--sgns-walks", type=int, default=8_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--window", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config4_1m.json"))
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        from embedding_b200 import parallel
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = abi.Context(local)
    if dist is not None:
        parallel.init_comm(ctx, dist, torch.device("cuda", local))
    L = 24
    res = dict(regions=a.regions, L=L, n_gpus=world)
    t0 = time.perf_counter()
    g = synth.powerlaw_flow_graph_layered(a.regions, L=L, seed=a.regions)
    res.update(host_generate_s=time.perf_counter() - t0, n_vertices=int(g["n_vertices"]), n_edges=int(len(g["src"])))
    if rank == 0:
        print("generated %d vertices, %d edges in %.1f s" % (res["n_vertices"], res["n_edges"], res["host_generate_s"]), flush=True)
    t0 = time.perf_counter()
    G = abi.Graph(ctx, g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"])
    res.update(graph_build_wall_s=time.perf_counter() - t0,
               phases_ms={k: ctx.phase_ms(k) for k in ("coo_h2d", "csr", "alias", "pack") if ctx.phase_ms(k) is not None})
    if rank == 0:
        print("graph build %.2f s wall, phases %s" % (res["graph_build_wall_s"], res["phases_ms"]), flush=True)
    del g
    walks = []
    for rep in range(3):
        c = G.walk(a.walks, L, seed=11 + rep, first_walk_id=rank * a.walks)
        ms, steps = ctx.phase_ms("walk"), c.count_tokens()
        walks.append(dict(ms=ms, steps=steps, steps_per_s=steps / ms * 1e3, algorithmic_gbs=steps * 28.0 / ms / 1e6))
        c.free()
        if rank == 0:
            print("walk rep %d: %.3f ms, %d steps, %.1f G steps/s, %.0f GB/s algorithmic" %
                  (rep, ms, steps, steps / ms / 1e6, steps * 28.0 / ms / 1e6), flush=True)
    res["walk"] = walks
    c = G.walk(a.sgns_walks, L, seed=99, first_walk_id=rank * a.sgns_walks)
    sg = []
    for rep in range(2):
        m = abi.Model.train(ctx, [c], abi.sgns_params(dim=a.dim, window=a.window, negative=5, min_count=2, seed=1 + rep))
        ms = ctx.phase_ms("sgns")
        sg.append(dict(ms=ms, pairs=m.pairs, V=m.V, pairs_per_s=m.pairs / ms * 1e3, algorithmic_gbs=m.pairs * 8.0 * a.dim * 7 / ms / 1e6,
                       sync_rounds=ctx.phase_ms("sgns_rounds"), sync_ms=ctx.phase_ms("sgns_sync")))
        sg[-1].update(m.stats())      # a diverging delta exchange shows up as exploding row norms
        if rank == 0:
            print("sgns rep %d: V=%d, %.1f ms, %d pairs, %.1f M pairs/s per GPU, sync %.1f ms in %s rounds, mean |syn0 row| %.3f, max |x| %.2f, non-finite %d" %
                  (rep, m.V, ms, m.pairs, m.pairs / ms / 1e3, sg[-1]["sync_ms"], sg[-1]["sync_rounds"], sg[-1]["mean_row_norm"],
                   sg[-1]["max_abs"], sg[-1]["nonfinite"]), flush=True)
        m.free()
    res["sgns"] = sg
    if rank == 0:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
