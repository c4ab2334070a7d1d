"""BASELINE.json configs[4]: SGNS sweep dim 16-256 x negatives 5-20 x window 5-10 on the 100K-region synthetic corpus.

    python scripts/sgns_sweep.py [--walks 2000000] [--regions 100000] [--out gpurun_out/sgns_sweep.json]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/sgns_sweep.py   (data-parallel)

Per cell: pairs/s (device time of dge_sgns_train's "sgns" phase, max over ranks; pairs summed over ranks) and the
fraction of the measured HBM peak at 8*D*(K+2) algorithmic bytes per pair.  One corpus per rank
(walk ids sharded by rank, weak scaling), reused by every cell.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from embedding_b200 import abi, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--walks", type=int, default=2_000_000)
    ap.add_argument("--regions", type=int, default=100_000)
    ap.add_argument("--dims", default="16,32,64,128,256")
    ap.add_argument("--negatives", default="5,10,20")
    ap.add_argument("--windows", default="5,10")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sgns_sweep.json"))
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
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    L = 24
    g = synth.powerlaw_flow_graph(a.regions, L=L, seed=100000)
    G = abi.Graph(ctx, g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"])
    corpus = G.walk(a.walks, L, seed=7, first_walk_id=rank * a.walks)
    cells = []
    for dim in [int(x) for x in a.dims.split(",")]:
        for neg in [int(x) for x in a.negatives.split(",")]:
            for win in [int(x) for x in a.windows.split(",")]:
                best = None
                for rep in range(2):                                   # second repetition = warm
                    m = abi.Model.train(ctx, [corpus], abi.sgns_params(dim=dim, window=win, negative=neg, min_count=2, seed=1 + rep))
                    ms, pairs = ctx.phase_ms("sgns"), m.pairs
                    m.free()
                    if dist is not None:
                        t = torch.tensor([ms, 0.0], dtype=torch.float64, device="cuda")
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                        p = torch.tensor([float(pairs)], dtype=torch.float64, device="cuda")
                        dist.all_reduce(p, op=dist.ReduceOp.SUM)
                        ms, pairs = float(t[0].item()), float(p.item())
                    best = (ms, pairs)
                ms, pairs = best
                bpp = 8.0 * dim * (neg + 2)
                rate = pairs / (ms * 1e-3)
                cell = dict(dim=dim, negative=neg, window=win, n_gpus=world, pairs=pairs, sgns_ms=ms, pairs_per_s=rate,
                            bytes_per_pair=bpp, algorithmic_gbs_per_gpu=rate * bpp / 1e9 / world,
                            frac_of_hbm_peak=rate * bpp / 1e9 / world / peak,
                            sync_rounds=ctx.phase_ms("sgns_rounds"), sync_ms=ctx.phase_ms("sgns_sync"))
                cells.append(cell)
                if rank == 0:
                    print("D=%3d K=%2d w=%2d  %8.1f M pairs/s  %6.1f ms  %5.1f%% of HBM peak/GPU  sync %.1f ms"
                          % (dim, neg, win, rate / 1e6, ms, 100 * cell["frac_of_hbm_peak"], cell["sync_ms"]), flush=True)
    if rank == 0:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(dict(workload="synth %d regions x %d slices, %d walks x %d per GPU, %d GPU(s)" % (a.regions, L, a.walks, L, world),
                       hbm_peak_gbs=peak, cells=cells), open(a.out, "w"), indent=1)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
