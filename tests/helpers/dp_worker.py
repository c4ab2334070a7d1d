"""One rank of the 2-GPU data-parallel skip-gram test (tests/test_comm_gpu.py): `python dp_worker.py rank world dir [rounds]`.
The NCCL id travels through a file, as a JNI host without torch would do it."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from embedding_b200 import abi, parallel, synth  # noqa: E402


def main():
    rank, world, d = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 6
    transport = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    ctx = abi.Context(rank)
    idf = os.path.join(d, "nccl_id.bin")
    if rank == 0:
        uid = abi.Context.comm_unique_id()
        with open(idf + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(idf + ".tmp", idf)
    else:
        t0 = time.time()
        while not os.path.exists(idf):
            if time.time() - t0 > 120:
                raise SystemExit("rank 0 never wrote the NCCL id")
            time.sleep(0.05)
        uid = open(idf, "rb").read()
    ctx.comm_init(rank, world, uid)
    assert ctx.comm_shape() == (rank, world)
    g = synth.powerlaw_flow_graph(300, L=8, seed=5, mean_degree=8, cap=64)
    G = abi.Graph(ctx, g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"])
    n_walks = 40_000
    first, count = parallel.walk_shard(n_walks, rank, world)
    corpus = G.walk(count, 8, seed=11, first_walk_id=first)          # this rank's shard of the walk ids
    m = abi.Model.train(ctx, [corpus], abi.sgns_params(dim=32, window=5, negative=5, min_count=2, seed=3, sync_rounds=rounds, transport=transport))
    syn0, syn1, ids = m.vectors(want_syn1neg=True)
    np.savez(os.path.join(d, "rank%d.npz" % rank), syn0=syn0, syn1=syn1, ids=ids, pairs=m.pairs,
             rounds=ctx.phase_ms("sgns_rounds"), sync_ms=ctx.phase_ms("sgns_sync"), transport=ctx.phase_ms("sgns_transport"),
             tok=corpus.tokens())
    ctx.close()


if __name__ == "__main__":
    main()
