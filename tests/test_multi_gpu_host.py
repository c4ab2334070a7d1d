"""N>1 host logic on CPU: two gloo processes shard walk ids, each draws its shard (CPU oracle stands in for the
GPU kernel -- same counter-based stream), and the gathered union equals the single-process result."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_walks, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from embedding_b200 import parallel, synth
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = synth.powerlaw_flow_graph(40, L=6, seed=3, mean_degree=5, cap=16)
    G = O.Graph(g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"])
    first, count = parallel.walk_shard(n_walks, rank, world)
    tok = G.walk(count, 6, seed=9, first_walk_id=first)
    # gather shard sizes, then tokens
    sizes = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([first, count]))
    steps = torch.tensor([int((tok >= 0).sum())])
    dist.all_reduce(steps)                       # whole-job step count, as bench.py sums over ranks
    np.save(os.path.join(out_dir, "tok%d.npy" % rank), tok)
    if rank == 0:
        np.save(os.path.join(out_dir, "sizes.npy"), torch.stack(sizes).numpy())
        np.save(os.path.join(out_dir, "steps.npy"), steps.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_walk_shards_union_equals_single_run(oracle, tmp_path):
    from embedding_b200 import parallel, synth
    n_walks, world = 1001, 2
    mp.spawn(_worker, args=(world, _free_port(), n_walks, str(tmp_path)), nprocs=world, join=True)
    sizes = np.load(tmp_path / "sizes.npy")
    assert sizes[0, 0] == 0 and sizes[:, 1].sum() == n_walks and sizes[1, 0] == sizes[0, 1]
    g = synth.powerlaw_flow_graph(40, L=6, seed=3, mean_degree=5, cap=16)
    G = oracle.Graph(g["n_vertices"], g["src"], g["dst"], g["w"], g["sources"])
    whole = G.walk(n_walks, 6, seed=9)
    parts = np.concatenate([np.load(tmp_path / ("tok%d.npy" % r)) for r in range(world)])
    assert np.array_equal(whole, parts)
    assert int(np.load(tmp_path / "steps.npy")[0]) == int((whole >= 0).sum())


def _bcast_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from embedding_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    payload = bytes(range(128)) if rank == 0 else b""
    got = parallel.broadcast_bytes(dist, payload, 128)
    with open(os.path.join(out_dir, "id%d.bin" % rank), "wb") as f:
        f.write(got)
    dist.barrier()
    dist.destroy_process_group()


def test_comm_id_broadcast_reaches_every_rank(tmp_path):
    """The host's only job for dge_comm_init: carry the opaque 128-byte id from rank 0 to all ranks."""
    mp.spawn(_bcast_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / ("id%d.bin" % r), "rb").read() == bytes(range(128))


def test_shard_arithmetic():
    from embedding_b200 import parallel
    for n in (0, 1, 7, 8, 15_600_000):
        for world in (1, 2, 3, 8):
            spans = [parallel.walk_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert parallel.weak_shard(100, 3) == (300, 100)
