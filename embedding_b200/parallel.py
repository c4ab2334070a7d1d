"""Multi-GPU partitioning of the hot path: one process per GPU.

Stage 1 shards by WALK ID with no collective: walk i draws from Philox(seed, i), so the union of the shards is
bit-identical to a single-GPU run whatever the GPU count.  Stage 2 on the small-vocabulary configs (CA / tract) is
"replicas only"; the large synthetic configs exchange embedding deltas.
"""


def walk_shard(n_walks, rank, world):
    """Contiguous, balanced range of walk ids owned by `rank`: returns (first_walk_id, count)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_walks), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def weak_shard(n_walks_per_gpu, rank):
    """Weak scaling (bench.py): every rank samples the full per-GPU count, ids offset by rank."""
    return rank * int(n_walks_per_gpu), int(n_walks_per_gpu)


def broadcast_bytes(dist, payload, n, device=None, src=0):
    """Moves `n` opaque bytes from rank `src` to every rank over an initialised torch.distributed group (gloo: CPU
    tensor, nccl: a tensor on `device`).  This is the only thing the host has to do for dge_comm_init."""
    import torch
    t = torch.zeros(n, dtype=torch.uint8, device=device if device is not None else "cpu")
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def init_comm(ctx, dist, device=None):
    """Gives `ctx` a NCCL communicator spanning the torch.distributed group: rank 0 creates the id
    (dge_comm_unique_id), the group broadcasts it, every rank calls dge_comm_init."""
    from . import abi
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = abi.Context.comm_unique_id() if rank == 0 else b""
    uid = broadcast_bytes(dist, uid, abi.COMM_ID_BYTES, device)
    ctx.comm_init(rank, world, uid)
    return rank, world
