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
