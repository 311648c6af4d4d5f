"""Multi-GPU plumbing: independent queries shard across ranks with no data-path collective
(SURVEY.md section 8e).  Each rank owns one GPU, holds a replica of every model and solves a
contiguous slice of the batch; the only exchange is the final gather of per-query results."""
import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous, balanced split of n queries: the first n % world ranks get one extra."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_results(local, n_total, rank, world, dist=None, dst=0):
    """Gather per-rank result dicts (field -> ndarray over the local slice) on rank ``dst``.
    Returns the full dict on ``dst`` and None elsewhere.  ``dist`` is torch.distributed (any backend);
    with world == 1 nothing is exchanged."""
    if world == 1:
        return local
    parts = [None] * world if rank == dst else None
    dist.gather_object(local, parts, dst=dst)
    if rank != dst:
        return None
    out = {}
    for k in parts[0]:
        out[k] = np.concatenate([p[k] for p in parts], 0)
        assert out[k].shape[0] == n_total
    return out
