"""Multi-GPU plumbing for one rank per GPU (bench.py under torchrun): independent queries shard across ranks with no
data-path collective.  Each rank holds a replica of the models and solves every world-th entry
of the batch's cost-sorted claim order -- the same split the library's multi-device entry
(c2a_b200_solve_batch_multi) makes over the devices of one process -- so that the queries expected to run long are
spread over the GPUs; the only exchange is the final gather of per-query results, scattered back to batch order."""
import numpy as np


def shard_indices(order, rank, world):
    """Queries of rank ``rank``: order[rank], order[rank + world], ...  (order = api.schedule_order, or arange)."""
    return np.ascontiguousarray(np.asarray(order)[int(rank)::int(world)])


def shard_bounds(n, rank, world):
    """Contiguous, balanced split of n queries: the first n % world ranks get one extra."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_results(local, idx, n_total, rank, world, dist=None, dst=0):
    """Gather per-rank result dicts (field -> ndarray over the rank's queries ``idx``) on rank ``dst`` and scatter them
    to batch order.  Returns the full dict on ``dst`` and None elsewhere.  ``dist`` is torch.distributed (any
    backend); with world == 1 nothing is exchanged."""
    parts = [(np.asarray(idx), local)]
    if world > 1:
        got = [None] * world if rank == dst else None
        dist.gather_object((np.asarray(idx), local), got, dst=dst)
        if rank != dst:
            return None
        parts = got
    out = {}
    seen = np.zeros(n_total, dtype=bool)
    for ix, res in parts:
        assert not seen[ix].any()
        seen[ix] = True
        for k, v in res.items():
            if k not in out:
                out[k] = np.zeros((n_total,) + v.shape[1:], dtype=v.dtype)
            out[k][ix] = v
    assert seen.all()
    return out
