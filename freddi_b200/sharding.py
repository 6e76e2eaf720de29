"""Multi-GPU plumbing: models are independent, so the ensemble is sharded by interleaved model index and nothing is
exchanged on the data path; torch.distributed only carries the barriers and the reduction of timing scalars."""
import numpy as np


def shard_indices(n_models, rank, world_size):
    """Model indices of one rank: k -> rank k mod world_size (balances cost correlated with the sweep parameters)."""
    return np.arange(rank, n_models, world_size)


def gather_rows(dist, local_rows, n_models, rank, world_size):
    """Rank 0 reassembles the (n_models, n_rows, n_cols) array from the per-rank shards (host-side gather)."""
    if dist is None or world_size == 1:
        return local_rows
    parts = [None] * world_size if rank == 0 else None
    dist.gather_object(local_rows, parts, dst=0)
    if rank != 0:
        return None
    out = np.full((n_models,) + local_rows.shape[1:], np.nan)
    for r, part in enumerate(parts):
        out[shard_indices(n_models, r, world_size)] = part
    return out


def reduce_timing(dist, torch, seconds, work, device=None):
    """(max over ranks of the time scalars, sum over ranks of the work counters)."""
    t = torch.tensor(list(seconds), dtype=torch.float64, device=device)
    w = torch.tensor(list(work), dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return t.tolist(), w.tolist()
