"""Multi-GPU plumbing: contiguous cell-block partitioning and the one collective the path needs.

Columns are independent (no lateral routing in any SVAT / oneD configuration), so the grid is split
into contiguous blocks of the flattened interior cell index, one block per process / GPU; this replaces
the reference's 2-D MPI domain decomposition (roger/distributed.py:121-187).  The only data-path
exchange is the bitwise OR of the scheduler's flag words (grid-wide any()/all() of
adaptive_time_stepping.py:38-81 and the global_and of numerics.sanity_check), a few bytes per step.
"""
import numpy as np


def cell_block(n_cells, rank, world):
    """[lo, hi) of this rank's contiguous block; sizes differ by at most one cell."""
    base, rem = divmod(int(n_cells), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def make_flag_reducer(dist, device=None):
    """Returns f(words) -> words OR-ed over all ranks.  NCCL has no bitwise reduction, so every word is
    expanded to 32 0/1 lanes and reduced with MAX; works unchanged on gloo (CPU tests)."""
    import torch

    def reduce(words):
        w = np.asarray(words, dtype=np.uint32)
        bits = ((w[:, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1).astype(np.int32)
        t = torch.from_numpy(bits.reshape(-1).copy())
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out = t.cpu().numpy().reshape(len(w), 32).astype(np.uint32)
        return [int(v) for v in (out << np.arange(32, dtype=np.uint32)[None, :]).sum(axis=1)]

    return reduce
