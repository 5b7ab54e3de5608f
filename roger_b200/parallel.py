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


def bind_to_gpu_numa_node(device_index):
    """Pins this process to the host cores of the NUMA node its GPU hangs off, BEFORE it allocates page-locked buffers: pinned
    memory is then first-touched on that node, so the eight ranks of a box do not push their output slabs through one socket's
    memory controllers (round 1: end-to-end scaled 3.96x from 1 to 8 GPUs, the device-timed kernels 7.97x).  Returns a short
    description for the bench record; never raises (a missing sysfs entry just leaves the affinity as it is)."""
    import os

    try:
        import torch

        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        addr = "%04x:%02x:%02x.0" % (dom, bus, dev)
        node = int(open(f"/sys/bus/pci/devices/{addr}/numa_node").read().strip())
        if node < 0:
            return f"gpu {addr}: no NUMA node reported"
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = os.sched_getaffinity(0)
        use = sorted(allowed & set(cpus))
        if not use:
            return f"gpu {addr}: node {node} has no allowed cpu"
        os.sched_setaffinity(0, use)
        return f"gpu {addr}: node {node}, {len(use)} cpus"
    except Exception as e:  # pragma: no cover - depends on the box
        return f"not bound ({type(e).__name__}: {e})"


def _parse_cpulist(text):
    out = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.extend(range(int(lo), int(hi or lo) + 1))
    return out
