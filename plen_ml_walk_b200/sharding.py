"""Multi-GPU plumbing of the PLEN path: envs are independent (no shared world state, static ground), so they shard by
env index across ranks with NO collective on the physics path.  torch.distributed is used only
for the barrier and for the max-over-ranks of the device time (bench.py) -- backend "nccl" on GPUs, "gloo" in the CPU
tests.
"""
from __future__ import annotations


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous block [lo, hi) of env indices owned by `rank`; blocks differ by at most one env."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_seed(seed: int, rank: int) -> int:
    """Action-stream seed of a rank: distinct per rank, reproducible for a fixed world size."""
    return int(seed) * 1000003 + int(rank)


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """MAX all-reduce of a host float (device time in ms); identity without a process group."""
    if dist is None or not dist.is_initialized():
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, dist=None, device=None) -> float:
    if dist is None or not dist.is_initialized():
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def bind_host_to_gpu(index: int) -> str:
    """One process per GPU: run this process (and therefore allocate its pinned host buffers, first touch) on the CPU cores
    NVML reports as local to GPU `index` -- the end-to-end call copies 24 MB per step and 131,072 robots through pinned
    memory, and eight ranks whose buffers sit on the other socket share one inter-socket link.  Returns what was done (for the
    bench line); any failure leaves the affinity alone."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "unchanged (NVML reports no usable local CPUs)"
        if cpus == os.sched_getaffinity(0):
            return "unchanged (all %d allowed CPUs are local to GPU %d)" % (len(cpus), index)
        os.sched_setaffinity(0, cpus)
        return "bound to the %d CPUs local to GPU %d" % (len(cpus), index)
    except Exception as e:      # no NVML, no permission, container cpuset ...: not fatal
        return "unchanged (%s: %s)" % (type(e).__name__, e)
