"""Host-side multi-rank logic on CPU: world_size 2 over gloo (the GPU path uses the same code over NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from plen_ml_walk_b200.sharding import env_seed, max_over_ranks, shard_range, sum_over_ranks


def test_shard_range_partitions_every_env_once():
    for n, w in ((1048576, 8), (65536, 4), (10, 3), (7, 8), (0, 2)):
        blocks = [shard_range(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)
    assert len({env_seed(0, r) for r in range(8)}) == 8


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(4097, rank, world)
    owned = torch.zeros(4097, dtype=torch.int64)
    owned[lo:hi] = 1
    dist.all_reduce(owned)                                  # every env owned exactly once across ranks
    t_max = max_over_ranks(10.0 + rank, dist)               # bench.py: step time = slowest rank
    total = sum_over_ranks(hi - lo, dist)
    dist.barrier()
    if rank == 0:
        torch.save({"owned_ok": bool((owned == 1).all()), "t_max": t_max, "total": total}, out)
    dist.destroy_process_group()


def test_two_ranks_over_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["owned_ok"] and r["t_max"] == 11.0 and r["total"] == 4097.0


def test_bind_host_to_gpu_never_raises_and_reports():
    """One process per GPU binds itself to the CPUs local to its GPU (NVML); without NVML / a GPU the call must leave the
    affinity alone and say so."""
    import os
    from plen_ml_walk_b200.sharding import bind_host_to_gpu
    before = os.sched_getaffinity(0)
    msg = bind_host_to_gpu(0)
    assert isinstance(msg, str) and (msg.startswith("unchanged") or msg.startswith("bound to"))
    if msg.startswith("unchanged"):
        assert os.sched_getaffinity(0) == before
    else:
        assert os.sched_getaffinity(0) <= before
        os.sched_setaffinity(0, before)
