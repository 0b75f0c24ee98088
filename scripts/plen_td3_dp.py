#!/usr/bin/env python
"""Data-parallel batched plen_td3 (BASELINE config 5's optional collective): one process per GPU, envs sharded by index
with no communication on the physics path, every rank collecting into its own device replay ring; the only collective is
an NCCL all-reduce (mean) of the flat TD3 gradient vectors -- 155,138 critic + 77,330 actor floats -- between the CUDA
gradient kernels and the CUDA Adam step of every update (TD3Agent.train(grad_hook=...)).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/plen_td3_dp.py [--envs-per-gpu 16384] [--vector-steps 32] [--updates-per-step 8]

Rank 0 prints one JSON line; all ranks must end with bit-identical parameters (checked with an all-gather of checksums).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from plen_ml_walk_b200.sharding import env_seed
from plen_ml_walk_b200.td3 import ReplayBuffer, TD3Agent
from plen_ml_walk_b200.vec_env import PlenVecEnv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs-per-gpu", type=int, default=16384)
    ap.add_argument("--vector-steps", type=int, default=32)
    ap.add_argument("--updates-per-step", type=int, default=8)
    ap.add_argument("--batch-size", type=int, default=100)
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.pop("NCCL_DEBUG", None)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)                                   # identical initial networks on every rank
    agent = TD3Agent(device=dev, seed=env_seed(0, rank))    # ... different minibatch / noise streams
    env = PlenVecEnv(a.envs_per_gpu, device=dev)
    rb = ReplayBuffer(2 * a.envs_per_gpu * a.vector_steps, device=dev, seed=env_seed(0, rank))
    gen = torch.Generator(device=dev); gen.manual_seed(env_seed(0, rank))

    def allreduce_mean(flat_grad):
        if world > 1:
            dist.all_reduce(flat_grad)
            flat_grad.mul_(1.0 / world)

    state = env.reset().clone()
    updates = 0
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.vector_steps):
        action = torch.empty((a.envs_per_gpu, 18), device=dev).uniform_(-1, 1, generator=gen)
        obs, reward, done, info = env.step(action)
        rb.add(state, action, torch.where(done[:, None], info["terminal_obs"], obs), reward, done & ~info["timeout"])
        state.copy_(obs)
        for _ in range(a.updates_per_step):
            agent.train(rb, a.batch_size, grad_hook=allreduce_mean)
            updates += 1
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    chk = torch.stack([agent._flat[k].double().sum() for k in ("actor", "critic", "actor_target", "critic_target")])
    same = True
    if world > 1:
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        same = all(bool(torch.equal(c, allc[0])) for c in allc)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "envs_per_gpu": a.envs_per_gpu, "vector_steps": a.vector_steps, "updates": updates,
                          "batch_per_gpu": a.batch_size, "allreduce_floats_per_update": 155138 + 77330 // 2,
                          "seconds": dt, "env_steps_per_s": world * a.envs_per_gpu * a.vector_steps / dt,
                          "updates_per_s": updates / dt, "parameters_identical_across_ranks": same}))
    if world > 1:
        dist.destroy_process_group()
    assert same, "ranks diverged"


if __name__ == "__main__":
    main()
