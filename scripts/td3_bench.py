"""Dev tool / config-4 evidence: TD3 updates per second, CUDA learner (plen_td3_train) vs the PyTorch reference rule
(autograd + torch.optim.Adam, eager) on the same device, batch 100 (td3.py:259) and larger minibatches.
    python scripts/td3_bench.py [updates]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plen_ml_walk_b200.td3 import ReplayBuffer, TD3Agent

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dev = torch.device("cuda:0")
torch.manual_seed(0)
rb = ReplayBuffer(max_size=200000, device=dev)
s = torch.randn(200000, 26, device=dev); a = torch.rand(200000, 18, device=dev) * 2 - 1
rb.add(s, a, s + 0.01, -(a ** 2).sum(1), torch.zeros(200000, dtype=torch.bool, device=dev))
out = {}
BATCHES = (100, 256, 1024, 4096, 16384)
for B in BATCHES:
    agent = TD3Agent(device=dev, max_batch=max(BATCHES))
    for _ in range(20):
        agent.train(rb, B)
    torch.cuda.synchronize()
    l0 = agent.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(K):
        agent.train(rb, B)
    e1.record(); torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ref = TD3Agent(device=dev)
    kr = max(50, K // 10)
    batch = rb.sample(B)
    for _ in range(10):
        ref.train_torch(batch)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(kr):
        ref.train_torch(rb.sample(B))
    torch.cuda.synchronize()
    wall_ref = time.perf_counter() - t1
    # the same update with the products of >= 128-row tiles on the tcgen05 tensor cores (TF32 operands)
    tc = TD3Agent(device=dev, max_batch=max(BATCHES), precision="tf32")
    for _ in range(20):
        tc.train(rb, B)
    torch.cuda.synchronize()
    t0, t1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(K):
        tc.train(rb, B)
    t1_.record(); torch.cuda.synchronize()
    tc_us = 1e3 * t0.elapsed_time(t1_) / K
    out["batch_%d" % B] = {"tf32_device_us_per_update": tc_us, "tf32_samples_per_s": B / tc_us * 1e6,
                           "fp32_samples_per_s": B * K / (1e-3 * e0.elapsed_time(e1)), "cuda_updates_per_s": K / wall, "cuda_device_us_per_update": 1e3 * e0.elapsed_time(e1) / K,
                           "launches_per_update": (agent.kernel_launches() - l0) / K,
                           "torch_eager_updates_per_s": kr / wall_ref, "speedup": (K / wall) / (kr / wall_ref)}
print(json.dumps(out))
