"""Dev tool: time the end-to-end call (plen_step_host: pinned host buffers in, host buffers out) of the in-tree library.
    [PLEN_HOST_RANGES=R] python scripts/ab_host.py [envs] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plen_ml_walk_b200.vec_env import PlenVecEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
env = PlenVecEnv(E)
env.reset()
g = torch.Generator(); g.manual_seed(0)
acts = [torch.empty((E, 18)).uniform_(-1, 1, generator=g).pin_memory() for _ in range(2)]
obs = torch.empty((E, 26)).pin_memory(); rew = torch.empty(E).pin_memory(); done = torch.empty(E, dtype=torch.uint8).pin_memory()
for w in range(5):
    env.step_host(acts[w % 2], obs, rew, done)
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(K):
    env.step_host(acts[k % 2], obs, rew, done)
dt = (time.perf_counter() - t0) / K
print("step_host ranges=%s: %d robots %.3f ms/step  %.3f M env-steps/s  (reward mean %.4f)" % (os.environ.get("PLEN_HOST_RANGES", "auto"), E, dt * 1e3, E / dt / 1e6, float(rew.mean())))
