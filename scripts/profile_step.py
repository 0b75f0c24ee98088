"""Tiny driver for ncu: a few env steps of E robots with random actions (python scripts/profile_step.py E STEPS)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plen_ml_walk_b200.vec_env import PlenVecEnv
E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
env = PlenVecEnv(E)
g = torch.Generator(device="cuda"); g.manual_seed(0)
env.reset()
for s in range(S):
    a = torch.empty((E, 18), device="cuda").uniform_(-1, 1, generator=g)
    env.step(a)
torch.cuda.synchronize()
print("ok")
