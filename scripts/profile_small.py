"""Dev tool: per-kernel device time (CUDA events inside the library) of a step at small batch sizes.
python scripts/profile_small.py [E ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plen_ml_walk_b200.vec_env import PlenVecEnv
for E in [int(x) for x in sys.argv[1:]] or [1024, 4096, 16384]:
    env = PlenVecEnv(E)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    env.reset()
    acts = [torch.empty((E, 18), device="cuda").uniform_(-1, 1, generator=g) for _ in range(8)]
    for s in range(60):
        env.step(acts[s % 8])
    torch.cuda.synchronize()
    env.profile_enable(100)
    for s in range(100):
        env.step(acts[s % 8])
    torch.cuda.synchronize()
    r = env.profile_read()
    print(E, {k: (round(1e3 * v / r["steps"], 1) if k != "steps" else v) for k, v in r.items()}, "us per step")
