"""Dev tool: the same robots in two contexts, the second one holding them in a random PERMUTATION (so every robot shares its
solver warp with different neighbours): after every step each robot must agree bit for bit with its twin.
    python scripts/stress_permutation.py [envs] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plen_ml_walk_b200.vec_env import PlenVecEnv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
S = int(sys.argv[2]) if len(sys.argv) > 2 else 200
a, b = PlenVecEnv(N), PlenVecEnv(N)
g = torch.Generator(device="cuda"); g.manual_seed(1)
perm = torch.randperm(N, device="cuda", generator=g)
a.reset(); b.reset()
bad_steps, bad_robots = 0, 0
for s in range(S):
    act = torch.empty((N, 18), device="cuda").uniform_(-1, 1, generator=g)
    pre = a.debug_records().clone()
    a.step(act); b.step(act[perm])
    ra, rb = a.debug_records()[perm], b.debug_records()
    same = (ra[:, :80].view(torch.int32) == rb[:, :80].view(torch.int32)).all(1)
    if not bool(same.all()):
        idx = (~same).nonzero()[:, 0]
        bad_steps += 1; bad_robots += idx.numel()
        i = int(idx[0]); j = int(perm[i])
        w = (ra[i, :80].view(torch.int32) != rb[i, :80].view(torch.int32)).nonzero()[:, 0].tolist()
        print("step %d: %d robots differ; robot %d: words %s manifold before %d iters %d/%d" % (
            s, idx.numel(), j, w[:8], int(pre[j, 35]), int(ra[i, 79]), int(rb[i, 79])))
        b.set_state(*[t[perm].clone() for t in a.get_state()])
print("permutation stress: %d steps x %d robots, %d steps / %d robots with a mismatch" % (S, N, bad_steps, bad_robots))
