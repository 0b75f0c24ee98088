"""Dev tool: two contexts, same state, same actions, plen_step on one stream -- every record word must agree bit for bit
after every step.  On a mismatch prints the robot, the differing words and its contact / limit situation, re-aligns, goes on.
    python scripts/stress_determinism.py [envs] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plen_ml_walk_b200.vec_env import PlenVecEnv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
S = int(sys.argv[2]) if len(sys.argv) > 2 else 300
OVR = {"sole_manifold": 1} if os.environ.get("PLEN_AB_MANIFOLD") else None      # the persistent sole manifold option
a, b = PlenVecEnv(N, config_overrides=OVR), PlenVecEnv(N, config_overrides=OVR)
g = torch.Generator(device="cuda"); g.manual_seed(0)
a.reset(); b.reset()
bad_steps = 0
for s in range(S):
    act = torch.empty((N, 18), device="cuda").uniform_(-1, 1, generator=g)
    pre = a.debug_records().clone()
    a.step(act); b.step(act)
    ra, rb = a.debug_records(), b.debug_records()
    same = (ra.view(torch.int32) == rb.view(torch.int32)).all(1)
    if not bool(same.all()):
        bad_steps += 1
        idx = (~same).nonzero()[:, 0]
        print("step %d: %d robots differ: %s" % (s, idx.numel(), idx[:6].tolist()))
        for i in idx[:3].tolist():
            w = (ra[i].view(torch.int32) != rb[i].view(torch.int32)).nonzero()[:, 0].tolist()
            print("  robot %d: words %s | manifold before %d after %d/%d | iters %d/%d | max|q| before %.3f | a[0..3] %s" % (
                i, w[:12], int(pre[i, 35]), int(ra[i, 35]), int(rb[i, 35]), int(ra[i, 79]), int(rb[i, 79]),
                float(pre[i, 38:56].abs().max()), [round(float(x), 4) for x in act[i, :4]]))
            print("    diff of words: %s" % [(k, float(ra[i, k]), float(rb[i, k])) for k in w[:4]])
        b.set_state(*[t.clone() for t in a.get_state()])
        rec = a.debug_records()
print("determinism stress: %d steps x %d robots, %d steps with a mismatch" % (S, N, bad_steps))
