"""Dev tool: hammer plen_step_host (four ranges on four streams) against plen_step (one stream) for bit-equality.
    python scripts/stress_step_host.py [rounds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plen_ml_walk_b200.vec_env import PlenVecEnv

R = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bad_total = 0
for n in (5000, 12345, 40000):
    a, b = PlenVecEnv(n), PlenVecEnv(n)
    g = torch.Generator(device="cuda"); g.manual_seed(n)
    a.reset()
    for _ in range(10):
        a.step(torch.empty((n, 18), device="cuda").uniform_(-1, 1, generator=g))
    b.reset(); b.set_state(*[t.clone() for t in a.get_state()])
    torch.cuda.synchronize()
    h_obs = torch.empty((n, 26)).pin_memory(); h_rew = torch.empty(n).pin_memory()
    h_done = torch.empty(n, dtype=torch.uint8).pin_memory()
    for r in range(R):
        act = torch.empty((n, 18), device="cuda").uniform_(-1, 1, generator=g)
        obs, rew, done, _ = a.step(act)
        torch.cuda.synchronize()
        h_act = act.cpu().pin_memory()
        b.step_host(h_act, h_obs, h_rew, h_done)
        bad = (torch.nan_to_num(obs.cpu(), nan=1234.5) != torch.nan_to_num(h_obs, nan=1234.5)).any(1)
        if bool(bad.any()):
            bad_total += int(bad.sum())
            print("n %d round %d: %d robots differ, first %s" % (n, r, int(bad.sum()), bad.nonzero()[:8, 0].tolist()))
            b.set_state(*[t.clone() for t in a.get_state()])      # re-align and go on
            torch.cuda.synchronize()
print("stress done: %d mismatching robots in total" % bad_total)
