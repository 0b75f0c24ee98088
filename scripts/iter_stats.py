"""Dev tool: distribution of PGS iterations / active contact points on the bench workload, and how much of a solver
warp's loop is spent waiting for its slowest robot under different grouping keys (python scripts/iter_stats.py [E] [STEPS])."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from plen_ml_walk_b200.vec_env import PlenVecEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
S = int(sys.argv[2]) if len(sys.argv) > 2 else 80
env = PlenVecEnv(E)
g = torch.Generator(device="cuda"); g.manual_seed(0)
env.reset()
prev = None
rows = []
for s in range(S):
    a = torch.empty((E, 18), device="cuda").uniform_(-1, 1, generator=g)
    env.step(a)
    rec = env.debug_records().cpu().numpy()
    it = rec[:, 79].astype(int); man = rec[:, 35].astype(int)
    n0 = np.array([bin(m & 15).count("1") for m in man]); n1 = np.array([bin((m >> 4) & 15).count("1") for m in man])
    if s >= S - 20 and prev is not None:
        rows.append((it, n0, n1, prev))
    prev = it
it = np.concatenate([r[0] for r in rows]); n0 = np.concatenate([r[1] for r in rows]); n1 = np.concatenate([r[2] for r in rows])
pv = np.concatenate([r[3] for r in rows])
print("iterations: mean %.2f median %d p90 %d frac at cap(50) %.3f" % (it.mean(), np.median(it), np.percentile(it, 90), (it >= 50).mean()))
print("hist iters (bins of 5):", np.histogram(it, bins=range(0, 56, 5))[0] / len(it))
for k in range(0, 9):
    m = (n0 + n1) == k
    if m.any():
        print("points %d: frac %.3f mean iters %.1f" % (k, m.mean(), it[m].mean()))
print("corr(iters, prev step iters) = %.3f" % np.corrcoef(it, pv)[0, 1])

def cost(key, it, n0, n1, tile=1024, grp=8):
    """mean over warps of max-iters x (18 + 6 * (nmax0 + nmax1)) row-slots, relative to the per-robot ideal"""
    tot = 0.0; ideal = 0.0
    for t0 in range(0, len(it) - tile + 1, tile):
        sl = slice(t0, t0 + tile)
        o = np.argsort(-key[sl], kind="stable")
        i_, a_, b_ = it[sl][o].reshape(-1, grp), n0[sl][o].reshape(-1, grp), n1[sl][o].reshape(-1, grp)
        tot += (i_.max(1) * (18 + 6 * (a_.max(1) + b_.max(1)))).sum() * grp
        ideal += (it[sl] * (18 + 6 * (n0[sl] + n1[sl]))).sum()
    return tot / ideal

nm = np.maximum(n0, n1)
print("row-slot inflation, key = load (current):        %.3f" % cost(nm * 25 + n0 * 5 + n1, it, n0, n1))
print("row-slot inflation, key = load, then prev iters: %.3f" % cost((nm * 25 + n0 * 5 + n1) * 64 + pv, it, n0, n1))
print("row-slot inflation, key = prev iters, then load: %.3f" % cost(pv * 128 + (nm * 25 + n0 * 5 + n1), it, n0, n1))
print("row-slot inflation, key = oracle (iters, load):  %.3f" % cost(it * 128 + (nm * 25 + n0 * 5 + n1), it, n0, n1))
print("row-slot inflation, no sort:                     %.3f" % cost(np.zeros_like(it), it, n0, n1))
