#!/usr/bin/env python
"""Generate tests/golden/manifold_golden.npz: a short closed trajectory of the ORACLE with the persistent sole manifold
(manifold_mode = 1, support_tie = 1e-7; profiles/r2_physics_pin.md sections 5-6) -- the reset, then 80 env steps of two robots
under seeded actions (one swaying, one falling), with the observations, rewards, done flags and the manifolds after every step.

What this is: a REGRESSION fixture of the restated procedure (support search with the tie tolerance, cache merge / reduction /
refresh, impulses travelling with the points).  It is produced by this repository's own oracle, not by PyBullet, so it pins
nothing about Bullet; it keeps the oracle -- the yardstick of the kernels' sole_manifold option -- from drifting unnoticed.

    python scripts/make_manifold_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import PlenOracle  # noqa: E402

N, STEPS = 2, 80


def rollout():
    o = PlenOracle(N)
    o.cfg.manifold_mode = 1
    o.cfg.support_tie = 1e-7
    obs0 = o.reset()
    man0 = o.get_manifold()
    rng = np.random.default_rng(2026)
    amp = np.array([0.12, 0.9])[:, None]
    acts, obs, rew, done, man = [], [], [], [], []
    for t in range(STEPS):
        a = rng.uniform(-1, 1, (N, 18)) * amp
        ob, r, d, _ = o.step(a)
        acts.append(a); obs.append(ob.copy()); rew.append(np.nan_to_num(r, nan=-1e9)); done.append(d.copy()); man.append(o.get_manifold())
        for e in np.where(d)[0]:
            o.reset_one(int(e))
    return dict(obs0=obs0, man0=man0, actions=np.array(acts), obs=np.array(obs), reward=np.array(rew), done=np.array(done),
                manifold=np.array(man))


if __name__ == "__main__":
    g = rollout()
    out = os.path.join(ROOT, "tests", "golden", "manifold_golden.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, {k: v.shape for k, v in g.items()}, "point counts seen:", sorted(set(g["manifold"][:, :, 48:50].ravel().astype(int))),
          "dones:", int(g["done"].sum()))
