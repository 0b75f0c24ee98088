#!/usr/bin/env python
"""Instrumented FLOP count of the REFERENCE ALGORITHM on the bench workload (SURVEY.md 8d asks for it to be frozen in
BASELINE.md): oracle/plen_oracle.c counts the adds + multiplies of its 33-link ABA, its unit-impulse responses and its
velocity-space projected Gauss-Seidel (plen_oracle_state.flops).  Random actions U(-1,1)^18, auto-reset, as bench.py.

    python scripts/oracle_flops.py [--envs 256] [--steps 100]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle.oracle import PlenOracle

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=256)
ap.add_argument("--steps", type=int, default=100)
a = ap.parse_args()
o = PlenOracle(a.envs, n_threads=min(8, os.cpu_count() or 1))
o.reset()
rng = np.random.default_rng(0)
for _ in range(20):
    o.step(rng.uniform(-1, 1, (a.envs, 18)), auto_reset=True)
f0 = sum(o.states[e].flops for e in range(a.envs))
its, rows = [], []
for _ in range(a.steps):
    o.step(rng.uniform(-1, 1, (a.envs, 18)), auto_reset=True)
    its += [o.states[e].last_iterations for e in range(a.envs)]
    rows += [o.states[e].last_rows for e in range(a.envs)]
f1 = sum(o.states[e].flops for e in range(a.envs))
print(json.dumps({"flop_per_env_step": (f1 - f0) / (a.envs * a.steps), "envs": a.envs, "steps": a.steps,
                  "mean_pgs_iterations": float(np.mean(its)), "mean_rows": float(np.mean(rows)),
                  "note": "includes the 8 reset ticks of the auto-resets (as the bench workload does)"}))
