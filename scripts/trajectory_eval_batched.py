#!/usr/bin/env python
"""Batched trajectory_eval (BASELINE config 3): analytic sinewave-gait IK + open-loop joint replay for N robots.

Mirrors plen_bullet/src/trajectory_eval.py:154-300: TrajectoryGenerator().main(), sign map, 20 x step(bend_legs), then
800 x step(row) with joint_act=True; `done` is ignored there and no metric is computed, so fall = any
`dead` (done without timeout) within the 820 steps and distance = final torso_x.  Env 0 uses the default gait; the others a seeded +-10 %
jitter of height / stride / body_sway / fwd_bias.

    python scripts/trajectory_eval_batched.py [--envs 65536] [--seed 0]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from plen_ml_walk_b200.gait import DEFAULTS, TrajectoryGenerator
from plen_ml_walk_b200.vec_env import PlenVecEnv


def gait_params(n, seed):
    base = np.array([DEFAULTS[k] for k in ("height", "stride", "bend_distance", "body_sway", "fwd_bias")])
    p = np.tile(base, (n, 1))
    rng = np.random.default_rng(seed)
    p[1:, [0, 1, 3, 4]] *= rng.uniform(0.9, 1.1, (n - 1, 4))
    return p


def run(n, seed, device="cuda:0"):
    p = gait_params(n, seed)
    gen = TrajectoryGenerator(height=p[:, 0], stride=p[:, 1], bend_distance=p[:, 2], body_sway=p[:, 3], fwd_bias=p[:, 4],
                              device=device).main()
    env = PlenVecEnv(n, device=device, joint_act=True, auto_reset=False)
    env.reset()
    fell = torch.zeros(n, dtype=torch.bool, device=device)
    bend = gen.bend_legs.float().contiguous()
    cyc = gen.cycle.float().contiguous()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):                                   # trajectory_eval.py:276-279
        _, _, done, info = env.step(bend)
        fell |= done & ~info["timeout"]          # dead, not the TimeLimit(500) truncation
    for t in range(800):                                  # :282-300
        _, _, done, info = env.step(cyc[:, t % 40].contiguous())
        fell |= done & ~info["timeout"]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    qpos, _, _ = env.get_state()
    x = qpos[:, 0].double()
    return {"envs": n, "ik_unreachable": int(gen.status.sum().item()), "fall_rate": float(fell.float().mean().item()),
            "torso_x_mean": float(x.mean().item()), "torso_x_std": float(x.std().item()),
            "torso_x_env0": float(x[0].item()), "fell_env0": bool(fell[0].item()),
            "env_steps_per_s": n * 820 / dt}, gen, env


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    res, _, _ = run(a.envs, a.seed)
    print(json.dumps(res))
