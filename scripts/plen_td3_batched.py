#!/usr/bin/env python
"""Batched plen_td3 (BASELINE config 4): TD3 with N device-resident PLEN envs and an on-GPU replay ring.

Mirrors plen_bullet/src/plen_td3.py:16-159 with the constants of :22-30 and td3.py:211-219 (start_timesteps 1e4,
expl_noise 0.1, batch 100, discount 0.99, tau 0.005, policy noise 0.2 clip 0.5, delay 2, Adam 3e-4, replay 1e6):
random actions until start_timesteps transitions are stored, then actor + N(0, 0.1) noise clipped to +-1;
done_bool = done and not timeout (plen_td3.py:109-110); the env auto-resets, so next_state of a finished episode is
info["terminal_obs"].  The reference does ONE gradient step per env step on ONE env; with N envs per vector step the
update-to-data ratio is a parameter: --updates-per-step U gradient steps per vector step (U2D = U / N).

    python scripts/plen_td3_batched.py [--envs 16384] [--env-steps 1048576] [--updates-per-step 8]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from plen_ml_walk_b200.td3 import ReplayBuffer, TD3Agent
from plen_ml_walk_b200.vec_env import PlenVecEnv


def run(n_envs, total_env_steps, updates_per_step, start_timesteps=10000, expl_noise=0.1, batch_size=100, seed=0,
        device="cuda:0", replay_size=1000000, learner=True, actor_precision="fp32", learner_precision="fp32"):
    dev = torch.device(device)
    torch.manual_seed(seed)
    env = PlenVecEnv(n_envs, device=dev, seed=seed)
    agent = TD3Agent(26, 18, 1.0, device=dev, max_batch=max(4096, batch_size), precision=learner_precision)
    rb = ReplayBuffer(replay_size, device=dev, seed=seed)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    state = env.reset().clone()
    ep_ret = torch.zeros(n_envs, device=dev)
    ep_len = torch.zeros(n_envs, dtype=torch.int32, device=dev)
    done_count = torch.zeros((), device=dev)           # episode statistics stay on the device: no host sync in the loop
    ret_sum = torch.zeros((), device=dev)
    updates, t = 0, 0
    vec_steps = (total_env_steps + n_envs - 1) // n_envs
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(vec_steps):
        if t < start_timesteps:
            action = torch.empty((n_envs, 18), device=dev).uniform_(-1, 1, generator=gen)        # action_space.sample()
        else:
            action = agent.select_action(state, expl_noise=expl_noise, precision=actor_precision)  # plen_td3.py:101-104
        obs, reward, done, info = env.step(action)
        # plen_td3.py:109-110: done_bool = float(done) if episode_timesteps < env._max_episode_steps else 0 -- also 0 for a
        # robot that falls exactly on the last step of the time limit (info["timeout"] alone is 0 there: TimeLimit semantics)
        ep_len += 1
        done_bool = done & (ep_len < env._max_episode_steps)
        ep_len *= ~done
        next_state = torch.where(done[:, None], info["terminal_obs"], obs)
        rb.add(state, action, next_state, reward, done_bool)
        ep_ret += reward
        done_count += done.sum()
        ret_sum += (ep_ret * done).sum()
        ep_ret *= ~done
        state.copy_(obs)
        t += n_envs
        if learner and t >= start_timesteps:
            for _ in range(updates_per_step):
                agent.train(rb, batch_size)
                updates += 1
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    return {"envs": n_envs, "env_steps": vec_steps * n_envs, "vector_steps": vec_steps, "updates": updates,
            "update_to_data": updates / max(1, vec_steps * n_envs), "seconds": dt, "env_steps_per_s": vec_steps * n_envs / dt,
            "episodes": int(done_count), "mean_episode_return": float(ret_sum) / max(1, int(done_count)), "replay_len": len(rb),
            "learner": learner, "batch_size": batch_size, "actor_precision": actor_precision, "learner_precision": learner_precision,
            "samples_per_env_step": updates * batch_size / max(1, vec_steps * n_envs),
            "td3_kernel_launches": agent.kernel_launches()}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--env-steps", type=int, default=1048576)
    ap.add_argument("--updates-per-step", type=int, default=8)
    ap.add_argument("--no-learner", action="store_true")
    ap.add_argument("--batch-size", type=int, default=100)
    ap.add_argument("--start-timesteps", type=int, default=10000)
    ap.add_argument("--actor-precision", default="fp32", choices=["fp32", "fp16"])
    ap.add_argument("--learner-precision", default="fp32", choices=["fp32", "tf32"],
                    help="tf32: the learner's products on tcgen05 (3xTF32), for minibatches >= 512")
    a = ap.parse_args()
    print(json.dumps(run(a.envs, a.env_steps, a.updates_per_step, learner=not a.no_learner, batch_size=a.batch_size,
                         start_timesteps=a.start_timesteps, actor_precision=a.actor_precision,
                         learner_precision=a.learner_precision)))
