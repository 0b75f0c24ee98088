"""Generate tests/golden/td3_golden.npz from the REFERENCE (authoring container only).

Imports plen_ros/src/plen_ros_helpers/td3.py unmodified (with a stub `gym` module, td3.py:9), loads the shipped
checkpoint plen_bullet/models/plen_walk_gazebo_3229999_{actor,critic} and records Actor.forward / Critic.forward on seeded
observations (fp32 CPU), together with the actor's weights so the GPU kernel can be fed the same parameters on a box
that has no reference checkout.  Also replays ReplayBuffer.add on a small ring (the storage order known-answer).
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "td3_golden.npz")


def main():
    sys.modules["gym"] = types.ModuleType("gym")
    sys.path.insert(0, os.path.join(REF, "plen_ros/src/plen_ros_helpers"))
    import td3 as ref
    actor = ref.Actor(26, 18, 1.0)
    critic = ref.Critic(26, 18)
    base = os.path.join(REF, "plen_bullet/models/plen_walk_gazebo_3229999")
    actor.load_state_dict(torch.load(base + "_actor", map_location="cpu"))
    critic.load_state_dict(torch.load(base + "_critic", map_location="cpu"))
    rng = np.random.default_rng(0)
    obs = np.zeros((64, 26), dtype=np.float32)
    obs[1, 18], obs[1, 24], obs[1, 25] = 0.160178937611, 1, 1                 # SURVEY.md section 4 known answers
    obs[2:, :18] = rng.uniform(-1, 1, (62, 18))
    obs[2:, 18] = rng.uniform(0.08, 0.2, 62)
    obs[2:, 19:24] = rng.normal(size=(62, 5)) * 0.3
    obs[2:, 24:] = rng.integers(0, 2, (62, 2))
    with torch.no_grad():
        a = actor(torch.from_numpy(obs))
        q1, q2 = critic(torch.from_numpy(obs), a)
    # ReplayBuffer.add order (td3.py:136-147): 5 adds into a ring of 3
    rb = ref.ReplayBuffer(max_size=3)
    for k in range(5):
        rb.add((np.full(26, k), np.full(18, k), np.full(26, k + 0.5), float(k), 0.0))
    ring_rewards = np.array([t[3] for t in rb.storage])
    sd = {k: v.numpy() for k, v in actor.state_dict().items()}
    np.savez_compressed(OUT, obs=obs, actor_out=a.numpy(), q1=q1.numpy(), q2=q2.numpy(), ring_rewards=ring_rewards,
                        ring_ptr=np.array(rb.ptr), **{"actor_" + k.replace(".", "_"): v for k, v in sd.items()})
    print("wrote", OUT, os.path.getsize(OUT), "bytes; actor(zeros)[:4] =", a[0, :4].numpy(), "ring", ring_rewards, rb.ptr)


if __name__ == "__main__":
    main()
