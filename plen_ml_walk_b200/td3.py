"""Device-resident TD3 for the batched PLEN env -- host mirror of plen_ros/src/plen_ros_helpers/td3.py (Actor :19-57,
Critic :60-117, ReplayBuffer :122-193, TD3Agent :196-376) as driven by plen_bullet/src/plen_td3.py.

Same class / method / state_dict names as the reference so its checkpoints (`plen_bullet/models/*_actor`, `*_critic`:
keys fc1..fc3 / fc1..fc6) load unchanged.  What runs where:

  * ReplayBuffer      hand-written CUDA ring in libplen_b200.so (plen_replay_*): [capacity, 72] floats in HBM,
                      vectorised add of N transitions per call, uniform sampling with replacement on the device
                      (the reference rebuilds O(B^2) host tensors per sample, td3.py:178-191).
  * select_action     hand-written CUDA fused 3-layer MLP (plen_actor_forward), fp32, N observations per launch,
                      optional exploration noise + clip (plen_td3.py:101-104) in the same kernel.
  * train             the reference's update rule (td3.py:259-356) with torch autograd / Adam on the device: library
                      GEMMs (cuBLAS), stated as such -- the batch-100 critic GEMMs are not on the env-step critical path.

No CPU fallback for the two CUDA pieces.
"""
from __future__ import annotations

import copy
import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _abi

STATE_DIM, ACTION_DIM = 26, 18


class Actor(nn.Module):                                    # td3.py:19-57
    def __init__(self, state_dim=STATE_DIM, action_dim=ACTION_DIM, max_action=1.0):
        super().__init__()
        self.fc1 = nn.Linear(state_dim, 256)
        self.fc2 = nn.Linear(256, 256)
        self.fc3 = nn.Linear(256, action_dim)
        self.max_action = max_action

    def forward(self, state):
        a = F.relu(self.fc1(state))
        a = F.relu(self.fc2(a))
        return self.max_action * torch.tanh(self.fc3(a))


class Critic(nn.Module):                                   # td3.py:60-117
    def __init__(self, state_dim=STATE_DIM, action_dim=ACTION_DIM):
        super().__init__()
        self.fc1 = nn.Linear(state_dim + action_dim, 256)
        self.fc2 = nn.Linear(256, 256)
        self.fc3 = nn.Linear(256, 1)
        self.fc4 = nn.Linear(state_dim + action_dim, 256)
        self.fc5 = nn.Linear(256, 256)
        self.fc6 = nn.Linear(256, 1)

    def forward(self, state, action):
        sa = torch.cat([state, action], 1)
        q1 = self.fc3(F.relu(self.fc2(F.relu(self.fc1(sa)))))
        q2 = self.fc6(F.relu(self.fc5(F.relu(self.fc4(sa)))))
        return q1, q2

    def Q1(self, state, action):
        sa = torch.cat([state, action], 1)
        return self.fc3(F.relu(self.fc2(F.relu(self.fc1(sa)))))


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def actor_forward(actor: Actor, obs: torch.Tensor, noise_std: float = 0.0, seed: int = 0, out: torch.Tensor | None = None):
    """Actor.forward for obs [N,26] (float32, CUDA) through the fused CUDA kernel; optional exploration noise + clip."""
    if not obs.is_cuda:
        raise RuntimeError("plen_actor_forward needs CUDA tensors; there is no CPU fallback")
    lib = _abi.load_library()
    obs = obs.contiguous().float()
    n = obs.shape[0]
    if out is None:
        out = torch.empty((n, ACTION_DIM), dtype=torch.float32, device=obs.device)
    w = [actor.fc1.weight, actor.fc1.bias, actor.fc2.weight, actor.fc2.bias, actor.fc3.weight, actor.fc3.bias]
    w = [t.detach().contiguous() for t in w]
    if tuple(w[0].shape) != (256, STATE_DIM) or tuple(w[2].shape) != (256, 256) or tuple(w[4].shape) != (ACTION_DIM, 256):
        raise ValueError("plen_actor_forward implements the reference architecture 26-256-256-18 (td3.py:37-41)")
    dev = obs.device
    with torch.cuda.device(dev):
        rc = lib.plen_actor_forward(dev.index if dev.index is not None else torch.cuda.current_device(), *[_p(t) for t in w],
                                    _p(obs), n, float(actor.max_action), float(noise_std), int(seed) & (2 ** 64 - 1), _p(out),
                                    C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise RuntimeError("plen_actor_forward: %s" % lib.plen_td3_last_error().decode())
    return out


class ReplayBuffer:
    """Device ring of (state, action, next_state, reward, done) tuples -- ReplayBuffer of td3.py:122-193, batched."""

    def __init__(self, max_size=1000000, device="cuda:0", seed=0):
        if not torch.cuda.is_available():
            raise RuntimeError("the replay ring lives in HBM; there is no CPU fallback")
        self.lib = _abi.load_library()
        self.device = torch.device(device)
        self.max_size = int(max_size)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._rb = self.lib.plen_replay_create(self.max_size, idx)
        if not self._rb:
            raise RuntimeError("plen_replay_create: %s" % self.lib.plen_td3_last_error().decode())
        self._seed, self._draws = int(seed), 0

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def __len__(self):
        return int(self.lib.plen_replay_size(self._rb))

    @property
    def ptr(self):
        return int(self.lib.plen_replay_ptr(self._rb))

    def add(self, state, action, next_state, reward, done):
        """N tuples at once: state/next_state [N,26], action [N,18], reward [N], done [N] (bool / uint8: done_bool of
        plen_td3.py:109-110, i.e. False on TimeLimit truncation)."""
        f = lambda t, w: torch.as_tensor(t, device=self.device).float().reshape(-1, w).contiguous()
        s, a, s2 = f(state, STATE_DIM), f(action, ACTION_DIM), f(next_state, STATE_DIM)
        r = torch.as_tensor(reward, device=self.device).float().reshape(-1).contiguous()
        d = torch.as_tensor(done, device=self.device).to(torch.uint8).reshape(-1).contiguous()
        n = s.shape[0]
        if not (a.shape[0] == s2.shape[0] == r.shape[0] == d.shape[0] == n):
            raise ValueError("ReplayBuffer.add: inconsistent batch sizes")
        with torch.cuda.device(self.device):
            rc = self.lib.plen_replay_add(self._rb, _p(s), _p(a), _p(s2), _p(r), _p(d), n, self._stream())
        if rc != 0:
            raise RuntimeError("plen_replay_add: %s" % self.lib.plen_td3_last_error().decode())

    def sample(self, batch_size, return_index=False):
        """-> state, action, next_state, reward [B,1], not_done [B,1]   (td3.py:166-193)"""
        dev, b = self.device, int(batch_size)
        s = torch.empty((b, STATE_DIM), device=dev); a = torch.empty((b, ACTION_DIM), device=dev)
        s2 = torch.empty((b, STATE_DIM), device=dev); r = torch.empty((b, 1), device=dev); nd = torch.empty((b, 1), device=dev)
        idx = torch.empty(b, dtype=torch.int32, device=dev) if return_index else None
        self._draws += 1
        with torch.cuda.device(dev):
            rc = self.lib.plen_replay_sample(self._rb, b, (self._seed * 1000003 + self._draws) & (2 ** 64 - 1), _p(s), _p(a), _p(s2),
                                             _p(r), _p(nd), _p(idx), self._stream())
        if rc != 0:
            raise RuntimeError("plen_replay_sample: %s" % self.lib.plen_td3_last_error().decode())
        return (s, a, s2, r, nd, idx) if return_index else (s, a, s2, r, nd)

    def storage(self):
        """[len, 72] copy of the stored transitions in ring order (tests, checkpointing)."""
        n = len(self)
        if n == 0:
            return torch.empty((0, 72), device=self.device)
        src = self.lib.plen_replay_storage(self._rb)
        return _from_device_ptr(src, n * 72, self.device).view(n, 72).clone()

    def close(self):
        if getattr(self, "_rb", None):
            self.lib.plen_replay_destroy(self._rb)
            self._rb = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _from_device_ptr(ptr, numel, device):
    """torch float32 view of raw device memory owned by the library (no copy)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device=device)


class TD3Agent:                                            # td3.py:196-376
    def __init__(self, state_dim=STATE_DIM, action_dim=ACTION_DIM, max_action=1.0, discount=0.99, tau=0.005,
                 policy_noise=0.2, noise_clip=0.5, policy_freq=2, device="cuda:0", lr=3e-4):
        self.device = torch.device(device)
        self.actor = Actor(state_dim, action_dim, max_action).to(self.device)
        self.actor_target = copy.deepcopy(self.actor)
        self.actor_optimizer = torch.optim.Adam(self.actor.parameters(), lr=lr)
        self.critic = Critic(state_dim, action_dim).to(self.device)
        self.critic_target = copy.deepcopy(self.critic)
        self.critic_optimizer = torch.optim.Adam(self.critic.parameters(), lr=lr)
        self.max_action, self.discount, self.tau = max_action, discount, tau
        self.policy_noise, self.noise_clip, self.policy_freq = policy_noise, noise_clip, policy_freq
        self.total_it = 0
        self._draws = 0

    def select_action(self, state, expl_noise=0.0):
        """Batched: state [N,26] on the device -> action [N,18]; expl_noise = plen_td3.py:101-104 (std = max_action * expl_noise)."""
        self._draws += 1
        return actor_forward(self.actor, torch.as_tensor(state, device=self.device).reshape(-1, STATE_DIM),
                             noise_std=self.max_action * expl_noise, seed=self._draws)

    def train(self, replay_buffer, batch_size=100):
        """One TD3 update, td3.py:259-356 verbatim semantics."""
        self.total_it += 1
        state, action, next_state, reward, not_done = replay_buffer.sample(batch_size)
        with torch.no_grad():
            noise = (torch.randn_like(action) * self.policy_noise).clamp(-self.noise_clip, self.noise_clip)
            next_action = (self.actor_target(next_state) + noise).clamp(-self.max_action, self.max_action)
            target_Q1, target_Q2 = self.critic_target(next_state, next_action)
            target_Q = reward + not_done * self.discount * torch.min(target_Q1, target_Q2)
        current_Q1, current_Q2 = self.critic(state, action)
        critic_loss = F.mse_loss(current_Q1, target_Q) + F.mse_loss(current_Q2, target_Q)
        self.critic_optimizer.zero_grad()
        critic_loss.backward()
        self.critic_optimizer.step()
        actor_loss = None
        if self.total_it % self.policy_freq == 0:
            actor_loss = -self.critic.Q1(state, self.actor(state)).mean()
            self.actor_optimizer.zero_grad()
            actor_loss.backward()
            self.actor_optimizer.step()
            with torch.no_grad():
                for param, target_param in zip(self.critic.parameters(), self.critic_target.parameters()):
                    target_param.mul_(1 - self.tau).add_(param, alpha=self.tau)
                for param, target_param in zip(self.actor.parameters(), self.actor_target.parameters()):
                    target_param.mul_(1 - self.tau).add_(param, alpha=self.tau)
        return actor_loss, critic_loss

    def save(self, filename):                              # same four files as td3.py:358-365
        torch.save(self.critic.state_dict(), filename + "_critic")
        torch.save(self.critic_optimizer.state_dict(), filename + "_critic_optimizer")
        torch.save(self.actor.state_dict(), filename + "_actor")
        torch.save(self.actor_optimizer.state_dict(), filename + "_actor_optimizer")

    def load(self, filename, optimizers=True):
        self.critic.load_state_dict(torch.load(filename + "_critic", map_location=self.device))
        self.actor.load_state_dict(torch.load(filename + "_actor", map_location=self.device))
        if optimizers:   # the reference's optimizer files are python-2 pickles that need weights_only=False
            self.critic_optimizer.load_state_dict(torch.load(filename + "_critic_optimizer", map_location=self.device, weights_only=False))
            self.actor_optimizer.load_state_dict(torch.load(filename + "_actor_optimizer", map_location=self.device, weights_only=False))
        self.actor_target = copy.deepcopy(self.actor)
        self.critic_target = copy.deepcopy(self.critic)
