"""PlenVecEnv -- the reference's Gym surface (PlenWalkEnv, plen_bullet/src/plen_bullet/plen_env.py:22-1097),
vectorised over N device-resident robots and driven through the C ABI of libplen_b200.so.

Same verbs and attributes as the reference env so plen_td3.py / walk_eval.py / trajectory_eval.py style loops port
one-to-one:

    env = PlenVecEnv(num_envs=4096, device="cuda:0")          # gym.make("PlenWalkEnv-v1")        plen_td3.py:43
    obs = env.reset()                                          # [N,26]                            plen_env.py:558
    obs, reward, done, info = env.step(actions)                # [N,26], [N], [N] bool, dict       plen_env.py:638

Differences forced by batching: `step` auto-resets finished envs in the same kernel launch (the reference leaves the
reset to the caller, plen_td3.py:122-129); `info["terminal_obs"]` holds the last observation of the finished
episodes, `info["timeout"]` the TimeLimit truncation flag (plen_td3.py:109-110 uses it as done_bool = 0).

torch is used only for device memory and streams; every computation is in the CUDA library.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi
from .urdf_loader import PlenModel, packaged_model

OBS_DIM, ACT_DIM = 26, 18

# plen_env.py:148-167 (env_ranges) and :170-189 (real_ranges)
ENV_RANGES = [[-1.57, 1.57], [-0.15, 1.5], [-0.95, 0.75], [-0.9, 0.3], [-0.95, 1.2], [-0.8, 0.4],
              [-1.57, 1.57], [-1.5, 0.15], [-0.75, 0.95], [-0.3, 0.9], [-1.2, 0.95], [-0.4, 0.8],
              [-1.57, 1.57], [-0.15, 1.57], [-0.2, 0.35], [-1.57, 1.57], [-0.15, 1.57], [-0.2, 0.35]]
REAL_RANGES = [[-1.57, 1.57], [-0.15, 1.5], [-0.95, 1.2], [-1.0, 1.57], [-0.95, 1.2], [-0.8, 0.4],
               [-1.57, 1.57], [-1.5, 0.15], [-1.2, 0.95], [-1.0, 1.57], [-1.2, 1.2], [-0.4, 0.8],
               [-1.57, 1.57], [-0.15, 1.57], [-0.2, 0.35], [-1.57, 1.57], [-0.15, 1.57], [-0.2, 0.35]]


class Box:
    """Minimal stand-in for gym.spaces.Box (gym is not a dependency): low/high/shape/dtype/sample."""

    def __init__(self, low, high, dtype=np.float32, seed=None):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)
        self._rng = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)


def _spaces():
    # plen_env.py:141-144 (action space) and :246-263 (observation space)
    lo = [r[0] for r in ENV_RANGES] + [0, -np.inf, -np.pi, -np.pi, -np.pi, -np.inf, 0, 0]
    hi = [r[1] for r in ENV_RANGES] + [0.25, np.inf, np.pi, np.pi, np.pi, np.inf, 1, 1]
    return Box(-np.ones(ACT_DIM), np.ones(ACT_DIM)), Box(lo, hi)


class PlenVecEnv:
    metadata = {"render.modes": []}

    def __init__(self, num_envs, device="cuda:0", joint_act=False, auto_reset=True, model: PlenModel | None = None,
                 config_overrides: dict | None = None, seed=None):
        if not torch.cuda.is_available():
            raise RuntimeError("PlenVecEnv needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _abi.load_library()
        self.device = torch.device(device)
        self.device_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.num_envs = int(num_envs)
        self.joint_act = bool(joint_act)
        self.model = model if model is not None else packaged_model()
        self._cmodel = _abi.model_to_c(self.model)
        self.cfg = _abi.PlenConfigC()
        self.lib.plen_default_config(C.byref(self.cfg), int(joint_act))
        self.cfg.auto_reset = int(auto_reset)
        if model is not None:
            self.cfg.hull_margin = float(getattr(model, "foot_margin", self.cfg.hull_margin))      # box feet carry no margin
        for k, v in (config_overrides or {}).items():
            setattr(self.cfg, k, v)
        self._ctx = self.lib.plen_create(C.byref(self.cfg), C.byref(self._cmodel), self.num_envs, self.device_index)
        if not self._ctx:
            raise RuntimeError("plen_create failed: %s" % self.lib.plen_last_error(None).decode())
        self.action_space, self.observation_space = _spaces()
        self.action_space.seed(seed)
        self.env_ranges, self.real_ranges = ENV_RANGES, REAL_RANGES
        self._max_episode_steps = int(self.cfg.max_episode_steps)         # TimeLimit attribute, plen_td3.py:110
        n, dev = self.num_envs, self.device
        self._obs = torch.empty((n, OBS_DIM), dtype=torch.float32, device=dev)
        self._terminal_obs = torch.zeros((n, OBS_DIM), dtype=torch.float32, device=dev)
        self._reward = torch.empty(n, dtype=torch.float32, device=dev)
        self._done = torch.empty(n, dtype=torch.uint8, device=dev)
        self._timeout = torch.empty(n, dtype=torch.uint8, device=dev)
        self._scales = {}                                                  # per-env scales set so far (set_env_scales)

    # ---- plumbing
    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("libplen_b200: %s" % self.lib.plen_last_error(self._ctx).decode())

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def _dev_f32(self, x, shape):
        t = torch.as_tensor(x, dtype=torch.float32, device=self.device).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError("expected shape %s, got %s" % (shape, tuple(t.shape)))
        return t

    # ---- Gym surface
    def seed(self, seed=None):
        """The reference env is effectively unseeded (it only defines _seed, plen_env.py:28-32); seeds action_space."""
        self.action_space.seed(seed)
        return [seed]

    def reset(self, mask=None):
        """PlenWalkEnv.reset (plen_env.py:558-614) for all envs, or for those where mask is True."""
        m = None
        if mask is not None:
            m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            self._check(self.lib.plen_reset(self._ctx, self._p(m), self._p(self._obs), self._stream()))
        return self._obs

    @property
    def launches(self):
        """Kernels of ours launched by step()/step_host()/reset()/tick() so far, counted where they are launched (bench.py reports it)."""
        return int(self.lib.plen_kernel_launches(self._ctx))

    def step(self, actions):
        """PlenWalkEnv.step (plen_env.py:638-692) for all N envs in one launch.  Returned tensors are reused buffers."""
        a = self._dev_f32(actions, (self.num_envs, ACT_DIM))
        with torch.cuda.device(self.device):
            self._check(self.lib.plen_step(self._ctx, self._p(a), self._p(self._obs), self._p(self._reward),
                                           self._p(self._done), self._p(self._timeout), self._p(self._terminal_obs),
                                           self._stream()))
        info = {"terminal_obs": self._terminal_obs, "timeout": self._timeout.bool()}
        return self._obs, self._reward, self._done.bool(), info

    def step_host(self, actions_host, obs_host, reward_host, done_host, timeout_host=None):
        """End-to-end call with HOST (ideally pinned) numpy/torch CPU buffers: H2D + step + D2H + sync."""
        def hp(x):
            if x is None:
                return None
            return C.c_void_p(x.data_ptr() if isinstance(x, torch.Tensor) else x.ctypes.data)
        self._check(self.lib.plen_step_host(self._ctx, hp(actions_host), hp(obs_host), hp(reward_host), hp(done_host),
                                            hp(timeout_host)))

    def fault_count(self):
        """Robots retired by the numeric guard (non-finite physical state -> done, reward -100, forced reset) so far."""
        c = C.c_ulonglong()
        self._check(self.lib.plen_fault_count(self._ctx, C.byref(c)))
        return int(c.value)

    def profile_enable(self, max_steps):
        """Record CUDA events around every kernel of the next max_steps step() calls (bench.py's roofline leg)."""
        self._check(self.lib.plen_profile_enable(self._ctx, int(max_steps)))

    def profile_read(self):
        """-> dict(ms_dyn, ms_solve, ms_post, steps): summed device time per kernel family since the last read."""
        d, s, p, n = C.c_float(), C.c_float(), C.c_float(), C.c_int()
        self._check(self.lib.plen_profile_read(self._ctx, C.byref(d), C.byref(s), C.byref(p), C.byref(n)))
        return {"ms_dyn": d.value, "ms_solve": s.value, "ms_post": p.value, "steps": n.value}

    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.plen_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state access (parity tests, checkpointing)
    def get_state(self):
        n, dev = self.num_envs, self.device
        qpos = torch.empty((n, 25), dtype=torch.float32, device=dev)
        qvel = torch.empty((n, 24), dtype=torch.float32, device=dev)
        aux = torch.empty((n, _abi.AUX_WORDS), dtype=torch.float32, device=dev)
        with torch.cuda.device(self.device):
            self._check(self.lib.plen_get_state(self._ctx, self._p(qpos), self._p(qvel), self._p(aux), self._stream()))
        return qpos, qvel, aux

    def set_state(self, qpos=None, qvel=None, aux=None):
        n = self.num_envs
        qpos = None if qpos is None else self._dev_f32(qpos, (n, 25))
        qvel = None if qvel is None else self._dev_f32(qvel, (n, 24))
        aux = None if aux is None else self._dev_f32(aux, (n, _abi.AUX_WORDS))
        with torch.cuda.device(self.device):
            self._check(self.lib.plen_set_state(self._ctx, self._p(qpos), self._p(qvel), self._p(aux), self._stream()))

    def get_manifold(self):
        """Persistent sole manifolds [N, 52] (config_overrides={"sole_manifold": 1} only; layout: include/plen_b200.h)."""
        man = torch.empty((self.num_envs, _abi.MAN_WORDS), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.plen_get_manifold(self._ctx, self._p(man), self._stream()))
        return man

    def set_manifold(self, man):
        man = self._dev_f32(man, (self.num_envs, _abi.MAN_WORDS))
        with torch.cuda.device(self.device):
            self._check(self.lib.plen_set_manifold(self._ctx, self._p(man), self._stream()))

    def set_env_scales(self, friction=None, motor_force=None, motor_gain=None):
        """Per-env domain randomisation: scale factors [N] of the foot friction coefficients, the servo force
        limit and the servo position gain; None leaves a quantity unchanged, 1.0 is the reference's constant.

        Approximation: a reset (explicit or in-kernel auto-reset) restores the ONE post-reset snapshot that was settled for 8
        ticks under unit scales at creation, whatever the robot's own scales are; its first steps then run under its scales."""
        arr = [None if x is None else self._dev_f32(x, (self.num_envs,)) for x in (friction, motor_force, motor_gain)]
        for k, x in zip(("friction", "motor_force", "motor_gain"), arr):
            if x is not None:
                self._scales[k] = x.clone()               # mirror kept for save_state (the library has no getter)
        with torch.cuda.device(self.device):
            self._check(self.lib.plen_set_env_scales(self._ctx, *[self._p(x) for x in arr], self._stream()))

    def save_state(self, path):
        """Snapshot of every env (pose, velocities, contact cache, env bookkeeping) to one .pt file."""
        qpos, qvel, aux = self.get_state()
        torch.save({"num_envs": self.num_envs, "joint_act": self.joint_act, "qpos": qpos.cpu(), "qvel": qvel.cpu(), "aux": aux.cpu(),
                    "scales": {k: v.cpu() for k, v in self._scales.items()},
                    "manifold": self.get_manifold().cpu() if int(self.cfg.sole_manifold) else None}, path)

    def load_state(self, path):
        """Restore a save_state snapshot; stepping on from it reproduces the original run bit for bit."""
        d = torch.load(path, map_location="cpu")
        if int(d["num_envs"]) != self.num_envs or bool(d["joint_act"]) != self.joint_act:
            raise ValueError("snapshot is for %d envs (joint_act=%s)" % (d["num_envs"], d["joint_act"]))
        self.set_state(d["qpos"], d["qvel"], d["aux"])
        if d.get("manifold") is not None:
            self.set_manifold(d["manifold"])
        sc = d.get("scales", {})
        if sc or self._scales:      # per-env scales travel with the snapshot; a snapshot without any restores the unit scales
            ones = torch.ones(self.num_envs)
            self.set_env_scales(*[sc.get(k, ones) for k in ("friction", "motor_force", "motor_gain")])

    def tick(self, targets, n_ticks=1):
        """Raw physics: n_ticks of 1/240 s with joint targets in radians, no env logic (move_joints + stepSimulation)."""
        t = self._dev_f32(targets, (self.num_envs, ACT_DIM))
        with torch.cuda.device(self.device):
            self._check(self.lib.plen_tick(self._ctx, self._p(t), int(n_ticks), self._stream()))

    def debug_records(self):
        """Raw per-env state records [N,96] (word 79 = PGS iterations of the last tick)."""
        rec = torch.empty((self.num_envs, _abi.STATE_WORDS), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.plen_debug_records(self._ctx, self._p(rec), self._stream()))
        return rec

    def debug_dynamics(self):
        n, dev = self.num_envs, self.device
        minv = torch.zeros((n, 24, 24), dtype=torch.float32, device=dev)
        pos = torch.zeros((n, 24, 3), dtype=torch.float32, device=dev)
        rot = torch.zeros((n, 24, 3, 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(self.device):
            self._check(self.lib.plen_debug_dynamics(self._ctx, self._p(minv), self._p(pos), self._p(rot), self._stream()))
        return minv, pos, rot


class PlenWalkEnv:
    """1-env adapter with the reference's exact call shapes: reset() -> np[26]; step(a[18]) -> (obs, r, done, {}).

    Mirrors `gym.make("PlenWalkEnv-v1", render=False, joint_act=...)` (plen_env.py:15-19, :34) including the
    TimeLimit(500) wrapper; like the reference it does NOT auto-reset (the caller does, plen_td3.py:122-129).
    """

    def __init__(self, render=False, realtime=False, joint_act=False, device="cuda:0", config_overrides=None):
        if render or realtime:
            raise NotImplementedError("the B200 env is headless (PyBullet GUI / realtime modes are out of scope)")
        # config_overrides: plen_config fields, e.g. {"sole_manifold": 1} (gym.make passes keyword arguments through)
        self.vec = PlenVecEnv(1, device=device, joint_act=joint_act, auto_reset=False, config_overrides=config_overrides)
        self.action_space, self.observation_space = self.vec.action_space, self.vec.observation_space
        self.env_ranges, self.real_ranges = ENV_RANGES, REAL_RANGES
        self._max_episode_steps = self.vec._max_episode_steps
        self.joint_act = joint_act

    def seed(self, seed=None):
        return self.vec.seed(seed)

    def reset(self):
        return self.vec.reset()[0].cpu().numpy().astype(np.float64)

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1, ACT_DIM)
        obs, r, d, info = self.vec.step(a)
        return obs[0].cpu().numpy().astype(np.float64), float(r[0]), bool(d[0]), {"timeout": bool(info["timeout"][0])}

    def close(self):
        self.vec.close()
