"""ORACLE SIDE (test infrastructure, NOT product code): the REAL reference -- the unmodified PlenWalkEnv of
plen_bullet/src/plen_bullet/plen_env.py on a real PyBullet -- wherever both are reachable.

PyBullet is an un-vendored, un-pinned third-party dependency of the reference and is NOT in this
image (`import pybullet` fails, no wheel in /opt/wheelhouse), so in this round nothing below has ever executed against
a real PyBullet: `available()` is False here and on the GPU boxes, the head-to-head tests skip, and bench.py's
reference arm falls back to the oracle port.  The moment a `pybullet` module is importable (site-packages or
`baseline/_ref/`) and the reference checkout is present, the same code turns "physics parity unpinned" into a measured
statement without any further change:

  * tests/test_pybullet_head_to_head.py drives PyBullet and oracle/plen_oracle.c from identical states
    (reference call sites: plen_env.py:276-315 world set-up, :558-614 reset, :638-692 step);
  * bench.py --impl reference times the reference env itself, one process per host core (kind "pybullet").

Only tests/ and bench.py's reference arm may import this module.
"""
from __future__ import annotations

import importlib
import multiprocessing as mp
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CANDIDATES = ("/root/reference", os.path.join(ROOT, "baseline", "_ref", "reference"))
MOVING = [5, 6, 7, 9, 10, 11, 13, 14, 15, 17, 18, 19, 20, 21, 24, 26, 27, 30]        # plen_env.py:318-320


def _probe_paths():
    extra = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(extra) and extra not in sys.path:
        sys.path.append(extra)


def reference_src():
    """Directory holding the reference's plen.urdf + plen_bullet package (its scripts run with this as cwd), or None."""
    for r in REF_CANDIDATES:
        d = os.path.join(r, "plen_bullet", "src")
        if os.path.exists(os.path.join(d, "plen.urdf")) and os.path.exists(os.path.join(d, "plen_bullet", "plen_env.py")):
            return d
    return None


def available():
    _probe_paths()
    try:
        importlib.import_module("pybullet")
        importlib.import_module("pybullet_data")
    except Exception:
        return False
    return reference_src() is not None


def _gym_shim():
    """The reference env subclasses gym.Env and builds spaces.Box (plen_env.py:1-4, :22, :141-144); when gym is not
    installed a few inert stand-ins are enough for the UNMODIFIED class to import and run."""
    try:
        importlib.import_module("gym")
        return
    except ImportError:
        pass
    gym = types.ModuleType("gym")
    gym.Env = type("Env", (), {})
    spaces = types.ModuleType("gym.spaces")

    class Box:
        def __init__(self, low, high, dtype=np.float32, **kw):
            self.low, self.high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
            self.shape, self.dtype = self.low.shape, dtype

        def sample(self):
            return np.random.uniform(-1, 1, self.shape).astype(self.dtype)

    spaces.Box = Box
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = lambda seed=None: (np.random.RandomState(seed), seed)
    utils.seeding = seeding
    envs = types.ModuleType("gym.envs")
    reg = types.ModuleType("gym.envs.registration")
    reg.register = lambda **kw: None
    envs.registration = reg
    gym.spaces, gym.utils, gym.envs = spaces, utils, envs
    for name, mod in (("gym", gym), ("gym.spaces", spaces), ("gym.utils", utils), ("gym.utils.seeding", seeding),
                      ("gym.envs", envs), ("gym.envs.registration", reg)):
        sys.modules[name] = mod


class BulletPlen:
    """The reference's own env object plus state exchange with the oracle's flat layout (qpos[25] = pos, quat xyzw, q18;
    qvel[24] = v_lin world, omega world, qd18 -- the layout of include/plen_b200.h)."""

    def __init__(self, joint_act=False):
        if not available():
            raise RuntimeError("PyBullet and/or the reference checkout are not reachable")
        _gym_shim()
        src = reference_src()
        self._cwd = os.getcwd()
        os.chdir(src)                       # the reference loads "plen.urdf" and its package:// meshes relative to the cwd
        if src not in sys.path:
            sys.path.insert(0, src)
        self.p = importlib.import_module("pybullet")
        mod = importlib.import_module("plen_bullet.plen_env")
        self.env = mod.PlenWalkEnv(render=False, joint_act=joint_act)      # plen_env.py:34, DIRECT mode :278
        self.robot = self.env.robotId
        self.version = str(getattr(self.p, "getAPIVersion", lambda: "?")())

    def close(self):
        try:
            self.p.disconnect()
        finally:
            os.chdir(self._cwd)

    def reset(self):
        return np.asarray(self.env.reset(), dtype=np.float64)              # plen_env.py:558-614

    def step(self, action):
        obs, r, d, _ = self.env.step(np.asarray(action))                   # plen_env.py:638-692 (no TimeLimit wrapper here)
        return np.asarray(obs, dtype=np.float64), float(r), bool(d)

    def get_state(self):
        p = self.p
        pos, quat = p.getBasePositionAndOrientation(self.robot)
        lin, ang = p.getBaseVelocity(self.robot)
        js = p.getJointStates(self.robot, MOVING)
        qpos = np.concatenate([pos, quat, [s[0] for s in js]])
        qvel = np.concatenate([lin, ang, [s[1] for s in js]])
        return qpos, qvel

    def set_state(self, qpos, qvel):
        """Teleport (this also empties the contact manifolds, as every resetBasePositionAndOrientation does)."""
        p = self.p
        p.resetBasePositionAndOrientation(self.robot, list(qpos[0:3]), list(qpos[3:7]))
        p.resetBaseVelocity(self.robot, list(qvel[0:3]), list(qvel[3:6]))
        for k, j in enumerate(MOVING):
            p.resetJointState(self.robot, j, float(qpos[7 + k]), float(qvel[6 + k]))

    def tick(self, targets, n=1):
        self.env.move_joints(np.asarray(targets, dtype=np.float64))        # plen_env.py:746-753
        for _ in range(n):
            self.p.stepSimulation()                                        # plen_env.py:665-667


def _worker(args):
    steps, seed = args
    b = BulletPlen()
    rng = np.random.default_rng(seed)
    b.reset()
    t, n = 0, 0
    t0 = time.perf_counter()
    for _ in range(steps):
        _, _, done = b.step(rng.uniform(-1, 1, 18).astype(np.float32))
        t += 1
        n += 1
        if done or t >= 500:                                               # plen_td3.py:122-129 resets on done / TimeLimit
            b.reset()
            t = 0
    dt = time.perf_counter() - t0
    ver = b.version
    b.close()
    return n, dt, ver


def timed_throughput(steps, n_procs):
    """BASELINE config 1 scaled out: one process per host core, each 1 env x `steps` random-action steps with resets
    counted in the wall time.  Returns env-steps/s over all processes."""
    ctx = mp.get_context("spawn")
    with ctx.Pool(n_procs) as pool:
        t0 = time.perf_counter()
        res = pool.map(_worker, [(steps, s) for s in range(n_procs)])
        wall = time.perf_counter() - t0
    per_proc = sum(n / dt for n, dt, _ in res)
    return {"value": per_proc, "seconds": max(dt for _, dt, _ in res), "wall": wall, "envs": n_procs, "steps": steps,
            "version": res[0][2]}
