#!/usr/bin/env python
"""Generate tests/golden/env_logic_golden.npz by running the REFERENCE'S OWN PlenWalkEnv class, unmodified, imported
from the read-only checkout, on top of stub `gym` / `pybullet` / `pybullet_data` modules whose physics answers come
from the float64 oracle (oracle/plen_oracle.c).

What this pins: everything the reference computes in Python around the physics -- agent_to_env, the 26-d observation
assembly, the gait-shaped reward with its growing-array histories, the one-sided done rule, the counters and their
ordering quirks (plen_env.py:638-1093) -- i.e. rows A1, O1, O2, D1, R1, E1-E5, S1 of SURVEY.md section 8a.  The
oracle's C restatement of that logic must reproduce these vectors to float64 round-off (tests/test_env_logic_golden.py).
What it does NOT pin: the physics itself (PyBullet is not installable here; "parity unpinned").

Run here (needs /root/reference):  python scripts/make_env_logic_golden.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"

from oracle.oracle import PlenOracle  # noqa: E402
from oracle import urdf_tree  # noqa: E402


def install_stubs(phys: PlenOracle):
    tree = phys.tree
    # ---- gym
    gym = types.ModuleType("gym")

    class Env:
        pass

    class Box:
        def __init__(self, low, high, dtype=np.float32):
            self.low, self.high = np.asarray(low), np.asarray(high)
            self.shape = self.low.shape

    spaces = types.ModuleType("gym.spaces")
    spaces.Box = Box
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = lambda seed=None: (np.random.RandomState(seed), seed)
    utils.seeding = seeding
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.register = lambda **kw: None
    envs.registration = registration
    gym.Env, gym.spaces, gym.utils, gym.envs = Env, spaces, utils, envs
    for name, mod in (("gym", gym), ("gym.spaces", spaces), ("gym.utils", utils), ("gym.utils.seeding", seeding),
                      ("gym.envs", envs), ("gym.envs.registration", registration)):
        sys.modules[name] = mod

    # ---- pybullet backed by the oracle physics (client 0, plane = body 0, robot = body 1)
    p = types.ModuleType("pybullet")
    p.GUI, p.DIRECT, p.POSITION_CONTROL = 1, 2, 2
    noop = lambda *a, **k: None
    for f in ("setAdditionalSearchPath", "resetDebugVisualizerCamera", "setRealTimeSimulation", "resetSimulation",
              "setGravity", "changeDynamics", "setCollisionFilterPair", "disconnect"):
        setattr(p, f, noop)
    p.connect = lambda *a, **k: 0
    ids = iter([0, 1])
    p.loadURDF = lambda *a, **k: next(ids)
    p.getNumJoints = lambda body: tree["n_links"]
    p.getBodyInfo = lambda body: (tree["base_name"].encode(), b"plen")

    def getJointInfo(body, i):
        info = [None] * 17
        info[0], info[1], info[12] = i, tree["joint_names"][i].encode(), tree["link_names"][i].encode()
        return tuple(info)

    p.getJointInfo = getJointInfo
    p.getQuaternionFromEuler = lambda e: (0.0, 0.0, 0.0, 1.0) if tuple(e) == (0, 0, 0) else None
    moving = tree["moving_joints"]

    def resetBasePositionAndOrientation(body, posObj, ornObj):
        st = phys.get_state()
        st["qpos"][0, 0:3] = posObj
        st["qpos"][0, 3:7] = ornObj
        st["qvel"][0, 0:6] = 0.0                       # teleport zeroes the base velocity
        st["lam_n"][:] = 0                             # and invalidates the cached manifold points
        st["in_manifold"][:] = 0
        phys.set_state(st)

    def resetJointState(body, joint, value):
        st = phys.get_state()
        d = moving.index(joint)
        st["qpos"][0, 7 + d] = value
        st["qvel"][0, 6 + d] = 0.0
        phys.set_state(st)

    def setJointMotorControlArray(bodyUniqueId, jointIndices, controlMode, targetPositions, forces):
        assert list(jointIndices) == moving and controlMode == p.POSITION_CONTROL
        assert np.allclose(forces, 0.15)
        phys.set_state({"target": np.asarray(targetPositions, dtype=np.float64)[None]})

    p.resetBasePositionAndOrientation = resetBasePositionAndOrientation
    p.resetJointState = resetJointState
    p.setJointMotorControlArray = setJointMotorControlArray
    p.stepSimulation = lambda: phys.tick(0)

    def getBasePositionAndOrientation(body):
        s = phys.states[0]
        out = np.empty(2, dtype=object)      # the reference wraps this in np.array(...) (plen_env.py:771), which was a
        out[0], out[1] = tuple(s.pos[:]), tuple(s.quat[:])   # ragged object array under the NumPy of 2020
        return out

    p.getBasePositionAndOrientation = getBasePositionAndOrientation
    p.getJointStates = lambda body, joints: [(phys.states[0].q[moving.index(j)], phys.states[0].qd[moving.index(j)],
                                              (0,) * 6, 0.0) for j in joints]
    p.getBaseVelocity = lambda body: (tuple(phys.states[0].vel[:]), tuple(phys.states[0].omega[:]))

    def getContactPoints(bodyA, bodyB, link):
        f = {urdf_tree.RIGHT_FOOT_LINK: 0, urdf_tree.LEFT_FOOT_LINK: 1}[link]
        return [("pt",)] * int(sum(phys.states[0].in_manifold[f][:]))

    p.getContactPoints = getContactPoints

    def mat_to_quat(R):
        tr = R[0, 0] + R[1, 1] + R[2, 2]
        if tr > 0:
            s = np.sqrt(tr + 1.0) * 2
            return ((R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s)
        if R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
            s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
            return (0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s, (R[2, 1] - R[1, 2]) / s)
        if R[1, 1] > R[2, 2]:
            s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
            return ((R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s, (R[0, 2] - R[2, 0]) / s)
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        return ((R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s, (R[1, 0] - R[0, 1]) / s)

    def getLinkState(body, link):
        pos, rot = phys.fk(0)
        return (tuple(pos[link + 1]), mat_to_quat(rot[link + 1]))

    p.getLinkState = getLinkState

    def getEulerFromQuaternion(q):
        x, y, z, w = q
        sarg = -2.0 * (x * z - w * y)
        if sarg <= -0.99999:
            return (0.0, -0.5 * np.pi, 2 * np.arctan2(x, -y))
        if sarg >= 0.99999:
            return (0.0, 0.5 * np.pi, 2 * np.arctan2(-x, y))
        return (np.arctan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z), np.arcsin(sarg),
                np.arctan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z))

    p.getEulerFromQuaternion = getEulerFromQuaternion
    sys.modules["pybullet"] = p
    pd = types.ModuleType("pybullet_data")
    pd.getDataPath = lambda: ""
    sys.modules["pybullet_data"] = pd


def run_episode_set(joint_act, seed, n_steps, action_fn):
    phys = PlenOracle(1, joint_act=joint_act)
    install_stubs(phys)
    sys.path.insert(0, os.path.join(REF, "plen_bullet", "src"))
    for k in [k for k in sys.modules if k.startswith("plen_bullet")]:
        del sys.modules[k]
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        from plen_bullet.plen_env import PlenWalkEnv
        env = PlenWalkEnv(render=False, joint_act=joint_act)
        obs0 = env.reset()
    rng = np.random.default_rng(seed)
    A, O, R, D, RO = [], [], [], [], []
    ep_t = 0
    for t in range(n_steps):
        # float32-representable values handed over as float64: under NumPy >= 2 a np.float32 scalar would drag
        # `m * agent_val + b` (plen_env.py:704) down to float32, which the NumPy of 2020 did not do
        a = action_fn(rng, t).astype(np.float64)
        with contextlib.redirect_stdout(io.StringIO()):
            obs, r, done, _ = env.step(a)
        ep_t += 1
        A.append(a); O.append(obs); R.append(r); D.append(bool(done))
        if done or ep_t >= 500:          # TimeLimit wrapper of the registration (plen_env.py:15-19) + caller reset
            with contextlib.redirect_stdout(io.StringIO()):
                RO.append(env.reset())
            ep_t = 0
        else:
            RO.append(np.full(26, np.nan))
    return dict(actions=np.array(A, dtype=np.float64), rewards=np.array(R, dtype=np.float64), dones=np.array(D),
                obs=np.array(O), reset_obs=np.array(RO), obs0=np.array(obs0))


if __name__ == "__main__":
    out = {}
    # (a) random agent-space actions incl. exact +-1 (the 1e-3 inset branch) and long survivals via small actions
    def act_random(rng, t):
        a = rng.uniform(-1, 1, 18).astype(np.float32)
        if t % 7 == 0:
            a[rng.integers(0, 18)] = 1.0
        if t % 11 == 0:
            a[rng.integers(0, 18)] = -1.0
        return a

    # (b) gentle standing-ish actions so gait counters pass 80 / 120 and double support passes 16
    def act_gentle(rng, t):
        return rng.normal(0, 0.002, 18).astype(np.float32)

    for name, ja, seed, n, fn in (("random", False, 0, 400, act_random), ("gentle", False, 1, 300, act_gentle)):
        res = run_episode_set(ja, seed, n, fn)
        for k, v in res.items():
            out["%s_%s" % (name, k)] = v
        print(name, "steps", n, "episodes", int(res["dones"].sum()), "reward range", res["rewards"].min(), res["rewards"].max(),
              "nan rewards", int(np.isnan(res["rewards"]).sum()))
    path = os.path.join(ROOT, "tests", "golden", "env_logic_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
