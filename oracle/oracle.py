"""ORACLE (test infrastructure, NOT product code): ctypes front end of oracle/plen_oracle.c.

`PlenOracle` exposes the float64 CPU restatement of PlenWalkEnv (reference:
plen_bullet/src/plen_bullet/plen_env.py:558-692) for N independent envs; `PlenOracleEnv` is the 1-env
Gym-shaped view used by the parity tests so they read like the reference's callers
(plen_td3.py:108, trajectory_eval.py:279).

Physics parity is UNPINNED (no PyBullet in reach); the env logic is pinned to plen_env.py.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import urdf_tree

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libplen_oracle.so")

MAXL, NDOF, NJ, NFEET, NPTS, MAXBOX, MAXHULL = 32, 24, 18, 2, 4, 32, 256


class Model(C.Structure):
    _fields_ = [
        ("n_links", C.c_int), ("parent", C.c_int * MAXL), ("jtype", C.c_int * MAXL), ("dof", C.c_int * MAXL),
        ("axis", (C.c_double * 3) * MAXL), ("R_pj", (C.c_double * 9) * MAXL), ("p_pj", (C.c_double * 3) * MAXL),
        ("com", (C.c_double * 3) * MAXL), ("mass", C.c_double * MAXL), ("inertia", (C.c_double * 3) * MAXL),
        ("lower", C.c_double * MAXL), ("upper", C.c_double * MAXL),
        ("base_mass", C.c_double), ("base_com", C.c_double * 3), ("base_inertia", C.c_double * 3),
        ("foot_link", C.c_int * NFEET), ("foot_pts", ((C.c_double * 3) * NPTS) * NFEET),
        ("foot_break", C.c_double * NFEET),
        ("n_boxes", C.c_int), ("box_link", C.c_int * MAXBOX), ("box_center", (C.c_double * 3) * MAXBOX),
        ("box_rot", (C.c_double * 9) * MAXBOX), ("box_half", (C.c_double * 3) * MAXBOX),
        ("n_hull", C.c_int * NFEET), ("foot_hull", ((C.c_double * 3) * MAXHULL) * NFEET),
    ]


class Config(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("substeps", C.c_int), ("reset_ticks", C.c_int), ("gravity_z", C.c_double),
        ("start_pos", C.c_double * 3), ("motor_max_force", C.c_double), ("joint_act", C.c_int),
        ("linear_damping", C.c_double), ("angular_damping", C.c_double), ("mu_lateral", C.c_double),
        ("mu_spinning", C.c_double), ("mu_rolling", C.c_double), ("restitution", C.c_double),
        ("env_lo", C.c_double * NJ), ("env_hi", C.c_double * NJ), ("max_episode_steps", C.c_int),
        ("motor_kp", C.c_double), ("motor_kd", C.c_double), ("solver_iterations", C.c_int),
        ("residual_threshold", C.c_double), ("erp_contact", C.c_double), ("erp_joint", C.c_double),
        ("linear_slop", C.c_double), ("warmstart_factor", C.c_double), ("restitution_vel_threshold", C.c_double),
        ("hull_margin", C.c_double), ("max_coord_velocity", C.c_double), ("implicit_cone", C.c_int),
        ("link_contacts", C.c_int), ("mu_link", C.c_double), ("restitution_base", C.c_double), ("max_contact_points", C.c_int),
        ("manifold_mode", C.c_int), ("support_tie", C.c_double),
    ]


class State(C.Structure):
    _fields_ = [
        ("pos", C.c_double * 3), ("quat", C.c_double * 4), ("omega", C.c_double * 3), ("vel", C.c_double * 3),
        ("q", C.c_double * NJ), ("qd", C.c_double * NJ), ("target", C.c_double * NJ),
        ("lam_n", (C.c_double * NPTS) * NFEET), ("in_manifold", (C.c_int * NPTS) * NFEET),
        ("cnt", C.c_int), ("ds", C.c_int), ("hist_len", C.c_int), ("ep_t", C.c_int), ("dead", C.c_int),
        ("last", C.c_double * 6), ("sums", C.c_double * 9), ("ep_ret", C.c_double),
        ("last_iterations", C.c_int), ("last_rows", C.c_int), ("last_box_points", C.c_int),
        ("last_boxes_touching", C.c_int), ("flops", C.c_longlong),
        ("man_n", C.c_int * NFEET), ("man_local", ((C.c_double * 3) * NPTS) * NFEET), ("man_world", ((C.c_double * 3) * NPTS) * NFEET),
    ]


def build(force=False):
    """Compile oracle/plen_oracle.c -> oracle/_build/libplen_oracle.so (gcc; no reference sources involved)."""
    src = [os.path.join(HERE, f) for f in ("plen_oracle.c", "plen_oracle.h")]
    if (not force) and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src):
        return LIB_PATH
    subprocess.check_call(["make", "-C", HERE, "-s", "CC=gcc"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        assert L.plen_oracle_sizeof_state() == C.sizeof(State), (L.plen_oracle_sizeof_state(), C.sizeof(State))
        assert L.plen_oracle_sizeof_model() == C.sizeof(Model)
        assert L.plen_oracle_sizeof_config() == C.sizeof(Config)
        L.plen_oracle_minv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.plen_oracle_fk.argtypes = [C.c_void_p] * 4
        L.plen_oracle_observe.argtypes = [C.c_void_p] * 5
        L.plen_oracle_tick.argtypes = [C.c_void_p] * 3
        L.plen_oracle_reset.argtypes = [C.c_void_p] * 4
        L.plen_oracle_step.argtypes = [C.c_void_p] * 8
        L.plen_oracle_step_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_int]
        L.plen_oracle_reset_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        for f in ("plen_oracle_minv", "plen_oracle_fk", "plen_oracle_observe", "plen_oracle_tick", "plen_oracle_reset",
                  "plen_oracle_step", "plen_oracle_step_batch", "plen_oracle_reset_batch", "plen_oracle_default_config",
                  "plen_oracle_init_state"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def model_from_tree(tree) -> Model:
    m = Model()
    n = tree["n_links"]
    m.n_links = n
    d = 0
    for i in range(n):
        m.parent[i] = tree["parent"][i]
        m.jtype[i] = tree["jtype"][i]
        m.dof[i] = -1
    for d, link in enumerate(tree["moving_joints"]):
        m.dof[link] = d
    for i in range(n):
        for k in range(3):
            m.axis[i][k] = tree["axis"][i][k]
            m.p_pj[i][k] = tree["p_pj"][i][k]
            m.com[i][k] = tree["com"][i][k]
            m.inertia[i][k] = tree["inertia"][i][k]
        for k in range(9):
            m.R_pj[i][k] = tree["R_pj"][i][k // 3][k % 3]
        m.mass[i] = tree["mass"][i]
        m.lower[i] = tree["lower"][i]
        m.upper[i] = tree["upper"][i]
    m.base_mass = tree["base_mass"]
    for k in range(3):
        m.base_com[k] = tree["base_com"][k]
        m.base_inertia[k] = tree["base_inertia"][k]
    for f, foot in enumerate(tree["feet"]):
        m.foot_link[f] = foot["link"]
        m.foot_break[f] = foot["breaking_threshold"]
        for p in range(NPTS):
            for k in range(3):
                m.foot_pts[f][p][k] = foot["points"][p][k]
        hull = foot.get("hull", [])                      # manifold_mode 1 only
        m.n_hull[f] = min(len(hull), MAXHULL)
        for i in range(m.n_hull[f]):
            for k in range(3):
                m.foot_hull[f][i][k] = hull[i][k]
    boxes = tree.get("boxes", [])
    m.n_boxes = len(boxes)
    for b, bx in enumerate(boxes):
        m.box_link[b] = bx["link"]
        for k in range(3):
            m.box_center[b][k] = bx["center"][k]
            m.box_half[b][k] = bx["half"][k]
        for k in range(9):
            m.box_rot[b][k] = bx["rot"][k // 3][k % 3]
    return m


class PlenOracle:
    """N independent float64 PLEN envs (vectorised view of the reference's 1-env PlenWalkEnv)."""

    def __init__(self, num_envs=1, joint_act=False, tree=None, n_threads=1, max_contact_points=4):
        """max_contact_points: cap on the box-vs-ground contact points per tick (deepest first).  The wrapper defaults to 4,
        the cap of the CUDA path it is the checker of (PLEN_MAX_BOX_POINTS); -1 = uncapped, what Bullet does (the C default).
        Under random actions 99.98 % of the robot-ticks with box contacts have <= 4 points (profiles/r2_box_contacts.md)."""
        self.L = lib()
        self.tree = tree if tree is not None else urdf_tree.load_tree()
        self.model = model_from_tree(self.tree)
        self.cfg = Config()
        self.L.plen_oracle_default_config(C.byref(self.cfg), int(joint_act))
        self.cfg.max_contact_points = int(max_contact_points)
        self.n = int(num_envs)
        self.states = (State * self.n)()
        for e in range(self.n):
            self.L.plen_oracle_init_state(C.byref(self.cfg), C.byref(self.states[e]))
        self.n_threads = max(1, int(n_threads))
        self._pool = ThreadPoolExecutor(self.n_threads) if self.n_threads > 1 else None

    # ---- slices over threads (ctypes releases the GIL)
    def _slices(self):
        per = (self.n + self.n_threads - 1) // self.n_threads
        return [(a, min(self.n, a + per)) for a in range(0, self.n, per)]

    def _sptr(self, e):
        return C.addressof(self.states) + e * C.sizeof(State)

    def reset(self):
        obs = np.zeros((self.n, 26))

        def run(sl):
            a, b = sl
            self.L.plen_oracle_reset_batch(C.byref(self.model), C.byref(self.cfg), self._sptr(a), b - a,
                                           obs[a:].ctypes.data, 0)

        if self._pool:
            list(self._pool.map(run, self._slices()))
        else:
            run((0, self.n))
        return obs

    def reset_one(self, e):
        obs = np.zeros(26)
        self.L.plen_oracle_reset(C.byref(self.model), C.byref(self.cfg), self._sptr(e), obs.ctypes.data)
        return obs

    def step(self, actions, auto_reset=False):
        actions = np.ascontiguousarray(actions, dtype=np.float64).reshape(self.n, NJ)
        obs = np.zeros((self.n, 26))
        rew = np.zeros(self.n)
        done = np.zeros(self.n, dtype=np.int32)
        tmo = np.zeros(self.n, dtype=np.int32)

        def run(sl):
            a, b = sl
            self.L.plen_oracle_step_batch(C.byref(self.model), C.byref(self.cfg), self._sptr(a), b - a,
                                          actions[a:].ctypes.data, obs[a:].ctypes.data, rew[a:].ctypes.data,
                                          done[a:].ctypes.data, tmo[a:].ctypes.data, int(auto_reset), 0)

        if self._pool:
            list(self._pool.map(run, self._slices()))
        else:
            run((0, self.n))
        return obs, rew, done.astype(bool), tmo.astype(bool)

    def tick(self, e=0):
        self.L.plen_oracle_tick(C.byref(self.model), C.byref(self.cfg), self._sptr(e))

    def observe(self, e=0):
        obs = np.zeros(26)
        fe = np.zeros(6)
        self.L.plen_oracle_observe(C.byref(self.model), C.byref(self.cfg), self._sptr(e), obs.ctypes.data, fe.ctypes.data)
        return obs, fe

    def fk(self, e=0):
        n = self.model.n_links + 1
        pos = np.zeros((n, 3))
        rot = np.zeros((n, 3, 3))
        self.L.plen_oracle_fk(C.byref(self.model), self._sptr(e), pos.ctypes.data, rot.ctypes.data)
        return pos, rot

    def minv(self, e=0):
        M = np.zeros((NDOF, NDOF))
        self.L.plen_oracle_minv(C.byref(self.model), self._sptr(e), M.ctypes.data)
        return M

    # ---- flat state exchange with the CUDA path (layout = include/plen_b200.h)
    def get_state(self):
        """qpos[N,25]=[pos3, quat_xyzw4, q18], qvel[N,24]=[v_lin_world3, omega_world3, qd18], aux dict."""
        qpos = np.zeros((self.n, 25))
        qvel = np.zeros((self.n, 24))
        lam = np.zeros((self.n, 8))
        man = np.zeros((self.n, 8), dtype=np.int32)
        target = np.zeros((self.n, 18))
        book_i = np.zeros((self.n, 5), dtype=np.int32)
        book_f = np.zeros((self.n, 16))
        for e in range(self.n):
            s = self.states[e]
            qpos[e, 0:3] = s.pos[:]
            qpos[e, 3:7] = s.quat[:]
            qpos[e, 7:] = s.q[:]
            qvel[e, 0:3] = s.vel[:]
            qvel[e, 3:6] = s.omega[:]
            qvel[e, 6:] = s.qd[:]
            lam[e] = np.array([list(r) for r in s.lam_n]).ravel()
            man[e] = np.array([list(r) for r in s.in_manifold]).ravel()
            target[e] = s.target[:]
            book_i[e] = [s.cnt, s.ds, s.hist_len, s.ep_t, s.dead]
            book_f[e, :6] = s.last[:]
            book_f[e, 6:15] = s.sums[:]
            book_f[e, 15] = s.ep_ret
        return dict(qpos=qpos, qvel=qvel, lam_n=lam, in_manifold=man, target=target, book_i=book_i, book_f=book_f)

    def get_manifold(self):
        """Persistent sole manifolds (manifold_mode 1) in the C ABI's layout (PLEN_MAN_WORDS = 52 per robot): per foot the local
        xyz of 4 points, then the plane xyz of 4 points; then the two point counts."""
        out = np.zeros((self.n, 52))
        for e in range(self.n):
            s = self.states[e]
            for f in range(NFEET):
                out[e, 24 * f:24 * f + 12] = np.array([list(r) for r in s.man_local[f]]).ravel()
                out[e, 24 * f + 12:24 * f + 24] = np.array([list(r) for r in s.man_world[f]]).ravel()
                out[e, 48 + f] = s.man_n[f]
        return out

    def set_manifold(self, man):
        for e in range(self.n):
            s = self.states[e]
            for f in range(NFEET):
                s.man_n[f] = int(round(float(man[e, 48 + f])))
                for k in range(NPTS):
                    for c in range(3):
                        s.man_local[f][k][c] = float(man[e, 24 * f + 3 * k + c])
                        s.man_world[f][k][c] = float(man[e, 24 * f + 12 + 3 * k + c])

    def set_state(self, st):
        for e in range(self.n):
            s = self.states[e]
            if "qpos" in st:
                for k in range(3):
                    s.pos[k] = st["qpos"][e, k]
                for k in range(4):
                    s.quat[k] = st["qpos"][e, 3 + k]
                for k in range(NJ):
                    s.q[k] = st["qpos"][e, 7 + k]
            if "qvel" in st:
                for k in range(3):
                    s.vel[k] = st["qvel"][e, k]
                    s.omega[k] = st["qvel"][e, 3 + k]
                for k in range(NJ):
                    s.qd[k] = st["qvel"][e, 6 + k]
            if "lam_n" in st:
                for f in range(NFEET):
                    for p in range(NPTS):
                        s.lam_n[f][p] = st["lam_n"][e, f * NPTS + p]
            if "in_manifold" in st:
                for f in range(NFEET):
                    for p in range(NPTS):
                        s.in_manifold[f][p] = int(st["in_manifold"][e, f * NPTS + p])
            if "target" in st:
                for k in range(NJ):
                    s.target[k] = st["target"][e, k]
            if "book_i" in st:
                s.cnt, s.ds, s.hist_len, s.ep_t, s.dead = (int(v) for v in st["book_i"][e])
            if "book_f" in st:
                for k in range(6):
                    s.last[k] = st["book_f"][e, k]
                for k in range(9):
                    s.sums[k] = st["book_f"][e, 6 + k]
                s.ep_ret = st["book_f"][e, 15]


class PlenOracleEnv:
    """1-env Gym-shaped oracle: reset() -> obs[26]; step(a[18]) -> (obs, reward, done, {}) (plen_env.py:558, :638)."""

    def __init__(self, joint_act=False):
        self.o = PlenOracle(1, joint_act=joint_act)

    def reset(self):
        return self.o.reset()[0]

    def step(self, action):
        obs, r, d, t = self.o.step(np.asarray(action, dtype=np.float64)[None])
        return obs[0], float(r[0]), bool(d[0]), {"timeout": bool(t[0])}
