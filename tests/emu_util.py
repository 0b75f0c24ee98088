"""TEST-ONLY helpers: build and drive tests/emu/libplen_emu.so (the product's device code compiled for the host,
32 threads emulating one warp) so CPU tests can compare the real kernel source against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np

from plen_ml_walk_b200._abi import PlenConfigC, PlenModelC, model_to_c
from plen_ml_walk_b200.urdf_loader import packaged_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_SRC = os.path.join(ROOT, "tests", "emu", "emu_warp.cpp")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libplen_emu.so")
_DEPS = [EMU_SRC] + [os.path.join(ROOT, "plen_ml_walk_b200", "csrc", f)
                     for f in ("plen_device.cuh", "plen_solve.cuh", "plen_env.cuh", "plen_host_tables.h")] + \
        [os.path.join(ROOT, "include", "plen_b200.h")]


def build_emu(double=False):
    lib = EMU_LIB.replace(".so", "_f64.so") if double else EMU_LIB
    if os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in _DEPS):
        return lib
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-fPIC", "-shared", "-ffp-contract=off",
                           "-I", ROOT, "-o", lib, EMU_SRC] + (["-DPLEN_EMU_DOUBLE"] if double else []))
    return lib


class Emu:
    def __init__(self, joint_act=False, double=False):
        self.L = C.CDLL(build_emu(double))
        self.real = np.float64 if double else np.float32
        assert self.L.emu_real_bytes() == np.dtype(self.real).itemsize
        self.model = model_to_c(packaged_model())
        self.cfg = PlenConfigC()
        self.L.emu_default_config(C.byref(self.cfg), int(joint_act))
        assert self.L.emu_sizeof_config() == C.sizeof(PlenConfigC)
        assert self.L.emu_sizeof_model() == C.sizeof(PlenModelC)

    def set_manifold(self, man):
        """man: [n,52] array of self.real kept alive by the caller (the harness updates it in place); None switches the option's
        storage off."""
        self._man = man
        self.L.emu_set_manifold(man.ctypes.data_as(C.c_void_p) if man is not None else None)

    def init_record(self, n=1):
        rec = np.zeros((n, 96), dtype=self.real)
        for e in range(n):
            self.L.emu_init_record(C.byref(self.cfg), rec[e:].ctypes.data_as(C.c_void_p))
        return rec

    def tick(self, rec, targets, n_ticks=1, debug=False):
        n = rec.shape[0]
        targets = np.ascontiguousarray(targets, dtype=self.real).reshape(n, 18)
        minv = np.zeros((24, 24), dtype=self.real)
        pos = np.zeros((24, 3), dtype=self.real)
        rot = np.zeros((24, 3, 3), dtype=self.real)
        iters = np.zeros(n, dtype=np.int32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self.L.emu_tick(C.byref(self.model), C.byref(self.cfg), vp(rec), vp(targets), n, n_ticks,
                        vp(minv) if debug else None, vp(pos) if debug else None, vp(rot) if debug else None, vp(iters))
        return (minv, pos, rot, iters) if debug else iters

    def step(self, rec, actions, snapshot=None):
        n = rec.shape[0]
        actions = np.ascontiguousarray(actions, dtype=self.real).reshape(n, 18)
        obs = np.zeros((n, 26), dtype=self.real)
        tobs = np.zeros((n, 26), dtype=self.real)
        rew = np.zeros(n, dtype=self.real)
        done = np.zeros(n, dtype=np.uint8)
        tmo = np.zeros(n, dtype=np.uint8)
        if snapshot is None:
            snapshot = np.zeros(96 + 26, dtype=self.real)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self.L.emu_step(C.byref(self.model), C.byref(self.cfg), vp(rec), vp(actions), n, vp(obs), vp(rew), vp(done),
                        vp(tmo), vp(tobs), vp(snapshot))
        return obs, rew, done.astype(bool), tmo.astype(bool), tobs

    def observe(self, rec):
        n = rec.shape[0]
        obs = np.zeros((n, 26), dtype=self.real)
        self.L.emu_observe(C.byref(self.model), C.byref(self.cfg), rec.ctypes.data_as(C.c_void_p), n,
                           obs.ctypes.data_as(C.c_void_p))
        return obs


# ---- record <-> oracle state (layout: plen_ml_walk_b200/csrc/plen_device.cuh W_* enum)
def record_from_oracle_state(st, e=0, dtype=np.float32):
    rec = np.zeros(96, dtype=dtype)
    qpos, qvel = st["qpos"][e], st["qvel"][e]
    rec[0:3] = qvel[3:6]      # omega
    rec[3:6] = qvel[0:3]      # v lin
    rec[6:24] = qvel[6:24]
    rec[24:32] = st["lam_n"][e]
    rec[32:35] = qpos[0:3]
    man = 0
    for i in range(8):
        man |= int(st["in_manifold"][e, i]) << i
    rec[35] = man
    rec[36] = st["book_f"][e, 15]
    rec[38:56] = qpos[7:25]
    rec[56:60] = qpos[3:7]
    rec[60:64] = st["book_i"][e, 0:4]
    rec[64:70] = st["book_f"][e, 0:6]
    rec[70:79] = st["book_f"][e, 6:15]
    return rec


def oracle_state_from_record(rec):
    """rec [n,96] -> dict accepted by PlenOracle.set_state"""
    n = rec.shape[0]
    st = dict(qpos=np.zeros((n, 25)), qvel=np.zeros((n, 24)), lam_n=np.zeros((n, 8)),
              in_manifold=np.zeros((n, 8), dtype=np.int32), book_i=np.zeros((n, 5), dtype=np.int32),
              book_f=np.zeros((n, 16)))
    for e in range(n):
        r = rec[e].astype(np.float64)
        st["qvel"][e, 3:6] = r[0:3]
        st["qvel"][e, 0:3] = r[3:6]
        st["qvel"][e, 6:24] = r[6:24]
        st["lam_n"][e] = r[24:32]
        st["qpos"][e, 0:3] = r[32:35]
        st["in_manifold"][e] = [(int(r[35]) >> i) & 1 for i in range(8)]
        st["book_f"][e, 15] = r[36]
        st["qpos"][e, 7:25] = r[38:56]
        st["qpos"][e, 3:7] = r[56:60]
        st["book_i"][e, 0:4] = r[60:64]
        st["book_f"][e, 0:6] = r[64:70]
        st["book_f"][e, 6:15] = r[70:79]
    return st
