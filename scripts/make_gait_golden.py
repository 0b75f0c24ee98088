"""Generate tests/golden/gait_golden.npz from the REFERENCE (run in the authoring container only; /root/reference does
not exist on the GPU box).

* imports plen_bullet/src/plen_bullet/trajectory_generator.py unmodified, with stub modules for matplotlib.pyplot and
  plen_bullet.plen_env (the generator instantiates PlenWalkEnv(joint_act=True) only to read `real_ranges`,
  trajectory_generator.py:52, :203-222; the stub carries the literal table of plen_env.py:170-189);
* runs TrajectoryGenerator(...).main() for the default parameters and for seeded +-10 % jitters of
  height / stride / body_sway / fwd_bias (BASELINE config 3), plus two unreachable parameter sets (math.acos raises
  ValueError, trajectory_generator.py:187-189);
* applies the sign map / assembly of trajectory_eval.py:180-261 exactly as that script does;
* also stores the shipped goldens plen_bullet/trajectories/*_traj.npy (18 x 800) and bend_traj.npy, and the recorded
  policy episode *_cmd.npy (18 x 500).
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gait_golden.npz")

JOINT_NAMES = ["rb_servo_r_hip", "r_hip_r_thigh", "r_thigh_r_knee", "r_knee_r_shin", "r_shin_r_ankle", "r_ankle_r_foot",
               "lb_servo_l_hip", "l_hip_l_thigh", "l_thigh_l_knee", "l_knee_l_shin", "l_shin_l_ankle", "l_ankle_l_foot",
               "torso_r_shoulder", "r_shoulder_rs_servo", "re_servo_r_elbow", "torso_l_shoulder", "l_shoulder_ls_servo",
               "le_servo_l_elbow"]   # plen_env.py:547-554 / trajectory_eval.py joint_names

REAL_RANGES = [[-1.57, 1.57], [-0.15, 1.5], [-0.95, 1.2], [-1.0, 1.57], [-0.95, 1.2], [-0.8, 0.4],
               [-1.57, 1.57], [-1.5, 0.15], [-1.2, 0.95], [-1.0, 1.57], [-1.2, 1.2], [-0.4, 0.8],
               [-1.57, 1.57], [-0.15, 1.57], [-0.2, 0.35], [-1.57, 1.57], [-0.15, 1.57], [-0.2, 0.35]]   # plen_env.py:170-189


def import_reference_generator():
    plt = types.ModuleType("matplotlib.pyplot")
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    pkg = types.ModuleType("plen_bullet")
    pkg.__path__ = [os.path.join(REF, "plen_bullet/src/plen_bullet")]
    envmod = types.ModuleType("plen_bullet.plen_env")

    class PlenWalkEnv:
        def __init__(self, *a, **k):
            self.real_ranges = REAL_RANGES
    envmod.PlenWalkEnv = PlenWalkEnv
    sys.modules["plen_bullet"], sys.modules["plen_bullet.plen_env"] = pkg, envmod
    import importlib
    return importlib.import_module("plen_bullet.trajectory_generator").TrajectoryGenerator


def assemble(traj):
    """trajectory_eval.py:180-261: one gait cycle (20 rfwd + 20 lfwd rows) in action order, and bend_legs."""
    rows = []
    for tab in (traj.foot_walk_rfwd, traj.foot_walk_lfwd):
        for i in range(np.size(tab, 0)):
            r = tab[i]
            rows.append([-r[0], -r[1], -r[2], -r[3], r[4], r[5], r[6], r[7], r[8], r[9], -r[10], r[11],
                         np.pi / 5, np.pi / 8, 0, -np.pi / 5, np.pi / 8, 0])
    bend_legs = traj.bend[:][0]
    for i in range(6):
        bend_legs = np.append(bend_legs, 0)
    bend_legs[13] = 0.5
    bend_legs[16] = 0.5
    for i in range(4):
        bend_legs[i] = -bend_legs[i]
    bend_legs[10] = -bend_legs[10]
    return np.array(rows, dtype=np.float64), np.array(bend_legs, dtype=np.float64)


def main():
    TG = import_reference_generator()
    rng = np.random.default_rng(3)
    base = np.array([30.0, 30.0, 10.0, 5.0, 10.0])          # height, stride, bend_distance, body_sway, fwd_bias
    params = [base.copy()]
    for _ in range(31):
        p = base.copy()
        p[[0, 1, 3, 4]] *= rng.uniform(0.9, 1.1, 4)
        params.append(p)
    params.append(np.array([30.0, 200.0, 10.0, 5.0, 10.0]))   # stride far out of reach
    params.append(np.array([30.0, 30.0, -80.0, 5.0, 10.0]))   # leg longer than l1 + l2
    params = np.array(params)
    cyc, bend, status = [], [], []
    for p in params:
        t = TG(height=p[0], stride=p[1], bend_distance=p[2], body_sway=p[3], fwd_bias=p[4])
        try:
            t.main()
            c, b = assemble(t)
            ok = 0
        except ValueError:
            c, b, ok = np.full((40, 18), np.nan), np.full(18, np.nan), 1
        cyc.append(c); bend.append(b); status.append(ok)
    t0 = TG()
    t0.main()
    shipped = np.stack([np.load(os.path.join(REF, "plen_bullet/trajectories", n + "_traj.npy")) for n in JOINT_NAMES], 1)
    cmd = np.stack([np.load(os.path.join(REF, "plen_bullet/trajectories", n + "_cmd.npy")) for n in JOINT_NAMES], 1)
    shipped_bend = np.load(os.path.join(REF, "plen_bullet/trajectories/bend_traj.npy"))
    full = np.tile(cyc[0], (20, 1))
    assert np.abs(full - shipped).max() == 0.0 and np.abs(bend[0] - shipped_bend).max() == 0.0, "reference no longer reproduces its goldens"
    np.savez_compressed(OUT, params=params, cycle=np.array(cyc), bend=np.array(bend), status=np.array(status, dtype=np.uint8),
                        shipped_traj=shipped, shipped_bend=shipped_bend, shipped_cmd=cmd.astype(np.float32),
                        foot_walk_rfwd=t0.foot_walk_rfwd, foot_walk_lfwd=t0.foot_walk_lfwd, bend_rows=t0.bend)
    print("wrote", OUT, "params", params.shape, "unreachable", int(np.sum(status)), "shipped", shipped.shape, cmd.shape)


if __name__ == "__main__":
    main()
