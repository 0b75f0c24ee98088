"""ORACLE (test infrastructure, NOT product code): numpy float64 restatement of the reference's sinewave gait
generator and leg IK for a batch of parameter sets.

Follows plen_bullet/src/plen_bullet/trajectory_generator.py -- foot_path :54-152, assemble_trajectories :154-167,
IK :169-233, joint_space_trajectories :235-270 -- and the trajectory assembly of plen_bullet/src/trajectory_eval.py
:180-261 (sign map, arm constants, bend_legs).  Pinned: tests/test_gait_oracle.py checks it against
tests/golden/gait_golden.npz, which scripts/make_gait_golden.py produced by running the reference generator itself
(and which contains the shipped plen_bullet/trajectories/*_traj.npy goldens).
Only tests/ may import this.
"""
import numpy as np

L1, L2 = 25.0, 40.0                      # trajectory_generator.py:35-37
NDS, NSS = 5, 10                         # :10-11 defaults
# real_ranges rows used by IK (plen_env.py:170-189): [2] R thigh, [3] R knee, [8] L thigh, [9] L knee
R_THIGH, R_KNEE, L_THIGH, L_KNEE = (-0.95, 1.2), (-1.0, 1.57), (-1.2, 0.95), (-1.0, 1.57)


def foot_paths(height, stride, bend, sway, bias):
    """-> (rfwd_r, lfwd_r), each [N,3,20]: right-foot cartesian path when the right / the left foot leads."""
    height, stride, bend, sway, bias = (np.asarray(a, dtype=np.float64)[:, None] for a in (height, stride, bend, sway, bias))
    ts = (np.arange(NSS) / (NSS - 1.0))[None]
    td = (np.arange(2 * NDS) / (2 * NDS - 1.0))[None]
    one = np.ones_like(height)
    ss_dom = np.stack([ts * stride, np.sin(-np.pi * ((1 / 3.0) * (1 + ts))) * sway, np.sin(ts * np.pi) * height + bend], 1)
    ds_dom = np.stack([-stride * (td / 2.0), np.sin(-np.pi * ((-1.0 / 3.0) + (2 / 3.0) * td)) * sway, bend * one * np.ones_like(td)], 1)
    ss_sup = np.stack([stride * ((1.0 / 2.0) - ts) / 2.0, np.sin(np.pi * ((1 / 3.0) * (1 + ts))) * sway, bend * np.ones_like(ts)], 1)
    ds_sup = np.stack([stride * (1.0 - td) / 2.0, np.sin(-np.pi * ((2.0 / 3.0) + (2 / 3.0) * td)) * sway, bend * np.ones_like(td)], 1)
    nz = (bias != 0)[:, :, None]
    for a in (ss_dom, ds_dom, ss_sup, ds_sup):                         # :128-132
        a[:, 0:1] = np.where(nz, a[:, 0:1] - bias[:, :, None], a[:, 0:1])
    rfwd_r = np.concatenate([ds_dom[:, :, NDS:], ss_dom, ds_sup[:, :, :NDS]], 2)      # :135-139
    lfwd_r = np.concatenate([ds_sup[:, :, NDS:], ss_sup, ds_dom[:, :, :NDS]], 2)      # :142-145
    return rfwd_r, lfwd_r


def leg_ik(point, right):
    """point [N,3,T] -> (angles [N,T,6], unreachable [N] bool)   (IK :169-233; math.acos raises ValueError)"""
    zx, zy, zz = point[:, 0], point[:, 1], L1 + L2 - point[:, 2]
    th1 = np.arctan2(-zy, zz) if right else np.arctan2(zy, zz)
    arg = (zx ** 2 + zy ** 2 + zz ** 2 - L1 ** 2 - L2 ** 2) / (2.0 * L1 * L2)
    bad = ~((arg >= -1.0) & (arg <= 1.0))
    with np.errstate(invalid="ignore"):
        th3 = np.arccos(arg)
    sq, hok = np.sqrt(zy ** 2 + zz ** 2), L1 / L2
    th2 = -np.arctan2(sq * np.sin(th3) + zx * np.cos(th3) + zx * hok, sq * np.cos(th3) + sq * hok - zx * np.sin(th3))
    kn, thg = (R_KNEE, R_THIGH) if right else (L_KNEE, L_THIGH)
    th3 = np.clip(th3, kn[0], kn[1])
    th2 = np.clip(th2, thg[0], thg[1])
    th4 = -(th2 + th3)
    th5 = -th1 if right else th1
    out = np.stack([np.zeros_like(th1), th1, th2, th3, th4, th5], -1)
    return out, bad.any(1)


def gait(params):
    """params [N,5] = height, stride, bend_distance, body_sway, fwd_bias (mm) ->
    cycle [N,40,18] (20 right-forward + 20 left-forward rows, action order, sign map applied), bend_legs [N,18],
    status [N] (1 = unreachable, rows NaN)."""
    p = np.atleast_2d(np.asarray(params, dtype=np.float64))
    h, s, b, sw, bias = p.T
    rfwd_r, lfwd_r = foot_paths(h, s, b, sw, bias)
    flip = np.array([1.0, -1.0, 1.0])[None, :, None]
    lfwd_l, rfwd_l = rfwd_r * flip, lfwd_r * flip                       # :160-167
    rows, bad = [], np.zeros(len(p), dtype=bool)
    for pr, pl in ((rfwd_r, rfwd_l), (lfwd_r, lfwd_l)):
        a, ba = leg_ik(pr, True)
        c, bc = leg_ik(pl, False)
        rows.append(np.concatenate([a, c], -1))
        bad |= ba | bc
    r = np.concatenate(rows, 1)                                         # [N,40,12]
    n = len(p)
    cyc = np.empty((n, 40, 18))
    cyc[..., 0:4] = -r[..., 0:4]; cyc[..., 4:6] = r[..., 4:6]           # trajectory_eval.py:180-205
    cyc[..., 6:10] = r[..., 6:10]; cyc[..., 10] = -r[..., 10]; cyc[..., 11] = r[..., 11]
    cyc[..., 12:18] = np.array([np.pi / 5, np.pi / 8, 0, -np.pi / 5, np.pi / 8, 0])
    bp = np.stack([np.zeros(n), np.zeros(n), b], 1)[:, :, None]
    a, ba = leg_ik(bp, True)
    c, bc = leg_ik(bp, False)
    bad |= ba | bc
    bend = np.zeros((n, 18))
    bend[:, 0:12] = np.concatenate([a[:, 0], c[:, 0]], -1)              # :251-261
    bend[:, 13] = bend[:, 16] = 0.5
    bend[:, 0:4] *= -1
    bend[:, 10] *= -1
    cyc[bad] = np.nan
    bend[bad] = np.nan
    return cyc, bend, bad.astype(np.uint8)
