"""CPU tests of the PRODUCT device code (plen_device.cuh + plen_solve.cuh + plen_env.cuh) through the warp-emulation
harness (tests/emu, test-only: 32 host threads + barriers stand in for one warp) against the float64 oracle.

The emulator runs (k_dyn, k_solve) x ticks exactly as plen_b200.cu launches them -- one emulated warp per robot for the
dynamics, one emulated warp per 4 robots (8 lanes each) for the projected Gauss-Seidel -- so the lane mappings, the
solve-record layout and the 4-robots-per-warp freezing logic are all exercised without a GPU.
"""
import numpy as np
import pytest

from emu_util import Emu, oracle_state_from_record, record_from_oracle_state


def _forced_ticks(oracle_lib, emu, n, ticks, seed, prep=None, warm_steps=6):
    o = oracle_lib.PlenOracle(n)
    o.reset()
    rng = np.random.default_rng(seed)
    for _ in range(warm_steps):
        o.step(rng.uniform(-1, 1, (n, 18)))
    if prep is not None:
        st = o.get_state()
        prep(st, rng)
        o.set_state(st)
    qd_err, q_err, it_pairs, rows = [], [], [], []
    for _ in range(ticks):
        st = o.get_state()
        rec = np.stack([record_from_oracle_state(st, e, dtype=emu.real) for e in range(n)])
        tg = rng.uniform(-1, 1, (n, 18))
        for e in range(n):
            for k in range(18):
                o.states[e].target[k] = tg[e, k]
        emu.tick(rec, tg, 1)
        for e in range(n):
            o.tick(e)
        got, ref = oracle_state_from_record(rec), o.get_state()
        qd_err.append(np.abs(got["qvel"] - ref["qvel"]).max(1))
        q_err.append(np.abs(got["qpos"] - ref["qpos"]).max(1))
        it_pairs.append((rec[:, 79].astype(int), np.array([o.states[e].last_iterations for e in range(n)])))
        rows.append([o.states[e].last_rows for e in range(n)])
    return np.array(qd_err), np.array(q_err), it_pairs, np.array(rows)


def test_f64_emulation_matches_oracle_through_contacts(oracle_lib):
    """Same algorithm check: the device source evaluated in float64 (model tables stay float32) follows the oracle to
    ~1e-5 rad/s per tick through contact phases, with identical PGS iteration counts; the rare impact ticks where the
    50-iteration Gauss-Seidel amplifies round-off by > 1e7 ('PGS sensitivity') are excluded by quantile."""
    emu = Emu(double=True)
    qd, q, its, rows = _forced_ticks(oracle_lib, emu, n=6, ticks=6, seed=0)   # n not a multiple of 4: padded solver warp
    assert rows.max() > 18                       # contact rows were exercised
    assert np.median(qd) < 2e-5 and np.quantile(qd, 0.8) < 1e-3
    assert np.median(q) < 1e-6
    same = np.mean([np.mean(a == b) for a, b in its])
    assert same > 0.9


def test_f64_emulation_joint_limit_rows(oracle_lib):
    """Joints pushed beyond the +-1.7 rad URDF limits: the rare limit-row path of the solver (other robots of the same
    emulated warp have no limit rows)."""
    def prep(st, rng):
        st["qpos"][0, 7 + 3] = 1.75
        st["qpos"][0, 7 + 17] = -1.73
        st["qpos"][1, 7 + 0] = -1.8
        st["qpos"][3, 7 + 8] = 1.71
        st["qpos"][3, 7 + 9] = 1.9
        st["qpos"][3, 7 + 16] = 1.72
        st["qpos"][:, 2] += 0.5          # airborne: limits + servos only, no chaotic impacts
        st["lam_n"][:] = 0
        st["in_manifold"][:] = 0

    emu = Emu(double=True)
    qd, q, its, rows = _forced_ticks(oracle_lib, emu, n=5, ticks=2, seed=5, prep=prep, warm_steps=3)
    assert rows[0, 0] == 20 and rows[0, 1] == 19 and rows[0, 3] == 21 and rows[0, 2] == 18
    assert qd.max() < 2e-4 and q.max() < 1e-6


def test_f32_emulation_free_flight_one_step(oracle_lib):
    """north_star tolerance: joint angles / base pose within 1e-4 rad / 1e-4 m after one env step in free flight."""
    from parity_util import random_flight_state
    emu = Emu(double=False)
    emu.cfg.auto_reset = 0
    n = 4
    rng = np.random.default_rng(1)
    o = oracle_lib.PlenOracle(n)
    st = o.get_state()
    st["qpos"], st["qvel"] = random_flight_state(rng, n)
    o.set_state(st)
    st = o.get_state()
    rec = np.stack([record_from_oracle_state(st, e, dtype=np.float32) for e in range(n)])
    act = rng.uniform(-1, 1, (n, 18)).astype(np.float32)
    obs, rew, done, tmo, _ = emu.step(rec, act)
    oo, orw, od, _ = o.step(act.astype(np.float64))
    got, ref = oracle_state_from_record(rec), o.get_state()
    assert np.abs(got["qpos"][:, 7:] - ref["qpos"][:, 7:]).max() < 1e-4
    assert np.abs(got["qpos"][:, :7] - ref["qpos"][:, :7]).max() < 1e-4
    assert np.abs(obs[:, :24] - oo[:, :24]).max() < 1e-4
    assert (done == od).all()


def _fallen_states(o, rng, depth=0.002):
    """Random tilted poses lowered until the lowest collider point (box vertex or sole corner) is `depth` under the ground."""
    n = o.n
    st = o.get_state()
    st["qpos"][:, 7:] = rng.uniform(-0.7, 0.7, (n, 18))
    ang = rng.uniform(-0.3, 0.3, (n, 3))
    ang[:, 2] = rng.uniform(-3, 3, n)
    lying = rng.integers(0, 2, n)                      # lying on the front / back (pitch) or on a side (roll)
    ang[np.arange(n), lying] = rng.choice([-1, 1], n) * rng.uniform(1.2, 1.9, n)
    for e in range(n):
        cr, sr, cp, sp, cy, sy = (f(a / 2) for a in ang[e] for f in (np.cos, np.sin))
        st["qpos"][e, 3:7] = [sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy]
    st["qpos"][:, 2] = 0.5
    st["qvel"][:, :6] = rng.normal(size=(n, 6)) * 0.1
    st["qvel"][:, 6:] = rng.normal(size=(n, 18)) * 0.5
    st["lam_n"][:] = 0
    st["in_manifold"][:] = 0
    o.set_state(st)
    for e in range(n):
        pos, rot = o.fk(e)
        low = np.inf
        for b in o.tree["boxes"]:
            R, p = rot[b["link"] + 1], pos[b["link"] + 1]
            Rb, c, h = R @ np.array(b["rot"]), p + R @ np.array(b["center"]), np.array(b["half"])
            low = min(low, c[2] - np.abs(Rb[2]) @ h)
        lowf = min((pos[f["link"] + 1] + np.array(f["points"]) @ rot[f["link"] + 1].T)[:, 2].min() - 0.001 for f in o.tree["feet"])
        st["qpos"][e, 2] -= max(low, lowf - 0.004) + depth      # a box on the ground, the soles at most ~6 mm under it
    o.set_state(st)
    return st


def test_f64_emulation_box_contacts_match_oracle(oracle_lib):
    """knees / hands / torso boxes against the ground.  The device source in float64 follows the oracle (same
    4-deepest-points cap) through one tick from fallen poses; robots of the same solver warp without box contacts are
    unaffected by sharing it."""
    emu = Emu(double=True)
    n = 11
    o = oracle_lib.PlenOracle(n)
    o.reset()
    rng = np.random.default_rng(3)
    st = _fallen_states(o, rng)
    st["qpos"][n - 1, 2] += 0.3                       # one robot of the second warp stays airborne
    st["qpos"][n - 2, 3:7] = [0, -np.sqrt(0.5), 0, np.sqrt(0.5)]      # and one lies flat on its back, joints at zero:
    st["qpos"][n - 2, 7:] = 0                                          # many boxes touch, the 4-point cap binds
    st["qpos"][n - 2, 2] = 0.0235
    o.set_state(st)
    worst_qd, worst_q, pts = [], [], []
    for tick in range(3):
        st = o.get_state()
        rec = np.stack([record_from_oracle_state(st, e, dtype=emu.real) for e in range(n)])
        tg = rng.uniform(-1, 1, (n, 18))
        for e in range(n):
            for k in range(18):
                o.states[e].target[k] = tg[e, k]
        emu.tick(rec, tg, 1)
        for e in range(n):
            o.tick(e)
        got, ref = oracle_state_from_record(rec), o.get_state()
        worst_qd.append(np.abs(got["qvel"] - ref["qvel"]).max(1))
        worst_q.append(np.abs(got["qpos"] - ref["qpos"]).max(1))
        pts.append([o.states[e].last_box_points for e in range(n)])
    pts, worst_qd, worst_q = np.array(pts), np.array(worst_qd), np.array(worst_q)
    assert (pts[0, :n - 1] > 0).mean() > 0.5 and pts[0, n - 1] == 0 and pts[0, n - 2] == 4     # box rows exercised, incl. the cap
    assert np.median(worst_qd) < 2e-5 and np.quantile(worst_qd, 0.8) < 1e-3, (np.median(worst_qd), np.quantile(worst_qd, 0.8))
    assert np.median(worst_q) < 1e-6


def test_f32_emulation_box_contacts_keep_the_robot_above_the_floor(oracle_lib):
    """float32 build: a robot dropped on its hands and knees is held by the box contacts (round 1 let it sink through
    the floor until z < 0.08 tripped) and tracks the oracle's torso height."""
    emu = Emu(double=False)
    n = 8
    o = oracle_lib.PlenOracle(n)
    o.reset()
    rng = np.random.default_rng(4)
    st = _fallen_states(o, rng, depth=0.0005)
    st["qvel"][:] = 0
    o.set_state(st)
    st = o.get_state()
    rec = np.stack([record_from_oracle_state(st, e, dtype=np.float32) for e in range(n)])
    tg = np.array(st["qpos"][:, 7:])
    for e in range(n):
        for k in range(18):
            o.states[e].target[k] = tg[e, k]
    z0 = st["qpos"][:, 2].copy()
    for _ in range(24):
        emu.tick(rec, tg, 1)
        for e in range(n):
            o.tick(e)
    got, ref = oracle_state_from_record(rec), o.get_state()
    # 0.1 s of free fall would be 49 mm; resting on its colliders a robot settles / tips by far less
    assert (z0 - got["qpos"][:, 2]).max() < 0.03 and (z0 - ref["qpos"][:, 2]).max() < 0.03
    assert np.median(np.abs(got["qpos"][:, 2] - ref["qpos"][:, 2])) < 2e-3


def test_f64_emulation_persistent_sole_manifold_matches_oracle(oracle_lib):
    """plen_config.sole_manifold = 1 against the oracle's manifold_mode = 1 (profiles/r2_physics_pin.md section 5): from the
    reset pose, random actions, state AND manifolds teacher-forced from the oracle before every tick.  The float64 build of the
    device source must build the same manifolds (same support vertex, same merge / reduction / refresh decisions: point counts
    equal, cached points within 2e-8 m) and follow the oracle's velocities like the four-corner path does."""
    emu = Emu(double=True)
    emu.cfg.sole_manifold = 1
    n = 6
    o = oracle_lib.PlenOracle(n)
    o.cfg.manifold_mode = 1
    o.reset()
    rng = np.random.default_rng(9)
    worst_qd, counts, man_err, n_ticks, max_count = [], [], [], 160, 0
    amp = np.where(np.arange(n) % 2 == 0, 0.04, 0.3)[:, None]      # half of the robots only sway (their manifolds fill up and are
    tg = rng.uniform(-1, 1, (n, 18)) * amp                           # reduced), the others fall over
    for tick in range(n_ticks):
        if tick % 4 == 0:
            tg = np.clip(tg + rng.normal(0, 1.0, (n, 18)) * amp, -1.2, 1.2)
        st = o.get_state()
        rec = np.stack([record_from_oracle_state(st, e, dtype=emu.real) for e in range(n)])
        man = o.get_manifold().astype(emu.real)
        emu.set_manifold(man)
        for e in range(n):
            for k in range(18):
                o.states[e].target[k] = tg[e, k]
        emu.tick(rec, tg, 1)
        for e in range(n):
            o.tick(e)
        got, ref = oracle_state_from_record(rec), o.get_state()
        mref = o.get_manifold()
        max_count = max(max_count, int(mref[:, 48:50].max()))
        worst_qd.append(np.abs(got["qvel"] - ref["qvel"]).max(1))
        counts.append((man[:, 48:50] == mref[:, 48:50]).all(1))
        same = (man[:, 48:50] == mref[:, 48:50]).all(1)
        for e in np.where(same)[0]:
            for f in range(2):
                k = int(mref[e, 48 + f])
                man_err.append(np.abs(man[e, 24 * f:24 * f + 3 * k] - mref[e, 24 * f:24 * f + 3 * k]).max() if k else 0.0)
        assert (got["in_manifold"][same] == ref["in_manifold"][same]).all()
    emu.set_manifold(None)
    worst_qd, counts = np.array(worst_qd), np.array(counts)
    seen = np.array([o.states[e].man_n[f] for e in range(n) for f in range(2)])
    print("manifold: counts equal on %.3f of the robot-ticks, cached points max err %.2e, qd err median %.2e p90 %.2e, final counts %s, max %d"
          % (counts.mean(), max(man_err), np.median(worst_qd), np.quantile(worst_qd, 0.9), seen, max_count))
    assert max_count == 4                            # the cache filled up: merge, reduction and refresh all ran
    assert counts.mean() > 0.97                      # a support vertex chosen among near-ties may differ at the 1e-16 level
    assert max(man_err) < 2e-8                       # the hull vertices reach the device code as float32 (C ABI struct)
    assert np.median(worst_qd) < 2e-5 and np.quantile(worst_qd, 0.8) < 1e-3


def test_persistent_sole_manifold_reset_and_support_tie(oracle_lib):
    """The reset of sole_manifold = 1 (8 free-running settle ticks from the start pose, manifolds empty).  In the start pose the
    feet are flat: every sole vertex is equally low and the support-vertex search is an exact tie.  With the tie tolerance
    (plen_config.support_tie = oracle support_tie = 1e-7 m: the first vertex of the list among those within it of the lowest)
    the float32 build of the device source reproduces the oracle's reset like the float64 build does; with the plain first-minimum
    search (0) rounding decides, and float32 lands 7e-3 away from float64 after the 8 ticks."""
    for tie, bound32 in ((1e-7, 1e-5), (0.0, None)):
        o = oracle_lib.PlenOracle(1)
        o.cfg.manifold_mode = 1
        o.cfg.support_tie = tie
        ref = o.reset()
        err = {}
        for dbl in (True, False):
            emu = Emu(double=dbl)
            emu.cfg.sole_manifold = 1
            emu.cfg.support_tie = tie
            rec = emu.init_record(1)
            man = np.zeros((1, 52), dtype=emu.real)
            emu.set_manifold(man)
            emu.tick(rec, np.zeros((1, 18)), int(emu.cfg.reset_ticks))
            emu.set_manifold(None)
            err[dbl] = np.abs(emu.observe(rec) - ref).max()
            if dbl or bound32:
                assert (man[0, 48:50] == o.get_manifold()[0, 48:50]).all()
            assert 0 < man[0, 48:50].sum() <= 8
        print("manifold reset, support_tie %g: obs err f64 %.2e, f32 %.2e" % (tie, err[True], err[False]))
        assert err[True] < 1e-7
        assert err[False] < (bound32 or 2e-2)
        if not bound32:
            assert err[False] > 1e-4      # the ill-conditioning the tolerance removes (if this fails, drop the tolerance)
