"""Oracle loader and physics known answers."""
import os

import numpy as np
import pytest

REF = "/root/reference"


def test_tree_tables(oracle_lib):
    from oracle import urdf_tree
    t = urdf_tree.load_tree()
    assert t["n_links"] == 32 and abs(t["total_mass"] - 0.495834) < 1e-9      # plen_walk.py:350-351 "0.495Kg"
    assert [i for i in range(32) if t["jtype"][i] == 1] == sorted(urdf_tree.MOVING_JOINTS)   # plen_env.py:318-320
    assert t["link_names"][11] == "r_foot" and t["link_names"][19] == "l_foot"
    assert t["joint_names"][5] == "rb_servo_r_hip" and t["joint_names"][30] == "le_servo_l_elbow"
    for f in t["feet"]:
        assert f["n_sole_vertices"] == 32 and 0.0008 < f["breaking_threshold"] < 0.0012


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present (GPU box)")
def test_tree_regenerates_from_reference(oracle_lib):
    from oracle import urdf_tree
    t = urdf_tree.build_tree(os.path.join(REF, "plen_bullet/src/plen.urdf"), os.path.join(REF, "plen_ros/meshes_bin"))
    s = urdf_tree.load_tree()
    for k in ("parent", "jtype", "axis", "R_pj", "p_pj", "com", "mass", "inertia", "lower", "upper"):
        assert np.allclose(np.array(t[k], dtype=float), np.array(s[k], dtype=float), atol=1e-15), k


def test_zero_pose_fk_table(oracle_lib):
    """Joint origins at the zero pose with the base at (0,0,0.158) -- SURVEY.md Appendix B."""
    o = oracle_lib.PlenOracle(1)
    pos, rot = o.fk()
    expect = {5: (-0.0066, -0.0228, 0.119), 6: (-0.0046, -0.0173, 0.1062), 11: (-0.0018, -0.0161, 0.0189),
              14: (0.0104, 0.0173, 0.1064), 19: (-0.0007, 0.0162, 0.0189), 21: (-0.0072, -0.0506, 0.1835),
              30: (-0.0085, 0.0506, 0.1445)}
    for idx, p in expect.items():
        assert np.abs(pos[idx + 1] - np.array(p)).max() < 1.5e-4, idx
    # lowest sole points: left foot 2.85 mm below the ground, right 2.63 mm above
    t = o.tree
    low = []
    for f in t["feet"]:
        pts = np.array(f["points"])
        low.append((pos[f["link"] + 1] + pts @ rot[f["link"] + 1].T)[:, 2].min())
    assert abs(low[0] - 0.00263) < 1e-4 and abs(low[1] + 0.00285) < 1e-4


def test_inverse_mass_matrix_is_spd(oracle_lib):
    o = oracle_lib.PlenOracle(1)
    M = o.minv()
    assert np.abs(M - M.T).max() < 1e-9 * np.abs(M).max()
    assert np.linalg.eigvalsh((M + M.T) / 2).min() > 0
    # translational block: total mass 0.4958 kg => the base feels at least 1/m in every direction
    assert np.linalg.eigvalsh(np.linalg.inv(M))[-1] < 0.51


def test_free_flight_conserves_momentum(oracle_lib):
    """No gravity, servo torque 0: total linear momentum drifts only by the integrator's O(dt) truncation error --
    small at dt = 1/240 and shrinking with dt (ABA + integrator sanity)."""
    def drift(dt, ticks):
        o = oracle_lib.PlenOracle(1)
        o.cfg.gravity_z = 0.0
        o.cfg.motor_max_force = 0.0
        o.cfg.dt = dt
        rng = np.random.default_rng(0)
        st = o.get_state()
        st["qpos"][0, 2] = 1.0
        st["qpos"][0, 7:] = rng.uniform(-0.5, 0.5, 18)
        st["qvel"][0] = np.concatenate([rng.normal(size=6) * 0.2, rng.normal(size=18)])
        o.set_state(st)

        def lin_momentum():
            # generalized momentum M u; rows 3..5 (conjugate to the base linear velocity) = total linear momentum
            u = np.concatenate([o.states[0].omega[:], o.states[0].vel[:], o.states[0].qd[:]])
            return np.linalg.solve(o.minv(), u)[3:6]

        p0 = lin_momentum()
        for _ in range(ticks):
            o.tick()
        return np.abs(lin_momentum() - p0).max() / np.abs(p0).max()

    d1, d4 = drift(1.0 / 240.0, 50), drift(1.0 / 960.0, 200)
    assert d1 < 2e-4 and d4 < d1 / 3.0


def test_settle_height_and_standing(oracle_lib):
    """reset + stand with zero joint targets in the joint_act variant (rolling friction 0.01, linear damping 0.1,
    plen_env.py:439-481): the robot stands, ~4e-4 m below the RL-mode constant init_height = 0.160178937611 (plen_env.py:70).
    The 1e-4 comparison with that constant belongs to the RL-mode env: tests/test_physics_pins.py."""
    o = oracle_lib.PlenOracle(1, joint_act=True)
    o.reset()
    z = []
    for _ in range(60):
        ob, r, d, _ = o.step(np.zeros((1, 18)))
        assert not d[0]
        z.append(ob[0, 18])
    assert abs(np.mean(z[20:]) - 0.160178937611) < 1e-3


def test_loader_handles_the_box_foot_variant():
    """plen_new.urdf (box feet, +-1.0 rad limits): same 33-link tree and mass, sole = bottom face of the
    foot box with no collision margin.  Regenerated from the reference checkout when it is present, else the packaged copy."""
    import os
    import numpy as np
    from plen_ml_walk_b200.urdf_loader import load_plen_model, packaged_model
    new, old = packaged_model("plen_new"), packaged_model("plen")
    assert abs(new.total_mass - old.total_mass) < 1e-9 and new.foot_margin == 0.0 and old.foot_margin == 1e-3
    assert np.allclose(new.upper[6:24], 1.0) and np.allclose(old.upper[6:24], 1.7)
    for f in range(2):
        q = new.foot_pts[f]
        assert np.allclose(q[:, 2], q[0, 2])                                   # a flat rectangle ...
        assert np.allclose(q[0] + q[2], q[1] + q[3], atol=1e-9)                # ... (diagonals share their midpoint)
    ref = "/root/reference"
    if os.path.exists(os.path.join(ref, "plen_bullet/src/plen_new.urdf")):
        m = load_plen_model(os.path.join(ref, "plen_bullet/src/plen_new.urdf"), os.path.join(ref, "plen_ros/meshes_bin"))
        assert np.allclose(m.foot_pts, new.foot_pts) and np.allclose(m.inertia, new.inertia)


def test_persistent_manifold_regression_fixture():
    """tests/golden/manifold_golden.npz (scripts/make_manifold_golden.py): the oracle's manifold_mode = 1 with support_tie = 1e-7
    replays its own committed trajectory -- observations, rewards, done flags AND the manifolds (cached points, point counts) after
    every one of 80 steps of a swaying and a falling robot, resets included.  A regression pin of the restated procedure (it was
    produced by this oracle, not by PyBullet): the kernels' sole_manifold option is judged against this oracle."""
    import os
    from oracle.oracle import PlenOracle
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "manifold_golden.npz"))
    n = g["obs0"].shape[0]
    o = PlenOracle(n)
    o.cfg.manifold_mode = 1
    o.cfg.support_tie = 1e-7
    assert np.abs(o.reset() - g["obs0"]).max() < 1e-12
    assert np.abs(o.get_manifold() - g["man0"]).max() < 1e-12
    counts = set()
    for t in range(g["actions"].shape[0]):
        ob, r, d, _ = o.step(g["actions"][t])
        assert (d == g["done"][t]).all(), "done flags differ at step %d" % t
        assert np.abs(ob - g["obs"][t]).max() < 1e-9, "observation differs at step %d" % t
        assert np.abs(np.nan_to_num(r, nan=-1e9) - g["reward"][t]).max() < 1e-8
        m = o.get_manifold()
        assert (m[:, 48:50] == g["manifold"][t][:, 48:50]).all() and np.abs(m - g["manifold"][t]).max() < 1e-9
        counts |= set(m[:, 48:50].ravel().astype(int))
        for e in np.where(d)[0]:
            o.reset_one(int(e))
    assert counts == {0, 1, 2, 3, 4} and g["done"].sum() >= 3      # every cache size and a few resets occur in the fixture
