"""GPU: plen_gait_ik (through plen_ml_walk_b200.gait) vs the reference generator's goldens at 1e-6 rad, and the
config-3 replay (bend + open-loop gait) against the oracle on a small batch."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "gait_golden.npz")


def test_gait_ik_matches_reference_goldens():
    from plen_ml_walk_b200.gait import TrajectoryGenerator, gait_trajectories
    g = np.load(GOLD)
    cyc, bend, status = (t.cpu().numpy() for t in gait_trajectories(g["params"]))
    assert (status == g["status"]).all()
    ok = status == 0
    assert np.abs(cyc[ok] - g["cycle"][ok]).max() < 1e-6            # north_star: IK within 1e-6 rad of the numpy output
    assert np.abs(bend[ok] - g["bend"][ok]).max() < 1e-6
    assert np.isnan(cyc[~ok]).all()                                  # reference raises ValueError there
    gen = TrajectoryGenerator().main()                               # defaults = the shipped *_traj.npy goldens
    assert np.abs(gen.full_trajectory(20)[0].cpu().numpy() - g["shipped_traj"]).max() < 1e-6
    assert np.abs(gen.bend_legs[0].cpu().numpy() - g["shipped_bend"]).max() < 1e-6
    assert np.abs(gen.foot_walk_rfwd[0].cpu().numpy() - g["foot_walk_rfwd"]).max() < 1e-6
    assert np.abs(gen.foot_walk_lfwd[0].cpu().numpy() - g["foot_walk_lfwd"]).max() < 1e-6
    assert np.abs(gen.bend[0].cpu().numpy() - g["bend_rows"]).max() < 1e-6


def test_gait_ik_large_batch_is_consistent():
    """65,536 parameter sets (config 3 size): every env with the default parameters reproduces env 0 bit for bit."""
    from plen_ml_walk_b200.gait import gait_trajectories
    n = 65536
    p = np.tile(np.array([30.0, 30.0, 10.0, 5.0, 10.0]), (n, 1))
    p[1::2, 1] = 33.0
    cyc, bend, status = gait_trajectories(p)
    assert int(status.sum().item()) == 0
    assert bool((cyc[0::2] == cyc[0]).all().item()) and bool((cyc[1::2] == cyc[1]).all().item())
    assert not bool((cyc[0] == cyc[1]).all().item())


def test_open_loop_gait_replay_vs_oracle(oracle_lib):
    """trajectory_eval.py replay with joint_act=True (rolling friction 0.01, linear damping 0.1): 20 bend steps + 120 gait
    steps, teacher-forced from the oracle every step; medians within 1e-4."""
    from plen_ml_walk_b200.gait import TrajectoryGenerator
    from plen_ml_walk_b200.vec_env import PlenVecEnv
    from parity_util import abi_from_oracle
    n = 16
    rng = np.random.default_rng(0)
    p = np.tile(np.array([30.0, 30.0, 10.0, 5.0, 10.0]), (n, 1))
    p[1:, [0, 1, 3, 4]] *= rng.uniform(0.9, 1.1, (n - 1, 4))
    gen = TrajectoryGenerator(height=p[:, 0], stride=p[:, 1], bend_distance=p[:, 2], body_sway=p[:, 3], fwd_bias=p[:, 4]).main()
    bend, cyc = gen.bend_legs.cpu().numpy(), gen.cycle.cpu().numpy()
    o = oracle_lib.PlenOracle(n, joint_act=True, n_threads=8)
    env = PlenVecEnv(n, joint_act=True, auto_reset=False)
    oo, go = o.reset(), env.reset().cpu().numpy()
    assert np.abs(oo - go).max() < 1e-5
    errs, flags = [], 0
    for t in range(140):
        act = bend if t < 20 else cyc[:, (t - 20) % 40]
        env.set_state(*abi_from_oracle(o.get_state()))
        oo, orw, od, _ = o.step(act)
        go, grw, gd, _ = env.step(torch.from_numpy(act.astype(np.float32)).cuda())
        go = go.cpu().numpy()
        errs.append(np.abs(go[:, :24] - oo[:, :24]).max(1))
        flags += int((go[:, 24:] != oo[:, 24:]).sum())
        for e in np.where(od)[0]:
            o.reset_one(int(e))
    errs = np.concatenate(errs)
    print("gait replay: obs err median %.2e p90 %.2e max %.2e, contact-flag mismatches %d / %d" %
          (np.median(errs), np.quantile(errs, .9), errs.max(), flags, 2 * n * 140))
    assert np.median(errs) < 1e-4 and flags <= 0.03 * 2 * n * 140


def test_config3_free_running_gait_statistics_vs_oracle(oracle_lib):
    """BASELINE config 3 on a 256-env subset, FREE-RUNNING (no teacher forcing): 20 bend + 800 gait steps of the
    jittered sinewave gaits; fall rate and final torso_x statistics of the CUDA path vs the float64 oracle."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("teb", os.path.join(os.path.dirname(__file__), "..", "scripts", "trajectory_eval_batched.py"))
    teb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(teb)
    n = 256
    res, gen, env = teb.run(n, seed=0)
    bend, cyc = gen.bend_legs.cpu().numpy(), gen.cycle.cpu().numpy()
    o = oracle_lib.PlenOracle(n, joint_act=True, n_threads=16)
    o.reset()
    fell = np.zeros(n, dtype=bool)
    for t in range(820):
        act = bend if t < 20 else cyc[:, (t - 20) % 40]
        _, _, d, tmo = o.step(act)
        fell |= d & ~tmo
    x = o.get_state()["qpos"][:, 0]
    print("config 3 (256 envs): GPU fall %.3f x %.4f +- %.4f | oracle fall %.3f x %.4f +- %.4f" %
          (res["fall_rate"], res["torso_x_mean"], res["torso_x_std"], fell.mean(), x.mean(), x.std()))
    assert abs(res["fall_rate"] - fell.mean()) <= 0.02
    assert abs(res["torso_x_mean"] - x.mean()) < 0.02 and abs(res["torso_x_std"] - x.std()) < 0.02
