"""Gait IK oracle vs the reference's own output (tests/golden/gait_golden.npz, scripts/make_gait_golden.py):
default parameters reproduce the shipped plen_bullet/trajectories/*_traj.npy goldens; jittered parameter sets match the
reference generator; unreachable targets are flagged where the reference raises ValueError."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden", "gait_golden.npz")


def test_gait_oracle_matches_reference_generator():
    from oracle import gait_oracle
    g = np.load(GOLD)
    cyc, bend, status = gait_oracle.gait(g["params"])
    assert (status == g["status"]).all() and status.sum() == 2
    ok = status == 0
    assert np.abs(cyc[ok] - g["cycle"][ok]).max() < 1e-12
    assert np.abs(bend[ok] - g["bend"][ok]).max() < 1e-12
    assert np.isnan(cyc[~ok]).all() and np.isnan(bend[~ok]).all()
    # env 0 = default gait = the shipped goldens (20 cycles of 40 rows; trajectory_eval.py:180)
    assert np.abs(np.tile(cyc[0], (20, 1)) - g["shipped_traj"]).max() < 1e-12
    assert np.abs(bend[0] - g["shipped_bend"]).max() < 1e-12
    # known-answer rows of SURVEY.md Appendix D
    assert abs(g["foot_walk_rfwd"][0][1] - 0.01055351) < 1e-8 and abs(g["foot_walk_rfwd"][0][8] + 0.66491819) < 1e-8
    assert abs(bend[0][2] - 0.72957981) < 1e-8 and abs(bend[0][9] - 1.15927948) < 1e-8
