"""The panel path's fast paths against the oracle: a window with ~50 points per (first camera, camera span) class, so
that most panels of twelve points hold ONE camera interval (strips taken whole, ba_panel.cu strip_full) and a panel
buffer is refilled with the interval it held before.  The BASELINE windows reach these paths only at C4 size, where
the oracle cannot follow a whole trajectory; here it can: accept sequence, per-iteration cost, poses, flags."""
import numpy as np
import pytest

import b200ba  # noqa: F401
from oracle import lm_oracle as O
from visual_slam_core_b200 import LMConfig, LocalBASolver, synth

from helpers import rot_angle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_cam,n_pt,seed", [(64, 30000, 5), (90, 24000, 6)])
def test_dense_classes_match_the_oracle(n_cam, n_pt, seed):
    p = synth.make_problem(None, n_cam=n_cam, n_pt=n_pt, k_mean=6.0, k_max=10, seed=seed)
    s = LocalBASolver(LMConfig(schur="panel", max_iter=5))
    r = s.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, p.pt_offsets)
    s.close()
    ref = O.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, O.LMParams(max_iter=5))
    rel = np.abs(r.cost_hist - ref.cost_hist) / ref.cost_hist
    print(f"\n[{n_cam} cams / {n_pt} pts / {p.n_obs} obs] solver {r.extra['solver_used']}: accepted {r.accepted.tolist()}, "
          f"max rel cost diff {rel.max():.2e}, rot {rot_angle(r.R, ref.R).max():.2e}, flags differing {int((r.flags != ref.flags).sum())}")
    assert np.array_equal(r.accepted, ref.accepted)
    assert rel.max() <= 1e-7
    assert rot_angle(r.R, ref.R).max() <= 1e-5
    assert int((r.flags != ref.flags).sum()) == 0
