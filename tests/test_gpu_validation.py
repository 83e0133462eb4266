"""The host-pointer API refuses malformed index arrays (the kernels index with them unchecked).  The check runs on the
device once the arrays are there (k_validate, ba_local.cu) and the call returns BA_ERR_INVALID before any solver kernel
is launched; the inputs stay untouched."""
import numpy as np
import pytest

import b200ba  # noqa: F401
from visual_slam_core_b200 import LMConfig, LocalBASolver, synth
from visual_slam_core_b200._lib import B200BAError

pytestmark = pytest.mark.gpu


def _solve(s, p, **over):
    a = dict(R=p.R.copy(), t=p.t.copy(), fixed=p.fixed, X=p.X.copy(), obs_cam=p.obs_cam, obs_pt=p.obs_pt, obs_uv=p.obs_uv,
             pt_offsets=p.pt_offsets)
    a.update(over)
    return s.solve_local(p.intr, a["R"], a["t"], a["fixed"], a["X"], a["obs_cam"], a["obs_pt"], a["obs_uv"], a["pt_offsets"])


def test_malformed_indices_are_refused_and_the_solver_still_works():
    p = synth.make_problem("T1", seed=1)
    s = LocalBASolver(LMConfig(max_iter=3))
    cam = p.obs_cam.copy(); cam[len(cam) // 2] = p.n_cam                 # camera index out of range
    with pytest.raises(B200BAError, match="obs_cam"):
        _solve(s, p, obs_cam=cam)
    cam[len(cam) // 2] = -1
    with pytest.raises(B200BAError, match="obs_cam"):
        _solve(s, p, obs_cam=cam)
    off = p.pt_offsets.copy(); off[5], off[6] = off[6], off[5] - 1       # offsets fall
    with pytest.raises(B200BAError, match="pt_offsets"):
        _solve(s, p, pt_offsets=off)
    off = p.pt_offsets.copy(); off[-1] += 1                              # do not end at n_obs
    with pytest.raises(B200BAError, match="pt_offsets"):
        _solve(s, p, pt_offsets=off)
    good = _solve(s, p)                                                  # the handle is still usable
    assert good.cost_hist[-1] < good.cost_hist[0]
    s.close()
