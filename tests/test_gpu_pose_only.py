"""Batched pose-only LM (BASELINE config 3) against the oracle."""
import numpy as np
import pytest

from oracle import lm_oracle as O
from visual_slam_core_b200 import LMConfig, LocalBASolver, synth

from helpers import rot_angle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nf,nm,ragged,seed", [(5, 64, False, 1), (17, 333, True, 2), (64, 1000, False, 3)])
def test_pose_only_matches_oracle(nf, nm, ragged, seed):
    pp = synth.make_pose_only(n_frames=nf, n_match=nm, seed=seed, ragged=ragged)
    s = LocalBASolver(LMConfig())
    out = s.solve_pose_only(pp.intr, pp.R, pp.t, pp.X, pp.uv, pp.frame_offsets)
    s.close()
    ref = O.solve_pose_only(pp.intr, pp.R, pp.t, pp.X, pp.uv, pp.frame_offsets)
    np.testing.assert_array_equal(out["accepted"], ref["accepted"])
    np.testing.assert_allclose(out["cost_hist"], ref["cost_hist"], rtol=1e-5)
    assert rot_angle(out["R"], ref["R"]).max() < 1e-5
    assert np.abs(out["t"] - ref["t"]).max() < 1e-5
    np.testing.assert_array_equal(out["flags"], ref["flags"])


def test_pose_only_full_size_properties():
    pp = synth.make_pose_only(n_frames=256, n_match=1000, seed=0)
    s = LocalBASolver(LMConfig())
    out = s.solve_pose_only(pp.intr, pp.R, pp.t, pp.X, pp.uv, pp.frame_offsets)
    s.close()
    assert (np.diff(out["cost_hist"], axis=1) <= 0).all()
    assert rot_angle(out["R"], pp.R_true).max() < 5e-3
    assert np.abs(out["t"] - pp.t_true).max() < 0.2
    assert (out["flags"][pp.outlier] == 1).mean() > 0.9
    # empty batch and a frame with very few matches are accepted
    e = s2 = LocalBASolver(LMConfig())
    r = s2.solve_pose_only(pp.intr, np.zeros((0, 3, 3)), np.zeros((0, 3)), np.zeros((0, 3), np.float32),
                           np.zeros((0, 2), np.float32), np.zeros(1, np.int32))
    assert r["R"].shape == (0, 3, 3)
    s2.close()


def test_pose_only_inplace_with_caller_buffers_equals_the_copying_call():
    """`inplace=True` refines the caller's R / t and fills the caller's flags / cost_hist / accepted (the
    pinned-buffer path bench.py times end to end); results are those of the default call, bit for bit."""
    pp = synth.make_pose_only(n_frames=9, n_match=200, seed=4, ragged=True)
    s = LocalBASolver(LMConfig())
    a = s.solve_pose_only(pp.intr, pp.R, pp.t, pp.X, pp.uv, pp.frame_offsets)
    R, t = pp.R.copy(), pp.t.copy()
    nM, nF, mi = pp.X.shape[0], pp.R.shape[0], LMConfig().max_iter
    flags = np.full(nM, 7, np.uint8); hist = np.zeros((nF, mi + 1)); acc = np.zeros((nF, mi), np.uint8)
    b = s.solve_pose_only(pp.intr, R, t, pp.X, pp.uv, pp.frame_offsets, inplace=True, flags=flags,
                          cost_hist=hist, accepted=acc)
    assert b["R"].base is R or b["R"] is R
    np.testing.assert_array_equal(R, a["R"]); np.testing.assert_array_equal(t, a["t"])
    np.testing.assert_array_equal(flags, a["flags"]); np.testing.assert_array_equal(hist, a["cost_hist"])
    np.testing.assert_array_equal(acc, a["accepted"])
    with pytest.raises(ValueError):
        s.solve_pose_only(pp.intr, R.astype(np.float32), t, pp.X, pp.uv, pp.frame_offsets, inplace=True)
    with pytest.raises(ValueError):
        s.solve_pose_only(pp.intr, R, t, pp.X, pp.uv, pp.frame_offsets, flags=np.zeros(3, np.uint8))
    s.close()
