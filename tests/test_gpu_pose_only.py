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


# ---------------------------------------------------------------------------------------- no prior
def _frame(seed, n=400, outliers=0.15):
    rng = np.random.default_rng(seed)
    fx = fy = 700.0
    K = np.array([[fx, 0, 640.0], [0, fy, 360.0], [0, 0, 1.0]])
    w = rng.normal(0, 0.4, 3)
    th = np.linalg.norm(w)
    Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    R = np.eye(3) + np.sin(th) / th * Kx + (1 - np.cos(th)) / th ** 2 * Kx @ Kx
    Xc = np.stack([rng.uniform(-4, 4, n), rng.uniform(-2, 2, n), rng.uniform(5, 25, n)], axis=1)
    t = rng.normal(0, 1.0, 3)
    Xw = (Xc - t) @ R                                  # Xc = R Xw + t
    uv = Xc[:, :2] / Xc[:, 2:] * fx + np.array([640.0, 360.0]) + rng.normal(0, 0.5, (n, 2))
    out = rng.random(n) < outliers
    uv[out] += rng.normal(0, 40.0, (int(out.sum()), 2))
    return K, R, t, Xw.astype(np.float32), uv.astype(np.float32), out


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_p3p_ransac_bootstrap_finds_the_pose_without_a_prior(seed):
    K, R, t, Xw, uv, out = _frame(seed)
    s = LocalBASolver(LMConfig())
    got = s.pose_ransac([700.0, 700.0, 640.0, 360.0], Xw, uv, n_hypotheses=100, threshold_px=3.0, seed=42)
    s.close()
    assert got is not None
    ang = rot_angle(got["R"][None], R[None])[0]
    print(f"\n[p3p seed {seed}] inliers {got['n_inliers']} of {int((~out).sum())} clean, rot err {ang:.2e}, |dt| {np.abs(got['t'] - t).max():.2e}")
    assert got["n_inliers"] >= 0.8 * int((~out).sum())
    assert ang < 2e-2 and np.abs(got["t"] - t).max() < 0.2
    assert abs(np.linalg.det(got["R"]) - 1.0) < 1e-9


def test_call_site_without_prior_matches_the_references_opencv_result():
    """`refine_motion_from_2d3d(pts3d, pts2d, K, dist)` exactly as `Tracking._optimize_pose_mono` calls the function it
    replaces (tracking.py:350-356: no initial guess), on the frame the reference's own `estimate_motion_from_2d3d` was
    run on (fixture ref_pose_only.npz): same pose within the noise of the two estimators, same inliers but for
    threshold stragglers, same return types."""
    import os
    from helpers import GOLDEN
    from visual_slam_core_b200 import motion
    g = np.load(os.path.join(GOLDEN, "ref_pose_only.npz"))
    R, t, ok, inl = motion.refine_motion_from_2d3d(g["pts3d"], g["pts2d"], g["K"], np.zeros(5), True)
    assert ok is True and R.shape == (3, 3) and t.shape == (3, 1) and inl.dtype == np.int32 and inl.shape[1] == 1
    ang = rot_angle(R[None], g["ref_R"][None])[0]
    both = np.intersect1d(inl.ravel(), g["ref_inliers"].ravel()).size
    print(f"\n[call site] rot diff to cv2 {ang:.2e}, |dt| {np.abs(t - g['ref_t']).max():.2e}, inliers {inl.shape[0]} vs {g['ref_inliers'].shape[0]} (common {both})")
    assert ang < 3e-3 and np.abs(t - g["ref_t"]).max() < 3e-2
    assert both >= 0.95 * g["ref_inliers"].shape[0]
    # lens distortion at the call site: undistorted first, never an exception
    import cv2
    dist = np.array([0.1, -0.05, 0.001, 0.001, 0.0])
    Xc = g["pts3d"].astype(np.float64) @ g["R_true"].T + g["t_true"]
    uvd, _ = cv2.projectPoints(Xc, np.zeros(3), np.zeros(3), g["K"], dist)
    R2, t2, ok2, _ = motion.refine_motion_from_2d3d(g["pts3d"], uvd.reshape(-1, 2).astype(np.float32), g["K"], dist, True)
    assert ok2 is True and rot_angle(R2[None], g["R_true"][None])[0] < 2e-3 and np.abs(t2.ravel() - g["t_true"]).max() < 2e-2
