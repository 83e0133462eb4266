"""SURVEY 8f 'next' rows on the GPU: two-view initial refinement (N2), pose-only refinement behind
the tracking call site (N3), the local-mapping call site with the packer fast path and outlier
culling (A8, N1, N4)."""
import os

import numpy as np
import pytest

from oracle import lm_oracle as O
import refgraph
from visual_slam_core_b200 import B200Optimizer, LMConfig, LocalBASolver, handlers, motion, packing, synth

from helpers import GOLDEN, graph_from_golden, rot_angle, window_points

pytestmark = pytest.mark.gpu


def _two_view_graph(seed=5, n_pt=400):
    p = synth.make_problem(n_cam=2, n_pt=n_pt, k_mean=2.0, k_max=2, seed=seed)
    cam, kfs, mps = refgraph.graph_from_problem(p)
    # a point seen by the reference view only and an out-of-range keypoint index: both skipped (:240-258)
    lonely = refgraph.MapPoint([0.0, 0.0, 9.0])
    kfs[0].keypoints[0].append(refgraph.KeyPoint(640.0, 360.0))
    kfs[0].add_map_point(0, len(kfs[0].keypoints[0]) - 1, lonely)
    broken = refgraph.MapPoint([0.5, 0.0, 9.0])
    broken.add_observation(kfs[0].keyframe_id, 0, 0)
    broken.add_observation(kfs[1].keyframe_id, 0, 10_000_000)
    return p, kfs, [lonely] + mps + [broken]


def test_optimize_initial_matches_the_oracle_with_smooth_l1():
    p, kfs, mps = _two_view_graph()
    pi = packing.pack_initial(kfs, mps)
    assert pi.selected[0] == 1 and len(pi.selected) == len(mps) - 2
    pw = pi.window
    ref = O.solve_local(pw.intr, pw.R, pw.t, pw.fixed, pw.X, pw.obs_cam, pw.obs_pt, pw.obs_uv,
                        O.LMParams(huber_delta=1.0))        # SmoothL1(beta = 1), torch_optimizer.py:282
    untouched = [mps[0].position.copy(), mps[-1].position.copy(), kfs[0].R_w2c.copy(), kfs[0].t_w2c.copy()]
    opt = B200Optimizer(log_dir=None)
    assert opt.optimize_initial(kfs, mps) is True
    res = opt.last_result
    np.testing.assert_array_equal(res.accepted, ref.accepted)
    np.testing.assert_allclose(res.cost_hist, ref.cost_hist, rtol=1e-5)
    assert res.cost_hist[-1] < res.cost_hist[0]
    assert rot_angle(kfs[1].R_w2c[None], ref.R[1][None]).max() < 1e-5
    np.testing.assert_allclose(kfs[1].t_w2c, ref.t[1], atol=1e-5)
    np.testing.assert_allclose(kfs[1].R_w2c @ kfs[1].R_w2c.T, np.eye(3), atol=1e-6)   # stays a rotation (input is float32-rounded)
    for j, x in zip(pi.selected, ref.X):                                              # aligned write-back
        np.testing.assert_allclose(mps[int(j)].position, x, atol=1e-6)
    np.testing.assert_array_equal(mps[0].position, untouched[0])
    np.testing.assert_array_equal(mps[-1].position, untouched[1])
    np.testing.assert_array_equal(kfs[0].R_w2c, untouched[2])
    np.testing.assert_array_equal(kfs[0].t_w2c, untouched[3])
    # the default local-BA parameters are back after the call
    assert opt._get_solver().cfg.huber_delta == 5.0


def test_refine_motion_against_the_reference_pnp_result():
    g = np.load(os.path.join(GOLDEN, "ref_pose_only.npz"))
    rng = np.random.default_rng(3)
    w = rng.normal(0, 0.02, 3)
    R0 = synth.so3_exp(w[None])[0] @ g["R_true"]
    t0 = g["t_true"] + rng.normal(0, 0.1, 3)
    R, t, ok, inl = motion.refine_motion_from_2d3d(g["pts3d"], g["pts2d"], g["K"], None, True, R0=R0, t0=t0)
    assert ok and R.shape == (3, 3) and t.shape == (3, 1)
    assert inl.dtype == np.int32 and inl.ndim == 2 and inl.shape[1] == 1            # tracking.py:364,404-409
    # same pose as OpenCV's RANSAC + iterative PnP up to the noise level (it minimises plain L2 on its inliers)
    assert rot_angle(R[None], g["ref_R"][None]).max() < 2e-3
    assert np.abs(t - g["ref_t"]).max() < 2e-2
    ours, theirs = set(inl[:, 0].tolist()), set(g["ref_inliers"][:, 0].tolist())
    assert len(ours & theirs) >= 0.95 * len(theirs)
    assert not (ours & set(g["outliers"].tolist()))
    # and the oracle's pose-only LM with the same threshold, bit-exact flags
    intr = np.array([g["K"][0, 0], g["K"][1, 1], g["K"][0, 2], g["K"][1, 2]])
    ref = O.solve_pose_only(intr, R0[None], t0[None], g["pts3d"], g["pts2d"], np.array([0, len(g["pts3d"])], np.int32),
                            O.LMParams(chi2_thr=9.0))
    np.testing.assert_array_equal(np.flatnonzero(ref["flags"] == 0), inl[:, 0])
    assert rot_angle(R[None], ref["R"]).max() < 1e-5 and np.abs(t[:, 0] - ref["t"][0]).max() < 1e-5
    none = motion.refine_motion_from_2d3d(g["pts3d"], g["pts2d"], g["K"], None, False, R0=R0, t0=t0)
    assert none[2] is True and none[3] is None                                       # use_ransac=False: no inlier list


def test_local_mapping_call_site_with_fast_packer_and_culling():
    p = synth.make_problem("T2", seed=6)
    table = packing.ObservationTable()
    table.install(refgraph.KeyFrame)
    try:
        cam, kfs, mps = refgraph.graph_from_problem(p)
        cfg = type("Cfg", (), {"optimization": type("Opt", (), {"huber_delta": 5.0})()})()
        opt = B200Optimizer(cfg, log_dir=None)
        opt.table = table
        pts = handlers.collect_window_points(kfs)
        pw0 = packing.pack_window(kfs, pts)
        before = handlers.mean_reprojection_error(pw0)
        ref = O.solve_local(pw0.intr, pw0.R, pw0.t, pw0.fixed, pw0.X, pw0.obs_cam, pw0.obs_pt, pw0.obs_uv)
        assert handlers.local_optimize(opt, kfs, cull=True, table=table) is True
        res = opt.last_result
        np.testing.assert_array_equal(res.flags, ref.flags)
        np.testing.assert_allclose(res.cost_hist, ref.cost_hist, rtol=1e-5)
        after = handlers.mean_reprojection_error(opt.last_packed, res.R, res.t, res.X)
        assert after < before
        # flagged observations left the graph (and the table): repacking drops exactly those
        pw1 = table.pack(kfs, pts)
        assert pw1.n_obs == pw0.n_obs - int(res.flags.sum())
        assert packing.pack_window(kfs, pts).n_obs == pw1.n_obs
    finally:
        table.uninstall()
