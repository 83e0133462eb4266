"""Pose-only (motion-only) refinement behind the reference's tracking call site.

The reference estimates the current frame's pose in `Tracking._optimize_pose_mono`
(`visual_slam/tracking.py:332-413`) with `estimate_motion_from_2d3d`
(`visual_slam/utils/motion_estimation.py:57-132`: `cv2.solvePnPRansac`, SOLVEPNP_ITERATIVE,
reprojection error 3 px).  `refine_motion_from_2d3d` keeps that function's argument meaning and
return tuple -- `(R (3,3), t (3,1), success, inliers (M,1) int32 indices)`, `< 4` points ->
`(None, None, False, None)` -- and replaces the RANSAC + OpenCV LM by the batched pose-only
Levenberg-Marquardt of the CUDA library (`ba_solve_pose_only`), which needs the initial pose the
tracker already has from its motion model (`Tracking._predict_pose`, `tracking.py:446-451`).
The objective is the local-BA objective with the points fixed (per-component Huber,
`torch_optimizer.py:135-177`); inliers are the matches whose squared reprojection error at the
refined pose is at most `reprojection_error**2` (the RANSAC threshold's meaning).
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .solver import LMConfig, LocalBASolver


def _intr_from_K(K: np.ndarray) -> np.ndarray:
    K = np.asarray(K, dtype=np.float64)
    return np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], dtype=np.float64)


def refine_motion_batched(solver: LocalBASolver, pts3d: Sequence[np.ndarray], pts2d: Sequence[np.ndarray],
                          K: np.ndarray, R0: np.ndarray, t0: np.ndarray,
                          reprojection_error: float = 3.0):
    """Refine F frames in one launch.  pts3d[f] (N_f,3), pts2d[f] (N_f,2); R0 (F,3,3), t0 (F,3) world->camera.
    Returns (R (F,3,3), t (F,3), flags per match (concatenated), frame_offsets, cost_hist (F, it+1))."""
    F = len(pts3d)
    counts = [int(np.asarray(p).shape[0]) for p in pts3d]
    fo = np.zeros(F + 1, dtype=np.int32)
    np.cumsum(counts, out=fo[1:])
    X = np.concatenate([np.asarray(p, dtype=np.float32).reshape(-1, 3) for p in pts3d]) if F else np.zeros((0, 3), np.float32)
    uv = np.concatenate([np.asarray(p, dtype=np.float32).reshape(-1, 2) for p in pts2d]) if F else np.zeros((0, 2), np.float32)
    cfg = dataclasses.replace(solver.cfg, chi2_thr=float(reprojection_error) ** 2)
    old = solver.cfg
    solver.set_config(cfg)
    try:
        out = solver.solve_pose_only(_intr_from_K(K), np.asarray(R0, dtype=np.float64).reshape(F, 3, 3),
                                     np.asarray(t0, dtype=np.float64).reshape(F, 3), X, uv, fo)
    finally:
        solver.set_config(old)
    return out["R"], out["t"], out["flags"], fo, out["cost_hist"]


def refine_motion_from_2d3d(
    pts3d: np.ndarray,
    pts2d: np.ndarray,
    K: np.ndarray,
    dist: Optional[np.ndarray] = None,
    use_ransac: bool = True,
    iterations_count: int = 100,
    reprojection_error: float = 3.0,
    confidence: float = 0.99,
    *,
    R0: Optional[np.ndarray] = None,
    t0: Optional[np.ndarray] = None,
    solver: Optional[LocalBASolver] = None,
) -> Tuple[Optional[np.ndarray], Optional[np.ndarray], bool, Optional[np.ndarray]]:
    """Same signature and return tuple as `estimate_motion_from_2d3d` plus the initial pose
    (`R0`, `t0`: world->camera, default identity / zero as a stationary-camera prior).
    `dist` must be None or all zeros (the reference's optimizer ignores distortion too,
    `torch_optimizer.py:170-171`); `iterations_count` / `confidence` are RANSAC knobs with no
    counterpart here; with `use_ransac=False` the inlier list is None as in the reference (`:126`)."""
    pts3d = np.asarray(pts3d)
    pts2d = np.asarray(pts2d)
    if pts3d.shape[0] < 4 or pts2d.shape[0] < 4:               # motion_estimation.py:100-101
        return None, None, False, None
    if dist is not None and np.any(np.asarray(dist) != 0):
        raise ValueError("refine_motion_from_2d3d: lens distortion is not modelled; undistort the keypoints first")
    own = solver is None
    s = solver or LocalBASolver(LMConfig())
    try:
        R_init = np.eye(3) if R0 is None else np.asarray(R0, dtype=np.float64).reshape(3, 3)
        t_init = np.zeros(3) if t0 is None else np.asarray(t0, dtype=np.float64).reshape(3)
        R, t, flags, _, hist = refine_motion_batched(s, [pts3d], [pts2d], K, R_init[None], t_init[None],
                                                     reprojection_error)
    finally:
        if own:
            s.close()
    ok = bool(np.isfinite(R).all() and np.isfinite(t).all() and np.isfinite(hist[0, -1]))
    if not ok:
        return None, None, False, None
    inliers = np.flatnonzero(flags == 0).astype(np.int32).reshape(-1, 1) if use_ransac else None
    return R[0].copy(), t[0].reshape(3, 1).copy(), True, inliers
