"""Pose-only (motion-only) refinement behind the reference's tracking call site.

The reference estimates the current frame's pose in `Tracking._optimize_pose_mono`
(`visual_slam/tracking.py:332-413`) with `estimate_motion_from_2d3d`
(`visual_slam/utils/motion_estimation.py:57-132`: `cv2.solvePnPRansac`, SOLVEPNP_ITERATIVE,
reprojection error 3 px).  `refine_motion_from_2d3d` keeps that function's argument meaning and
return tuple -- `(R (3,3), t (3,1), success, inliers (M,1) int32 indices)`, `< 4` points ->
`(None, None, False, None)` -- and replaces the RANSAC + OpenCV LM by the batched pose-only
Levenberg-Marquardt of the CUDA library (`ba_solve_pose_only`), which needs the initial pose the
tracker already has from its motion model (`Tracking._predict_pose`, `tracking.py:446-451`, stores it in
`current_frame.pose`; read it back as `current_frame.R_w2c / t_w2c`).  Like the function it replaces it never
raises: whatever it cannot do comes back as the failure tuple `(None, None, False, None)`.
The objective is the local-BA objective with the points fixed (per-component Huber,
`torch_optimizer.py:135-177`); inliers are the matches whose squared reprojection error at the
refined pose is at most `reprojection_error**2` (the RANSAC threshold's meaning).
"""
from __future__ import annotations

import logging
import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .solver import LMConfig, LocalBASolver


def _intr_from_K(K: np.ndarray) -> np.ndarray:
    K = np.asarray(K, dtype=np.float64)
    return np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], dtype=np.float64)


_log = logging.getLogger("b200ba.motion")
_solvers = {}                    # device index -> LocalBASolver, created once (a handle owns a stream, pinned memory, graphs)
_solvers_lock = threading.Lock()


def default_solver(device: int = 0) -> LocalBASolver:
    """One cached solver per device for call sites that do not bring their own."""
    with _solvers_lock:
        s = _solvers.get(device)
        if s is None:
            s = _solvers[device] = LocalBASolver(LMConfig(), device=device)
        return s


def refine_motion_batched(solver: LocalBASolver, pts3d: Sequence[np.ndarray], pts2d: Sequence[np.ndarray],
                          K: np.ndarray, R0: np.ndarray, t0: np.ndarray,
                          reprojection_error: float = 3.0):
    """Refine F frames in one launch.  pts3d[f] (N_f,3), pts2d[f] (N_f,2); R0 (F,3,3), t0 (F,3) world->camera.
    Returns (R (F,3,3), t (F,3), flags per match (concatenated), frame_offsets, cost_hist (F, it+1)).
    The inlier threshold travels with the call (`chi2_thr`), applied under the solver's own lock, so a solver
    shared with the local-BA thread is never reconfigured while it runs."""
    F = len(pts3d)
    counts = [int(np.asarray(p).shape[0]) for p in pts3d]
    fo = np.zeros(F + 1, dtype=np.int32)
    np.cumsum(counts, out=fo[1:])
    X = np.concatenate([np.asarray(p, dtype=np.float32).reshape(-1, 3) for p in pts3d]) if F else np.zeros((0, 3), np.float32)
    uv = np.concatenate([np.asarray(p, dtype=np.float32).reshape(-1, 2) for p in pts2d]) if F else np.zeros((0, 2), np.float32)
    out = solver.solve_pose_only(_intr_from_K(K), np.asarray(R0, dtype=np.float64).reshape(F, 3, 3),
                                 np.asarray(t0, dtype=np.float64).reshape(F, 3), X, uv, fo,
                                 chi2_thr=float(reprojection_error) ** 2)
    return out["R"], out["t"], out["flags"], fo, out["cost_hist"]


def bootstrap_pose_ransac(solver: LocalBASolver, pts3d: np.ndarray, pts2d: np.ndarray, K: np.ndarray,
                          iterations_count: int = 100, reprojection_error: float = 3.0,
                          confidence: float = 0.99):
    """Pose without a prior: RANSAC over minimal three-point samples on the device (`ba_pose_ransac`), each
    sample solved by P3P and scored by its `reprojection_error` inlier count over all matches -- the consensus
    step of `cv2.solvePnPRansac` (`utils/motion_estimation.py:107-116`).  Returns (R (3,3), t (3,)) of the best
    hypothesis, or None when no hypothesis reaches four inliers."""
    out = solver.pose_ransac(_intr_from_K(K), np.asarray(pts3d, dtype=np.float32).reshape(-1, 3),
                             np.asarray(pts2d, dtype=np.float32).reshape(-1, 2),
                             n_hypotheses=max(int(iterations_count), 1), threshold_px=float(reprojection_error))
    if out is None or out["n_inliers"] < 4:
        return None
    return out["R"], out["t"]


def _undistort(pts2d: np.ndarray, K: np.ndarray, dist: np.ndarray) -> Optional[np.ndarray]:
    """Pixels of the ideal pinhole camera for distorted pixels (what `cv2.solvePnP*` does internally with
    `distCoeffs`); None when OpenCV is not there."""
    try:
        import cv2
        K = np.asarray(K, dtype=np.float64)
        out = cv2.undistortPoints(np.asarray(pts2d, dtype=np.float64).reshape(-1, 1, 2), K,
                                  np.asarray(dist, dtype=np.float64).reshape(-1), P=K)
        return out.reshape(-1, 2)
    except Exception as e:
        _log.warning(f"refine_motion_from_2d3d: cannot undistort ({e!r})")
        return None


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
    (`R0`, `t0`: world->camera -- the tracker's motion-model prediction).  Without a prior the pose is
    bootstrapped on the device (`bootstrap_pose_ransac`: RANSAC over minimal P3P samples scored by the 3 px
    inlier count, the same consensus `cv2.solvePnPRansac` forms) before the LM refinement.  Non-zero `dist`:
    the pixels are undistorted first (the objective is the ideal pinhole's, `torch_optimizer.py:170-171`).
    `iterations_count` / `confidence` bound the bootstrap's sample count; with `use_ransac=False` the inlier
    list is None as in the reference (`:126`).  Never raises: failure is `(None, None, False, None)`."""
    try:
        pts3d = np.asarray(pts3d)
        pts2d = np.asarray(pts2d)
        if pts3d.ndim != 2 or pts2d.ndim != 2 or pts3d.shape[0] < 4 or pts2d.shape[0] < 4 \
                or pts3d.shape[0] != pts2d.shape[0]:              # motion_estimation.py:100-101
            return None, None, False, None
        if dist is not None and np.any(np.asarray(dist) != 0):
            pts2d = _undistort(pts2d, K, dist)
            if pts2d is None:
                return None, None, False, None
        s = solver or default_solver()
        if R0 is None or t0 is None:
            boot = bootstrap_pose_ransac(s, pts3d, pts2d, K, iterations_count, reprojection_error, confidence)
            if boot is None:
                return None, None, False, None
            R_init, t_init = boot
        else:
            R_init = np.asarray(R0, dtype=np.float64).reshape(3, 3)
            t_init = np.asarray(t0, dtype=np.float64).reshape(3)
        R, t, flags, _, hist = refine_motion_batched(s, [pts3d], [pts2d], K, R_init[None], t_init[None],
                                                     reprojection_error)
        ok = bool(np.isfinite(R).all() and np.isfinite(t).all() and np.isfinite(hist[0, -1]))
        if not ok:
            return None, None, False, None
        inliers = np.flatnonzero(flags == 0).astype(np.int32).reshape(-1, 1) if use_ransac else None
        return R[0].copy(), t[0].reshape(3, 1).copy(), True, inliers
    except Exception as e:                                       # the tracking thread must survive
        _log.error(f"refine_motion_from_2d3d failed: {e!r}")
        return None, None, False, None
