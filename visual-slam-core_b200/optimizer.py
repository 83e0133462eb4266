"""`B200Optimizer`: the drop-in for the reference's optimizer plugin.

Mirrors `BaseOptimizer` (`visual_slam/optimization/base_optimizer.py:11-50`) with the signatures
its callers actually use (`map/map.py:148,171,178`): `optimize_local(keyframes, map_points) -> bool`,
`optimize_global(...)`, `optimize_initial(...)`; constructor `(config, logger, log_dir, device)` as
`TorchOptimizer` (`torch_optimizer.py:18-26`).  Results are delivered by mutating the passed
objects; nothing is raised to the caller: any failure is logged and reported as `False`.
"""
from __future__ import annotations

import dataclasses
import logging
import os
import threading
from abc import ABC, abstractmethod
from typing import Any, List, Optional

import numpy as np

from . import packing
from .solver import LMConfig, LocalBASolver, SolveResult


class BaseOptimizer(ABC):
    """Interface mirror of the reference ABC (base_optimizer.py:11-50)."""

    def __init__(self, config: Optional[Any] = None, logger: Optional[Any] = None,
                 log_dir: Optional[str] = "logs") -> None:
        self.config = config
        self.logger = logger or _make_logger(self.__class__.__name__, log_dir)
        self.logger.info("*" * 80)

    @abstractmethod
    def optimize_local(self, keyframes: List[Any], map_points: List[Any]) -> bool: ...

    @abstractmethod
    def optimize_global(self, keyframes: List[Any], map_points: List[Any]) -> bool: ...

    @abstractmethod
    def optimize_initial(self, keyframes: List[Any], map_points: List[Any]) -> bool: ...


def _make_logger(name: str, log_dir: Optional[str]) -> logging.Logger:
    logger = logging.getLogger(f"b200ba.{name}")
    logger.setLevel(logging.INFO)
    if log_dir and not any(isinstance(h, logging.FileHandler) for h in logger.handlers):
        try:
            from logging.handlers import RotatingFileHandler
            os.makedirs(log_dir, exist_ok=True)
            fh = RotatingFileHandler(os.path.join(log_dir, f"{name.lower()}.log"),
                                     maxBytes=5 * 1024 * 1024, backupCount=3)
            fh.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
            logger.addHandler(fh)
        except OSError:
            pass
    return logger


def lm_config_from(config: Optional[Any]) -> LMConfig:
    """Read LM fields the way the reference reads its own (`getattr(config.optimization, ...)`,
    torch_optimizer.py:106-108), so a stock `Config` works and extra fields are optional."""
    opt = getattr(config, "optimization", None) if config is not None else None
    g = (lambda k, d: getattr(opt, k, d)) if opt is not None else (lambda k, d: d)
    return LMConfig(
        huber_delta=float(g("huber_delta", 5.0)),
        chi2_thr=float(g("chi2_threshold", 5.991)),
        lambda0=float(g("lm_lambda0", 1e-4)),
        lambda_up=float(g("lm_lambda_up", 10.0)),
        lambda_down=float(g("lm_lambda_down", 0.1)),
        lambda_min=float(g("lm_lambda_min", 1e-6)),
        lambda_max=float(g("lm_lambda_max", 1e8)),
        accept_rtol=float(g("lm_accept_rtol", 1e-8)),
        max_iter=int(g("lm_max_iter", 10)),
        solver=str(g("solver", "auto")),
        schur=str(g("schur", "auto")),
    )


class B200Optimizer(BaseOptimizer):
    """Levenberg-Marquardt local BA on a B200 behind the reference's optimizer API."""

    def __init__(self, config: Optional[Any] = None, logger: Optional[Any] = None,
                 log_dir: Optional[str] = "logs", device: str = "cuda") -> None:
        super().__init__(config, logger, log_dir=log_dir)
        self.device = device
        dev_index = 0
        if isinstance(device, str) and ":" in device:
            dev_index = int(device.split(":")[1])
        self._lm = lm_config_from(config)
        self._solver: Optional[LocalBASolver] = None
        self._dev_index = dev_index
        self._lock = threading.Lock()          # one shared instance, two caller threads (tracking.py:61)
        self.last_result: Optional[SolveResult] = None
        self.last_packed: Optional[packing.PackedWindow] = None
        self.mark_bad_points = bool(getattr(getattr(config, "optimization", None), "mark_bad_points", False))
        # optional incremental SoA mirror of the observation graph (packing.ObservationTable):
        # when set, windows are packed from it instead of walking the object graph
        self.table: Optional[packing.ObservationTable] = None

    def _get_solver(self) -> LocalBASolver:
        if self._solver is None:
            self._solver = LocalBASolver(self._lm, device=self._dev_index)   # raises without a GPU
        return self._solver

    # ------------------------------------------------------------------------------ local BA
    def optimize_local(self, keyframes: List[Any], map_points: List[Any]) -> bool:
        try:
            self.logger.info(f"[B200Optimizer] keyframes - {len(keyframes)}, points - {len(map_points)}.")
            if not keyframes:
                return False
            if self.table is not None:
                pw = self.table.pack(keyframes, map_points)
            else:
                pw = packing.pack_window(keyframes, map_points)  # snapshot up front, as :46-95
            self.logger.info(f"[B200Optimizer] observations - {pw.n_obs}.")
            if pw.n_obs < packing.MIN_OBSERVATIONS:              # :99-101: False, nothing mutated
                self.logger.warning("[B200Optimizer] too few observations for local optimisation.")
                return False
            with self._lock:
                res = self._get_solver().solve_local(pw.intr, pw.R, pw.t, pw.fixed, pw.X, pw.obs_cam,
                                                     pw.obs_pt, pw.obs_uv, pw.pt_offsets)
            for i in range(res.n_iter):
                self.logger.info(
                    f"[Local BA] iter={i}, loss={res.cost_hist[i + 1]:.3f}, lambda={res.lambda_hist[i]:.1e}, "
                    f"accepted={int(res.accepted[i])}")
            if not (np.isfinite(res.R).all() and np.isfinite(res.t).all() and np.isfinite(res.X).all()):
                self.logger.error("[B200Optimizer] non-finite result; nothing written back.")
                return False
            packing.write_back(keyframes, map_points, res.R, res.t, res.X)
            self.last_result, self.last_packed = res, pw
            if self.mark_bad_points:
                self._mark_bad(map_points, pw, res.flags)
            self.logger.info("[Local BA] finished.")
            return True
        except Exception as e:  # the caller thread must survive (handlers/local_handler.py:28-48)
            self.logger.error(f"[B200Optimizer] optimize_local failed: {e!r}")
            return False

    @staticmethod
    def _mark_bad(map_points, pw, flags):
        """A point whose in-window observations are all flagged is marked bad (map_point.py:75-78)."""
        n_pt = len(map_points)
        tot = np.diff(pw.pt_offsets)
        bad = np.bincount(pw.obs_pt, weights=flags.astype(np.float64), minlength=n_pt)
        for j in np.nonzero((tot > 0) & (bad == tot))[0]:
            map_points[int(j)].mark_as_bad()

    @property
    def outlier_mask(self) -> Optional[np.ndarray]:
        """uint8 per packed observation of the last successful `optimize_local`."""
        return None if self.last_result is None else self.last_result.flags

    # ------------------------------------------------------------------------------ global / initial
    def optimize_global(self, keyframes: List[Any], map_points: List[Any]) -> bool:
        """Same solver over every keyframe (the reference's is an empty stub, torch_optimizer.py:319-320)."""
        return self.optimize_local(keyframes, map_points)

    def optimize_initial(self, keyframes: List[Any], map_points: List[Any]) -> bool:
        """Two-view refinement (reference: torch_optimizer.py:209-317) on the LM solver.

        Kept from the reference: `kf_ref, kf_cur = keyframes[:2]`, reference view fixed; only points
        seen in BOTH views with `cam_id == 0` and in-range keypoints take part (`:240-258`); fewer
        than 10 of them -> `False`, nothing mutated (`:265-267`); the objective is the sum over both
        views of SmoothL1(beta = 1) = per-component Huber with delta = 1 px (`:282,296-298`; the
        reference's mean reduction only rescales it).  Deliberately different: the
        reference optimises the nine entries of `R_cur` freely, so its result is not a rotation
        (`:228,284`), and its write-back pairs the selected points with the FIRST `len(selected)`
        entries of `map_points` (`zip`, `:313-314`); here `R_cur` stays on SO(3) and every selected
        point gets its own refined position.  The reference writes back `kf_cur` only (`:308-311`).
        """
        try:
            if len(keyframes) < 2 or len(map_points) == 0:     # :215-217
                self.logger.warning("[B200Optimizer] not enough data for optimize_initial")
                return False
            pi = packing.pack_initial(keyframes, map_points)
            if len(pi.selected) < packing.MIN_OBSERVATIONS:    # :265-267
                self.logger.warning("[B200Optimizer] too few observations for optimize_initial.")
                return False
            pw = pi.window
            lm = dataclasses.replace(self._lm, huber_delta=1.0)
            with self._lock:
                solver = self._get_solver()
                solver.set_config(lm)
                try:
                    res = solver.solve_local(pw.intr, pw.R, pw.t, pw.fixed, pw.X, pw.obs_cam, pw.obs_pt,
                                             pw.obs_uv, pw.pt_offsets)
                finally:
                    solver.set_config(self._lm)
            if not (np.isfinite(res.R).all() and np.isfinite(res.t).all() and np.isfinite(res.X).all()):
                self.logger.error("[B200Optimizer] non-finite result; nothing written back.")
                return False
            n = 2 * len(pi.selected)
            self.logger.info(f"[Initial BA] points={len(pi.selected)}, loss {res.cost_hist[0] / n:.4f} -> "
                             f"{res.cost_hist[-1] / n:.4f} (SmoothL1 mean per view pair)")
            keyframes[1].set_pose_Rt(np.array(res.R[1], dtype=np.float64),
                                     np.array(res.t[1], dtype=np.float64).reshape(3, 1))
            for j, pos in zip(pi.selected, res.X):
                map_points[int(j)].update_position(np.asarray(pos))
            self.last_result, self.last_packed = res, pw
            return True
        except Exception as e:
            self.logger.error(f"[B200Optimizer] optimize_initial failed: {e!r}")
            return False

    # ------------------------------------------------------------------------------ outliers -> map
    def cull_bad_points(self, keyframes: List[Any], map_points: List[Any], min_observations: int = 2) -> int:
        """What the stub `LocalHandler.cull_bad_points` (`handlers/local_handler.py:79-80`) was waiting
        for: drop the observations flagged by the last `optimize_local` over the same window
        (`KeyFrame.remove_map_point`, `map/keyframe.py:93-100`, which also removes the point's
        observation) and `mark_as_bad()` (`map/map_point.py:75-78`) the points left with fewer than
        `min_observations` in-window observations.  Returns the number of points marked bad."""
        res, pw = self.last_result, self.last_packed
        if res is None or pw is None or res.flags is None or len(map_points) != pw.X.shape[0] \
                or len(keyframes) != pw.R.shape[0] or pw.obs_cam_id.shape[0] != res.flags.shape[0]:
            return 0
        flagged = np.flatnonzero(res.flags)
        for o in flagged:
            kf, mp, cam_id = keyframes[int(pw.obs_cam[o])], map_points[int(pw.obs_pt[o])], int(pw.obs_cam_id[o])
            ob = mp.get_observations().get(kf.keyframe_id, {})
            if cam_id in ob and hasattr(kf, "remove_map_point"):
                kf.remove_map_point(cam_id, ob[cam_id])
        left = np.bincount(pw.obs_pt, weights=(res.flags == 0).astype(np.float64), minlength=len(map_points))
        had = np.diff(pw.pt_offsets)
        n_bad = 0
        for j in np.flatnonzero((had > 0) & (left < min_observations)):
            map_points[int(j)].mark_as_bad()
            n_bad += 1
        self.logger.info(f"[B200Optimizer] culled {len(flagged)} observations, marked {n_bad} points bad.")
        return n_bad
