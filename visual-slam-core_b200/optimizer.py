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
        # incremental SoA mirror of the observation graph (packing.ObservationTable): windows are
        # packed from it instead of walking the object graph.  On by default (`config.optimization.observation_table
        # = False` turns it off); the hooks go onto the map-point class on the first call; whatever the table cannot
        # represent falls back to the reference's walk (`packing.pack_window`).
        self.use_table = bool(getattr(getattr(config, "optimization", None), "observation_table", True))
        self.table: Optional[packing.ObservationTable] = None
        self.packer_name = "pack_window"
        # sharded global BA: set by `enable_sharding` or picked up from torch.distributed
        self._dist = None                      # (rank, world, group)

    def _get_solver(self) -> LocalBASolver:
        if self._solver is None:
            self._solver = LocalBASolver(self._lm, device=self._dev_index)   # raises without a GPU
        return self._solver

    def pack(self, keyframes: List[Any], map_points: List[Any]) -> packing.PackedWindow:
        """Snapshot of the window as structure-of-arrays (`torch_optimizer.py:43-95`): from the observation table
        when it can represent the window, by the reference's walk otherwise.  Same arrays either way."""
        if self.use_table and map_points:
            try:
                if self.table is None:
                    self.table = packing.shared_table()
                cls = type(map_points[0])
                if hasattr(cls, "add_observation") and not self.table.installed_on(cls):
                    self.table.install(cls)
                pw = self.table.pack(keyframes, map_points)
                if pw is not None:
                    self.packer_name = "ObservationTable"
                    return pw
            except Exception as e:             # never costs the solve: the walk is always there
                self.logger.warning(f"[B200Optimizer] observation table unavailable ({e!r}); walking the graph.")
        self.packer_name = "pack_window"
        return packing.pack_window(keyframes, map_points)

    # ------------------------------------------------------------------------------ local BA
    def optimize_local(self, keyframes: List[Any], map_points: List[Any]) -> bool:
        try:
            self.logger.info(f"[B200Optimizer] keyframes - {len(keyframes)}, points - {len(map_points)}.")
            if not keyframes:
                return False
            pw = self.pack(keyframes, map_points)               # snapshot up front, as :46-95
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
    def enable_sharding(self, rank: int, world: int, group=None) -> None:
        """Global BA sharded by map point over `world` processes, one GPU each.  Every
        process calls `optimize_global` with the same keyframes and map points; `torch.distributed` only carries
        the NCCL id and the refined points of the other shards."""
        self._dist = (int(rank), int(world), group)

    def _dist_state(self):
        if self._dist is not None:
            return self._dist
        try:
            import torch.distributed as td
            if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
                return (td.get_rank(), td.get_world_size(), None)
        except Exception:
            pass
        return None

    def optimize_global(self, keyframes: List[Any], map_points: List[Any]) -> bool:
        """Global BA: the same LM solver over every keyframe (the reference's is an empty stub,
        `torch_optimizer.py:319-320`, reached from `Map.optimize_global`, `map/map.py:173-178`).  In a
        `torch.distributed` job of more than one process the window is sharded by map point: every rank solves
        its contiguous point range (balanced by observation count) with the cameras replicated, the reduced
        camera system and the costs are NCCL all-reduced inside the CUDA library, and the refined points of all
        shards are gathered before the write-back, so every process ends with the same map."""
        st = self._dist_state()
        if st is None or st[1] <= 1:
            return self.optimize_local(keyframes, map_points)
        rank, world, group = st
        try:
            from . import dist as badist
            if not keyframes:
                return False
            pw = self.pack(keyframes, map_points)
            if pw.n_obs < packing.MIN_OBSERVATIONS:
                self.logger.warning("[B200Optimizer] too few observations for global optimisation.")
                return False
            lo, hi = badist.shard_point_range(pw.pt_offsets, rank, world)
            o0, o1 = int(pw.pt_offsets[lo]), int(pw.pt_offsets[hi])
            with self._lock:
                solver = self._get_solver()
                if getattr(solver, "_world", 1) != world:
                    badist.init_solver(solver, rank, world, group=group, device=f"cuda:{self._dev_index}")
                    solver._world = world
                res = solver.solve_local(pw.intr, pw.R, pw.t, pw.fixed, pw.X[lo:hi], pw.obs_cam[o0:o1],
                                         pw.obs_pt[o0:o1] - lo, pw.obs_uv[o0:o1], pw.pt_offsets[lo:hi + 1] - o0)
            X = badist.gather_points(res.X, (lo, hi), pw.X.shape[0], group=group)
            flags = badist.gather_flags(res.flags, (o0, o1), pw.n_obs, group=group)
            if not (np.isfinite(res.R).all() and np.isfinite(res.t).all() and np.isfinite(X).all()):
                self.logger.error("[B200Optimizer] non-finite result; nothing written back.")
                return False
            for i in range(res.n_iter):
                self.logger.info(f"[Global BA] iter={i}, loss={res.cost_hist[i + 1]:.3f}, lambda={res.lambda_hist[i]:.1e}, "
                                 f"accepted={int(res.accepted[i])}, shard={rank}/{world}")
            packing.write_back(keyframes, map_points, res.R, res.t, X)
            self.last_result = dataclasses.replace(res, X=X, flags=flags)
            self.last_packed = pw
            return True
        except Exception as e:
            self.logger.error(f"[B200Optimizer] optimize_global (sharded) failed: {e!r}")
            return False

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
        observation) and `mark_as_bad()` (`map/map_point.py:75-78`) the points that lost an observation here
        and are left with fewer than `min_observations` observations in the whole map.  Returns the number of
        points marked bad."""
        res, pw = self.last_result, self.last_packed
        if res is None or pw is None or res.flags is None or len(map_points) != pw.X.shape[0] \
                or len(keyframes) != pw.R.shape[0] or pw.obs_cam_id.shape[0] != res.flags.shape[0]:
            return 0
        flagged = np.flatnonzero(res.flags)
        touched = set()
        for o in flagged:
            j = int(pw.obs_pt[o])
            kf, mp, cam_id = keyframes[int(pw.obs_cam[o])], map_points[j], int(pw.obs_cam_id[o])
            ob = mp.get_observations().get(kf.keyframe_id, {})
            if cam_id in ob and hasattr(kf, "remove_map_point"):
                kf.remove_map_point(cam_id, ob[cam_id])
                touched.add(j)
        # only points that LOST an observation in this call can starve, and what counts is everything the point
        # still has in the map (keyframes outside the window included), not its in-window share
        n_bad = 0
        for j in sorted(touched):
            mp = map_points[j]
            left = sum(len(cams) for cams in mp.get_observations().values())
            if left < min_observations:
                mp.mark_as_bad()
                n_bad += 1
        self.logger.info(f"[B200Optimizer] culled {len(flagged)} observations, marked {n_bad} points bad.")
        return n_bad
