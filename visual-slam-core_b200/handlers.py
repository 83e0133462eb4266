"""The local-mapping call site: what `LocalHandler.local_optimize` and
`LocalHandler.postprocessing` do around the optimizer (`visual_slam/handlers/local_handler.py:54-80`),
restated over the packed window so that the two O(N_obs) object walks of
`Map.compute_mean_reprojection_error` (`map/map.py:95-141`, `utils/geometry.py:209-247`) become two
vectorised passes over arrays the packer has built anyway.

The reference code path needs no change to use `B200Optimizer`; this module is the
call site for hosts that also want the outlier flags consumed (`cull_bad_points`).
"""
from __future__ import annotations

from typing import Any, List, Optional

import numpy as np

from . import packing


def mean_reprojection_error(pw: packing.PackedWindow, R: Optional[np.ndarray] = None,
                            t: Optional[np.ndarray] = None, X: Optional[np.ndarray] = None) -> float:
    """Mean Euclidean pixel error over the window's observations with the depth clipped at 1e-8,
    as `compute_reprojection_error` (`utils/geometry.py:232-247`) computes it per keyframe and
    `Map.compute_mean_reprojection_error` averages it weighted by the observation count."""
    R = pw.R if R is None else R
    t = pw.t if t is None else t
    X = pw.X if X is None else X
    if pw.n_obs == 0:
        return 0.0
    Xc = np.einsum("nij,nj->ni", R[pw.obs_cam], X[pw.obs_pt]) + t[pw.obs_cam]
    z = np.clip(Xc[:, 2], 1e-8, None)
    u = pw.intr[0] * Xc[:, 0] / z + pw.intr[2]
    v = pw.intr[1] * Xc[:, 1] / z + pw.intr[3]
    return float(np.mean(np.hypot(u - pw.obs_uv[:, 0], v - pw.obs_uv[:, 1])))


def collect_window_points(keyframes: List[Any]) -> List[Any]:
    """`Map.optimize_local`'s collection (`map/map.py:159-169`): points of every keyframe in order,
    deduplicated by identity, first occurrence kept."""
    fp = packing._fastpack
    if fp is not None and hasattr(fp, "collect_points") and isinstance(keyframes, list):
        out = fp.collect_points(keyframes)               # the same walk in C (csrc/fastpack.c)
        if out is not None:
            return out
    unique = {}
    for kf in keyframes:
        for mp in kf.get_all_mappoints().values():
            unique[id(mp)] = mp
    return list(unique.values())


def local_optimize(optimizer, keyframes: List[Any], logger=None, cull: bool = False,
                   table: Optional[packing.ObservationTable] = None) -> bool:
    """`LocalHandler.local_optimize` + `postprocessing` for one window: mean reprojection error
    before, the solve, mean error after (`local_handler.py:56-74`), then optionally the outlier
    culling the reference left as a stub (`:76-80`).  More than two keyframes are required by the
    caller (`:37-40`); with fewer this returns False without touching the map."""
    log = logger or optimizer.logger
    if len(keyframes) <= 2:
        return False
    if table is not None and hasattr(optimizer, "table"):
        optimizer.table = table                          # pack from this table instead of the process-wide one
    points = collect_window_points(keyframes)
    ok = optimizer.optimize_local(keyframes, points)
    log.info(f"[Local Optimize] finished: success={ok}")
    if ok and optimizer.last_result is not None:
        # the window is packed once, by the optimizer; its snapshot holds the poses and points from before the solve
        res, pw = optimizer.last_result, optimizer.last_packed
        log.info(f"[Local Optimize] mean reprojection error before: {mean_reprojection_error(pw):.3f}px, "
                 f"after: {mean_reprojection_error(pw, res.R, res.t, res.X):.3f}px")
        if cull:
            optimizer.cull_bad_points(keyframes, points)
    return ok
