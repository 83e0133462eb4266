"""Run the UNMODIFIED reference's local BA on a packed window  --  TEST / BASELINE INFRASTRUCTURE ONLY.

Used by `bench.py --impl reference` and by tests; never imported by the product package.

The reference is looked for in `baseline/_ref` (an offline `pip install --target` of an unchanged copy
of the reference tree, made by `__graft_entry__.build()` when `/root/reference` is present; git-ignored,
it travels to the GPU box with the snapshot) and then in `/root/reference`.  When neither is there,
`find_reference()` returns None and the caller falls back to the restated loop `oracle/ref_adam.py`
(pinned to the reference's own output by tests/test_ref_adam.py).

What runs is the reference's own call stack, nothing of ours inside it:
  Map.optimize_local                      visual_slam/map/map.py:150-171
    -> TorchOptimizer.optimize_local      visual_slam/optimization/torch_optimizer.py:28-206
on a graph of the reference's own KeyFrame / MapPoint / Pose / PinholeCamera objects and cv2.KeyPoint,
built from the packed arrays exactly as SURVEY.md 8c-C2 describes.  The only shim is the
`matplotlib.pylab` stub for the stray import at visual_slam/map/keyframe.py:4.
"""
from __future__ import annotations

import logging
import os
import re
import sys
import tempfile
import time
import types
from typing import Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")


def find_reference() -> Optional[str]:
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "visual_slam", "optimization", "torch_optimizer.py")):
            return c
    return None


def import_reference(path: Optional[str] = None) -> dict:
    """Import the reference's optimizer + map types from `path`.  Raises ImportError if it cannot be had."""
    path = path or find_reference()
    if path is None:
        raise ImportError("reference tree not found (baseline/_ref, /root/reference)")
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        pylab = types.ModuleType("matplotlib.pylab")
        pylab.ndarray = np.ndarray
        mpl.pylab = pylab
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pylab"] = pylab
    if path not in sys.path:
        sys.path.insert(0, path)
    import cv2  # noqa: F401
    from visual_slam.camera import PinholeCamera
    from visual_slam.config import Config, OptimizationConfig
    from visual_slam.map.keyframe import KeyFrame
    from visual_slam.map.map import Map
    from visual_slam.map.map_point import MapPoint
    from visual_slam.map.pose import Pose
    from visual_slam.optimization.torch_optimizer import TorchOptimizer
    return dict(path=path, Config=Config, OptimizationConfig=OptimizationConfig, PinholeCamera=PinholeCamera,
                KeyFrame=KeyFrame, MapPoint=MapPoint, Map=Map, Pose=Pose, TorchOptimizer=TorchOptimizer)


def graph_from_arrays(ref: dict, intr, R, t, X, obs_cam, obs_pt, obs_uv):
    """Reference object graph of a packed window: one keypoint per observation, linked with
    `KeyFrame.add_map_point` (keyframe.py:82-86), points created in index order."""
    import cv2
    cam = ref["PinholeCamera"](1280, 720, *[float(v) for v in intr])
    n_cam = int(np.asarray(R).shape[0])
    obs_cam = np.asarray(obs_cam)
    obs_pt = np.asarray(obs_pt)
    obs_uv = np.asarray(obs_uv, dtype=np.float32)
    per_cam = [np.nonzero(obs_cam == i)[0] for i in range(n_cam)]
    kp_index = np.zeros(len(obs_cam), dtype=np.int64)
    kfs = []
    for i in range(n_cam):
        idx = per_cam[i]
        kp_index[idx] = np.arange(len(idx))
        kps = [cv2.KeyPoint(float(obs_uv[o, 0]), float(obs_uv[o, 1]), 1.0) for o in idx]
        kfs.append(ref["KeyFrame"](timestamp=float(i), camera=cam, images=[], keypoints=[kps], descriptors=[None],
                                   pose=ref["Pose"](R=np.asarray(R[i], dtype=np.float64),
                                                    t=np.asarray(t[i], dtype=np.float64))))
    mps = [ref["MapPoint"](position=np.asarray(X[j], dtype=np.float64)) for j in range(int(np.asarray(X).shape[0]))]
    for o in range(len(obs_cam)):
        kfs[int(obs_cam[o])].add_map_point(0, int(kp_index[o]), mps[int(obs_pt[o])])
    return cam, kfs, mps


class _Capture(logging.Handler):
    def __init__(self):
        super().__init__()
        self.lines = []

    def emit(self, record):
        self.lines.append(record.getMessage())


def run_local_ba(ref: dict, kfs, n_iter: int, lr: float = 1e-4, delta: float = 5.0):
    """`Map().optimize_local(TorchOptimizer(cfg, device="cpu"), kfs)`; returns (ok, seconds, logged losses)."""
    import torch
    tmp = tempfile.mkdtemp(prefix="ref_ba_")
    cfg = ref["Config"](optimization=ref["OptimizationConfig"](lr=lr, n_iter=n_iter, huber_delta=delta))
    opt = ref["TorchOptimizer"](cfg, device="cpu", log_dir=tmp)
    cap = _Capture()
    opt.logger.addHandler(cap)
    m = ref["Map"]()
    t0 = time.perf_counter()
    ok = m.optimize_local(opt, kfs)
    dt = time.perf_counter() - t0
    losses = []
    for ln in cap.lines:
        mm = re.search(r"iter=(\d+), loss=([-0-9.eE+naif]+)", ln)
        if mm:
            losses.append((int(mm.group(1)), float(mm.group(2))))
    return bool(ok), dt, losses, dict(torch=torch.__version__, torch_threads=torch.get_num_threads())
