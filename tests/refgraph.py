"""Minimal stand-ins for the reference's map object graph.

`B200Optimizer` is duck-typed against the attributes the reference's optimizer reads
and the methods it calls to deliver results; these classes provide exactly those, so
tests and benchmarks can build a window without `/root/reference` (absent on the GPU
box).  With the reference importable, its own classes are used unchanged.

Mirrored surface (reference file:line):
  KeyFrame.keyframe_id                      map/keyframe.py:41-43  (process-global counter)
  KeyFrame.map_points / add_map_point       map/keyframe.py:45,82-86
  KeyFrame.get_all_mappoints                map/keyframe.py:102-105
  Frame.keypoints  (list per camera of objects with .pt)   map/frame.py:291
  Frame.R_w2c / t_w2c / set_pose_Rt         map/frame.py:78-90,386-388
  Frame._camera.K                           camera.py:70-72
  MapPoint.position / get_observations / update_position / mark_as_bad / is_valid
                                            map/map_point.py:28,45-47,61-63,75-89
  Observations.add semantics (first write wins)            map/observation.py:10-22
  KeyFrame.remove_map_point / MapPoint.remove_observation  map/keyframe.py:93-100, map/observation.py:32-50
  Map.optimize_local (collect + dedup by id)               map/map.py:150-171
"""
from __future__ import annotations

from threading import RLock
from typing import Dict, List, Optional, Tuple

import numpy as np


class Camera:
    def __init__(self, width: int, height: int, fx: float, fy: float, cx: float, cy: float):
        self.width, self.height = width, height
        self.K = np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=np.float64)


class KeyPoint:
    """cv2.KeyPoint stand-in: `.pt` is a pair of float32-valued Python floats."""

    __slots__ = ("pt",)

    def __init__(self, u: float, v: float):
        self.pt = (float(np.float32(u)), float(np.float32(v)))


class MapPoint:
    _next_id = 0
    _id_lock = RLock()

    def __init__(self, position):
        with MapPoint._id_lock:
            self.id = MapPoint._next_id
            MapPoint._next_id += 1
        self._lock = RLock()
        self.position = np.asarray(position, dtype=float)
        self._obs: Dict[int, Dict[int, int]] = {}
        self._is_bad = False

    def add_observation(self, kf_id: int, cam_id: int, kp_idx: int) -> bool:
        with self._lock:
            cams = self._obs.setdefault(kf_id, {})
            if cam_id in cams:
                return False
            cams[cam_id] = kp_idx
            return True

    def get_observations(self) -> Dict[int, Dict[int, int]]:
        with self._lock:
            return {kf: cams.copy() for kf, cams in self._obs.items()}

    def remove_observation(self, kf_id: int, cam_id: Optional[int] = None) -> bool:
        with self._lock:                                   # map/observation.py:32-50
            if kf_id not in self._obs:
                return False
            if cam_id is None:
                self._obs.pop(kf_id, None)
                return True
            if cam_id in self._obs[kf_id]:
                del self._obs[kf_id][cam_id]
                if not self._obs[kf_id]:
                    del self._obs[kf_id]
                return True
            return False

    def update_position(self, new_position):
        with self._lock:
            self.position = np.asarray(new_position, dtype=float)

    def mark_as_bad(self):
        with self._lock:
            self._is_bad = True

    def is_valid(self) -> bool:
        with self._lock:
            return (not self._is_bad) and bool(np.isfinite(self.position).all()) and self.position[2] > 0


class KeyFrame:
    _next_id = 0
    _id_lock = RLock()

    def __init__(self, camera: Camera, keypoints: List[List[KeyPoint]], R, t, timestamp: float = 0.0):
        with KeyFrame._id_lock:
            self.keyframe_id = KeyFrame._next_id
            KeyFrame._next_id += 1
        self._lock = RLock()
        self._camera = camera
        self.timestamp = timestamp
        self.keypoints = keypoints
        self.map_points: Dict[Tuple[int, int], MapPoint] = {}
        self._T = np.eye(4)
        self.set_pose_Rt(np.asarray(R, dtype=float), np.asarray(t, dtype=float))

    @property
    def R_w2c(self) -> np.ndarray:
        return self._T[:3, :3].copy()

    @property
    def t_w2c(self) -> np.ndarray:
        return self._T[:3, 3].copy()

    def set_pose_Rt(self, R: np.ndarray, t: np.ndarray):
        assert R.shape == (3, 3)
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = np.asarray(t).reshape(3)
        self._T = T

    def add_map_point(self, cam_id: int, kp_idx: int, mp: MapPoint):
        with self._lock:
            self.map_points[(cam_id, kp_idx)] = mp
            mp.add_observation(self.keyframe_id, cam_id, kp_idx)

    def remove_map_point(self, cam_id: int, kp_idx: int):
        with self._lock:                                   # map/keyframe.py:93-100
            key = (cam_id, kp_idx)
            if key in self.map_points:
                self.map_points[key].remove_observation(self.keyframe_id, cam_id)
                del self.map_points[key]

    def get_all_mappoints(self) -> Dict[Tuple[int, int], MapPoint]:
        with self._lock:
            return dict(self.map_points)


class Map:
    """Only the local-BA dispatch of the reference's Map (map/map.py:150-171)."""

    def optimize_local(self, optimizer, keyframes: List[KeyFrame]) -> bool:
        if not keyframes:
            return False
        unique = {}
        for kf in keyframes:
            for mp in kf.get_all_mappoints().values():
                unique[id(mp)] = mp
        return optimizer.optimize_local(keyframes, list(unique.values()))


def graph_from_problem(prob, kp_cls=KeyPoint):
    """Build a stand-in KeyFrame/MapPoint window from a packed `synth.BAProblem`.

    Point j is linked to its cameras in ascending camera order, so the packer's
    (point, keyframe-insertion) order reproduces the problem's CSR order.
    """
    fx, fy, cx, cy = [float(v) for v in prob.intr]
    cam = Camera(1280, 720, fx, fy, cx, cy)
    n_cam = prob.n_cam
    per_cam: List[List[KeyPoint]] = [[] for _ in range(n_cam)]
    kfs = [KeyFrame(cam, [per_cam[i]], prob.R[i], prob.t[i], float(i)) for i in range(n_cam)]
    mps = [MapPoint(prob.X[j].astype(np.float64)) for j in range(prob.n_pt)]
    oc, op, uv = prob.obs_cam, prob.obs_pt, prob.obs_uv
    for o in range(prob.n_obs):
        i = int(oc[o])
        kps = per_cam[i]
        kps.append(kp_cls(float(uv[o, 0]), float(uv[o, 1])))
        kfs[i].add_map_point(0, len(kps) - 1, mps[int(op[o])])
    return cam, kfs, mps
