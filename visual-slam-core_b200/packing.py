"""Map-point / observation packing: reference object graph -> structure-of-arrays.

Reproduces, order and skip rules included, what the reference's optimizer builds before it
iterates (`visual_slam/optimization/torch_optimizer.py:43-95`):

  * gauge: `keyframes[0]` is the fixed keyframe (`:43`, identity test `kf is kf_fixed` `:54`)
  * intrinsics from `keyframes[0]._camera.K` only, through float32 (`:46`, used `:170-171`)
  * per keyframe R_w2c / t_w2c through float32 (`:51-52`); poses are keyed by `keyframe_id` (`:61`)
  * points `mp.position` through float32 (`:64-68`)
  * observations: for point j, for (kf_id, cams) in `mp.get_observations()` insertion order,
    skip kf_id not in the window (`:82-83`), for (cam_id, kp_idx) skip cam_id >= len(kf.keypoints)
    and kp_idx >= len(kf.keypoints[cam_id]) (`:90-93`); uv = `kf.keypoints[cam_id][kp_idx].pt` (`:94`)

The outer loop is over points, so the list is already sorted by point: `pt_offsets` is its CSR.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List

import numpy as np

MIN_OBSERVATIONS = 10   # torch_optimizer.py:99-101


@dataclass
class PackedWindow:
    intr: np.ndarray        # (4,) f64 (float32-valued)
    R: np.ndarray           # (n_cam,3,3) f64 (float32-valued)
    t: np.ndarray           # (n_cam,3)
    fixed: np.ndarray       # (n_cam,) u8
    X: np.ndarray           # (n_pt,3) f64 (float32-valued)
    obs_cam: np.ndarray     # (n_obs,) i32  index into `keyframes`
    obs_pt: np.ndarray      # (n_obs,) i32
    obs_uv: np.ndarray      # (n_obs,2) f32
    obs_cam_id: np.ndarray  # (n_obs,) i32  the reference's cam_id (sensor index inside the keyframe)
    pt_offsets: np.ndarray  # (n_pt+1,) i32

    @property
    def n_obs(self) -> int:
        return int(self.obs_cam.shape[0])


def pack_window(keyframes: List[Any], map_points: List[Any]) -> PackedWindow:
    kf_fixed = keyframes[0]
    K = np.asarray(kf_fixed._camera.K, dtype=np.float32)
    intr = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], dtype=np.float64)

    n_cam = len(keyframes)
    R = np.empty((n_cam, 3, 3), dtype=np.float64)
    t = np.empty((n_cam, 3), dtype=np.float64)
    fixed = np.zeros(n_cam, dtype=np.uint8)
    pose_index = {}                       # keyframe_id -> row (a later duplicate id overwrites, as the dict at :61 does)
    for i, kf in enumerate(keyframes):
        R[i] = np.asarray(kf.R_w2c, dtype=np.float32)
        t[i] = np.asarray(kf.t_w2c, dtype=np.float32).reshape(3)
        fixed[i] = 1 if kf is kf_fixed else 0
        pose_index[kf.keyframe_id] = i

    n_pt = len(map_points)
    if n_pt:
        X = np.array([mp.position for mp in map_points], dtype=np.float32).reshape(n_pt, 3).astype(np.float64)
    else:
        X = np.zeros((0, 3), dtype=np.float64)

    obs_cam: List[int] = []
    obs_pt: List[int] = []
    obs_cid: List[int] = []
    obs_uv: List[Any] = []
    counts = np.zeros(n_pt, dtype=np.int64)
    kps_of = [keyframes[i].keypoints for i in range(n_cam)]
    for j, mp in enumerate(map_points):
        c0 = len(obs_cam)
        for kf_id, cams in mp.get_observations().items():
            i = pose_index.get(kf_id)
            if i is None:
                continue
            kps = kps_of[i]
            for cam_id, kp_idx in cams.items():
                if cam_id >= len(kps):
                    continue
                if kp_idx >= len(kps[cam_id]):
                    continue
                obs_uv.append(kps[cam_id][kp_idx].pt)
                obs_cam.append(i)
                obs_pt.append(j)
                obs_cid.append(cam_id)
        counts[j] = len(obs_cam) - c0

    pt_offsets = np.zeros(n_pt + 1, dtype=np.int32)
    np.cumsum(counts, out=pt_offsets[1:])
    return PackedWindow(
        intr=intr, R=R, t=t, fixed=fixed, X=X,
        obs_cam=np.asarray(obs_cam, dtype=np.int32), obs_pt=np.asarray(obs_pt, dtype=np.int32),
        obs_uv=np.asarray(obs_uv, dtype=np.float32).reshape(-1, 2),
        obs_cam_id=np.asarray(obs_cid, dtype=np.int32), pt_offsets=pt_offsets)


def write_back(keyframes: List[Any], map_points: List[Any], R: np.ndarray, t: np.ndarray,
               X: np.ndarray) -> None:
    """Deliver results the way the reference does (`torch_optimizer.py:188-203`): every keyframe
    (the fixed one included) gets `set_pose_Rt(R 3x3, t 3x1)`, every point `update_position`."""
    for i, kf in enumerate(keyframes):
        kf.set_pose_Rt(np.array(R[i], dtype=np.float64), np.array(t[i], dtype=np.float64).reshape(3, 1))
    for mp, pos in zip(map_points, X):
        mp.update_position(np.asarray(pos))
