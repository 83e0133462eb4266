"""Shared test helpers: rebuild golden fixtures as stand-in graphs, pose metrics."""
import os

import numpy as np

import refgraph

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, f"ref_local_ba_{name}.npz"))


def graph_from_golden(g):
    """Stand-in KeyFrame/MapPoint graph with the same insertion orders as the fixture."""
    fx, fy, cx, cy = [float(v) for v in g["intr"]]
    cam = refgraph.Camera(1280, 720, fx, fy, cx, cy)
    n_all = g["R"].shape[0]
    kfs = []
    for i in range(n_all):
        kp = [refgraph.KeyPoint(u, v) for (u, v) in g["kp_flat"][g["kp_off"][i]:g["kp_off"][i + 1]]]
        kfs.append(refgraph.KeyFrame(cam, [kp], g["R"][i], g["t"][i], float(i)))
    mps = []
    for j in range(g["X"].shape[0]):
        mp = refgraph.MapPoint(g["X"][j])
        for (i, cam_id, kp_idx) in g["ob_flat"][g["ob_off"][j]:g["ob_off"][j + 1]]:
            i, cam_id, kp_idx = int(i), int(cam_id), int(kp_idx)
            if cam_id == 0 and kp_idx < len(kfs[i].keypoints[0]):
                kfs[i].add_map_point(cam_id, kp_idx, mp)
            else:
                mp.add_observation(kfs[i].keyframe_id, cam_id, kp_idx)
        mps.append(mp)
    return cam, kfs, mps


def window_points(kfs_window):
    """Map.optimize_local's collection (map/map.py:159-169)."""
    unique = {}
    for kf in kfs_window:
        for mp in kf.get_all_mappoints().values():
            unique[id(mp)] = mp
    return list(unique.values())


def rot_angle(Ra, Rb):
    dR = np.einsum("nij,nkj->nik", Ra, Rb)
    c = np.clip((np.trace(dR, axis1=1, axis2=2) - 1.0) / 2.0, -1.0, 1.0)
    # small-angle accurate: use the skew part
    s = 0.5 * np.sqrt((dR[:, 2, 1] - dR[:, 1, 2]) ** 2 + (dR[:, 0, 2] - dR[:, 2, 0]) ** 2
                      + (dR[:, 1, 0] - dR[:, 0, 1]) ** 2)
    return np.arctan2(s, c)


def scale_norm_trans_err(Ra, ta, Rb, tb):
    """||t_a - t_b|| / median camera-centre baseline."""
    cb = -np.einsum("nji,nj->ni", Rb, tb)
    base = np.linalg.norm(cb[1:] - cb[:-1], axis=1) if cb.shape[0] > 1 else np.array([1.0])
    med = float(np.median(base)) if base.size and np.median(base) > 0 else 1.0
    return np.linalg.norm(ta - tb, axis=1) / med


def gauge_aligned_trans_err(Ra, ta, Rb, tb):
    """Translation error after removing the free monocular scale (one similarity dof about the
    fixed first camera), divided by the median baseline.  Used for long trajectories, where the
    scale gauge is so weakly damped that no float64 solver pins it to 1e-5 of a 0.4 m baseline."""
    ca = -np.einsum("nji,nj->ni", Ra, ta)
    cb = -np.einsum("nji,nj->ni", Rb, tb)
    da, db = ca - ca[0], cb - cb[0]
    s = float((da * db).sum() / max((db * db).sum(), 1e-300))
    base = np.linalg.norm(cb[1:] - cb[:-1], axis=1)
    med = float(np.median(base)) if base.size and np.median(base) > 0 else 1.0
    return np.linalg.norm(da - s * db, axis=1) / med, s
