"""Generate golden fixtures by running the UNMODIFIED reference in this container.

Run from the repo root (needs /root/reference, CPU only):

    python tests/golden/make_golden.py

Writes tests/golden/ref_local_ba_*.npz.  Each fixture holds
  * a small KeyFrame/MapPoint graph (poses, keypoints, observation dicts in insertion
    order, including the irregular cases the reference's packing loop skips:
    observations in keyframes outside the window, out-of-range kp_idx / cam_id),
  * the observation list the reference built for it (captured from the live frame of
    `TorchOptimizer.optimize_local`, torch_optimizer.py:74-95),
  * the loss the reference logged at iter 0 (`:182-183`, 3 decimals),
  * the poses / points the reference wrote back after n_iter Adam steps (`:188-203`).

The reference is imported with the matplotlib.pylab stub of SURVEY.md Appendix A.
Nothing here is imported by the product; tests only read the .npz files.
"""
from __future__ import annotations

import logging
import os
import re
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    mpl = types.ModuleType("matplotlib")
    pylab = types.ModuleType("matplotlib.pylab")
    pylab.ndarray = np.ndarray
    mpl.pylab = pylab
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pylab", pylab)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import cv2  # noqa: F401
    from visual_slam.config import Config, OptimizationConfig
    from visual_slam.camera import PinholeCamera
    from visual_slam.map.keyframe import KeyFrame
    from visual_slam.map.map_point import MapPoint
    from visual_slam.map.map import Map
    from visual_slam.map.pose import Pose
    from visual_slam.optimization.torch_optimizer import TorchOptimizer
    return dict(Config=Config, OptimizationConfig=OptimizationConfig, PinholeCamera=PinholeCamera,
                KeyFrame=KeyFrame, MapPoint=MapPoint, Map=Map, Pose=Pose, TorchOptimizer=TorchOptimizer)


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def build_case(seed, n_kf, n_pt, n_extra_kf=1, irregular=True):
    """Plain-array description of a window (no reference types)."""
    rng = np.random.default_rng(seed)
    fx = fy = 700.0
    cx, cy = 640.0, 360.0
    n_all = n_kf + n_extra_kf
    R_all, t_all, C = [], [], []
    for i in range(n_all):
        R_c2w = rot_y(0.01 * i)
        c = np.array([0.4 * i, 0.0, 0.05 * i])
        R = R_c2w.T
        R_all.append(R)
        t_all.append(-R @ c)
        C.append(c)
    Xw = np.stack([rng.uniform(-3, 5, n_pt), rng.uniform(-2, 2, n_pt), rng.uniform(6, 25, n_pt)], axis=1)
    kps = [[] for _ in range(n_all)]           # per KF: list of (u,v)
    obs = [[] for _ in range(n_pt)]            # per point: [(kf_local, cam_id, kp_idx)] insertion order
    for j in range(n_pt):
        k = int(rng.integers(2, n_all + 1))
        cams = rng.permutation(n_all)[:k]      # insertion order deliberately not sorted
        for i in cams:
            Xc = R_all[i] @ Xw[j] + t_all[i]
            u = fx * Xc[0] / Xc[2] + cx + rng.normal(0, 1.0)
            v = fy * Xc[1] / Xc[2] + cy + rng.normal(0, 1.0)
            if rng.random() < 0.08:
                u += rng.normal(0, 30.0)
                v += rng.normal(0, 30.0)
            kps[i].append((np.float32(u), np.float32(v)))
            obs[j].append((int(i), 0, len(kps[i]) - 1))
    if irregular:
        # out-of-range keypoint index and out-of-range camera id: both must be skipped
        obs[0].append((1, 0, 10_000))          # kp_idx >= len(keypoints[0])   (:92-93)
        obs[1].append((2, 3, 0))               # cam_id >= len(keypoints)      (:90-91)
    X0 = Xw + rng.normal(0, 0.05, Xw.shape)
    R0, t0 = [], []
    for i in range(n_all):
        if i == 0:
            R0.append(R_all[i])
            t0.append(t_all[i])
        else:
            w = rng.normal(0, 0.01, 3)
            th = np.linalg.norm(w)
            K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
            dR = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * K @ K
            R0.append(dR @ R_all[i])
            t0.append(t_all[i] + rng.normal(0, 0.05, 3))
    return dict(intr=np.array([fx, fy, cx, cy]), R=np.array(R0), t=np.array(t0), X=X0,
                kps=kps, obs=obs, n_kf=n_kf, n_all=n_all)


def to_reference_graph(ref, case):
    import cv2
    cam = ref["PinholeCamera"](1280, 720, *[float(v) for v in case["intr"]])
    kfs = []
    for i in range(case["n_all"]):
        kp = [cv2.KeyPoint(float(u), float(v), 1.0) for (u, v) in case["kps"][i]]
        kf = ref["KeyFrame"](timestamp=float(i), camera=cam, images=[], keypoints=[kp],
                             descriptors=[None], pose=ref["Pose"](R=case["R"][i], t=case["t"][i]))
        kfs.append(kf)
    mps = []
    for j in range(case["X"].shape[0]):
        mp = ref["MapPoint"](position=case["X"][j])
        for (i, cam_id, kp_idx) in case["obs"][j]:
            if cam_id == 0 and kp_idx < len(case["kps"][i]):
                kfs[i].add_map_point(cam_id, kp_idx, mp)       # also records the observation
            else:
                mp.add_observation(kfs[i].keyframe_id, cam_id, kp_idx)
        mps.append(mp)
    return cam, kfs, mps


class _Capture(logging.Handler):
    def __init__(self):
        super().__init__()
        self.lines = []

    def emit(self, record):
        self.lines.append(record.getMessage())


def run_reference(ref, case, n_iter, lr=1e-4, delta=5.0):
    cam, kfs, mps = to_reference_graph(ref, case)
    window = kfs[: case["n_kf"]]
    tmp = tempfile.mkdtemp()
    cfg = ref["Config"](optimization=ref["OptimizationConfig"](lr=lr, n_iter=n_iter, huber_delta=delta))
    opt = ref["TorchOptimizer"](cfg, device="cpu", log_dir=tmp)
    cap = _Capture()
    opt.logger.addHandler(cap)
    grabbed = {}

    def prof(frame, event, arg):
        if event == "return" and frame.f_code.co_name == "optimize_local" and "obs" in frame.f_locals:
            grabbed["obs"] = list(frame.f_locals["obs"])
            grabbed["n_points"] = len(frame.f_locals["map_points"])
            grabbed["pt_ids"] = [id(m) for m in frame.f_locals["map_points"]]

    sys.setprofile(prof)
    try:
        ok = ref["Map"]().optimize_local(opt, window)
    finally:
        sys.setprofile(None)
    loss0 = None
    for ln in cap.lines:
        m = re.search(r"iter=0, loss=([-0-9.eE+]+)", ln)
        if m:
            loss0 = float(m.group(1))
    kfid_to_local = {kf.keyframe_id: i for i, kf in enumerate(kfs)}
    id_to_j = {id(m): j for j, m in enumerate(mps)}
    ref_obs = np.array([[kfid_to_local[k], j, c, uv[0], uv[1]] for (k, j, c, uv) in grabbed["obs"]],
                       dtype=np.float64)
    # point order handed to the optimizer by Map.optimize_local (map.py:159-169)
    pt_order = np.array([id_to_j[i] for i in grabbed["pt_ids"]], dtype=np.int64)
    R_out = np.array([kf.R_w2c for kf in window])
    t_out = np.array([kf.t_w2c for kf in window])
    X_out = np.array([m.position for m in mps])
    return dict(ok=ok, loss0=loss0, ref_obs=ref_obs, pt_order=pt_order,
                R_out=R_out, t_out=t_out, X_out=X_out,
                kf_ids=np.array([kf.keyframe_id for kf in kfs], dtype=np.int64))


def save_case(name, case, out):
    n_all = case["n_all"]
    kp_off = np.zeros(n_all + 1, dtype=np.int64)
    kp_off[1:] = np.cumsum([len(k) for k in case["kps"]])
    kp_flat = np.array([p for k in case["kps"] for p in k], dtype=np.float32).reshape(-1, 2)
    ob_off = np.zeros(len(case["obs"]) + 1, dtype=np.int64)
    ob_off[1:] = np.cumsum([len(o) for o in case["obs"]])
    ob_flat = np.array([e for o in case["obs"] for e in o], dtype=np.int64).reshape(-1, 3)
    path = os.path.join(HERE, f"ref_local_ba_{name}.npz")
    np.savez_compressed(
        path, intr=case["intr"], R=case["R"], t=case["t"], X=case["X"], n_kf=case["n_kf"],
        kp_off=kp_off, kp_flat=kp_flat, ob_off=ob_off, ob_flat=ob_flat,
        kf_ids=out["kf_ids"], ref_ok=out["ok"], ref_loss0=out["loss0"], ref_obs=out["ref_obs"],
        ref_pt_order=out["pt_order"], ref_R_out=out["R_out"], ref_t_out=out["t_out"],
        ref_X_out=out["X_out"], n_iter=out["n_iter"], lr=out["lr"], delta=out["delta"])
    print("wrote", path, "obs", out["ref_obs"].shape[0], "loss0", out["loss0"])


def run_reference_initial(ref, case, n_iter, lr):
    """`TorchOptimizer.optimize_initial` (torch_optimizer.py:209-317) on the first two keyframes and
    ALL points of the case, the way `Map.optimize_initial` (map/map.py:143-148) hands them over."""
    cam, kfs, mps = to_reference_graph(ref, case)
    tmp = tempfile.mkdtemp()
    cfg = ref["Config"](optimization=ref["OptimizationConfig"](lr=lr, n_iter=n_iter))
    opt = ref["TorchOptimizer"](cfg, device="cpu", log_dir=tmp)
    ok = opt.optimize_initial(kfs, mps)
    return dict(ok=ok, R_out=np.array([kf.R_w2c for kf in kfs]), t_out=np.array([kf.t_w2c for kf in kfs]),
                X_out=np.array([m.position for m in mps]),
                kf_ids=np.array([kf.keyframe_id for kf in kfs], dtype=np.int64))


def save_initial(name, case, out, n_iter, lr):
    n_all = case["n_all"]
    kp_off = np.zeros(n_all + 1, dtype=np.int64)
    kp_off[1:] = np.cumsum([len(k) for k in case["kps"]])
    kp_flat = np.array([p for k in case["kps"] for p in k], dtype=np.float32).reshape(-1, 2)
    ob_off = np.zeros(len(case["obs"]) + 1, dtype=np.int64)
    ob_off[1:] = np.cumsum([len(o) for o in case["obs"]])
    ob_flat = np.array([e for o in case["obs"] for e in o], dtype=np.int64).reshape(-1, 3)
    path = os.path.join(HERE, f"ref_initial_{name}.npz")
    np.savez_compressed(
        path, intr=case["intr"], R=case["R"], t=case["t"], X=case["X"], n_kf=case["n_kf"],
        kp_off=kp_off, kp_flat=kp_flat, ob_off=ob_off, ob_flat=ob_flat, kf_ids=out["kf_ids"],
        ref_ok=out["ok"], ref_R_out=out["R_out"], ref_t_out=out["t_out"], ref_X_out=out["X_out"],
        n_iter=n_iter, lr=lr)
    print("wrote", path, "ok", out["ok"])


def make_pose_only_golden():
    """`estimate_motion_from_2d3d` (utils/motion_estimation.py:57-132) on synthetic frames: the
    return tuple (shapes, dtypes, failure cases) and the pose OpenCV finds, as a sanity
    cross-check of the pose-only LM (OpenCV minimises plain L2 after RANSAC, so only loosely comparable)."""
    import cv2
    from visual_slam.utils.motion_estimation import estimate_motion_from_2d3d
    rng = np.random.default_rng(21)
    fx = fy = 700.0
    cx, cy = 640.0, 360.0
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    n = 300
    Xw = np.stack([rng.uniform(-4, 4, n), rng.uniform(-2, 2, n), rng.uniform(6, 25, n)], axis=1)
    R = rot_y(0.07).T
    t = np.array([0.3, -0.05, 0.2])
    Xc = Xw @ R.T + t
    uv = np.stack([fx * Xc[:, 0] / Xc[:, 2] + cx, fy * Xc[:, 1] / Xc[:, 2] + cy], axis=1)
    uv += rng.normal(0, 0.5, uv.shape)
    out_idx = rng.choice(n, 30, replace=False)
    uv[out_idx] += rng.normal(0, 40.0, (30, 2))
    pts3d = Xw.astype(np.float32)
    pts2d = uv.astype(np.float32)
    cv2.setRNGSeed(42)
    R_cv, t_cv, ok, inl = estimate_motion_from_2d3d(pts3d, pts2d, K, None, True)
    few = estimate_motion_from_2d3d(pts3d[:3], pts2d[:3], K)
    path = os.path.join(HERE, "ref_pose_only.npz")
    np.savez_compressed(path, K=K, pts3d=pts3d, pts2d=pts2d, R_true=R, t_true=t, outliers=np.sort(out_idx),
                        ref_R=R_cv, ref_t=t_cv, ref_ok=ok, ref_inliers=inl,
                        ref_inliers_dtype=str(inl.dtype), ref_few_is_none=np.array([x is None for x in few[:2]] + [few[2] is False, few[3] is None]))
    print("wrote", path, "ok", ok, "inliers", None if inl is None else inl.shape, inl.dtype)


def main():
    ref = import_reference()
    # two-view initial refinement (torch_optimizer.py:209-317): irregular observations included so the
    # selection rule (:240-258) and the zip() write-back (:313-314) are exercised
    for name, cs, n_iter, lr in [("a", dict(seed=31, n_kf=2, n_pt=60, n_extra_kf=1), 3, 1e-3),
                                 ("few", dict(seed=32, n_kf=2, n_pt=6, n_extra_kf=0, irregular=False), 2, 1e-3)]:
        case = build_case(**cs)
        out = run_reference_initial(ref, case, n_iter, lr)
        save_initial(name, case, out, n_iter, lr)
    make_pose_only_golden()
    specs = [
        ("a", dict(seed=11, n_kf=4, n_pt=40, n_extra_kf=1), dict(n_iter=1, lr=1e-4, delta=5.0)),
        ("b", dict(seed=12, n_kf=6, n_pt=120, n_extra_kf=2), dict(n_iter=3, lr=1e-3, delta=5.0)),
        ("c", dict(seed=13, n_kf=3, n_pt=25, n_extra_kf=0), dict(n_iter=2, lr=1e-3, delta=2.0)),
        # fewer than 10 observations -> reference returns False and mutates nothing (:99-101)
        ("few", dict(seed=14, n_kf=2, n_pt=3, n_extra_kf=0, irregular=False), dict(n_iter=1, lr=1e-3, delta=5.0)),
    ]
    for name, cs, rs in specs:
        case = build_case(**cs)
        out = run_reference(ref, case, **rs)
        out.update(rs)
        if out["loss0"] is None:
            out["loss0"] = np.nan
        save_case(name, case, out)


if __name__ == "__main__":
    main()
