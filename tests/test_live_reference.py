"""Packers against the UNMODIFIED reference's own objects (runs only where /root/reference is mounted, i.e. in
the build container; the GPU box and any other checkout skip it).  The golden fixtures pin the same order
rules offline; this one makes sure the real `KeyFrame` / `MapPoint` / `cv2.KeyPoint` types take the fast paths."""
import os
import sys
import types

import numpy as np
import pytest

from visual_slam_core_b200 import packing, synth

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "visual_slam")),
                                reason="the reference tree is not mounted here")


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
    table.install(KeyFrame)
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
