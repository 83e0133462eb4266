"""SURVEY 8f 'next' rows on the CPU: packer fast path (N1), two-view initial refinement (N2),
pose-only call-site format (N3), outlier consumption (N4), the local-mapping call site (A8)."""
import os

import numpy as np
import pytest

from oracle import ref_adam
import refgraph
from visual_slam_core_b200 import B200Optimizer, handlers, motion, packing, synth
from visual_slam_core_b200.solver import SolveResult

from helpers import GOLDEN, graph_from_golden, load_golden, window_points


def _assert_same_window(a, b):
    for f in ("intr", "R", "t", "fixed", "X", "obs_cam", "obs_pt", "obs_uv", "obs_cam_id", "pt_offsets"):
        np.testing.assert_array_equal(getattr(a, f), getattr(b, f), err_msg=f)


# ---------------------------------------------------------------------------------------- N1
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_observation_table_packs_like_the_reference_loop(name):
    table = packing.ObservationTable(capacity=8)             # small: exercises growth
    table.install(refgraph.MapPoint)
    try:
        g = load_golden(name)
        cam, kfs, mps = graph_from_golden(g)
    finally:
        table.uninstall()
    window = kfs[: int(g["n_kf"])]
    pts = window_points(window)
    _assert_same_window(table.pack(window, pts), packing.pack_window(window, pts))
    # a different window over the same table (keyframes 1..), gauge moves with it
    if len(kfs) > 3:
        w2 = kfs[1:]
        p2 = window_points(w2)
        _assert_same_window(table.pack(w2, p2), packing.pack_window(w2, p2))


def test_observation_table_multi_sensor_order_and_removal():
    table = packing.ObservationTable()
    table.install(refgraph.KeyFrame)
    try:
        cam = refgraph.Camera(1280, 720, 700, 700, 640, 360)
        kps = lambda n: [refgraph.KeyPoint(10.0 * i, 5.0 * i) for i in range(n)]
        kfs = [refgraph.KeyFrame(cam, [kps(8), kps(8)], np.eye(3), np.zeros(3)) for _ in range(3)]
        mps = [refgraph.MapPoint([0.1 * j, 0.0, 5.0]) for j in range(4)]
        # interleaved insertion: the reference groups a point's observations by keyframe
        kfs[0].add_map_point(0, 1, mps[0]); kfs[1].add_map_point(0, 2, mps[0]); kfs[0].add_map_point(1, 3, mps[0])
        kfs[2].add_map_point(1, 0, mps[1]); kfs[0].add_map_point(0, 4, mps[1])
        kfs[1].add_map_point(0, 5, mps[2]); kfs[1].add_map_point(0, 6, mps[2])      # second write: first wins
        kfs[2].add_map_point(0, 7, mps[3]); kfs[1].add_map_point(1, 7, mps[3])
        _assert_same_window(table.pack(kfs, mps), packing.pack_window(kfs, mps))
        kfs[0].remove_map_point(0, 1)                                               # map/keyframe.py:93-100
        _assert_same_window(table.pack(kfs, mps), packing.pack_window(kfs, mps))
    finally:
        table.uninstall()


def test_observation_table_readd_after_removal_goes_to_the_end_like_the_dict():
    """`remove_map_point` then re-adding the same (keyframe, point) must sort the observation where the
    reference's dictionary now has it (at the end), not at its old row."""
    table = packing.ObservationTable()
    table.install(refgraph.MapPoint)
    try:
        cam = refgraph.Camera(1280, 720, 700, 700, 640, 360)
        kps = lambda n: [refgraph.KeyPoint(10.0 * i, 5.0 * i) for i in range(n)]
        kfs = [refgraph.KeyFrame(cam, [kps(8)], np.eye(3), np.zeros(3)) for _ in range(3)]
        mp = refgraph.MapPoint([0.0, 0.0, 5.0])
        for i in range(3):
            kfs[i].add_map_point(0, i, mp)
        kfs[0].remove_map_point(0, 0)
        kfs[0].add_map_point(0, 3, mp)
        pw = table.pack(kfs, [mp])
        _assert_same_window(pw, packing.pack_window(kfs, [mp]))
        np.testing.assert_array_equal(pw.obs_cam, [1, 2, 0])
        # two sensors of one keyframe: removing one keeps the keyframe's place, the re-added sensor goes last in it
        kfs2 = [refgraph.KeyFrame(cam, [kps(8), kps(8)], np.eye(3), np.zeros(3)) for _ in range(2)]
        mp2 = refgraph.MapPoint([0.1, 0.0, 5.0])
        kfs2[0].add_map_point(0, 1, mp2); kfs2[0].add_map_point(1, 2, mp2); kfs2[1].add_map_point(0, 3, mp2)
        kfs2[0].remove_map_point(0, 1); kfs2[0].add_map_point(0, 4, mp2)
        _assert_same_window(table.pack(kfs2, [mp2]), packing.pack_window(kfs2, [mp2]))
        mp2.remove_observation(kfs2[0].keyframe_id)             # whole keyframe entry (map/observation.py:36-38)
        _assert_same_window(table.pack(kfs2, [mp2]), packing.pack_window(kfs2, [mp2]))
    finally:
        table.uninstall()


def test_observation_table_refuses_duplicate_points_and_the_plugin_walks_instead():
    """a point handed in twice needs its rows twice; the table says so and the walk takes over."""
    table = packing.ObservationTable()
    table.install(refgraph.MapPoint)
    try:
        p = synth.make_problem("T0", seed=5)
        cam, kfs, mps = refgraph.graph_from_problem(p)
    finally:
        table.uninstall()
    dup = mps + [mps[0]]
    assert table.pack(kfs, dup) is None and table.stats["fallbacks"] == 1
    opt = B200Optimizer(log_dir=None)
    opt.table = table
    pw = opt.pack(kfs, dup)
    assert opt.packer_name == "pack_window"
    _assert_same_window(pw, packing.pack_window(kfs, dup))
    assert opt.pack(kfs, mps) is not None and opt.packer_name == "ObservationTable"


def test_observation_table_ingests_a_graph_built_before_the_hooks():
    """The plugin installs the hooks on its first call: points that already exist are mirrored by one walk over
    their own dictionaries, points created later through the hooks, and both kinds mix in one window."""
    g = load_golden("b")
    cam, kfs, mps = graph_from_golden(g)                      # built with no table anywhere
    window = kfs[: int(g["n_kf"])]
    pts = window_points(window)
    opt = B200Optimizer(log_dir=None)
    opt.table = packing.ObservationTable()
    try:
        pw = opt.pack(window, pts)
        assert opt.packer_name == "ObservationTable" and opt.table.stats["ingested_points"] == len(pts)
        _assert_same_window(pw, packing.pack_window(window, pts))
        # the map grows: a new point (hook from its first observation) and a new observation of an old point
        new = refgraph.MapPoint([0.3, 0.1, 7.0])
        window[1].keypoints[0].append(refgraph.KeyPoint(11.0, 12.0))
        window[2].keypoints[0].append(refgraph.KeyPoint(13.0, 14.0))
        window[1].add_map_point(0, len(window[1].keypoints[0]) - 1, new)
        window[2].add_map_point(0, len(window[2].keypoints[0]) - 1, new)
        window[0].keypoints[0].append(refgraph.KeyPoint(15.0, 16.0))
        old = pts[3]
        old.remove_observation(window[0].keyframe_id)
        window[0].add_map_point(0, len(window[0].keypoints[0]) - 1, old)
        pts2 = window_points(window)
        n_ing = opt.table.stats["ingested_points"]
        pw2 = opt.pack(window, pts2)
        assert opt.packer_name == "ObservationTable" and opt.table.stats["ingested_points"] == n_ing
        _assert_same_window(pw2, packing.pack_window(window, pts2))
    finally:
        opt.table.uninstall()


def test_observation_table_is_safe_under_concurrent_map_growth():
    """Hooks fire from the tracking / local-mapping threads while pack() runs in the local-BA thread."""
    import threading
    table = packing.ObservationTable(capacity=16)
    table.install(refgraph.MapPoint)
    try:
        cam = refgraph.Camera(1280, 720, 700, 700, 640, 360)
        kfs = [refgraph.KeyFrame(cam, [[refgraph.KeyPoint(float(i), float(i)) for i in range(4000)]], np.eye(3), np.zeros(3))
               for _ in range(3)]
        mps = [refgraph.MapPoint([0.01 * j, 0.0, 5.0]) for j in range(3000)]
        errors = []

        def grow():
            try:
                for j, mp in enumerate(mps):
                    for i in range(3):
                        kfs[i].add_map_point(0, j, mp)
            except Exception as e:          # pragma: no cover
                errors.append(e)

        th = threading.Thread(target=grow)
        th.start()
        while th.is_alive():
            pw = table.pack(kfs, mps[:500])
            assert pw is not None and int(pw.pt_offsets[-1]) == pw.n_obs
        th.join()
        assert not errors
        _assert_same_window(table.pack(kfs, mps), packing.pack_window(kfs, mps))
    finally:
        table.uninstall()


def test_observation_table_key_lookup_paths_and_empty_windows():
    """Small integer ids go through look-up tables, huge ids (object identities, sparse counters) through a
    sorted search; an empty point list and a window without recorded observations pack to nothing."""
    cam = refgraph.Camera(1280, 720, 700, 700, 640, 360)

    class Point:                                        # no `id`: the object identity is the key
        def __init__(self, x):
            self.position = np.asarray(x, dtype=np.float64)

        def get_observations(self):
            return self.obs

    kps = [refgraph.KeyPoint(1.0, 2.0), refgraph.KeyPoint(3.0, 4.0)]
    kfs = [refgraph.KeyFrame(cam, [list(kps)], np.eye(3), np.zeros(3)) for _ in range(2)]
    kfs[1].keyframe_id = 900_000_000_000                # far outside any look-up table
    pts = [Point([0, 0, 1]), Point([1, 0, 2])]
    table = packing.ObservationTable()
    table.add(kfs[0], 0, 0, pts[0]); table.add(kfs[1], 0, 1, pts[0]); table.add(kfs[1], 0, 0, pts[1])
    pts[0].obs = {kfs[0].keyframe_id: {0: 0}, kfs[1].keyframe_id: {0: 1}}
    pts[1].obs = {kfs[1].keyframe_id: {0: 0}}
    pw = table.pack(kfs, pts)
    _assert_same_window(pw, packing.pack_window(kfs, pts))
    np.testing.assert_array_equal(pw.obs_cam, [0, 1, 1])
    np.testing.assert_array_equal(pw.pt_offsets, [0, 2, 3])
    assert table.pack(kfs, []).n_obs == 0
    stranger = [refgraph.KeyFrame(cam, [[]], np.eye(3), np.zeros(3))]
    empty = table.pack(stranger, pts)
    assert empty.n_obs == 0 and empty.pt_offsets.tolist() == [0, 0, 0]


def test_c_packer_loop_equals_the_python_loop(monkeypatch):
    """`csrc/fastpack.c` walks the object graph exactly as `pack_window`'s Python loop does (iteration order,
    the two skip rules, negative indices as Python indexes them); containers that are not exact dicts / lists
    send the call back to the Python loop."""
    from visual_slam_core_b200 import _build
    path = _build.fastpack_path()
    before = os.path.getmtime(path) if os.path.exists(path) else None
    try:
        _build.build_fastpack()
    except RuntimeError as e:                              # no compiler / no Python.h on this machine
        pytest.skip(str(e))
    if packing._fastpack is not None and os.path.getmtime(path) != before:
        pytest.skip("packer extension rebuilt from newer sources; this process still holds the old one - re-run")
    fp = packing._load_fastpack()
    assert fp is not None

    def both(kfs, pts):
        monkeypatch.setattr(packing, "_fastpack", fp)
        a = packing.pack_window(kfs, pts)
        monkeypatch.setattr(packing, "_fastpack", None)
        b = packing.pack_window(kfs, pts)
        _assert_same_window(a, b)
        for f in ("obs_cam", "obs_pt", "obs_uv", "obs_cam_id", "pt_offsets"):
            assert getattr(a, f).dtype == getattr(b, f).dtype and getattr(a, f).flags["WRITEABLE"]
        return a

    for name in ("a", "b", "c"):
        g = load_golden(name)
        cam, kfs, mps = graph_from_golden(g)
        window = kfs[: int(g["n_kf"])]
        both(window, window_points(window))
        both(kfs[1:], window_points(kfs[1:]))              # keyframes outside the window are skipped
    # two sensors, out-of-range sensor / keypoint indices (skipped), a negative keypoint index (last keypoint)
    cam = refgraph.Camera(1280, 720, 700, 700, 640, 360)
    kps = lambda n: [refgraph.KeyPoint(10.0 * i, 5.0 * i) for i in range(n)]
    kfs = [refgraph.KeyFrame(cam, [kps(4), kps(3)], np.eye(3), np.zeros(3)) for _ in range(3)]
    mps = [refgraph.MapPoint([0.1 * j, 0.0, 5.0]) for j in range(3)]
    mps[0].add_observation(kfs[0].keyframe_id, 0, 1); mps[0].add_observation(kfs[1].keyframe_id, 1, 2)
    mps[0].add_observation(kfs[0].keyframe_id, 1, 0)
    mps[1].add_observation(kfs[2].keyframe_id, 2, 0)        # sensor 2 does not exist
    mps[1].add_observation(kfs[2].keyframe_id, 1, 3)        # keypoint 3 of sensor 1 does not exist
    mps[1].add_observation(kfs[1].keyframe_id, 0, -1)       # Python indexes from the end
    mps[2].add_observation(12345678, 0, 0)                  # a keyframe that is not in the window
    pw = both(kfs, mps)
    assert pw.n_obs == 4 and pw.pt_offsets.tolist() == [0, 3, 4, 4]
    np.testing.assert_array_equal(pw.obs_uv[3], [30.0, 15.0])

    # cv2's detectors return TUPLES of cv2.KeyPoint (feature/feature.py:33,47): same path, same result
    cv2 = pytest.importorskip("cv2")
    real = lambda n: tuple(cv2.KeyPoint(10.5 * i, 5.25 * i, 1.0) for i in range(n))
    kfs_cv = [refgraph.KeyFrame(cam, [real(4), real(3)], np.eye(3), np.zeros(3)) for _ in range(3)]
    mps_cv = [refgraph.MapPoint([0.1 * j, 0.0, 5.0]) for j in range(2)]
    mps_cv[0].add_observation(kfs_cv[0].keyframe_id, 0, 1); mps_cv[0].add_observation(kfs_cv[2].keyframe_id, 1, -1)
    mps_cv[1].add_observation(kfs_cv[1].keyframe_id, 1, 5); mps_cv[1].add_observation(kfs_cv[1].keyframe_id, 0, 3)
    index = {kf.keyframe_id: i for i, kf in enumerate(kfs_cv)}
    assert fp.pack_observations(mps_cv, index, [kf.keypoints for kf in kfs_cv]) is not None
    pw_cv = both(kfs_cv, mps_cv)
    np.testing.assert_array_equal(pw_cv.obs_uv, np.array([[10.5, 5.25], [21.0, 10.5], [31.5, 15.75]], np.float32))

    class Broken(refgraph.MapPoint):                         # an exception inside the walk surfaces unchanged
        def get_observations(self):
            raise KeyError("lost")

    for impl in (fp, None):
        monkeypatch.setattr(packing, "_fastpack", impl)
        with pytest.raises(KeyError):
            packing.pack_window(kfs, mps + [Broken([0.0, 0.0, 1.0])])

    class Observed(dict):                                   # not an exact dict: the C loop declines
        pass

    orig = mps[0].get_observations
    mps[0].get_observations = lambda: Observed(orig())
    assert fp.pack_observations(mps, {kf.keyframe_id: i for i, kf in enumerate(kfs)}, [kf.keypoints for kf in kfs]) is None
    both(kfs, mps)
    assert both(kfs, []).n_obs == 0


def test_fast_path_is_faster_on_a_real_window(monkeypatch):
    """The table beats the interpreter's walk of the object graph (best of three after one warm-up call each,
    so that first-use imports are not what is timed)."""
    import time
    p = synth.make_problem("T2", seed=3)
    table = packing.ObservationTable()
    table.install(refgraph.KeyFrame)
    try:
        cam, kfs, mps = refgraph.graph_from_problem(p)
    finally:
        table.uninstall()
    pts = window_points(kfs)
    monkeypatch.setattr(packing, "_fastpack", None)          # the Python loop

    def best(fn):
        out, t_best = fn(), float("inf")
        for _ in range(3):
            t0 = time.perf_counter(); out = fn(); t_best = min(t_best, time.perf_counter() - t0)
        return out, t_best

    a, t_loop = best(lambda: packing.pack_window(kfs, pts))
    b, t_table = best(lambda: table.pack(kfs, pts))
    _assert_same_window(b, a)
    assert t_table < t_loop


# ---------------------------------------------------------------------------------------- N2
def _initial_golden(name):
    return np.load(os.path.join(GOLDEN, f"ref_initial_{name}.npz"))


def test_initial_selection_and_adam_restatement_match_the_reference():
    g = _initial_golden("a")
    cam, kfs, mps = graph_from_golden(g)
    pi = packing.pack_initial(kfs, mps)
    assert bool(g["ref_ok"]) and len(pi.selected) >= packing.MIN_OBSERVATIONS
    out = ref_adam.adam_initial(pi.window.intr, pi.window.R[0], pi.window.t[0], pi.window.R[1], pi.window.t[1],
                                pi.window.X, pi.uv_ref, pi.uv_cur, n_iter=int(g["n_iter"]), lr=float(g["lr"]))
    # the reference writes back kf_cur only (:308-311) ...
    np.testing.assert_allclose(out["R_cur"], g["ref_R_out"][1], atol=2e-6)
    np.testing.assert_allclose(out["t_cur"], g["ref_t_out"][1], atol=2e-6)
    np.testing.assert_array_equal(g["ref_R_out"][0], g["R"][0])
    # ... and pairs the selected points with the FIRST len(selected) map points (zip, :313-314)
    n = len(pi.selected)
    np.testing.assert_allclose(out["X"], g["ref_X_out"][:n], atol=2e-5)
    np.testing.assert_array_equal(g["ref_X_out"][n:], g["X"][n:])
    assert not np.array_equal(pi.selected, np.arange(n))      # the fixture does exercise the misalignment


def test_initial_returns_false_on_few_points_like_the_reference():
    g = _initial_golden("few")
    cam, kfs, mps = graph_from_golden(g)
    assert not bool(g["ref_ok"])
    before = [m.position.copy() for m in mps], kfs[1].t_w2c.copy()
    opt = B200Optimizer(log_dir=None)
    assert opt.optimize_initial(kfs, mps) is False            # early return: no GPU needed
    assert opt.optimize_initial(kfs[:1], mps) is False
    assert opt.optimize_initial(kfs, []) is False
    np.testing.assert_array_equal(kfs[1].t_w2c, before[1])
    for m, x in zip(mps, before[0]):
        np.testing.assert_array_equal(m.position, x)
    np.testing.assert_array_equal(g["ref_X_out"], g["X"])       # and the reference left everything alone too


# ---------------------------------------------------------------------------------------- N3
def test_pose_only_call_site_failure_tuple_matches_the_reference():
    g = np.load(os.path.join(GOLDEN, "ref_pose_only.npz"))
    assert g["ref_few_is_none"].all()                           # reference: (None, None, False, None)
    out = motion.refine_motion_from_2d3d(g["pts3d"][:3], g["pts2d"][:3], g["K"])
    assert out == (None, None, False, None)
    assert g["ref_inliers"].ndim == 2 and g["ref_inliers"].shape[1] == 1 and str(g["ref_inliers_dtype"]) == "int32"
    # the call site passes `self.camera.dist` (tracking.py:350-356) and the reference never raises: neither do we
    # (no GPU in this process: the solver cannot be created, which must come back as the failure tuple too)
    out = motion.refine_motion_from_2d3d(g["pts3d"], g["pts2d"], g["K"], dist=np.array([0.1, 0, 0, 0, 0]),
                                         R0=np.eye(3), t0=np.zeros(3))
    assert isinstance(out, tuple) and len(out) == 4 and (out[2] is True or out == (None, None, False, None))
    assert motion.refine_motion_from_2d3d(g["pts3d"], g["pts2d"][:-1], g["K"]) == (None, None, False, None)


def test_undistortion_of_the_call_sites_pixels_inverts_the_lens_model():
    cv2 = pytest.importorskip("cv2")
    K = np.array([[700.0, 0, 640.0], [0, 700.0, 360.0], [0, 0, 1.0]])
    dist = np.array([0.12, -0.05, 0.001, -0.002, 0.01])
    rng = np.random.default_rng(3)
    X = np.stack([rng.uniform(-2, 2, 50), rng.uniform(-1, 1, 50), rng.uniform(5, 12, 50)], axis=1)
    ideal = (X[:, :2] / X[:, 2:]) * 700.0 + np.array([640.0, 360.0])
    distorted, _ = cv2.projectPoints(X, np.zeros(3), np.zeros(3), K, dist)
    back = motion._undistort(distorted.reshape(-1, 2), K, dist)
    assert np.abs(back - ideal).max() < 1e-3


# ---------------------------------------------------------------------------------------- N4 / A8
def test_mean_reprojection_error_matches_the_reference_formula():
    p = synth.make_problem("T1", seed=4)
    cam, kfs, mps = refgraph.graph_from_problem(p)
    pts = window_points(kfs)
    pw = packing.pack_window(kfs, pts)
    # per keyframe as utils/geometry.py:232-247, weighted by count as map/map.py:137-141
    tot, cnt = 0.0, 0
    K = cam.K
    for i, kf in enumerate(kfs):
        sel = pw.obs_cam == i
        if not sel.any():
            continue
        P3, P2 = pw.X[pw.obs_pt[sel]], pw.obs_uv[sel].astype(float)
        pc = (pw.R[i] @ P3.T + pw.t[i].reshape(3, 1)).T
        pn = pc[:, :2] / pc[:, 2:].clip(min=1e-8)
        proj = (K @ np.concatenate([pn, np.ones((len(pn), 1))], axis=1).T).T[:, :2]
        tot += float(np.mean(np.linalg.norm(P2 - proj, axis=1))) * len(P3)
        cnt += len(P3)
    assert handlers.mean_reprojection_error(pw) == pytest.approx(tot / cnt, rel=1e-6)


def test_cull_bad_points_consumes_the_flags():
    p = synth.make_problem("T0", seed=2)
    cam, kfs, mps = refgraph.graph_from_problem(p)
    pts = window_points(kfs)
    pw = packing.pack_window(kfs, pts)
    flags = np.zeros(pw.n_obs, dtype=np.uint8)
    j_bad = int(np.argmax(np.diff(pw.pt_offsets) >= 2))
    flags[pw.pt_offsets[j_bad]:pw.pt_offsets[j_bad + 1]] = 1    # every observation of one point
    o_single = int(pw.pt_offsets[(j_bad + 1) % len(pts)])       # one observation of another
    flags[o_single] = 1
    opt = B200Optimizer(log_dir=None)
    opt.last_packed = pw
    opt.last_result = SolveResult(R=pw.R, t=pw.t, X=pw.X, flags=flags, cost_hist=np.zeros(1), cand_cost_hist=np.zeros(0),
                                  lambda_hist=np.zeros(0), accepted=np.zeros(0, np.uint8), solve_fail=np.zeros(0, np.int32),
                                  pcg_iters=np.zeros(0, np.int32), n_iter=0, gpu_ms=0.0, n_launches=0)
    n_obs_before = sum(len(c) for m in pts for c in m.get_observations().values())
    n_bad = opt.cull_bad_points(kfs, pts, min_observations=1)
    assert n_bad == 1 and not pts[j_bad].is_valid()
    assert pts[j_bad].get_observations() == {}
    n_obs_after = sum(len(c) for m in pts for c in m.get_observations().values())
    assert n_obs_before - n_obs_after == int(flags.sum())
    kf = kfs[int(pw.obs_cam[o_single])]
    assert pts[int(pw.obs_pt[o_single])] not in kf.get_all_mappoints().values()


def test_cull_bad_points_leaves_points_with_observations_outside_the_window_alone():
    """one in-window observation (unflagged) and two in keyframes outside the window is a healthy
    point; a point that loses its flagged in-window observation but keeps two outside is healthy too."""
    cam = refgraph.Camera(1280, 720, 700, 700, 640, 360)
    kps = lambda n: [refgraph.KeyPoint(10.0 * i, 5.0 * i) for i in range(n)]
    kfs = [refgraph.KeyFrame(cam, [kps(8)], np.eye(3), np.zeros(3)) for _ in range(5)]
    a, b, c = (refgraph.MapPoint([0.1 * j, 0.0, 5.0]) for j in range(3))
    for i, mp in ((0, a), (3, a), (4, a), (1, b), (3, b), (4, b), (0, c), (1, c)):
        kfs[i].add_map_point(0, [a, b, c].index(mp), mp)
    window, pts = kfs[:3], [a, b, c]
    pw = packing.pack_window(window, pts)
    assert pw.n_obs == 4
    flags = np.zeros(4, dtype=np.uint8)
    flags[int(pw.pt_offsets[1])] = 1                               # b's only in-window observation
    flags[int(pw.pt_offsets[2])] = 1                               # one of c's two (c has nothing outside)
    opt = B200Optimizer(log_dir=None)
    opt.last_packed = pw
    opt.last_result = SolveResult(R=pw.R, t=pw.t, X=pw.X, flags=flags, cost_hist=np.zeros(1), cand_cost_hist=np.zeros(0),
                                  lambda_hist=np.zeros(0), accepted=np.zeros(0, np.uint8), solve_fail=np.zeros(0, np.int32),
                                  pcg_iters=np.zeros(0, np.int32), n_iter=0, gpu_ms=0.0, n_launches=0)
    assert opt.cull_bad_points(window, pts, min_observations=2) == 1
    assert a.is_valid() and b.is_valid() and not c.is_valid()
    assert len(b.get_observations()) == 2 and len(a.get_observations()) == 3


def test_local_optimize_call_site_needs_more_than_two_keyframes():
    p = synth.make_problem("T0", seed=2)
    cam, kfs, mps = refgraph.graph_from_problem(p)
    assert handlers.local_optimize(B200Optimizer(log_dir=None), kfs[:2]) is False     # local_handler.py:37-40


# ---------------------------------------------------------------------------------------------
# C helpers of the host path: window collection, table selection, direct write-back
# ---------------------------------------------------------------------------------------------
def _python_collect(keyframes):
    unique = {}
    for kf in keyframes:
        for mp in kf.get_all_mappoints().values():
            unique[id(mp)] = mp
    return list(unique.values())


def test_collect_points_in_c_matches_the_reference_loop():
    fp = packing._fastpack
    if fp is None or not hasattr(fp, "collect_points"):
        pytest.skip("C extension not built")
    p = synth.make_problem("T2", seed=11)
    cam, kfs, mps = refgraph.graph_from_problem(p)
    got = fp.collect_points(kfs)
    want = _python_collect(kfs)
    assert len(got) == len(want) and all(a is b for a, b in zip(got, want))
    assert fp.collect_points(tuple(kfs)) is None                  # not an exact list: the Python loop runs
    assert handlers.collect_window_points(tuple(kfs))[0] is want[0]
    assert fp.collect_points([]) == []


def test_table_select_in_c_equals_the_numpy_selection():
    fp = packing._fastpack
    if fp is None or not hasattr(fp, "table_select"):
        pytest.skip("C extension not built")
    p = synth.make_problem("T2", seed=12)
    table = packing.ObservationTable()
    table.install(refgraph.MapPoint)
    try:
        cam, kfs, mps = refgraph.graph_from_problem(p)
        # tombstones, an out-of-range keypoint index, a keyframe outside the window
        kfs[1].remove_map_point(0, next(iter(kfs[1].map_points))[1])
        mps[3].add_observation(kfs[2].keyframe_id + 1000, 0, 5)
        win = kfs[1:]                                              # the first keyframe's observations drop out
        pts = _python_collect(win)
        ref = packing.pack_window(win, pts)
        got = table.pack(win, pts)
        saved, packing._fastpack = packing._fastpack, None         # the numpy path of the same method
        try:
            slow = table.pack(win, pts)
        finally:
            packing._fastpack = saved
        for a in ("R", "t", "fixed", "X", "obs_cam", "obs_pt", "obs_uv", "obs_cam_id", "pt_offsets"):
            np.testing.assert_array_equal(getattr(got, a), getattr(ref, a), err_msg=a)
            np.testing.assert_array_equal(getattr(slow, a), getattr(ref, a), err_msg=a)
        assert got.obs_cam.dtype == np.int32 and got.obs_uv.dtype == np.float32 and got.pt_offsets.dtype == np.int32
    finally:
        table.uninstall()


def test_direct_write_back_only_for_the_plain_update_position():
    class Plain:
        def __init__(self):
            import threading
            self._lock = threading.RLock()
            self.position = np.zeros(3)

        def update_position(self, new_position):
            with self._lock:
                self.position = np.asarray(new_position, dtype=float)

    class Recording(Plain):
        def update_position(self, new_position):
            super().update_position(new_position)
            self.touched = True

    assert packing._plain_update_position(Plain) and not packing._plain_update_position(Recording)
    assert packing._plain_update_position(refgraph.MapPoint)
    X = np.arange(12, dtype=np.float64).reshape(4, 3)
    a, b = [Plain() for _ in range(4)], [Recording() for _ in range(4)]
    packing.write_back_points(a, X)
    packing.write_back_points(b, X)
    mixed = [Plain(), Recording(), Plain(), Recording()]
    packing.write_back_points(mixed, X)
    for lst in (a, b, mixed):
        for j, mp in enumerate(lst):
            np.testing.assert_array_equal(mp.position, X[j])
            assert mp.position.dtype == np.float64
    assert all(getattr(mp, "touched", False) for mp in b) and mixed[1].touched and mixed[3].touched
