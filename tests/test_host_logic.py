"""Host-side logic that needs no GPU: packing, write-back, plugin error behaviour, sharding."""
import numpy as np
import pytest

from oracle import lm_oracle as O
import refgraph
from visual_slam_core_b200 import B200Optimizer, packing, synth

from helpers import graph_from_golden, load_golden, window_points


def test_pack_roundtrip_from_problem():
    p = synth.make_problem("T1", seed=5)
    cam, kfs, mps = refgraph.graph_from_problem(p)
    pts = window_points(kfs)
    pw = packing.pack_window(kfs, pts)
    # Map.optimize_local's collection order differs from the generator's point order: map back
    order = np.array([mps.index(m) for m in pts])
    assert sorted(order.tolist()) == list(range(p.n_pt))
    np.testing.assert_array_equal(pw.X, p.X[order])
    np.testing.assert_array_equal(pw.R, p.R)
    assert pw.fixed.tolist() == [1] + [0] * (p.n_cam - 1)
    # same multiset of observations, same objective
    c0 = O.reference_cost(p.intr, p.R, p.t, p.X, p.obs_cam, p.obs_pt, p.obs_uv)
    c1 = O.reference_cost(pw.intr, pw.R, pw.t, pw.X, pw.obs_cam, pw.obs_pt, pw.obs_uv)
    assert c0 == pytest.approx(c1, rel=1e-12)


def test_write_back_reaches_every_object():
    g = load_golden("a")
    cam, kfs, mps = graph_from_golden(g)
    window = kfs[: int(g["n_kf"])]
    pts = window_points(window)
    R = np.stack([np.eye(3)] * len(window)); t = np.arange(3 * len(window), dtype=float).reshape(-1, 3)
    X = np.arange(3 * len(pts), dtype=np.float64).reshape(-1, 3)
    packing.write_back(window, pts, R, t, X)
    for i, kf in enumerate(window):
        np.testing.assert_array_equal(kf.R_w2c, np.eye(3))
        np.testing.assert_array_equal(kf.t_w2c, t[i])
    for j, mp in enumerate(pts):
        assert mp.position.dtype == np.float64
        np.testing.assert_array_equal(mp.position, X[j].astype(np.float64))


def test_plugin_returns_false_and_mutates_nothing_on_few_observations(tmp_path):
    g = load_golden("few")
    cam, kfs, mps = graph_from_golden(g)
    before = [kf.t_w2c.copy() for kf in kfs], [m.position.copy() for m in mps]
    opt = B200Optimizer(log_dir=None)
    assert refgraph.Map().optimize_local(opt, kfs) is False       # needs no GPU: early return (:99-101)
    for kf, t0 in zip(kfs, before[0]):
        np.testing.assert_array_equal(kf.t_w2c, t0)
    for m, x0 in zip(mps, before[1]):
        np.testing.assert_array_equal(m.position, x0)
    assert refgraph.Map().optimize_local(opt, []) is False


def test_plugin_never_raises_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = synth.make_problem("T0", seed=1)
    cam, kfs, mps = refgraph.graph_from_problem(p)
    opt = B200Optimizer(log_dir=None)
    t_before = kfs[1].t_w2c.copy()
    assert refgraph.Map().optimize_local(opt, kfs) is False       # logged, not raised; no CPU fallback
    np.testing.assert_array_equal(kfs[1].t_w2c, t_before)


def test_lm_config_from_stock_config_object():
    class Opt: lr = 1e-3; n_iter = 150; batch_size = 1000; huber_delta = 2.5
    class Cfg: optimization = Opt()
    from visual_slam_core_b200 import lm_config_from
    c = lm_config_from(Cfg())
    assert c.huber_delta == 2.5 and c.max_iter == 10 and c.lambda0 == 1e-4
    assert lm_config_from(None).huber_delta == 5.0


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_by_point_partitions_observations(world):
    p = synth.make_problem("T2", seed=2)
    tot_obs, tot_pt = 0, 0
    for r in range(world):
        s = p.shard_by_point(r, world)
        tot_obs += s.n_obs; tot_pt += s.n_pt
        assert s.pt_offsets[0] == 0 and s.pt_offsets[-1] == s.n_obs
        assert (s.obs_pt >= 0).all() and (s.obs_pt < max(s.n_pt, 1)).all()
        if world > 1 and s.n_obs:
            assert abs(s.n_obs - p.n_obs / world) < 0.2 * p.n_obs / world + 64
    assert tot_obs == p.n_obs and tot_pt == p.n_pt


def test_generator_matches_spec():
    p = synth.make_problem("C1", seed=0)
    assert (p.n_cam, p.n_pt, p.n_obs) == (10, 2000, 20000)
    assert p.fixed.tolist() == [1] + [0] * 9
    assert np.array_equal(p.R, p.R.astype(np.float32).astype(np.float64))
    assert (np.diff(p.obs_pt) >= 0).all()
    assert 0.03 < p.outlier.mean() < 0.07
    q = synth.make_problem("C1", seed=0)
    np.testing.assert_array_equal(p.obs_uv, q.obs_uv)


def test_bench_reference_arm_for_pose_only_prints_one_contract_line():
    """`bench.py --config C3 --impl reference` times the reference's own cv2.solvePnPRansac call on the
    host (no GPU needed) and prints exactly one JSON line with the contract keys."""
    import json
    import os
    import subprocess
    import sys
    pytest.importorskip("cv2")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--config", "C3", "--impl", "reference",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "config", "cpu_baseline", "e2e"):
        assert key in d
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
