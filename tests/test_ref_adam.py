"""The restatement of the reference's Adam loop (oracle/ref_adam.py, used only to time the
reference's CPU path in bench.py) reproduces what the unmodified reference wrote back."""
import numpy as np
import pytest

from oracle import ref_adam
from visual_slam_core_b200 import packing

from helpers import graph_from_golden, load_golden, window_points


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_adam_restatement_matches_reference_output(name):
    g = load_golden(name)
    cam, kfs, mps = graph_from_golden(g)
    window = kfs[: int(g["n_kf"])]
    pts = window_points(window)
    pw = packing.pack_window(window, pts)
    out = ref_adam.adam_local_ba(pw.intr, pw.R, pw.t, pw.fixed, pw.X, pw.obs_cam, pw.obs_pt, pw.obs_uv,
                                 n_iter=int(g["n_iter"]), lr=float(g["lr"]), delta=float(g["delta"]))
    assert out["losses"][0] == pytest.approx(float(g["ref_loss0"]), rel=2e-6, abs=6e-4)
    np.testing.assert_allclose(out["t"], g["ref_t_out"], atol=2e-6)
    np.testing.assert_allclose(out["R"], g["ref_R_out"], atol=2e-6)
    np.testing.assert_allclose(out["X"], g["ref_X_out"][g["ref_pt_order"]], atol=2e-5)
