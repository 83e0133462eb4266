"""Diagnostic (prints): a large window through the band solvers / PCG against the committed oracle run."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import b200ba
from visual_slam_core_b200 import LMConfig, LocalBASolver
import make_large_golden as G
from helpers import rot_angle, gauge_aligned_trans_err
name = sys.argv[1] if len(sys.argv) > 1 else "L2"
g = np.load(os.path.join(ROOT, "tests", "golden", f"lm_{name}.npz"))
p = G.make_window(name)
np.set_printoptions(precision=6, linewidth=200)
print("oracle  acc", g["accepted"].tolist()); print("oracle cost", g["cost_hist"]); print("oracle cand", g["cand_cost_hist"])
for solver in sys.argv[2:] or ["auto", "band_seq", "pcg"]:
    s = LocalBASolver(LMConfig(solver=solver, pcg_tol=1e-12, pcg_max_iter=5000))
    r = s.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, p.pt_offsets)
    s.close()
    print(f"{solver:9s} acc", r.accepted.tolist(), "fail", r.solve_fail.tolist(), "pcg", r.pcg_iters.tolist(), f"gpu {r.gpu_ms:.1f} ms")
    print(f"{solver:9s} cost", r.cost_hist); print(f"{solver:9s} cand", r.cand_cost_hist)
    print(f"{solver:9s} rel cand diff", np.abs(r.cand_cost_hist - g["cand_cost_hist"]) / g["cand_cost_hist"])
