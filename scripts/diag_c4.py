"""Diagnostic (prints): run-to-run and solver-to-solver spread of the C4 trajectory."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200ba
from visual_slam_core_b200 import LMConfig, LocalBASolver, synth
np.set_printoptions(precision=9, linewidth=250)
p = synth.make_problem("C4", seed=0)
runs = []
for solver, schur in [("auto", "panel"), ("auto", "panel"), ("auto", "mma"), ("band_seq", "panel")]:
    s = LocalBASolver(LMConfig(solver=solver, schur=schur))
    r = s.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, p.pt_offsets)
    s.close()
    runs.append(r)
    print(f"{solver:8s} {schur:6s} acc {r.accepted.tolist()} lam {r.lambda_hist}")
    print("   cost", r.cost_hist)
base = runs[0]
for r in runs[1:]:
    print("rel diff vs run 0:", np.abs(r.cost_hist - base.cost_hist) / base.cost_hist)
