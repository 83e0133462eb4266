"""Parity figures of the sparse/PCG path vs the oracle's sparse direct solve (prints, no asserts)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import b200ba
from visual_slam_core_b200 import LMConfig, LocalBASolver, synth
from oracle import lm_oracle as O
from helpers import rot_angle, scale_norm_trans_err, gauge_aligned_trans_err
name = sys.argv[1] if len(sys.argv) > 1 else "T4"
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-12
p = synth.make_problem(name, seed=2)
s = LocalBASolver(LMConfig(pcg_tol=tol, pcg_max_iter=3000))
res = s.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, p.pt_offsets)
t0 = time.time(); res = s.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, p.pt_offsets)
print("gpu ms", res.gpu_ms, "pcg iters", res.pcg_iters.tolist(), "fail", res.solve_fail.tolist())
ref = O.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv)
print("acc", res.accepted.tolist(), ref.accepted.tolist())
print("cost rel", np.abs(res.cost_hist - ref.cost_hist) / ref.cost_hist)
c = -np.einsum("nji,nj->ni", ref.R, ref.t); extent = np.linalg.norm(c - c[0], axis=1).max()
te = np.linalg.norm(res.t - ref.t, axis=1)
print("rot", rot_angle(res.R, ref.R).max(), "trans/baseline", scale_norm_trans_err(res.R, res.t, ref.R, ref.t).max(),
      "aligned", gauge_aligned_trans_err(res.R, res.t, ref.R, ref.t)[0].max(), "trans/extent", te.max() / extent, "extent", extent)
print("flag mismatches", int((res.flags != ref.flags).sum()), "of", p.n_obs)
