"""Time one large config on one GPU (profile mode: per-kernel CUDA events)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ba
from visual_slam_core_b200 import LMConfig, LocalBASolver, synth
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-8
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
t0 = time.time(); p = synth.make_problem(name, seed=0); print("gen", time.time() - t0, p.n_cam, p.n_pt, p.n_obs, flush=True)
for prof in (True, False):
    s = LocalBASolver(LMConfig(pcg_tol=tol, pcg_max_iter=3000, max_iter=iters, profile=prof))
    for rep in range(2):
        t0 = time.time()
        res = s.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, p.pt_offsets)
        wall = time.time() - t0
    print("profile" if prof else "plain", "gpu ms", res.gpu_ms, "wall s", wall, "pcg iters", res.pcg_iters.tolist(), "fail", res.solve_fail.tolist())
    print("  per-iteration ms:", {k: v / iters for k, v in res.extra.items()})
    print("  cost", res.cost_hist, "acc", res.accepted.tolist(), flush=True)
    s.close()
