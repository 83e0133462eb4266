"""Time one large config on one GPU (profile mode: per-kernel CUDA events)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ba
from visual_slam_core_b200 import LMConfig, LocalBASolver, synth
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-8
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
schurs = sys.argv[4].split(",") if len(sys.argv) > 4 else ["stream"]
t0 = time.time(); p = synth.make_problem(name, seed=0); print("gen", time.time() - t0, p.n_cam, p.n_pt, p.n_obs, flush=True)
for schur in schurs:
    for prof in (True, False):
        s = LocalBASolver(LMConfig(pcg_tol=tol, pcg_max_iter=3000, max_iter=iters, profile=prof, schur=schur))
        for rep in range(3):
            t0 = time.time()
            res = s.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, p.pt_offsets)
            wall = time.time() - t0
        print(schur, "profile" if prof else "plain", "gpu ms", round(res.gpu_ms, 3), "wall s", round(wall, 4),
              "solver", res.extra["solver_used"], "bw", res.extra["band_width"], "fail", int(res.solve_fail.sum()))
        if prof:
            print("  per-iteration ms:", {k: round(v / iters, 4) for k, v in res.extra.items() if k.startswith("ms_")})
        print("  cost", np.array2string(res.cost_hist, precision=10), "acc", res.accepted.tolist(), flush=True)
        s.close()
