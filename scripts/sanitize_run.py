"""Small solves through every kernel family, for `compute-sanitizer --tool memcheck|racecheck|initcheck python
scripts/sanitize_run.py` (SURVEY 4 last row).  No torch: host buffers through the C ABI only, so the sanitizer
sees nothing but this library's kernels.  Prints one line per case and exits non-zero on a parity failure."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200ba  # noqa: E402,F401
from oracle import lm_oracle as O  # noqa: E402
from visual_slam_core_b200 import LMConfig, LocalBASolver, synth  # noqa: E402

cases = [("T1", "scatter", "auto"), ("T1", "tiled", "auto"), ("T1", "stream", "auto"), ("T1", "mma", "auto"),
         ("T2", "mma", "auto"), ("T2", "auto", "auto"), ("T3", "auto", "auto"), ("T3", "auto", "pcg"), ("T3", "auto", "band_seq")]
if "--quick" not in sys.argv:
    cases += [("T4", "auto", "auto")]
bad = 0
for name, schur, solver in cases:
    p = synth.make_problem(name, seed=1)
    s = LocalBASolver(LMConfig(schur=schur, solver=solver, max_iter=4, use_graph="--graph" in sys.argv))
    res = s.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, p.pt_offsets)
    s.close()
    ref = O.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, O.LMParams(max_iter=4))
    rel = float(np.max(np.abs(res.cost_hist - ref.cost_hist) / ref.cost_hist))
    tol = 1e-6 if solver == "pcg" else 1e-9
    ok = rel < tol and np.array_equal(res.accepted, ref.accepted)
    bad += 0 if ok else 1
    print(f"{name} schur={schur} solver={solver} -> {res.extra['solver_used']}: rel cost diff {rel:.2e} {'ok' if ok else 'MISMATCH'}", flush=True)
pp = synth.make_pose_only(n_frames=6, n_match=96, seed=1)
s = LocalBASolver(LMConfig(max_iter=4))
out = s.solve_pose_only(pp.intr, pp.R, pp.t, pp.X, pp.uv, pp.frame_offsets)
refp = O.solve_pose_only(pp.intr, pp.R, pp.t, pp.X, pp.uv, pp.frame_offsets, O.LMParams(max_iter=4))
ok = np.allclose(out["cost_hist"], refp["cost_hist"], rtol=1e-9)
bad += 0 if ok else 1
print(f"pose-only 6 x 16: {'ok' if ok else 'MISMATCH'}", flush=True)
s.close()
sys.exit(1 if bad else 0)
