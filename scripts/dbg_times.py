"""Latency hunting: BA_DBG_TIMES=1 python scripts/dbg_times.py [config] [schur] -> clock64 deltas of CTA 0 phases."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200ba
from visual_slam_core_b200 import LMConfig, LocalBASolver, synth
p = synth.make_problem(sys.argv[1] if len(sys.argv) > 1 else "C2", seed=0)
s = LocalBASolver(LMConfig(max_iter=2, use_graph=False, schur=sys.argv[2] if len(sys.argv) > 2 else "auto"))
for _ in range(2):
    r = s.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, p.pt_offsets)
print(r.gpu_ms)
