"""One profiled solve (no graph) of a config with a given Schur kind: for ncu launch lists."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200ba
from visual_slam_core_b200 import LMConfig, LocalBASolver, synth
name, kind = sys.argv[1], sys.argv[2]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
p = synth.make_problem(name, seed=0)
if len(sys.argv) > 4:
    which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    p = p.shard_by_point(which, int(sys.argv[4]))      # one of N shards, solved alone (no communicator)
    print('shard', which, 'of', sys.argv[4], p.n_pt, 'points', p.n_obs, 'observations')
s = LocalBASolver(LMConfig(schur=kind, profile=True, max_iter=iters))
for _ in range(2):
    r = s.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, p.pt_offsets)
print(name, kind, "lin ms/iter", r.extra["ms_linearize"] / iters, "cost", r.cost_hist[-1])
