"""Per-LM-iteration device times of the panel path's kernels on C4 (profile mode), after a warm-up solve; parity of the final cost
against the last known value is printed for a quick sanity check."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200ba  # noqa: F401
from visual_slam_core_b200 import LMConfig, LocalBASolver, synth
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
p = synth.make_problem(name, seed=0)
s = LocalBASolver(LMConfig(profile=True))
for _ in range(3):
    r = s.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, p.pt_offsets)
e, n = r.extra, r.n_iter
print(f"{name}: per LM iteration lin {e['ms_linearize'] / n:.4f} (point blocks {e['ms_point_blocks'] / n:.4f}, camera blocks {e['ms_cam_blocks'] / n:.4f}, "
      f"panel {(e['ms_linearize'] - e['ms_point_blocks'] - e['ms_cam_blocks']) / n:.4f}) solve {e['ms_solve'] / n:.4f} back {e['ms_backsub'] / n:.4f} ms, "
      f"other per solve {e['ms_other']:.4f}; cost {r.cost_hist[0]:.6f} -> {r.cost_hist[-1]:.6f}, accepted {r.accepted.tolist()}")
