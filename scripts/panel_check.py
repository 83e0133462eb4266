"""GPU check of the panel linearise path (prints): one linearisation against the oracle and the round-1 kernel on
dense windows, whole solves on banded windows, then kernel timings."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import b200ba
from oracle import lm_oracle as O
from visual_slam_core_b200 import LMConfig, LocalBASolver, synth
np.set_printoptions(precision=4, linewidth=200)
what = sys.argv[1:] or ["lin", "solve", "time"]
if "lin" in what:
    for name, seed in [("T0", 1), ("T1", 2), ("T2", 3), ("C1", 0), ("C2", 0)]:
        p = synth.make_problem(name, seed=seed)
        lam = 1e-4
        lin = O.linearize(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam.astype(np.int64), p.obs_pt.astype(np.int64), p.obs_uv, 5.0)
        S, b, _, _, _ = O.build_rcs(lin, lam, p.fixed, p.obs_cam.astype(np.int64), p.obs_pt.astype(np.int64), p.n_cam, p.n_pt)
        Ud = np.array([np.diag(lin["U"][i]) for i in range(p.n_cam)]).reshape(-1)
        # the GPU's S is undamped on the pose blocks (the solver adds lambda diagU): compare S - damping
        free = np.repeat(p.fixed == 0, 6)
        S_und = S - np.diag(np.where(free, lam * np.where(Ud > 0, Ud, 1.0), 0.0))
        out = {}
        for kind in ("mma", "panel"):
            s = LocalBASolver(LMConfig(schur=kind))
            Sg, bg, dUg, cg = s.debug_linearize(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, p.pt_offsets, lam)
            s.close()
            L = np.tril(np.ones_like(S)) > 0
            fr = free[:, None] & free[None, :] & L
            scale = np.abs(S_und[fr]).max()
            out[kind] = (np.abs(Sg[fr] - S_und[fr]).max() / scale, np.abs(bg[free] - b[free]).max() / np.abs(b).max(),
                         np.abs(dUg[free] - Ud[free]).max() / Ud.max(), abs(cg - lin["cost"]) / lin["cost"])
        print(f"{name}: rel err (S, b, diagU, cost) mma {np.array(out['mma'])} panel {np.array(out['panel'])}", flush=True)
if "solve" in what:
    import make_large_golden as G
    for name in ["T3", "T4", "L1"]:
        p = G.make_window(name) if name.startswith("L") else synth.make_problem(name, seed=2)
        res = {}
        for kind in ("mma", "panel"):
            s = LocalBASolver(LMConfig(schur=kind))
            r = s.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, p.pt_offsets)
            s.close()
            res[kind] = r
        a, b = res["mma"], res["panel"]
        print(f"{name}: solver {b.extra['solver_used']} band {b.extra['band_width']} acc {a.accepted.tolist()} {b.accepted.tolist()} "
              f"max rel cost diff {np.max(np.abs(a.cost_hist - b.cost_hist) / a.cost_hist):.2e} dR {np.abs(a.R - b.R).max():.2e} "
              f"flags diff {int((a.flags != b.flags).sum())} gpu ms {a.gpu_ms:.2f} {b.gpu_ms:.2f}", flush=True)
if "time" in what:
    import torch
    for name in ["C1", "C2", "C4"]:
        p = synth.make_problem(name, seed=0)
        dev = torch.device("cuda", 0)
        T = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
        R0, t0, X0 = T(p.R), T(p.t), T(p.X)
        fixed, oc, op, uv, off = T(p.fixed), T(p.obs_cam), T(p.obs_pt), T(p.obs_uv), T(p.pt_offsets)
        flags = torch.zeros(p.n_obs, dtype=torch.uint8, device=dev)
        for kind in ("mma", "panel"):
            for prof in (True, False):
                s = LocalBASolver(LMConfig(schur=kind, profile=prof))
                ms = []
                for i in range(5):
                    R, t, X = R0.clone(), t0.clone(), X0.clone()
                    rep = s.solve_local_device(p.intr, R, t, fixed, X, oc, op, uv, off, flags, sync=True)
                    ms.append(rep.gpu_ms)
                if prof:
                    print(f"{name} {kind}: per LM iteration lin {rep.ms_linearize / 10:.4f} solve {rep.ms_solve / 10:.4f} back {rep.ms_backsub / 10:.4f} ms "
                          f"other/step {rep.ms_other:.4f}; final cost {rep.cost_hist[10]:.6f}", flush=True)
                else:
                    print(f"{name} {kind}: graph solve {min(ms):.4f} ms -> {10e3 / min(ms):.0f} LM iter/s", flush=True)
                s.close()
