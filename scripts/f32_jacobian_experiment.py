"""Precision experiment (VERDICT r1 next #3e), CPU only: the oracle with its Jacobians rounded through float32 (residuals,
accumulation and state float64) against the pure float64 oracle.  Result (round 2): translations 2.6e-5 on C1 and
1.9e-5 on T2 (over the 1e-5 tolerance), and on the longer windows (T4, 1 000 cameras) the trajectories separate
completely (cost 3e-2 / 1.8e-1, thousands of flags).  float32 Jacobians are therefore out; see DESIGN.md section 2."""
import time, sys, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import b200ba
from visual_slam_core_b200 import synth
from oracle import lm_oracle as O
from helpers import rot_angle, scale_norm_trans_err
orig = O.linearize
def lin32(intr, R, t, fixed, X, obs_cam, obs_pt, obs_uv, delta):
    # Jacobians rounded to f32 (as if computed in f32 from f64-exact inputs), residual f64, products f64
    fx, fy, cx, cy = [float(v) for v in intr]
    d = orig(intr, R, t, fixed, X, obs_cam, obs_pt, obs_uv, delta)
    Jp = d["Jp"].astype(np.float32).astype(np.float64)
    Jl = d["Jl"].astype(np.float32).astype(np.float64)
    r, w = d["r"], d["w"]
    n_cam, n_pt = R.shape[0], X.shape[0]
    JpW = Jp * w[:, :, None]; JlW = Jl * w[:, :, None]
    Uo = np.einsum("nca,ncb->nab", JpW, Jp); Vo = np.einsum("nca,ncb->nab", JlW, Jl)
    Wo = np.einsum("nca,ncb->nab", JpW, Jl); gpo = np.einsum("nca,nc->na", JpW, r); glo = np.einsum("nca,nc->na", JlW, r)
    def seg_sum(idx, vals, m):
        flat = vals.reshape(vals.shape[0], -1); out = np.empty((m, flat.shape[1]))
        for k in range(flat.shape[1]): out[:, k] = np.bincount(idx, weights=flat[:, k], minlength=m)
        return out.reshape((m,) + vals.shape[1:])
    d.update(U=seg_sum(obs_cam,Uo,n_cam), gp=seg_sum(obs_cam,gpo,n_cam), V=seg_sum(obs_pt,Vo,n_pt), gl=seg_sum(obs_pt,glo,n_pt), W=Wo, Jp=Jp, Jl=Jl)
    return d
for name,kw in [("T2",{}),("C1",{}),("T4",{}),("C2",{}),(None,dict(n_cam=1000,n_pt=50000,k_mean=8.0,k_max=16))]:
    p = synth.make_problem(name, seed=2, **kw)
    O.linearize = orig
    a = O.solve_local(p.intr,p.R,p.t,p.fixed,p.X,p.obs_cam,p.obs_pt,p.obs_uv)
    O.linearize = lin32
    b = O.solve_local(p.intr,p.R,p.t,p.fixed,p.X,p.obs_cam,p.obs_pt,p.obs_uv)
    O.linearize = orig
    print(name or kw, "acc", a.accepted.tolist(), b.accepted.tolist())
    print("  cost rel", np.max(np.abs(a.cost_hist-b.cost_hist)/a.cost_hist), "rot", rot_angle(a.R,b.R).max(), "trans", scale_norm_trans_err(b.R,b.t,a.R,a.t).max(), "flag mism", int((a.flags!=b.flags).sum()), "of", p.n_obs, "X", np.abs(a.X-b.X).max())
