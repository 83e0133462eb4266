"""CPU oracle for the local bundle-adjustment hot path  --  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; the product path
(`visual-slam-core_b200/`) never does and fails loudly without its CUDA library.

What is restated from the reference (and pinned against the reference itself,
see tests/golden/make_golden.py and tests/test_oracle_golden.py):

* objective  -- `/root/reference/visual_slam/optimization/torch_optimizer.py:152-177`
  Xc = R X + t, Z' = Xc_z + 1e-8 (no clamp), u = fx Xc_x/Z' + cx, v = fy Xc_y/Z' + cy,
  r = predicted - observed, per-component Huber (`:135-143`), cost = sum rho(r_u)+rho(r_v).
* parameterisation and gauge -- `:43-61,159`: first keyframe fixed; other keyframes
  R = exp([w]x) R0 (left increment), t additive, all points free.

What the reference does NOT contain: a Levenberg-Marquardt
trajectory.  The LM below is the build's own algorithm specification; the CUDA
solver must follow it step for step.  **LM trajectory: parity unpinned by the
reference (it has no LM, no tests, no golden vectors); the objective, packing
order and write-back are pinned by reference-generated fixtures.**

Algorithm (float64 arithmetic and float64 state throughout; the *inputs* are
float32-representable because the packer rounds them as the reference does, `:51-52,65-68`):

  per observation: r, J_pi, J_p = J_pi [-[Y]x | I] (order dw, dt), J_l = J_pi R,
  IRLS weights w_c = 1 if |r_c| < delta else delta/|r_c|  
  U_i += J_p^T W J_p, V_j += J_l^T W J_l, W_ij = J_p^T W J_l, g_p, g_l
  damping  U_i + lam*diag*(U_i), V_j + lam*diag*(V_j), diag* = diag where > 0 else 1
  V_j^-1 through an explicit LDL^T (same operation order as the CUDA kernel)
  S = U - sum_j W V^-1 W^T, b = -g_p + sum_j W V^-1 g_l, fixed cameras -> identity rows
  S dp = b (dense Cholesky), dl_j = -V_j^-1 (g_l,j + sum_i W_ij^T dp_i)
  candidate: R <- exp(dw) R, t <- t + dt, X <- X + dl
  accept iff candidate cost is finite and < current cost * (1 - accept_rtol);
  lam <- max(lam*down, lam_min)
  else lam <- min(lam*up, lam_max); exactly max_iter iterations (rejected ones count)
  flags: (r_u^2 + r_v^2) > chi2_thr at the final accepted iterate, uint8.

Defaults: lam0 1e-4, up 10, down 0.1, lam in [1e-6, 1e8], accept_rtol 1e-8, 10 iterations.
lam_min = 1e-6 bounds the conditioning of rank-deficient point blocks (a point with a single
in-window observation has V of rank 2, so (V + lam diag V)^-1 ~ 1/lam along the viewing ray):
with 1e-9 two float64 implementations of this same algorithm drift apart by 1e-4 in cost on
such windows (measured on tests/golden/ref_local_ba_b), with 1e-6 they agree to 1e-9.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import scipy.linalg
import scipy.sparse as sp

DEPTH_EPS = 1e-8  # torch_optimizer.py:165


@dataclass
class LMParams:
    huber_delta: float = 5.0
    chi2_thr: float = 5.991
    lambda0: float = 1e-4
    lambda_up: float = 10.0
    lambda_down: float = 0.1
    lambda_min: float = 1e-6
    lambda_max: float = 1e8
    accept_rtol: float = 1e-8
    max_iter: int = 10


@dataclass
class LMResult:
    R: np.ndarray
    t: np.ndarray
    X: np.ndarray
    flags: np.ndarray
    cost_hist: np.ndarray       # (max_iter+1,) cost after each iteration; [0] = initial
    lambda_hist: np.ndarray     # (max_iter,) lambda used by each iteration
    accepted: np.ndarray        # (max_iter,) uint8
    cand_cost_hist: np.ndarray  # (max_iter,) candidate cost of each iteration
    extra: dict = field(default_factory=dict)


# ----------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------
def sinc(x):
    x = np.asarray(x, dtype=np.float64)
    x2 = x * x
    small = x < 1e-4
    xs = np.where(small, 1.0, x)
    return np.where(small, 1.0 - x2 / 6.0 + x2 * x2 / 120.0, np.sin(xs) / xs)


def so3_exp(w):
    """exp([w]x) = I + A [w]x + B [w]x^2, A = sinc(th), B = 0.5 sinc(th/2)^2."""
    w = np.asarray(w, dtype=np.float64)
    th = np.sqrt(np.sum(w * w, axis=-1))
    A = sinc(th)
    h = sinc(0.5 * th)
    B = 0.5 * h * h
    wx, wy, wz = w[..., 0], w[..., 1], w[..., 2]
    R = np.empty(w.shape[:-1] + (3, 3))
    # [w]x^2 = w w^T - |w|^2 I
    R[..., 0, 0] = 1.0 - B * (wy * wy + wz * wz)
    R[..., 1, 1] = 1.0 - B * (wx * wx + wz * wz)
    R[..., 2, 2] = 1.0 - B * (wx * wx + wy * wy)
    R[..., 0, 1] = B * wx * wy - A * wz
    R[..., 1, 0] = B * wx * wy + A * wz
    R[..., 0, 2] = B * wx * wz + A * wy
    R[..., 2, 0] = B * wx * wz - A * wy
    R[..., 1, 2] = B * wy * wz - A * wx
    R[..., 2, 1] = B * wy * wz + A * wx
    return R


def huber_rho(r, delta):
    a = np.abs(r)
    return np.where(a < delta, 0.5 * r * r, delta * (a - 0.5 * delta))


def huber_weight(r, delta):
    a = np.abs(r)
    return np.where(a < delta, 1.0, delta / np.where(a < delta, 1.0, a))


def residuals(intr, R, t, X, obs_cam, obs_pt, obs_uv):
    """Reprojection residual (predicted - observed), float64.  torch_optimizer.py:164-175."""
    fx, fy, cx, cy = [float(v) for v in intr]
    Rk = R[obs_cam]
    Xj = X[obs_pt].astype(np.float64)
    Y = np.einsum("nij,nj->ni", Rk, Xj)
    Xc = Y + t[obs_cam]
    iz = 1.0 / (Xc[:, 2] + DEPTH_EPS)
    ru = fx * Xc[:, 0] * iz + cx - obs_uv[:, 0].astype(np.float64)
    rv = fy * Xc[:, 1] * iz + cy - obs_uv[:, 1].astype(np.float64)
    return ru, rv


def reference_cost(intr, R, t, X, obs_cam, obs_pt, obs_uv, delta=5.0):
    """The reference objective (torch_optimizer.py:152-177) as one float64 number."""
    ru, rv = residuals(intr, R, t, X, obs_cam, obs_pt, obs_uv)
    return float(np.sum(huber_rho(ru, delta)) + np.sum(huber_rho(rv, delta)))


def outlier_flags(intr, R, t, X, obs_cam, obs_pt, obs_uv, chi2_thr=5.991):
    ru, rv = residuals(intr, R, t, X, obs_cam, obs_pt, obs_uv)
    return ((ru * ru + rv * rv) > chi2_thr).astype(np.uint8)


def sym3_inverse(V6):
    """Inverse of symmetric 3x3 [a b c; b d e; c e f] given as (n,6) = a,b,c,d,e,f.

    Explicit LDL^T; returns (inv6, ok).  Non-positive pivots -> inv = 0, ok = False.
    Same operation order as `sym3_inverse` in csrc/ba_device.cuh.
    """
    a, b, c, d, e, f = [V6[:, i] for i in range(6)]
    with np.errstate(divide="ignore", invalid="ignore"):
        i0 = 1.0 / a
        l10 = b * i0
        l20 = c * i0
        d1 = d - l10 * b
        e1 = e - l20 * b
        i1 = 1.0 / d1
        l21 = e1 * i1
        d2 = f - l20 * c - l21 * e1
        i2 = 1.0 / d2
        m10 = -l10
        m20 = l10 * l21 - l20
        m21 = -l21
        inv = np.stack([
            i0 + m10 * m10 * i1 + m20 * m20 * i2,
            m10 * i1 + m20 * m21 * i2,
            m20 * i2,
            i1 + m21 * m21 * i2,
            m21 * i2,
            i2,
        ], axis=1)
    ok = (a > 0) & (d1 > 0) & (d2 > 0) & np.isfinite(inv).all(axis=1)
    inv[~ok] = 0.0
    return inv, ok


def _sym6_to_mat(v6):
    a, b, c, d, e, f = [v6[:, i] for i in range(6)]
    M = np.empty((v6.shape[0], 3, 3))
    M[:, 0, 0], M[:, 0, 1], M[:, 0, 2] = a, b, c
    M[:, 1, 0], M[:, 1, 1], M[:, 1, 2] = b, d, e
    M[:, 2, 0], M[:, 2, 1], M[:, 2, 2] = c, e, f
    return M


# ----------------------------------------------------------------------------
# linearisation
# ----------------------------------------------------------------------------
def linearize(intr, R, t, fixed, X, obs_cam, obs_pt, obs_uv, delta):
    """Per-observation residuals/Jacobians/weights and the normal-equation blocks."""
    fx, fy, cx, cy = [float(v) for v in intr]
    n_cam, n_pt = R.shape[0], X.shape[0]
    Rk = R[obs_cam]
    Xj = X[obs_pt].astype(np.float64)
    Y = np.einsum("nij,nj->ni", Rk, Xj)
    Xc = Y + t[obs_cam]
    iz = 1.0 / (Xc[:, 2] + DEPTH_EPS)
    xz = Xc[:, 0] * iz
    yz = Xc[:, 1] * iz
    ru = fx * xz + cx - obs_uv[:, 0].astype(np.float64)
    rv = fy * yz + cy - obs_uv[:, 1].astype(np.float64)
    n = ru.shape[0]
    # J_pi rows
    Jt = np.zeros((n, 2, 3))
    Jt[:, 0, 0] = fx * iz
    Jt[:, 0, 2] = -fx * xz * iz
    Jt[:, 1, 1] = fy * iz
    Jt[:, 1, 2] = -fy * yz * iz
    # -[Y]x
    nYx = np.zeros((n, 3, 3))
    nYx[:, 0, 1] = Y[:, 2]
    nYx[:, 0, 2] = -Y[:, 1]
    nYx[:, 1, 0] = -Y[:, 2]
    nYx[:, 1, 2] = Y[:, 0]
    nYx[:, 2, 0] = Y[:, 1]
    nYx[:, 2, 1] = -Y[:, 0]
    Jw = Jt @ nYx
    Jp = np.concatenate([Jw, Jt], axis=2)          # (n,2,6)  [dw, dt]
    Jl = Jt @ Rk                                   # (n,2,3)
    free = (np.asarray(fixed)[obs_cam] == 0)
    Jp = Jp * free[:, None, None]
    r = np.stack([ru, rv], axis=1)
    w = huber_weight(r, delta)                     # (n,2)
    cost = float(np.sum(huber_rho(r, delta)))

    JpW = Jp * w[:, :, None]
    JlW = Jl * w[:, :, None]
    Uo = np.einsum("nca,ncb->nab", JpW, Jp)        # (n,6,6)
    Vo = np.einsum("nca,ncb->nab", JlW, Jl)        # (n,3,3)
    Wo = np.einsum("nca,ncb->nab", JpW, Jl)        # (n,6,3)
    gpo = np.einsum("nca,nc->na", JpW, r)          # (n,6)
    glo = np.einsum("nca,nc->na", JlW, r)          # (n,3)

    def seg_sum(idx, vals, m):
        flat = vals.reshape(vals.shape[0], -1)
        out = np.empty((m, flat.shape[1]))
        for k in range(flat.shape[1]):
            out[:, k] = np.bincount(idx, weights=flat[:, k], minlength=m)
        return out.reshape((m,) + vals.shape[1:])

    U = seg_sum(obs_cam, Uo, n_cam)
    gp = seg_sum(obs_cam, gpo, n_cam)
    V = seg_sum(obs_pt, Vo, n_pt)
    gl = seg_sum(obs_pt, glo, n_pt)
    return dict(cost=cost, U=U, gp=gp, V=V, gl=gl, W=Wo, r=r, w=w, Jp=Jp, Jl=Jl)


def damp_diag(d, lam):
    return lam * np.where(d > 0, d, 1.0)


def build_rcs(lin, lam, fixed, obs_cam, obs_pt, n_cam, n_pt, sparse_out=False):
    """Reduced camera system S, b and the per-point inverses."""
    U, gp, V, gl, W = lin["U"], lin["gp"], lin["V"], lin["gl"], lin["W"]
    Ud = U.copy()
    idx6 = np.arange(6)
    Ud[:, idx6, idx6] += damp_diag(U[:, idx6, idx6], lam)
    V6 = np.stack([V[:, 0, 0], V[:, 0, 1], V[:, 0, 2], V[:, 1, 1], V[:, 1, 2], V[:, 2, 2]], axis=1)
    V6[:, 0] += damp_diag(V[:, 0, 0], lam)
    V6[:, 3] += damp_diag(V[:, 1, 1], lam)
    V6[:, 5] += damp_diag(V[:, 2, 2], lam)
    Vinv6, ok = sym3_inverse(V6)
    Vinv = _sym6_to_mat(Vinv6)

    n_obs = obs_cam.shape[0]
    rows = (6 * obs_cam[:, None, None] + np.arange(6)[None, :, None]).repeat(3, axis=2)
    cols = (3 * obs_pt[:, None, None] + np.arange(3)[None, None, :]).repeat(6, axis=1)
    Wb = sp.csr_matrix((W.ravel(), (rows.ravel(), cols.ravel())), shape=(6 * n_cam, 3 * n_pt))
    pr = (3 * np.arange(n_pt)[:, None, None] + np.arange(3)[None, :, None]).repeat(3, axis=2)
    pc = (3 * np.arange(n_pt)[:, None, None] + np.arange(3)[None, None, :]).repeat(3, axis=1)
    Vb = sp.csr_matrix((Vinv.ravel(), (pr.ravel(), pc.ravel())), shape=(3 * n_pt, 3 * n_pt))
    WV = Wb @ Vb
    Ssub = WV @ Wb.T
    b = -gp.reshape(-1) + WV @ gl.reshape(-1)
    ur = (6 * np.arange(n_cam)[:, None, None] + np.arange(6)[None, :, None]).repeat(6, axis=2)
    uc = (6 * np.arange(n_cam)[:, None, None] + np.arange(6)[None, None, :]).repeat(6, axis=1)
    Ub = sp.csr_matrix((Ud.ravel(), (ur.ravel(), uc.ravel())), shape=(6 * n_cam, 6 * n_cam))
    S = (Ub - Ssub).tocsr()
    fx_idx = np.nonzero(np.repeat(np.asarray(fixed) != 0, 6))[0]
    if not sparse_out:
        S = S.toarray()
        S[fx_idx, :] = 0.0
        S[:, fx_idx] = 0.0
        S[fx_idx, fx_idx] = 1.0
    else:
        keep = np.ones(6 * n_cam)
        keep[fx_idx] = 0.0
        D = sp.diags(keep)
        S = (D @ S @ D + sp.diags(1.0 - keep)).tocsc()
    b[fx_idx] = 0.0
    return S, b, Wb, Vinv, ok


def solve_rcs(S, b):
    """Dense Cholesky solve; returns (dp, ok)."""
    if sp.issparse(S):
        import scipy.sparse.linalg as spla
        try:
            dp = spla.spsolve(S, b)
        except Exception:
            return np.zeros_like(b), False
        return dp, bool(np.isfinite(dp).all())
    try:
        c, low = scipy.linalg.cho_factor(S, lower=True, check_finite=True)
        dp = scipy.linalg.cho_solve((c, low), b)
    except (np.linalg.LinAlgError, ValueError):
        return np.zeros_like(b), False
    return dp, bool(np.isfinite(dp).all())


def lm_step(intr, R, t, fixed, X, obs_cam, obs_pt, obs_uv, lam, delta, sparse=False):
    """One linearise + Schur + solve + back-substitute.  Returns candidate state."""
    n_cam, n_pt = R.shape[0], X.shape[0]
    lin = linearize(intr, R, t, fixed, X, obs_cam, obs_pt, obs_uv, delta)
    S, b, Wb, Vinv, _ = build_rcs(lin, lam, fixed, obs_cam, obs_pt, n_cam, n_pt, sparse_out=sparse)
    dp, ok = solve_rcs(S, b)
    rhs = lin["gl"] + (Wb.T @ dp).reshape(n_pt, 3)
    dl = -np.einsum("nab,nb->na", Vinv, rhs)
    dpc = dp.reshape(n_cam, 6)
    R_new = so3_exp(dpc[:, :3]) @ R
    t_new = t + dpc[:, 3:]
    free = np.asarray(fixed) == 0
    R_new[~free] = R[~free]
    t_new[~free] = t[~free]
    X_new = X.astype(np.float64) + dl
    return dict(R=R_new, t=t_new, X=X_new, ok=ok, lin=lin, S=S, b=b, dp=dp, dl=dl)


def solve_local(intr, R, t, fixed, X, obs_cam, obs_pt, obs_uv,
                params: Optional[LMParams] = None, sparse: Optional[bool] = None,
                keep_steps: bool = False) -> LMResult:
    """The LM trajectory the CUDA solver must reproduce."""
    p = params or LMParams()
    R = np.array(R, dtype=np.float64)
    t = np.array(t, dtype=np.float64)
    X = np.array(X, dtype=np.float64)
    obs_cam = np.asarray(obs_cam, dtype=np.int64)
    obs_pt = np.asarray(obs_pt, dtype=np.int64)
    if sparse is None:
        sparse = R.shape[0] > 400
    lam = p.lambda0
    cost = reference_cost(intr, R, t, X, obs_cam, obs_pt, obs_uv, p.huber_delta)
    cost_hist = [cost]
    lam_hist, acc_hist, cand_hist, steps = [], [], [], []
    for _ in range(p.max_iter):
        st = lm_step(intr, R, t, fixed, X, obs_cam, obs_pt, obs_uv, lam, p.huber_delta, sparse)
        cand = reference_cost(intr, st["R"], st["t"], st["X"], obs_cam, obs_pt, obs_uv, p.huber_delta)
        ok = bool(st["ok"] and np.isfinite(cand) and cand < cost * (1.0 - p.accept_rtol))
        lam_hist.append(lam)
        acc_hist.append(1 if ok else 0)
        cand_hist.append(cand)
        if keep_steps:
            steps.append(dict(S=st["S"], b=st["b"], dp=st["dp"], dl=st["dl"]))
        if ok:
            R, t, X, cost = st["R"], st["t"], st["X"], cand
            lam = max(lam * p.lambda_down, p.lambda_min)
        else:
            lam = min(lam * p.lambda_up, p.lambda_max)
        cost_hist.append(cost)
    flags = outlier_flags(intr, R, t, X, obs_cam, obs_pt, obs_uv, p.chi2_thr)
    return LMResult(R, t, X, flags, np.array(cost_hist), np.array(lam_hist),
                    np.array(acc_hist, dtype=np.uint8), np.array(cand_hist),
                    extra=dict(steps=steps))


# ----------------------------------------------------------------------------
# pose-only (motion-only) LM: one free pose, fixed points, same objective
# ----------------------------------------------------------------------------
def solve_pose_only(intr, R, t, X, uv, frame_offsets, params: Optional[LMParams] = None):
    """Batched pose-only LM (BASELINE config 3).  Each frame is independent.

    Same residual / Huber / damping / accept rule as `solve_local`, with the 6x6
    system solved directly.  Returns per-frame histories stacked on axis 0.
    """
    p = params or LMParams()
    fx, fy, cx, cy = [float(v) for v in intr]
    nF = R.shape[0]
    R = np.array(R, dtype=np.float64)
    t = np.array(t, dtype=np.float64)
    Xd = np.asarray(X, dtype=np.float64)
    uvd = np.asarray(uv, dtype=np.float64)
    fo = np.asarray(frame_offsets, dtype=np.int64)
    fid = np.repeat(np.arange(nF), np.diff(fo))
    delta = p.huber_delta

    def evaluate(Rf, tf, need_jac):
        Y = np.einsum("nij,nj->ni", Rf[fid], Xd)
        Xc = Y + tf[fid]
        iz = 1.0 / (Xc[:, 2] + DEPTH_EPS)
        xz, yz = Xc[:, 0] * iz, Xc[:, 1] * iz
        r = np.stack([fx * xz + cx - uvd[:, 0], fy * yz + cy - uvd[:, 1]], axis=1)
        cost = np.bincount(fid, weights=huber_rho(r, delta).sum(axis=1), minlength=nF)
        if not need_jac:
            return cost, r
        n = r.shape[0]
        Jt = np.zeros((n, 2, 3))
        Jt[:, 0, 0] = fx * iz
        Jt[:, 0, 2] = -fx * xz * iz
        Jt[:, 1, 1] = fy * iz
        Jt[:, 1, 2] = -fy * yz * iz
        nYx = np.zeros((n, 3, 3))
        nYx[:, 0, 1], nYx[:, 0, 2] = Y[:, 2], -Y[:, 1]
        nYx[:, 1, 0], nYx[:, 1, 2] = -Y[:, 2], Y[:, 0]
        nYx[:, 2, 0], nYx[:, 2, 1] = Y[:, 1], -Y[:, 0]
        Jp = np.concatenate([Jt @ nYx, Jt], axis=2)
        w = huber_weight(r, delta)
        JpW = Jp * w[:, :, None]
        Ho = np.einsum("nca,ncb->nab", JpW, Jp).reshape(n, 36)
        go = np.einsum("nca,nc->na", JpW, r)
        H = np.stack([np.bincount(fid, weights=Ho[:, k], minlength=nF) for k in range(36)], axis=1)
        g = np.stack([np.bincount(fid, weights=go[:, k], minlength=nF) for k in range(6)], axis=1)
        return cost, r, H.reshape(nF, 6, 6), g

    lam = np.full(nF, p.lambda0)
    cost, _ = evaluate(R, t, False)
    cost_hist = [cost.copy()]
    acc_hist, lam_hist = [], []
    i6 = np.arange(6)
    for _ in range(p.max_iter):
        _, _, H, g = evaluate(R, t, True)
        Hd = H.copy()
        Hd[:, i6, i6] += lam[:, None] * np.where(H[:, i6, i6] > 0, H[:, i6, i6], 1.0)
        dp = np.zeros((nF, 6))
        okf = np.zeros(nF, dtype=bool)
        for f in range(nF):
            try:
                c = scipy.linalg.cho_factor(Hd[f], lower=True)
                dp[f] = scipy.linalg.cho_solve(c, -g[f])
                okf[f] = np.isfinite(dp[f]).all()
            except (np.linalg.LinAlgError, ValueError):
                okf[f] = False
        Rn = so3_exp(dp[:, :3]) @ R
        tn = t + dp[:, 3:]
        cand, _ = evaluate(Rn, tn, False)
        ok = okf & np.isfinite(cand) & (cand < cost * (1.0 - p.accept_rtol))
        lam_hist.append(lam.copy())
        acc_hist.append(ok.astype(np.uint8))
        R = np.where(ok[:, None, None], Rn, R)
        t = np.where(ok[:, None], tn, t)
        cost = np.where(ok, cand, cost)
        lam = np.where(ok, np.maximum(lam * p.lambda_down, p.lambda_min),
                       np.minimum(lam * p.lambda_up, p.lambda_max))
        cost_hist.append(cost.copy())
    _, r = evaluate(R, t, False)
    flags = ((r[:, 0] ** 2 + r[:, 1] ** 2) > p.chi2_thr).astype(np.uint8)
    return dict(R=R, t=t, flags=flags, cost_hist=np.stack(cost_hist, axis=1),
                lambda_hist=np.stack(lam_hist, axis=1), accepted=np.stack(acc_hist, axis=1))
