"""Self-consistency of the LM oracle (float64): Jacobians vs autograd, Schur vs the full
normal equations, LM behaviour.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import lm_oracle as O
from visual_slam_core_b200 import synth


def _torch_residual(intr, w, t, R0, X, uv):
    """One observation, reference parameterisation: R = exp(w) R0 (torch_optimizer.py:159-175)."""
    th = torch.sqrt((w * w).sum() + 1e-300)
    K = torch.zeros(3, 3, dtype=torch.float64)
    K[0, 1], K[0, 2], K[1, 0], K[1, 2], K[2, 0], K[2, 1] = -w[2], w[1], w[2], -w[0], -w[1], w[0]
    Rm = torch.eye(3, dtype=torch.float64) + torch.sin(th) / th * K + (1 - torch.cos(th)) / th**2 * (K @ K)
    Xc = (Rm @ R0) @ X + t
    z = Xc[2] + 1e-8
    return torch.stack([intr[0] * Xc[0] / z + intr[2] - uv[0], intr[1] * Xc[1] / z + intr[3] - uv[1]])


def test_jacobians_match_autograd():
    p = synth.make_problem("T0", seed=3)
    lin = O.linearize(p.intr, p.R, p.t, np.zeros_like(p.fixed), p.X, p.obs_cam, p.obs_pt, p.obs_uv, 5.0)
    intr = torch.tensor(p.intr)
    for o in range(0, p.n_obs, 7):
        c, j = int(p.obs_cam[o]), int(p.obs_pt[o])
        R0 = torch.tensor(p.R[c]); t0 = torch.tensor(p.t[c])
        X0 = torch.tensor(p.X[j].astype(np.float64)); uv = torch.tensor(p.obs_uv[o].astype(np.float64))
        w0 = torch.full((3,), 1e-9, dtype=torch.float64)
        Jw, Jt, Jx = torch.autograd.functional.jacobian(
            lambda w, t, X: _torch_residual(intr, w, t, R0, X, uv), (w0, t0, X0))
        np.testing.assert_allclose(lin["Jp"][o][:, :3], Jw.numpy(), rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(lin["Jp"][o][:, 3:], Jt.numpy(), rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(lin["Jl"][o], Jx.numpy(), rtol=1e-6, atol=1e-6)


def test_huber_matches_reference_definition():
    r = np.array([-7.0, -5.0, -4.999, 0.0, 3.0, 5.0, 5.001, 40.0])
    d = 5.0
    ref = np.where(np.abs(r) < d, 0.5 * r**2, d * (np.abs(r) - 0.5 * d))   # torch_optimizer.py:135-143
    np.testing.assert_array_equal(O.huber_rho(r, d), ref)
    w = O.huber_weight(r, d)
    np.testing.assert_allclose(w, np.where(np.abs(r) < d, 1.0, d / np.maximum(np.abs(r), 1e-300)))


def test_sym3_inverse():
    rng = np.random.default_rng(0)
    A = rng.normal(size=(200, 3, 3))
    V = A @ np.transpose(A, (0, 2, 1)) + 1e-3 * np.eye(3)
    V6 = np.stack([V[:, 0, 0], V[:, 0, 1], V[:, 0, 2], V[:, 1, 1], V[:, 1, 2], V[:, 2, 2]], axis=1)
    inv6, ok = O.sym3_inverse(V6)
    assert ok.all()
    np.testing.assert_allclose(O._sym6_to_mat(inv6) @ V, np.broadcast_to(np.eye(3), V.shape), atol=1e-8)
    bad, okb = O.sym3_inverse(np.array([[1.0, 2.0, 0.0, 1.0, 0.0, 1.0], [0.0, 0, 0, 0, 0, 0]]))
    assert not okb.any() and (bad == 0).all()


@pytest.mark.parametrize("name,seed", [("T0", 1), ("T1", 2)])
def test_schur_step_equals_full_system_step(name, seed):
    """dp, dl from the reduced camera system == solution of the full damped normal equations."""
    p = synth.make_problem(name, seed=seed)
    lam = 1e-3
    st = O.lm_step(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, lam, 5.0)
    lin = st["lin"]
    nc, npt = p.n_cam, p.n_pt
    N = 6 * nc + 3 * npt
    H = np.zeros((N, N)); g = np.zeros(N)
    for o in range(p.n_obs):
        c, j = int(p.obs_cam[o]), int(p.obs_pt[o])
        J = np.zeros((2, N))
        J[:, 6 * c:6 * c + 6] = lin["Jp"][o]
        J[:, 6 * nc + 3 * j:6 * nc + 3 * j + 3] = lin["Jl"][o]
        Wt = np.diag(lin["w"][o])
        H += J.T @ Wt @ J
        g += J.T @ Wt @ lin["r"][o]
    d = np.diag(H).copy()
    H[np.arange(N), np.arange(N)] += lam * np.where(d > 0, d, 1.0)
    fx = np.nonzero(np.repeat(p.fixed != 0, 6))[0]
    H[fx, :] = 0; H[:, fx] = 0; H[fx, fx] = 1; g[fx] = 0
    full = np.linalg.solve(H, -g)
    np.testing.assert_allclose(st["dp"], full[:6 * nc], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(st["dl"].ravel(), full[6 * nc:], rtol=1e-6, atol=1e-9)


def test_lm_monotone_and_flags():
    p = synth.make_problem("T2", seed=4)
    res = O.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv)
    assert (np.diff(res.cost_hist) <= 0).all()
    assert res.cost_hist[-1] < 0.3 * res.cost_hist[0]
    np.testing.assert_array_equal(res.R[0], p.R[0])           # gauge camera untouched
    np.testing.assert_array_equal(res.t[0], p.t[0])
    assert res.X.dtype == np.float64
    flags = O.outlier_flags(p.intr, res.R, res.t, res.X, p.obs_cam, p.obs_pt, p.obs_uv)
    np.testing.assert_array_equal(flags, res.flags)
    # injected gross outliers are (almost all) flagged
    assert (res.flags[p.outlier] == 1).mean() > 0.95


def test_rejected_step_leaves_state_and_raises_lambda():
    p = synth.make_problem("T0", seed=1)
    prm = O.LMParams(max_iter=6)
    res = O.solve_local(p.intr, p.R, p.t, p.fixed, p.X, p.obs_cam, p.obs_pt, p.obs_uv, prm)
    for i in range(6):
        if res.accepted[i] == 0:
            assert res.cost_hist[i + 1] == res.cost_hist[i]
            if i + 1 < 6:
                assert res.lambda_hist[i + 1] == pytest.approx(res.lambda_hist[i] * 10)
        else:
            assert res.cost_hist[i + 1] == res.cand_cost_hist[i] < res.cost_hist[i]


def test_pose_only_oracle_converges():
    pp = synth.make_pose_only(n_frames=6, n_match=200, seed=2)
    out = O.solve_pose_only(pp.intr, pp.R, pp.t, pp.X, pp.uv, pp.frame_offsets)
    assert (np.diff(out["cost_hist"], axis=1) <= 0).all()
    dR = np.einsum("nij,nkj->nik", out["R"], pp.R_true)
    ang = np.arccos(np.clip((np.trace(dR, axis1=1, axis2=2) - 1) / 2, -1, 1))
    assert ang.max() < 5e-3
    assert np.abs(out["t"] - pp.t_true).max() < 5e-2
    assert (out["flags"][pp.outlier] == 1).mean() > 0.9
