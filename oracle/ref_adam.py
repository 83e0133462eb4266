"""Restatement of the reference's own local-BA algorithm  --  TEST / BASELINE INFRASTRUCTURE ONLY.

`/root/reference/visual_slam/optimization/torch_optimizer.py:103-206`: torch.optim.Adam on the
per-component-Huber reprojection loss, built by a Python loop over every observation, float32.
This file exists so that `bench.py --impl reference` can time the reference's CPU implementation
of the path on a box where `/root/reference` is absent (the GPU box); it is validated against
the unmodified reference through tests/golden (tests/test_ref_adam.py).

Behaviour reproduced on purpose: the skew matrix of `so3_exp` is rebuilt from
Python scalars (`:124-128`), which cuts it out of the autograd graph, so rotations receive no
gradient and never move; translations and points take Adam steps of about `lr`.
"""
from __future__ import annotations

import time
from typing import Optional

import numpy as np
import torch


def _rot_from_w(w: torch.Tensor) -> torch.Tensor:
    theta = torch.norm(w) + 1e-8
    k = w / theta
    kx, ky, kz = float(k[0]), float(k[1]), float(k[2])          # detached, as in the reference
    Km = torch.tensor([[0.0, -kz, ky], [kz, 0.0, -kx], [-ky, kx, 0.0]], dtype=torch.float32)
    return torch.eye(3) + torch.sin(theta) * Km + (1 - torch.cos(theta)) * (Km @ Km)


def _huber(res: torch.Tensor, delta: float) -> torch.Tensor:
    a = torch.abs(res)
    return torch.where(a < delta, 0.5 * res ** 2, delta * (a - 0.5 * delta))


def adam_local_ba(intr, R, t, fixed, X, obs_cam, obs_pt, obs_uv, n_iter: int, lr: float,
                  delta: float = 5.0, time_budget_s: Optional[float] = None):
    """Run `n_iter` Adam iterations the way the reference does.  Returns dict with the final
    R, t, X (float32-valued), the loss of every iteration and the seconds each one took."""
    n_cam = R.shape[0]
    Kt = torch.tensor(np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1]]), dtype=torch.float32)
    R0 = [torch.tensor(R[i], dtype=torch.float32) for i in range(n_cam)]
    ws, ts = [], []
    for i in range(n_cam):
        free = not bool(fixed[i])
        ws.append(torch.zeros(3, dtype=torch.float32, requires_grad=free))
        ts.append(torch.tensor(t[i], dtype=torch.float32).clone().detach().requires_grad_(free))
    Xt = torch.tensor(np.asarray(X, dtype=np.float32), requires_grad=True)
    params = [Xt]
    for i in range(n_cam):
        if ws[i].requires_grad:
            params += [ws[i], ts[i]]
    opt = torch.optim.Adam(params, lr=lr)
    obs = [(int(obs_cam[o]), int(obs_pt[o]), (float(obs_uv[o, 0]), float(obs_uv[o, 1])))
           for o in range(len(obs_cam))]
    losses, secs = [], []
    t_start = time.perf_counter()
    for _ in range(n_iter):
        t0 = time.perf_counter()
        opt.zero_grad()
        total = 0.0
        for (c, j, uv) in obs:
            Rc = _rot_from_w(ws[c]) @ R0[c] if ws[c].requires_grad else R0[c]
            Xc = Rc @ Xt[j] + ts[c]
            z = Xc[2] + 1e-8
            Kc = torch.tensor(Kt.numpy(), dtype=torch.float32)          # re-tensorised per observation (:167-168)
            pred = torch.stack([Kc[0, 0] * (Xc[0] / z) + Kc[0, 2], Kc[1, 1] * (Xc[1] / z) + Kc[1, 2]])
            res = pred - torch.tensor(uv, dtype=torch.float32)
            total = total + _huber(res, delta).sum()
        total.backward()
        opt.step()
        losses.append(float(total.item()))
        secs.append(time.perf_counter() - t0)
        if time_budget_s is not None and time.perf_counter() - t_start > time_budget_s:
            break
    with torch.no_grad():
        R_out = np.stack([(_rot_from_w(ws[i]) @ R0[i]).numpy() if ws[i].requires_grad else R0[i].numpy()
                          for i in range(n_cam)])
        t_out = np.stack([ts[i].detach().numpy() for i in range(n_cam)])
        X_out = Xt.detach().numpy().copy()
    return dict(R=R_out, t=t_out, X=X_out, losses=losses, secs=secs)


def adam_initial(intr, R_ref, t_ref, R_cur, t_cur, X_sel, uv_ref, uv_cur, n_iter: int, lr: float):
    """Restatement of `TorchOptimizer.optimize_initial`'s loop (torch_optimizer.py:268-305): Adam on
    the UNCONSTRAINED 3x3 `R_cur`, `t_cur` and the selected points, loss = SmoothL1(mean, beta = 1)
    of the reference view + of the current view, depth clamped at 1e-8, float32.  Inputs are the
    arrays the reference builds at :222-272 (see `packing.pack_initial`)."""
    K = torch.tensor(np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1]]), dtype=torch.float32)
    Rr = torch.tensor(np.asarray(R_ref), dtype=torch.float32)
    tr = torch.tensor(np.asarray(t_ref), dtype=torch.float32)
    Rc = torch.tensor(np.asarray(R_cur), dtype=torch.float32, requires_grad=True)
    tc = torch.tensor(np.asarray(t_cur), dtype=torch.float32, requires_grad=True)
    X = torch.tensor(np.asarray(X_sel, dtype=np.float32), requires_grad=True)
    pr = torch.tensor(np.asarray(uv_ref, dtype=np.float32))
    pc = torch.tensor(np.asarray(uv_cur, dtype=np.float32))
    opt = torch.optim.Adam([Rc, tc, X], lr=lr)
    loss_fn = torch.nn.SmoothL1Loss()
    losses = []
    for _ in range(n_iter):
        opt.zero_grad()
        Xr = (Rr @ X.T + tr.reshape(3, 1)).T
        ur = (K @ (Xr / Xr[:, 2:].clamp(min=1e-8)).T).T[:, :2]
        Xc = (Rc @ X.T + tc.reshape(3, 1)).T
        uc = (K @ (Xc / Xc[:, 2:].clamp(min=1e-8)).T).T[:, :2]
        loss = loss_fn(ur, pr) + loss_fn(uc, pc)
        loss.backward()
        opt.step()
        losses.append(float(loss.item()))
    return dict(R_cur=Rc.detach().numpy(), t_cur=tc.detach().numpy(), X=X.detach().numpy(), losses=losses)
