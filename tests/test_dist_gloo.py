"""The shard-by-point multi-GPU design, exercised on CPU with world_size-2 (and 3) gloo groups:
each rank linearises only its point shard (oracle code), the ranks all-reduce exactly what the CUDA
path all-reduces (S | b | diagU | cost, then the candidate cost), solve the reduced system
redundantly and back-substitute their own points.  The trajectory must equal the single-process
oracle's.  Also covers the NCCL-id broadcast helper and the point gather."""
import os
import socket
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
import b200ba  # noqa: E402,F401  (spawned workers import this module without conftest)

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lm_oracle as O
from visual_slam_core_b200 import synth
from visual_slam_core_b200 import dist as badist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _allreduce(a):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    dist.all_reduce(t)
    return t.numpy()


def _sharded_lm(rank, world, name, seed, max_iter):
    """Sharded LM with the oracle's building blocks; mirrors enqueue_lm_loop in csrc/ba_api.cu."""
    full = synth.make_problem(name, seed=seed)
    sh = full.shard_by_point(rank, world)
    prm = O.LMParams(max_iter=max_iter)
    R, t, X = sh.R.copy(), sh.t.copy(), sh.X.copy()
    oc, op = sh.obs_cam.astype(np.int64), sh.obs_pt.astype(np.int64)
    n_cam, n = sh.n_cam, 6 * sh.n_cam
    lam = prm.lambda0
    cost = float(_allreduce(np.array([O.reference_cost(sh.intr, R, t, X, oc, op, sh.obs_uv)]))[0])
    hist, acc = [cost], []
    i6 = np.arange(6)
    for _ in range(max_iter):
        lin = O.linearize(sh.intr, R, t, sh.fixed, X, oc, op, sh.obs_uv, prm.huber_delta)
        # partial reduced system of this shard: undamped pose blocks, Schur term of the local points
        S0, b0, Wb, Vinv, _ = O.build_rcs(lin, lam, np.zeros_like(sh.fixed), oc, op, n_cam, sh.n_pt)
        dU = lin["U"][:, i6, i6].reshape(-1)
        S0[np.arange(n), np.arange(n)] -= O.damp_diag(dU, lam)          # keep S un-damped in the exchange
        red = _allreduce(np.concatenate([S0.ravel(), b0, dU]))
        S = red[: n * n].reshape(n, n)
        b = red[n * n: n * n + n]
        dUa = red[n * n + n:]
        S[np.arange(n), np.arange(n)] += O.damp_diag(dUa, lam)
        fx = np.nonzero(np.repeat(sh.fixed != 0, 6))[0]
        S[fx, :] = 0; S[:, fx] = 0; S[fx, fx] = 1.0; b[fx] = 0
        dp, ok = O.solve_rcs(S, b)
        rhs = lin["gl"] + (Wb.T @ dp).reshape(sh.n_pt, 3)
        dl = -np.einsum("nab,nb->na", Vinv, rhs)
        dpc = dp.reshape(n_cam, 6)
        Rn = O.so3_exp(dpc[:, :3]) @ R
        tn = t + dpc[:, 3:]
        free = sh.fixed == 0
        Rn[~free] = R[~free]; tn[~free] = t[~free]
        Xn = X + dl
        cand = float(_allreduce(np.array([O.reference_cost(sh.intr, Rn, tn, Xn, oc, op, sh.obs_uv)]))[0])
        good = bool(ok and np.isfinite(cand) and cand < cost * (1 - prm.accept_rtol))
        acc.append(int(good))
        if good:
            R, t, X, cost = Rn, tn, Xn, cand
            lam = max(lam * prm.lambda_down, prm.lambda_min)
        else:
            lam = min(lam * prm.lambda_up, prm.lambda_max)
        hist.append(cost)
    Xall = badist.gather_points(X, sh.meta["pt_range"], full.n_pt)
    return np.array(hist), np.array(acc), R, t, Xall


def _worker(rank, world, port, name, seed, max_iter, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = badist.broadcast_unique_id(lambda: bytes(range(128)), rank)
        hist, acc, R, t, X = _sharded_lm(rank, world, name, seed, max_iter)
        q.put((rank, uid, hist, acc, R, t, X))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_lm_equals_single_process_oracle(world):
    name, seed, max_iter = "T1", 4, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, seed, max_iter, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in range(world)], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = synth.make_problem(name, seed=seed)
    ref = O.solve_local(full.intr, full.R, full.t, full.fixed, full.X, full.obs_cam, full.obs_pt, full.obs_uv,
                        O.LMParams(max_iter=max_iter))
    for rank, uid, hist, acc, R, t, X in outs:
        assert uid == bytes(range(128))                      # id broadcast reached every rank
        np.testing.assert_array_equal(acc, ref.accepted)
        np.testing.assert_allclose(hist, ref.cost_hist, rtol=1e-10)
        np.testing.assert_allclose(R, ref.R, atol=1e-9)      # cameras identical on every rank
        np.testing.assert_allclose(t, ref.t, atol=1e-8)
        np.testing.assert_allclose(X, ref.X, atol=1e-6)      # gathered shards = full point set
