"""torchrun --nproc-per-node N scripts/dist_check.py <window> [--plugin]: the sharded solve on N GPUs against the
single-GPU solve of the same window; rank 0 prints ONE JSON line (tests/test_gpu_dist.py asserts on it).

<window>: a synth config (T2, T4, C4 ...) or L1 / L2 of tests/golden/make_large_golden.py.
--plugin: go through `B200Optimizer.optimize_global` on an object graph instead of the packed-array API."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import b200ba  # noqa: E402,F401
from helpers import gauge_aligned_trans_err, rot_angle, scale_norm_trans_err  # noqa: E402
from visual_slam_core_b200 import B200Optimizer, LMConfig, LocalBASolver, synth  # noqa: E402
from visual_slam_core_b200 import dist as badist  # noqa: E402

rank, lrank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
saved = os.dup(1)
os.dup2(2, 1)                              # NCCL banners go to stderr: stdout carries one JSON line
dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "T4"
plugin = "--plugin" in sys.argv
if name in ("L1", "L2"):
    import make_large_golden as G
    full = G.make_window(name)
else:
    full = synth.make_problem(name, seed=2)


def terr(Ra, ta, Rb, tb):
    if full.n_cam > 40:
        return float(gauge_aligned_trans_err(Ra, ta, Rb, tb)[0].max())
    return float(scale_norm_trans_err(Ra, ta, Rb, tb).max())


if not plugin:
    sh = full.shard_by_point(rank, world)
    s = LocalBASolver(LMConfig(), device=lrank)
    badist.init_solver(s, rank, world, device=dev)
    res = s.solve_local(sh.intr, sh.R, sh.t, sh.fixed, sh.X, sh.obs_cam, sh.obs_pt, sh.obs_uv, sh.pt_offsets)
    lo, hi = sh.meta["pt_range"]
    X = badist.gather_points(res.X, (lo, hi), full.n_pt)
    o0, o1 = int(full.pt_offsets[lo]), int(full.pt_offsets[hi])
    flags = badist.gather_flags(res.flags, (o0, o1), full.n_obs)
    allR = [None] * world
    dist.all_gather_object(allR, (res.R, res.t, res.cost_hist, res.solve_fail))
    out = None
    if rank == 0:
        one = LocalBASolver(LMConfig(), device=lrank)
        ref = one.solve_local(full.intr, full.R, full.t, full.fixed, full.X, full.obs_cam, full.obs_pt, full.obs_uv, full.pt_offsets)
        out = {"window": name, "world": world, "n_cam": full.n_cam, "n_pt": full.n_pt, "n_obs": full.n_obs,
               "solver": res.extra["solver_used"], "band_width": res.extra["band_width"],
               "gpu_ms_sharded": res.gpu_ms, "gpu_ms_single": ref.gpu_ms,
               "accept_sequence_equal": bool(np.array_equal(res.accepted, ref.accepted)),
               "accepted": res.accepted.tolist(),
               "max_rel_cost_diff": float(np.max(np.abs(res.cost_hist - ref.cost_hist) / ref.cost_hist)),
               "max_rot_err_rad": float(rot_angle(res.R, ref.R).max()),
               "max_trans_err": terr(res.R, res.t, ref.R, ref.t),
               "max_point_diff": float(np.abs(X - ref.X).max()),
               "flag_mismatches": int((flags != ref.flags).sum()),
               "solve_fail_any": int(max(int(np.max(a[3])) if len(a[3]) else 0 for a in allR)),
               "ranks_identical_poses": bool(all(np.array_equal(allR[0][0], a[0]) and np.array_equal(allR[0][1], a[1]) for a in allR)),
               "ranks_identical_cost_hist": bool(all(np.array_equal(allR[0][2], a[2]) for a in allR))}
else:
    import tempfile
    import refgraph
    cam, kfs, mps = refgraph.graph_from_problem(full)
    opt = B200Optimizer(log_dir=tempfile.mkdtemp(), device=f"cuda:{lrank}")
    ok = opt.optimize_global(kfs, mps)                      # torch.distributed is initialised: sharded by point
    R = np.array([kf.R_w2c for kf in kfs]); t = np.array([kf.t_w2c for kf in kfs]); X = np.array([m.position for m in mps])
    allM = [None] * world
    dist.all_gather_object(allM, (bool(ok), R, t, X, opt.last_result.cost_hist if ok else None))
    out = None
    if rank == 0:
        cam1, kfs1, mps1 = refgraph.graph_from_problem(full)
        one = B200Optimizer(log_dir=tempfile.mkdtemp(), device=f"cuda:{lrank}")
        one._dist = (0, 1, None)                            # force the single-process path
        ok1 = one.optimize_global(kfs1, mps1)
        R1 = np.array([kf.R_w2c for kf in kfs1]); t1 = np.array([kf.t_w2c for kf in kfs1])
        out = {"window": name, "world": world, "ok_all_ranks": bool(all(a[0] for a in allM) and ok1),
               "maps_identical_across_ranks": bool(all(np.array_equal(allM[0][1], a[1]) and np.array_equal(allM[0][2], a[2])
                                                       and np.array_equal(allM[0][3], a[3]) for a in allM)),
               "max_rot_err_rad": float(rot_angle(R, R1).max()), "max_trans_err": terr(R, t, R1, t1),
               "max_rel_cost_diff": float(np.max(np.abs(allM[0][4] - one.last_result.cost_hist) / one.last_result.cost_hist))}
dist.barrier()
os.dup2(saved, 1)
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
