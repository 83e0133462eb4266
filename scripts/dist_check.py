"""torchrun --nproc-per-node N scripts/dist_check.py [config]: sharded solve on N GPUs vs the
single-GPU solve of the same problem (rank 0 prints the differences)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import b200ba
from visual_slam_core_b200 import LMConfig, LocalBASolver, synth
from visual_slam_core_b200 import dist as badist
rank, lrank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "T4"
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-12
full = synth.make_problem(name, seed=2)
sh = full.shard_by_point(rank, world)
cfg = LMConfig(pcg_tol=tol, pcg_max_iter=3000)
s = LocalBASolver(cfg, device=lrank)
badist.init_solver(s, rank, world, device=dev)
for _ in range(2):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.time()
    res = s.solve_local(sh.intr, sh.R, sh.t, sh.fixed, sh.X, sh.obs_cam, sh.obs_pt, sh.obs_uv, sh.pt_offsets)
    dt = time.time() - t0
X = badist.gather_points(res.X, sh.meta["pt_range"], full.n_pt)
allR = [None] * world
dist.all_gather_object(allR, (res.R, res.t, res.cost_hist))
if rank == 0:
    one = LocalBASolver(cfg, device=lrank)
    ref = one.solve_local(full.intr, full.R, full.t, full.fixed, full.X, full.obs_cam, full.obs_pt, full.obs_uv, full.pt_offsets)
    print(f"world {world} {name}: shard obs {sh.n_obs}, wall {dt*1e3:.2f} ms, gpu {res.gpu_ms:.2f} ms (single GPU {ref.gpu_ms:.2f} ms)")
    print("accepted equal", np.array_equal(res.accepted, ref.accepted), res.accepted.tolist())
    print("cost rel diff", float(np.max(np.abs(res.cost_hist - ref.cost_hist) / ref.cost_hist)))
    print("pose diff R", float(np.abs(res.R - ref.R).max()), "t", float(np.abs(res.t - ref.t).max()), "X", float(np.abs(X - ref.X).max()))
    print("ranks identical poses", all(np.array_equal(allR[0][0], a[0]) and np.array_equal(allR[0][1], a[1]) for a in allR),
          "identical cost hist", all(np.array_equal(allR[0][2], a[2]) for a in allR))
dist.destroy_process_group()
