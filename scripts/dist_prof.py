"""torchrun --nproc-per-node N scripts/dist_prof.py [C4]: per-rank device time per kernel class of the sharded solve."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import b200ba
from visual_slam_core_b200 import LMConfig, LocalBASolver, synth
from visual_slam_core_b200 import dist as badist
rank, lrank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
dist.init_process_group("nccl", device_id=dev)
full = synth.make_problem(sys.argv[1] if len(sys.argv) > 1 else "C4", seed=0)
sh = full.shard_by_point(rank, world)
for prof in (True, False):
    s = LocalBASolver(LMConfig(profile=prof), device=lrank)
    badist.init_solver(s, rank, world, device=dev)
    for _ in range(3):
        dist.barrier(); torch.cuda.synchronize()
        r = s.solve_local(sh.intr, sh.R, sh.t, sh.fixed, sh.X, sh.obs_cam, sh.obs_pt, sh.obs_uv, sh.pt_offsets)
    e = r.extra
    print(f"rank {rank}/{world} profile={prof}: gpu {r.gpu_ms:.2f} ms; per iteration lin {e['ms_linearize']/10:.3f} solve {e['ms_solve']/10:.3f} "
          f"back {e['ms_backsub']/10:.3f} comm {e['ms_comm']/10:.3f} other/step {e['ms_other']:.3f}; shard obs {sh.n_obs}", flush=True)
    s.close()
dist.destroy_process_group()
