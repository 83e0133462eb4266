"""The sharded CUDA path (points sharded over the ranks, NCCL all-reduce of the reduced camera system, SURVEY
8e-E1) against the single-GPU solve of the same window: SURVEY 4 "Multi-GPU" row, asserted.  Needs >= 2 GPUs in
the box (`gpurun --gpus 2`); with one GPU the tests skip and the driver's scaling run carries the same parity
block in its bench lines."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _run(world, args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "dist_check.py"), *args]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    return json.loads(line)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("name,kind", [("T2", "dense"), ("T4", "band"), ("L1", "band")])
def test_sharded_solve_equals_the_single_gpu_solve(name, kind):
    out = _run(2, [name], 29531 + len(name) + len(kind))
    print("\n", out)
    assert out["accept_sequence_equal"] and out["ranks_identical_poses"] and out["ranks_identical_cost_hist"]
    assert out["max_rel_cost_diff"] <= 1e-6
    assert out["max_rot_err_rad"] <= 1e-5 and out["max_trans_err"] <= 1e-5
    assert out["flag_mismatches"] == 0
    assert out["solve_fail_any"] == 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_plugin_global_ba_shards_under_torch_distributed():
    """`B200Optimizer.optimize_global` in a 2-process job: same map in, same map out on both ranks, equal to the
    single-process result within the parity tolerances."""
    out = _run(2, ["T3", "--plugin"], 29561)
    print("\n", out)
    assert out["ok_all_ranks"] and out["maps_identical_across_ranks"]
    assert out["max_rot_err_rad"] <= 1e-5 and out["max_trans_err"] <= 1e-5 and out["max_rel_cost_diff"] <= 1e-6
