#!/bin/bash
# round 2, GPU call G (2 GPUs): sharded CUDA path asserted against the single-GPU solve, plugin global BA, bench N = 2
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L
timeout 1500 python -m pytest tests/test_gpu_dist.py -m gpu -q -s > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc $?"; grep -E "passed|failed|window|FAILED|Error|assert" gpurun_out/r2g_pytest.log | head -30
NCCL_DEBUG=WARN timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "bench rc $?"; tail -3 gpurun_out/r2g_bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2g_bench_n2.json"))
print("C4 N=2 value", d["value"], "e2e", d["e2e"]["value"], d["kernel_ms_per_lm_iteration"])
print(json.dumps(d["parity"], indent=1)[:3000])
PY
