#!/bin/bash
# round 2, GPU call S (2 GPUs): sharded CUDA path -- asserted parity tests, then the default bench line under torchrun
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1; echo "tests rc $?"; tail -5 gpurun_out/r2s_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 scripts/dist_check.py L1 2>/dev/null | tail -1 | cut -c1-900
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/r2s_bench_n2.json 2> gpurun_out/r2s_bench_n2.err; echo "bench rc $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s_bench_n2.json'))
print(d['value'], d['kernel_ms_per_lm_iteration'])
print(json.dumps(d['parity'])[:700])
PY
BA_PEER_REDUCE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NCCL all-reduce path:', d['value'], d['kernel_ms_per_lm_iteration'])"
