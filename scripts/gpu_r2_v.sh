#!/bin/bash
# round 2, GPU call V (8 GPUs): default bench line under torchrun, peer-memory exchange vs the NCCL all-reduce
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --no-cpu > gpurun_out/r2v_bench_n8.json 2> gpurun_out/r2v_bench_n8.err; echo "bench rc $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2v_bench_n8.json'))
print(d['value'], d['e2e']['value'], d['kernel_ms_per_lm_iteration'])
print(json.dumps(d['parity'])[:600])
PY
BA_PEER_REDUCE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 8 --no-cpu --no-sub 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NCCL all-reduce path:', d['value'], d['kernel_ms_per_lm_iteration'])"
tail -2 gpurun_out/r2v_bench_n8.err
