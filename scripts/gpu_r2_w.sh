#!/bin/bash
# round 2, GPU call W (N GPUs): default bench line under torchrun with the final kernels
N=$1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus $N > gpurun_out/r2w_bench_n$N.json 2> gpurun_out/r2w_bench_n$N.err; echo "bench rc $?"; python - <<PY
import json
d=json.load(open('gpurun_out/r2w_bench_n$N.json'))
print(d['value'], d['e2e']['value'], d['kernel_ms_per_lm_iteration'])
print(d['config']['parallelism'][:120])
print({k:(v.get('ok'), v.get('max_rel_cost_diff'), v.get('entries_differing_across_ranks')) for k,v in d['parity'].items() if isinstance(v,dict)}, d['parity'].get('ok'))
PY
tail -2 gpurun_out/r2w_bench_n$N.err
