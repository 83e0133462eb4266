#!/bin/bash
# one GPU iteration: parity tests, then A/B bench lines of the kernels changed this session
set -u
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/s_pytest.log 2>&1
echo "pytest rc=$?" ; tail -3 $out/s_pytest.log
line() { python -c "
import json,sys
t=sys.stdin.read().strip().splitlines()
try:
    d=json.loads(t[-1]); print('$1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v,4) for k,v in d['kernel_ms_per_lm_iteration'].items()}, d.get('parity_vs_oracle'))
except Exception as e:
    print('$1 FAILED', e, t[-3:])
"; }
python bench.py --steps 20 --warmup 3 --no-cpu 2>$out/s_c2.err | tee $out/s_bench_c2.json | line C2-default
BA_CHOL=v1 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | line C2-chol-v1
BA_BACKSUB_GROUP=32 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | line C2-group32
BA_BACKSUB_GROUP=8 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | line C2-group8
python bench.py --config C1 --steps 20 --warmup 3 --no-cpu 2>/dev/null | tee $out/s_bench_c1.json | line C1-default
BA_CHOL=v1 BA_BACKSUB_GROUP=32 python bench.py --config C1 --steps 20 --warmup 3 --no-cpu 2>/dev/null | line C1-old
python bench.py --config C4 --steps 3 --warmup 3 --no-cpu 2>$out/s_c4.err | tee $out/s_bench_c4.json | line C4-default
BA_BACKSUB_GROUP=16 python bench.py --config C4 --steps 3 --warmup 3 --no-cpu 2>/dev/null | line C4-group16
