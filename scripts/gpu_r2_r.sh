#!/bin/bash
# round 2, GPU call R: full GPU test suite + default bench line (N = 1)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1; echo "tests rc $?"; tail -4 gpurun_out/r2r_tests.log
timeout 900 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2r_bench.json'))
print(d['value'], d['e2e']['value'], d['kernel_ms_per_lm_iteration'])
print(json.dumps(d['parity'])[:400])
for c,v in d['configs'].items(): print(c, v['value'], v.get('e2e',{}).get('value'), v.get('parity',{}).get('ok'))
PY
tail -2 gpurun_out/r2r_bench.err
