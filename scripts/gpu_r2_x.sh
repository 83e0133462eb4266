#!/bin/bash
# round 2, GPU call X: final check -- full GPU test suite, default bench line, ncu capture of k_schur_panel and the launch list
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_tests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r2x_tests.log
timeout 900 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2x_bench.json'))
print(d['value'], d['e2e']['value'], d['kernel_ms_per_lm_iteration'])
print(d['roofline']['frac'], d['roofline']['fp64']['frac'], d['parity']['ok'])
for c,v in d['configs'].items(): print(c, v['value'], v.get('e2e',{}).get('value'), v.get('parity',{}).get('ok'))
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_schur_panel -s 12 -c 1 -o gpurun_out/r2x_k_schur_panel -f python scripts/one_solve.py C4 auto 10 > gpurun_out/r2x_ncu.log 2>&1; echo "ncu rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2x_launches.csv python scripts/one_solve.py C4 auto 10 > gpurun_out/r2x_ncu2.log 2>&1; echo "ncu rc $?"
python scripts/launch_shares.py gpurun_out/r2x_launches.csv | head -12
