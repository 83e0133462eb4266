#!/bin/bash
# one gpurun call: GPU tests, bench (C2), launch list, full ncu capture, large-path timing
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/a_pytest.log
tail -3 gpurun_out/a_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/a_bench_c2.json 2> gpurun_out/a_bench_c2.err; echo "bench rc $?"
tail -c 1500 gpurun_out/a_bench_c2.json
python scripts/big_run.py C4 1e-8 5 > gpurun_out/a_big_c4.log 2>&1; echo "big rc $?"; tail -12 gpurun_out/a_big_c4.log
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/a_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/a_ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/a_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_linearize_tiled|k_rcs_cholesky|k_backsub' -s 30 -c 6 -o gpurun_out/a_prof_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/a_ncu2.log 2>&1
ls -la gpurun_out
