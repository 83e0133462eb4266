#!/bin/bash
# profile round of the v8 kernels (1 GPU): parity tests, bench lines C2 / C1 / C3 / C4, launch list,
# full ncu capture of the LM kernels, smoke
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/y_pytest.log 2>&1; echo "pytest rc $?"; tail -1 gpurun_out/y_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/y_bench_c2.json 2> gpurun_out/y_bench_c2.err; echo "c2 rc $?"
python bench.py --config C1 --steps 20 --warmup 3 > gpurun_out/y_bench_c1.json 2> gpurun_out/y_bench_c1.err; echo "c1 rc $?"
python bench.py --config C3 --steps 20 --warmup 3 > gpurun_out/y_bench_c3.json 2> gpurun_out/y_bench_c3.err; echo "c3 rc $?"
python bench.py --config C4 --steps 3 --warmup 3 --no-cpu > gpurun_out/y_bench_c4.json 2> gpurun_out/y_bench_c4.err; echo "c4 rc $?"
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/y_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/y_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/y_ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/y_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_linearize_mma|k_rcs_cholesky|k_backsub' -s 30 -c 6 -o gpurun_out/y_prof_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/y_ncu2.log 2>&1
python __graft_entry__.py smoke > gpurun_out/y_smoke.log 2>&1; tail -1 gpurun_out/y_smoke.log
ls -la gpurun_out | grep " y_" | tail -14
