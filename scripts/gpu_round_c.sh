#!/bin/bash
# one gpurun call (1 GPU): bench C2 / C1 / C4, launch list, full ncu capture of the LM kernels, reference arm
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/k_bench_c2.json 2> gpurun_out/k_bench_c2.err; echo "bench rc $?"
python bench.py --config C1 --steps 20 --warmup 3 > gpurun_out/k_bench_c1.json 2> gpurun_out/k_bench_c1.err; echo "bench c1 rc $?"
python bench.py --config C4 --steps 3 --warmup 3 --no-cpu > gpurun_out/k_bench_c4.json 2> gpurun_out/k_bench_c4.err; echo "bench c4 rc $?"
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/k_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/k_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/k_ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/k_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_linearize_mma|k_rcs_cholesky|k_backsub' -s 30 -c 6 -o gpurun_out/k_prof_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/k_ncu2.log 2>&1
ls -la gpurun_out | tail -8
