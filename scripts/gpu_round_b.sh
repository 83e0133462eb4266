#!/bin/bash
# one gpurun call: bench (C2), launch list, full ncu capture of the LM kernels, C4 single GPU bench
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/h_bench_c2.json 2> gpurun_out/h_bench_c2.err; echo "bench rc $?"
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/h_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/h_ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/h_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_linearize_mma|k_rcs_cholesky|k_backsub' -s 30 -c 6 -o gpurun_out/h_prof_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/h_ncu2.log 2>&1
python bench.py --config C4 --steps 2 --warmup 3 --no-cpu > gpurun_out/h_bench_c4.json 2> gpurun_out/h_bench_c4.err; echo "bench c4 rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/h_bench_ref.json 2> gpurun_out/h_bench_ref.err; echo "ref rc $?"
ls -la gpurun_out | tail -12
