#!/bin/bash
# round 2, GPU call R: full GPU test suite + default bench line (N = 1) with the class-order panel path
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r2r_tests.log
timeout 900 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc $?"; head -c 600 gpurun_out/r2r_bench.json; echo; tail -3 gpurun_out/r2r_bench.err
