#!/bin/bash
# round 2, GPU call N (re-entry check): full GPU test suite, default bench line, sanitizer passes on the small solves
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2n_tests.log 2>&1; echo "tests rc $?"; tail -14 gpurun_out/r2n_tests.log
timeout 900 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc $?"; head -c 1500 gpurun_out/r2n_bench.json; tail -3 gpurun_out/r2n_bench.err
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r2n_sanitizer_$tool.log python scripts/sanitize_run.py --quick > gpurun_out/r2n_sanitizer_$tool.out 2>&1; echo "$tool rc $?"
  tail -4 gpurun_out/r2n_sanitizer_$tool.out; tail -3 gpurun_out/r2n_sanitizer_$tool.log
done
