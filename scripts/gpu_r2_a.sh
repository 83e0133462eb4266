#!/bin/bash
# round 2, GPU call A: state of the tree after the host-side rework (bench contract, parity tests, ABI fixes)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc $?"; tail -c 1500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_run.py --quick > gpurun_out/r2a_memcheck.log 2>&1; echo "memcheck rc $?"; tail -4 gpurun_out/r2a_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_run.py --quick > gpurun_out/r2a_racecheck.log 2>&1; echo "racecheck rc $?"; tail -4 gpurun_out/r2a_racecheck.log
timeout 400 compute-sanitizer --tool initcheck --error-exitcode 9 python scripts/sanitize_run.py --quick > gpurun_out/r2a_initcheck.log 2>&1; echo "initcheck rc $?"; tail -4 gpurun_out/r2a_initcheck.log
