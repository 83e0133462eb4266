#!/bin/bash
# mma kernel bring-up: parity tests first, then timings of C2 / C4 against the tiled kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mma" > gpurun_out/f_pytest_mma.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest_mma.log
timeout 300 python scripts/big_run.py C2 1e-8 10 tiled,mma > gpurun_out/f_big_C2.log 2>&1
timeout 600 python scripts/big_run.py C4 1e-8 10 tiled,mma > gpurun_out/f_big_C4.log 2>&1
tail -5 gpurun_out/f_pytest_mma.log; cat gpurun_out/f_big_C2.log gpurun_out/f_big_C4.log
