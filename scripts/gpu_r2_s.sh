#!/bin/bash
# round 2, GPU call S (2 GPUs): sharded CUDA path -- asserted parity tests, then the default bench line under torchrun
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1; echo "tests rc $?"; tail -5 gpurun_out/r2s_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/r2s_bench_n2.json 2> gpurun_out/r2s_bench_n2.err; echo "bench rc $?"; head -c 1500 gpurun_out/r2s_bench_n2.json; echo; tail -3 gpurun_out/r2s_bench_n2.err
