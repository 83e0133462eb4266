#!/bin/bash
# 2-GPU check of the v8 kernels: C2 replicas (the driver's scaling run) and C4 sharded by point
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/w_c2_n2.json 2> gpurun_out/w_c2_n2.err; echo "c2 n2 rc $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config C4 --steps 3 --warmup 3 --no-cpu > gpurun_out/w_c4_n2.json 2> gpurun_out/w_c4_n2.err; echo "c4 n2 rc $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 > gpurun_out/w_ref_n2.json 2> gpurun_out/w_ref_n2.err; echo "ref n2 rc $?"
tail -c 600 gpurun_out/w_c2_n2.json; echo; tail -c 900 gpurun_out/w_c4_n2.json; echo; tail -c 400 gpurun_out/w_ref_n2.json
