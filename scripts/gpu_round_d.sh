#!/bin/bash
# C3 bench (pose-only) + its reference arm; ncu capture of the C4 kernels (HBM-regime evidence)
mkdir -p gpurun_out
python bench.py --config C3 --steps 20 --warmup 3 > gpurun_out/m_bench_c3.json 2> gpurun_out/m_bench_c3.err; echo "c3 rc $?"
python bench.py --config C3 --impl reference --steps 3 --warmup 1 > gpurun_out/m_bench_c3_ref.json 2> gpurun_out/m_bench_c3_ref.err; echo "c3 ref rc $?"
timeout 300 python -m pytest tests/test_gpu_pose_only.py -x -q -m gpu > gpurun_out/m_pytest_pose.log 2>&1; tail -2 gpurun_out/m_pytest_pose.log
python scripts/big_run.py C4 1e-8 2 auto > gpurun_out/m_plain_c4.log 2>&1 &&
ncu --set full --clock-control none -k regex:'k_linearize_mma|k_rcs_band|k_backsub|k_finalize|k_init' -s 10 -c 5 -o gpurun_out/m_prof_c4 -f python scripts/big_run.py C4 1e-8 2 auto > gpurun_out/m_ncu_c4.log 2>&1
tail -c 1500 gpurun_out/m_bench_c3.json
