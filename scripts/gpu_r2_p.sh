#!/bin/bash
# round 2, GPU call P: k_schur_panel variants -- solves on band windows, kernel times, clock64 trace (diagnostic build)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python scripts/panel_check.py lin solve 2>&1 | tail -5
timeout 600 python scripts/parts_c4.py
BA_LIB_PATH=$PWD/visual-slam-core_b200/libb200ba_trace.so BA_DBG_TIMES=1 timeout 600 python scripts/one_solve.py C4 panel 2 > gpurun_out/r2p_trace.log 2>&1; echo "trace rc $?"
grep "^cta" gpurun_out/r2p_trace.log | tail -148 | sort -t' ' -k11 -n | tail -2
python scripts/parse_trace.py < gpurun_out/r2p_trace.log | head -8
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_gpu_lane_groups.py -m gpu -x -q 2>&1 | tail -2
