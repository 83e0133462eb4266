#!/bin/bash
# round 2, GPU call O: clock64 trace of k_schur_panel at C4 (diagnostic build, BA_PANEL_TRACE)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
BA_LIB_PATH=$PWD/visual-slam-core_b200/libb200ba_trace.so BA_DBG_TIMES=1 timeout 600 python scripts/one_solve.py C4 panel 2 > gpurun_out/r2o_trace.log 2>&1; echo "trace rc $?"
grep "^cta" gpurun_out/r2o_trace.log | tail -148 | sort -t' ' -k11 -n | tail -2
python scripts/parse_trace.py < gpurun_out/r2o_trace.log
