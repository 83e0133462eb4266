#!/bin/bash
export BA_NVCC_EXTRA="-DBA_PANEL_TRACE"
python -c "
import b200ba
from visual_slam_core_b200 import _build
_build.build(force=True)" 
for which in 0 1; do
echo "== shard $which"
BA_DBG_TIMES=1 python scripts/one_solve.py C4 panel 1 2 $which 2>&1 | grep "ba dbg" | tail -1 | python scripts/parse_trace.py
done
