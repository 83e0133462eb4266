#!/bin/bash
export BA_NVCC_EXTRA="-DBA_PANEL_TRACE"
python -c "
import b200ba
from visual_slam_core_b200 import _build
_build.build(force=True)" 
BA_DBG_TIMES=1 python scripts/one_solve.py ${1:-C4} panel 1 2>&1 | tail -2 | cut -c1-1500
