#!/bin/bash
export BA_NVCC_EXTRA="-DBA_PANEL_TRACE"
python -c "
import b200ba
from visual_slam_core_b200 import _build
_build.build(force=True)" 
for which in 3 7; do
echo "== shard $which of 8"
BA_DBG_TIMES=1 python scripts/one_solve.py C4 panel 1 8 $which 2>&1 | grep -E "^cta" | sort -u -t: -k1,1 | sort -k13 -n | awk '{print $2, "pts", $4, "panels", $6, "reb", $8, "fb", $11, "clk", $13}' | (head -3; echo ...; tail -6)
done
