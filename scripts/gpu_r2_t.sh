#!/bin/bash
# round 2, GPU call T: ncu launch list (gpu__time_duration) of one C4 solve with the current kernels
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t_launches.csv python scripts/one_solve.py C4 auto 10 > gpurun_out/r2t_ncu.log 2>&1; echo "ncu rc $?"
python scripts/launch_shares.py gpurun_out/r2t_launches.csv | head -30
