#!/bin/bash
mkdir -p gpurun_out
CFG=${1:-C4}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_schur_panel -s 2 -c 1 -o gpurun_out/r2c_schur_${CFG} -f python scripts/one_solve.py $CFG panel 2 > gpurun_out/r2c_ncu.log 2>&1
tail -2 gpurun_out/r2c_ncu.log
ncu -i gpurun_out/r2c_schur_${CFG}.ncu-rep --page raw --csv > gpurun_out/r2c_schur_${CFG}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2c_schur_${CFG}.ncu-rep --page source --csv > gpurun_out/r2c_schur_${CFG}_source.csv 2>/dev/null
