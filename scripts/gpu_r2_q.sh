#!/bin/bash
# round 2, GPU call Q: ncu --set full capture (with source) of one k_schur_panel launch at C4
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_schur_panel -s 2 -c 1 -o gpurun_out/r2q_panel -f python scripts/one_solve.py C4 panel 2 > gpurun_out/r2q_ncu.log 2>&1; echo "ncu rc $?"
tail -3 gpurun_out/r2q_ncu.log; ls -la gpurun_out/
