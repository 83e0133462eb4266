#!/bin/bash
# round 2, GPU call J: bench N = 1 (driver command), then the ncu evidence of the same command
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc $?"; tail -2 gpurun_out/r2j_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 4 --warmup 1 > gpurun_out/r2j_ref.json 2> gpurun_out/r2j_ref.err; echo "ref rc $?"; tail -c 600 gpurun_out/r2j_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2j_launches_c4.csv python bench.py --steps 2 --warmup 1 --no-sub --no-cpu > gpurun_out/r2j_ncu1.log 2>&1; echo "ncu list rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_schur_panel -s 3 -c 1 -o gpurun_out/r2j_schur_panel_c4 -f python scripts/one_solve.py C4 panel 2 > gpurun_out/r2j_ncu2.log 2>&1; echo "ncu full rc $?"
ncu -i gpurun_out/r2j_schur_panel_c4.ncu-rep --page raw --csv > gpurun_out/r2j_schur_panel_c4_raw.csv 2>/dev/null
ncu -i gpurun_out/r2j_schur_panel_c4.ncu-rep --page details --csv > gpurun_out/r2j_schur_panel_c4_details.csv 2>/dev/null
for k in k_point_blocks k_cam_blocks; do
timeout 600 ncu --set full --clock-control none -k regex:$k -s 3 -c 1 -o gpurun_out/r2j_${k}_c4 -f python scripts/one_solve.py C4 panel 2 > gpurun_out/r2j_ncu3.log 2>&1
ncu -i gpurun_out/r2j_${k}_c4.ncu-rep --page raw --csv > gpurun_out/r2j_${k}_c4_raw.csv 2>/dev/null
done
ls -la gpurun_out/r2j_*
