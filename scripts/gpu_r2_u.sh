#!/bin/bash
# round 2, GPU call U: ncu --set full of the LM iteration's kernels at C4 (one launch each, after a warm-up solve) + launch list
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for k in k_point_blocks k_cam_blocks k_backsub; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 12 -c 1 -o gpurun_out/r2u_$k -f python scripts/one_solve.py C4 auto 10 > gpurun_out/r2u_$k.log 2>&1; echo "$k rc $?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2t_launches.csv python scripts/one_solve.py C4 auto 10 > gpurun_out/r2t_ncu.log 2>&1; echo "ncu rc $?"
python scripts/launch_shares.py gpurun_out/r2t_launches.csv | head -24
