#!/bin/bash
mkdir -p gpurun_out
for which in 0 3 7; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_schur_panel -c 20 --csv --log-file gpurun_out/r2l_w${which}.csv python scripts/one_solve.py C4 panel 10 8 $which > gpurun_out/r2l_w${which}.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2l_w${which}.csv")) if len(r)>5]
hdr=rows[0]; vi=hdr.index("Metric Value")
print("shard ${which}/8 k_schur_panel us per launch:", [round(float(r[vi].replace(",",""))/1e3) for r in rows[1:]])
PY
done
