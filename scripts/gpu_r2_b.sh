#!/bin/bash
mkdir -p gpurun_out
for cfg in C2 C4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_launch_${cfg}.csv python scripts/one_solve.py $cfg panel 2 > gpurun_out/r2b_ncu_${cfg}.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2b_launch_${cfg}.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); 
agg={}
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    k=r[ki][:60]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:14]: print("${cfg}", f"{k:60s} n={n:3d} total={t/1e3:10.1f} us avg={t/n/1e3:9.1f} us")
PY
done
timeout 600 python scripts/panel_check.py time 2>&1 | tail -14
