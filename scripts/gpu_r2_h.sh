#!/bin/bash
mkdir -p gpurun_out
for which in 0 1; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2h_launch_w${which}.csv python scripts/one_solve.py C4 panel 2 2 $which > gpurun_out/r2h_ncu_w${which}.log 2>&1
  tail -2 gpurun_out/r2h_ncu_w${which}.log
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2h_launch_w${which}.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); 
agg={}
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    k=r[ki][:60]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:7]: print("shard ${which}/2", f"{k:60s} n={n:3d} avg={t/n/1e3:9.1f} us")
PY
done
