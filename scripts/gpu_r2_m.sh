#!/bin/bash
# round 2, GPU call M: the row-split cyclic-reduction kernels -- correctness (band windows vs the oracle fixtures),
# per-LM-iteration timing, launch list of one C4 solve
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python scripts/panel_check.py solve 2>&1 | tail -4
timeout 600 python scripts/panel_check.py time 2>&1 | grep panel
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_launches.csv python scripts/one_solve.py C4 panel 3 > gpurun_out/r2m_ncu.log 2>&1; echo "ncu rc $?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2m_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); gi = hdr.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[1:]:
    if "bcr" not in r[ki]: continue
    k = r[ki].split("(")[0][-16:] + " grid " + r[gi]
    agg.setdefault(k, []).append(float(r[vi].replace(",", "")) / 1e3)
for k, v in agg.items(): print(f"{k:44s} n={len(v):3d} avg {sum(v)/len(v):7.2f} us")
PY
