#!/bin/bash
# round 2, GPU call F: full GPU test suite with the panel path in, bench at N = 1
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc $?"; grep -E "passed|failed|\[L1\]|\[L2\]|\[C2\]|\[C3\]|p3p|call site|self-sens|FAILED|Error" gpurun_out/r2f_pytest.log | head -40
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc $?"; tail -3 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2f_bench.json"))
print("C4 value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], d["kernel_ms_per_lm_iteration"], "roofline", d["roofline"]["frac"], d["roofline"]["fp64"]["frac"], "parity", d["parity"].get("ok"))
for k,v in d.get("configs",{}).items():
    print(k, v.get("value"), "e2e", v.get("e2e",{}).get("value"), v.get("e2e",{}).get("ms_per_call"), "abi", v.get("e2e_abi",{}).get("value"), "parity", (v.get("parity") or {}).get("ok"), v.get("error"))
PY
