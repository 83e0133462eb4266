#!/bin/bash
# round 2, GPU call K (8 GPUs): the driver's scaling command at N = 4 and 8 on C4, C5 at N = 8
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | wc -l
for n in 4 8; do
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2k_c4_n$n.json 2> gpurun_out/r2k_c4_n$n.err; echo "C4 N=$n rc $?"; tail -2 gpurun_out/r2k_c4_n$n.err
done
NCCL_DEBUG=WARN timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --config C5 --steps 5 --warmup 3 > gpurun_out/r2k_c5_n8.json 2> gpurun_out/r2k_c5_n8.err; echo "C5 N=8 rc $?"; tail -2 gpurun_out/r2k_c5_n8.err
python - <<'PY'
import json
for f in ("r2k_c4_n4","r2k_c4_n8","r2k_c5_n8"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k:(round(v,3) if not isinstance(v,dict) else {a:round(b,3) for a,b in v.items()}) for k,v in d["kernel_ms_per_lm_iteration"].items()}, "parity ok", d["parity"].get("ok"), "stable", d["parity"].get("stable_window",{}).get("ok"), d["parity"].get("stable_window",{}).get("max_rel_cost_diff"), "headline", d["parity"].get("headline_window",{}).get("ok"))
    except Exception as e: print(f, "ERR", e)
PY
