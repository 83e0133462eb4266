for d in 0 1 2; do BA_TILED_DBG=$d python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dbg',$d, d['value'], d['kernel_ms_per_lm_iteration'])"; done
