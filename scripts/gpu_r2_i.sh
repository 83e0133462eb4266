#!/bin/bash
python -c "
import b200ba
from visual_slam_core_b200 import _build
_build.build()" 
for nb in 0 1; do
echo "== BA_NO_BCAST=$nb"
BA_NO_BCAST=$nb timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$nb scripts/dist_prof.py C4 2>&1 | grep "rank "
done
