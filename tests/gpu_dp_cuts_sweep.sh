#!/bin/bash
# Data-parallel step time against the backward's all-reduce cut points (run with gpurun --gpus N): one line per setting.
N=${1:-2}
run() { echo -n "$1: "; env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-moco --no-gpu-reference --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d.get('same_work_no_allreduce') or {}; print(round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'pairs/s | same work, no all-reduce', s.get('ms_per_step'), '| exposed', s.get('exposed_allreduce_ms'))
"; }
run "MFVIT_DP_SEGMENTS=3"
run "MFVIT_DP_CUTS=6,2"
run "MFVIT_DP_CUTS=6,1"
run "MFVIT_DP_CUTS=5,1"
run "MFVIT_DP_CUTS=7,3,1"
