#!/bin/bash
# Data-parallel step time under the trainer's / NCCL's knobs (run with gpurun --gpus N): one line per setting.
N=${1:-2}
run() { echo -n "$1: "; env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-moco --no-gpu-reference --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d.get('same_work_no_allreduce') or {}; print(round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'pairs/s | same work, no all-reduce', round(s.get('ms_per_step') or 0,3), '| exposed', round(s.get('exposed_allreduce_ms') or 0,3))
"; }
run "MFVIT_DUMMY=0"
run "MFVIT_OVERLAP_OPT=1"
run "NCCL_MAX_CTAS=8"
run "NCCL_MAX_CTAS=4"
run "MFVIT_ALLREDUCE=fp32"
