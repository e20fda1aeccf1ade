#!/bin/bash
# Step time of the 32-pair bench under the library's opt-in knobs (one line per setting): python bench.py trimmed to the timed step.
run() { echo -n "$1: "; env $1 python bench.py --no-moco --no-gpu-reference --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'pairs/s')
"; }
if [ $# -gt 0 ]; then for k in "$@"; do run "$k"; done; exit 0; fi
run "MFVIT_DUMMY=0"
run "MFVIT_WGRAD_PAIR=1"
run "MFVIT_ROWS96=0"
run "MFVIT_ROWS96=2"
run "MFVIT_PREZERO=1"
run "MFVIT_GELU_TWIN=0"
