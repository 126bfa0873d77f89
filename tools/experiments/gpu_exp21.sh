#!/bin/bash
set -u
for rep in 1 2; do for ps in 1 0; do
echo "== PRESAMPLE $ps"
NSDP_B200_PRESAMPLE=$ps python bench.py --steps 10 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric'):
        d = json.loads(l); print(d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
"; done; done
