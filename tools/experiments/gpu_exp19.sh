#!/bin/bash
set -u
for seg in 2072 4144 8288 12432; do
  echo "== seg $seg"; NSDP_VATTN_SEG=$seg python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric'):
        d = json.loads(l); r = d['roofline']; k = r['kernel_ms_per_step']; print(d['ms_per_step'], 'vbwd', r['launch_ms'])
"; done
