#!/bin/bash
set -u
for np in 7 4 7 4; do
echo "== FWD NPART $np"
NSDP_FWD_NPART_DEC=$np python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric'):
        d = json.loads(l); r = d['roofline']; k = r['kernel_ms_per_step']; print(d['ms_per_step'], 'vfwd', k['vattn_fwd_D200_K7_M50000'])
"; done
