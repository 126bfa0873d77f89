#!/bin/bash
set -u
for dbg in 0 8 16 32 64 128 248 249; do
  echo "== DBG $dbg"; NSDP_DBG=$dbg NSDP_BWD_NPART_DEC=4 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric'):
        d = json.loads(l); r = d['roofline']; k = r['kernel_ms_per_step']; print(d['ms_per_step'], 'vbwd', r['launch_ms'])
"; done
