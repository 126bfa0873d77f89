#!/bin/bash
set -u
mkdir -p gpurun_out
for dbg in 0 1 2 3 4 6; do for np in 4 7; do
  echo "== DBG $dbg NPART $np"; NSDP_DBG=$dbg NSDP_BWD_NPART_DEC=$np python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric'):
        d = json.loads(l); r = d['roofline']; k = r['kernel_ms_per_step']; print(d['ms_per_step'], 'vbwd', r['launch_ms'])
"; done; done
