#!/bin/bash
set -u
for seg in 592 1184 2368 3256; do
echo "== TAIL SEG $seg"
NSDP_TAIL_SEG=$seg python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric'):
        d = json.loads(l); r = d['roofline']; k = r['kernel_ms_per_step']; print(d['ms_per_step'], 'tail', k['resnet_tail_bwd'])
"; done
