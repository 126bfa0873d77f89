#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_vattn.py tests/test_gpu_tdnet.py -x -q 2>&1 | tail -3
for ls in 1 0; do
echo "== LOCKSTEP $ls"
NSDP_DW_LOCKSTEP=$ls python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric'):
        d = json.loads(l); r = d['roofline']; k = r['kernel_ms_per_step']; print(d['ms_per_step'], 'vbwd', r['launch_ms'], 'tail', k['resnet_tail_bwd'])
"; done
