#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_vattn.py tests/test_gpu_tdnet.py -x -q 2>&1 | tail -6
echo "--- recompute path (NSDP_B200_SAVE_ACTIVATIONS=0)"
NSDP_B200_SAVE_ACTIVATIONS=0 timeout 900 python -m pytest tests/test_gpu_vattn.py -x -q -k "backward or oh" 2>&1 | tail -3
for sv in 1 0; do
echo "== SAVE_ACTIVATIONS=$sv"
NSDP_B200_SAVE_ACTIVATIONS=$sv python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric'):
        d = json.loads(l); r = d['roofline']; k = r['kernel_ms_per_step']; print(d['ms_per_step'], 'vbwd', r['launch_ms'], 'vfwd', k['vattn_fwd_D200_K7_M50000'], 'tail', k['resnet_tail_bwd'], k['resnet_tail_fwd'])
"; done
