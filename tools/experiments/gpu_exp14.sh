#!/bin/bash
set -u
for v in "" nopoll ""; do
  if [ -n "$v" ]; then export NSDP_B200_LIB=$PWD/nsdp_b200/lib/libnsdp_b200_$v.so; else unset NSDP_B200_LIB; fi
  echo "== variant '$v'"
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric'):
        d = json.loads(l); r = d['roofline']; k = r['kernel_ms_per_step']; print(d['ms_per_step'], 'vbwd', r['launch_ms'], 'vfwd', k['vattn_fwd_D200_K7_M50000'], 'tail', k['resnet_tail_bwd'], k['resnet_tail_fwd'], 'exec TF', r['executed_mma_tflops'])
"; done
