#!/bin/bash
set -u
for v in "" poll; do
  if [ -n "$v" ]; then export NSDP_B200_LIB=$PWD/nsdp_b200/lib/libnsdp_b200_$v.so; fi
  echo "== variant '$v'"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric'):
        d = json.loads(l); r = d['roofline']; k = r['kernel_ms_per_step']; print(d['ms_per_step'], 'vbwd', r['launch_ms'], 'vfwd', k['vattn_fwd_D200_K7_M50000'], 'tail', k['resnet_tail_bwd'], k['resnet_tail_fwd'], 'enc bwd', k['vattn_bwd_D120_K10_M4096'], k['vattn_bwd_D256_K100_M100'])
"; done
