#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vattn.py tests/test_gpu_tdnet.py -x -q 2>&1 | tail -8
for oh in 1 0; do for np in 4 7; do
  echo "== NO_ONEHOT $oh NPART $np"; NSDP_NO_ONEHOT=$oh NSDP_FWD_NPART_DEC=$np NSDP_BWD_NPART_DEC=4 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric'):
        d = json.loads(l); r = d['roofline']; k = r['kernel_ms_per_step']; print(d['ms_per_step'], 'vbwd', r['launch_ms'], 'vfwd', k['vattn_fwd_D200_K7_M50000'], 'tail', k['resnet_tail_bwd'], k['resnet_tail_fwd'])
"; done; done
