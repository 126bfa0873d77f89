#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vattn.py tests/test_gpu_tdnet.py -x -q 2>&1 | tail -15
for lo in -1 1; do
  echo "== STAGE_LO $lo"; if [ $lo -ge 0 ]; then export NSDP_STAGE_LO=$lo; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric'):
        d = json.loads(l); r = d['roofline']; k = r['kernel_ms_per_step']; print(d['ms_per_step'], 'vbwd', r['launch_ms'], 'vfwd', k['vattn_fwd_D200_K7_M50000'], 'tail', k['resnet_tail_bwd'], k['resnet_tail_fwd'])
"; done
unset NSDP_STAGE_LO
python tools/grad_err.py 2>&1 | grep -v Warn | head -8
