#!/bin/bash
set -u
mkdir -p gpurun_out
for t in 3 2 1; do NSDP_DW_TERMS=$t python tools/grad_err.py 2>&1 | grep -v Warning > gpurun_out/grad_err_t$t.txt; cat gpurun_out/grad_err_t$t.txt | head -8; done
for t in 3 1; do for seg in 2048 296; do
  echo "== terms $t seg $seg"; NSDP_DW_TERMS=$t NSDP_VATTN_SEG=$seg python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric'):
        d = json.loads(l); r = d['roofline']; print(d['ms_per_step'], r['launch_ms'], r['kernel_ms_per_step']['resnet_tail_bwd'])
"; done; done
bash tools/gpu_ncu.sh
