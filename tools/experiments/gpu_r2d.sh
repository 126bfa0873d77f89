#!/bin/bash
set -u
mkdir -p gpurun_out
for fmt in bf16x2 fp16; do
echo "== big_grad_check $fmt"
NSDP_STAGE_FMT=$fmt timeout 600 python tools/big_grad_check.py 2>&1 | grep -E "vattn|tail" | head -30
done
timeout 900 python -m pytest tests/test_gpu_vattn.py tests/test_gpu_tdnet.py tests/test_gpu_mlp.py -m gpu -q -s > gpurun_out/pytest_d.log 2>&1; echo "pytest fp16 rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E |worst|stage 2" gpurun_out/pytest_d.log | tail -20
for fmt in fp16 bf16x2; do
NSDP_STAGE_FMT=$fmt timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_d_$fmt.json 2> gpurun_out/bench_d.err; echo "bench $fmt rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_d_$fmt.json').read().strip().splitlines()[-1])
k = d['roofline']['kernel_ms_per_step']
print(d['ms_per_step'], 'vbwd', k['vattn_bwd_D200_K7_M50000'], 'tailbwd', k['resnet_tail_bwd'])
PY
done
