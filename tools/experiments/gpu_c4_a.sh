#!/bin/bash
# C4 (fused MLP) round: parity tests, width sweep, ncu capture of the W = 256 kernel
set -u
mkdir -p gpurun_out/ncu
timeout 300 python -m pytest tests/test_gpu_mlp.py -x -q 2>&1 | tail -5
timeout 300 python tools/microbench_c4.py > gpurun_out/microbench_c4.json 2> gpurun_out/microbench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/microbench_c4.json'))
for r in d['sweep']:
    print(r['W'], round(r['ms'], 4), 'eager', round(r['torch_eager_fp32_ms'], 2), round(r['torch_eager_tf32_ms'], 2), 'exec', round(r['frac_tensor_peak_executed'], 3))
PY
if [ "${NCU:-0}" = "1" ]; then
  bash tools/gpu_ncu1.sh mlp256 fused_mlp_tc_kernel 2 "tools/run_mlp.py 256"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/ncu/mlp_launches.csv python tools/run_mlp.py 256 > /dev/null 2>&1
  tail -5 gpurun_out/ncu/mlp_launches.csv | cut -c1-300
fi
