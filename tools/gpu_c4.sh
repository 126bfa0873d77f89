#!/bin/bash
# BASELINE configs[3] (fused MLP) on one B200: parity tests through the C ABI, then the width sweep (forward and forward+backward).
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp.py -x -q 2>&1 | tail -3
timeout 400 python tools/microbench_c4.py > gpurun_out/microbench_c4.json 2> gpurun_out/microbench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/microbench_c4.json'))
for r in d['sweep']:
    print(r['W'], 'fwd ms', round(r['ms'], 4), 'executed/peak', round(r['frac_tensor_peak_executed'], 3), 'fwd+bwd ms', round(r['fwd_bwd_ms'], 3))
PY
