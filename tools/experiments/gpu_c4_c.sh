#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_vattn.py tests/test_gpu_tdnet.py tests/test_io.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/microbench_c4.py > gpurun_out/microbench_c4.json 2> gpurun_out/microbench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/microbench_c4.json'))
for r in d['sweep']:
    print(r['W'], round(r['ms'], 4), 'exec', round(r['frac_tensor_peak_executed'], 3))
PY
bash tools/gpu_quick.sh 2>&1 | tail -2
