#!/bin/bash
set -u
mkdir -p gpurun_out/ncu
timeout 300 python -m pytest tests/test_gpu_mlp.py -x -q 2>&1 | tail -3
timeout 300 python tools/microbench_c4.py > gpurun_out/microbench_c4.json 2> gpurun_out/microbench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/microbench_c4.json'))
for r in d['sweep']:
    print(r['W'], round(r['ms'], 4), 'eager', round(r['torch_eager_fp32_ms'], 2), round(r['torch_eager_tf32_ms'], 2), 'exec', round(r['frac_tensor_peak_executed'], 3))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/ncu/mlp_launches.csv python tools/run_mlp.py 64 > /dev/null 2>&1
grep fused_mlp gpurun_out/ncu/mlp_launches.csv | tail -2 | cut -c100-330
timeout 400 python bench.py --impl reference --ref-device cuda --steps 3 --warmup 1 > gpurun_out/bench_reference_gpu_eager.json 2> gpurun_out/bench_reference_gpu_eager.err; echo "ref-gpu rc=$?"
cat gpurun_out/bench_reference_gpu_eager.json; tail -3 gpurun_out/bench_reference_gpu_eager.err
