#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 0 1 2 3 4 5; do NSDP_FPS_VARIANT=$v timeout 120 python tools/microbench_fps.py 2>&1 | tail -1; done
timeout 1200 python -m pytest tests/test_gpu_vattn.py tests/test_gpu_mlp.py tests/test_gpu_graph.py tests/test_gpu_tdnet.py -m gpu -q > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E " gpurun_out/pytest_f.log | tail -20
for g in 0 1; do
NSDP_DW_GROUP=$((g==0?1:0)) timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_f$g.json 2> gpurun_out/bench_f.err; echo "bench group-default=$g rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_f$g.json').read().strip().splitlines()[-1])
k = d['roofline']['kernel_ms_per_step']
print(d['ms_per_step'], 'vbwd', k['vattn_bwd_D200_K7_M50000'], 'tailbwd', k['resnet_tail_bwd'])
PY
done
timeout 600 python bench.py --forward-only --steps 20 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('forward-only', d['ms_per_step'], d['value'])"
WARM=6 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_r2_one_step.csv python tools/one_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/launches_r2_one_step.csv 60 > gpurun_out/launches_r2_summary.txt 2>&1; head -12 gpurun_out/launches_r2_summary.txt
