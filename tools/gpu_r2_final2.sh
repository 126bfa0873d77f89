#!/bin/bash
# last pass of the round on one GPU: whole -m gpu suite, smoke, bench (cpu baseline + same-GPU reference), launch list of one step
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -5
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'gpu_ref', d['roofline'].get('gpu_reference', {}).get('ms_per_step'), d['roofline'].get('gpu_reference', {}).get('ours_over_reference'), 'frac', d['roofline']['frac'])
PY
timeout 600 python bench.py --forward-only --steps 20 --warmup 5 > gpurun_out/bench_fwd.json 2>> gpurun_out/bench.err; echo "fwd rc=$?"
timeout 600 python bench.py --workload c3 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c3.json 2>> gpurun_out/bench.err; echo "c3 rc=$?"
WARM=6 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_r2_one_step.csv python tools/one_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/launches_r2_one_step.csv 60 > gpurun_out/launches_r2_summary.txt 2>&1; head -6 gpurun_out/launches_r2_summary.txt; grep -i "adam" gpurun_out/launches_r2_summary.txt
