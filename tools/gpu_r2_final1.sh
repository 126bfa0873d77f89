#!/bin/bash
# final 1-GPU pass: whole -m gpu suite, smoke, bench (with cpu baseline + same-GPU reference), reference arm, c3, forward-only,
# then the profiling pass (launch list of one replayed step, full ncu captures, DRAM bytes per launch)
set -u
mkdir -p gpurun_out/ncu
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -8
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'gpu_ref', d.get('gpu_reference', {}).get('ms_per_step'), d.get('gpu_reference', {}).get('ours_over_reference'), 'frac', d['roofline']['frac'])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "reference rc=$?"
timeout 600 python bench.py --workload c3 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c3.json 2>> gpurun_out/bench.err; echo "c3 rc=$?"; tail -c 400 gpurun_out/bench_c3.json | head -c 400; echo
timeout 600 python bench.py --forward-only --steps 20 --warmup 5 > gpurun_out/bench_fwd.json 2>> gpurun_out/bench.err; echo "fwd rc=$?"
timeout 600 python bench.py --impl reference --ref-device cuda --forward-only --steps 10 --warmup 3 > gpurun_out/bench_ref_gpu_fwd.json 2>> gpurun_out/bench.err; echo "ref gpu fwd rc=$?"
timeout 400 python tools/microbench_c4.py > gpurun_out/microbench_c4.json 2> gpurun_out/microbench_c4.err; echo "c4 rc=$?"
WARM=6 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_r2_one_step.csv python tools/one_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/launches_r2_one_step.csv 60 > gpurun_out/launches_r2_summary.txt 2>&1; head -8 gpurun_out/launches_r2_summary.txt
bash tools/gpu_ncu1.sh vattn_bwd_oh vattn_bwd_oh_kernel 1 tools/run_decoder_bwd.py
bash tools/gpu_ncu1.sh dw_tc_vattn dw_tc_kernel 4 tools/run_decoder_bwd.py
bash tools/gpu_ncu1.sh vattn_fwd_oh vattn_fwd_oh_kernel 0 tools/run_decoder_fwd.py
bash tools/gpu_ncu1.sh tail_bwd_tc resnet_tail_bwd_tc_kernel 1 tools/run_decoder_bwd.py
REPS=2 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/ncu/decoder_bwd_dram.csv python tools/run_decoder_bwd.py > gpurun_out/ncu/decoder_bwd_dram.log 2>&1
echo "dram list rc=$?"
