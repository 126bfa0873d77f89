#!/bin/bash
# 2 GPUs: NCCL tests (sharded decode, DP training vs single process), DP bench with the graph split; then the C4 / C5 microbenchmarks
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_graph.py -m gpu -q > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E " gpurun_out/pytest_h.log | tail -12
P=29711
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config']['step_execution'][:60])
PY
grep -iE "error|Traceback" gpurun_out/bench_n2.err | tail -5
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'])
PY
CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/microbench_c5.py > gpurun_out/microbench_c5.json 2> gpurun_out/microbench_c5.err; echo "c5 rc=$?"; cat gpurun_out/microbench_c5.json | head -c 1500
CUDA_VISIBLE_DEVICES=1 timeout 600 python tools/microbench_c4.py > gpurun_out/microbench_c4.json 2> gpurun_out/microbench_c4.err; echo "c4 rc=$?"
