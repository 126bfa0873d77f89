#!/bin/bash
# BASELINE configs[2]: FlowArbitrary training step, global batch 32 on 8 GPUs (4 shapes x 4096 x 50k per GPU), plus the 1-GPU slice
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 8 --workload c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c3_n8.json 2> gpurun_out/bench_c3_n8.err; echo "c3 n8 rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --gpus 1 --workload c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err; echo "c3 n1 rc=$?"
python - <<PY
import json
for n in (8, 1):
    d = json.loads(open(f'gpurun_out/bench_c3_n{n}.json').read().strip().splitlines()[-1])
    print(n, d['ms_per_step'], d['value'], d['config']['step_execution'][:50])
PY
grep -iE "error|Traceback|graph" gpurun_out/bench_c3_n8.err | tail -3
