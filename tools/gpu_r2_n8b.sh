#!/bin/bash
# 8-GPU box: 2-rank GPU tests, then bench at N = 8, 4, 2 (the driver's launch line)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_reference_scripts.py -m gpu -q 2>&1 | tail -3
P=29811
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((P+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n$n rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])
PY
grep -iE "error|Traceback" gpurun_out/bench_n$n.err | tail -3
done
