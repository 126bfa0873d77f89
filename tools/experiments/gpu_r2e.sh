#!/bin/bash
# 2 GPUs: data-parallel bench (graph + bucketed all-reduce), reference arm under torchrun, sharded decode test
set -u
mkdir -p gpurun_out
P=29611
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['config']['step_execution'])
PY
grep -iE "error|graph|Traceback" gpurun_out/bench_n2.err | tail -5
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'])
PY
NSDP_B200_GRAPH=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n2_eager.json 2> gpurun_out/bench_n2_eager.err; echo "bench n2 eager rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n2_eager.json').read().strip().splitlines()[-1])
print('eager', d['n_gpus'], d['ms_per_step'], d['value'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "reference n2 rc=$?"
tail -c 300 gpurun_out/bench_ref_n2.json
