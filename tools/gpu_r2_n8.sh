#!/bin/bash
set -u
mkdir -p gpurun_out
P=29811
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((P+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n$n rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])
PY
grep -iE "error|Traceback" gpurun_out/bench_n$n.err | tail -3
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((P+20)) bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err; echo "reference n8 rc=$?"
tail -c 200 gpurun_out/bench_ref_n8.json
