#!/bin/bash
# usage: gpu_ngpu.sh N  — the driver's multi-GPU launch of bench.py on N GPUs of one box
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.out 2> gpurun_out/bench_${N}gpu.err
echo rc=$?
grep '^{"metric' gpurun_out/bench_${N}gpu.out | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
"
