#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.out 2> gpurun_out/bench_2gpu.err
echo rc=$?
grep '^{"metric' gpurun_out/bench_2gpu.out | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['gpu_launches'])
"
tail -5 gpurun_out/bench_2gpu.err
