#!/bin/bash
set -u
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{\"metric'):
        d = json.loads(l); r = d['roofline']; k = r['kernel_ms_per_step']; print(d['ms_per_step'], json.dumps(k))
"
