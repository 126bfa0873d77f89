#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vattn.py tests/test_gpu_tdnet.py -m gpu -q -x > gpurun_out/pytest_n.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E " gpurun_out/pytest_n.log | tail -12
for v in 0 1 0 1; do NSDP_FWD_PAIR=$v timeout 300 python tools/time_decode.py 2>&1 | grep forward | sed "s/^/pair=$v /"; done
