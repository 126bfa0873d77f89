#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vattn.py tests/test_gpu_tdnet.py -m gpu -q -x > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR|^E " gpurun_out/pytest_l.log | tail -12
for v in _base "" _base ""; do NSDP_B200_LIB=nsdp_b200/lib/libnsdp_b200$v.so timeout 300 python tools/time_decode.py 2>&1 | tail -1; done
