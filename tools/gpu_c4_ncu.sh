#!/bin/bash
set -u
mkdir -p gpurun_out/ncu
bash tools/gpu_ncu1.sh mlp256 fused_mlp_tc_kernel 1 "tools/run_mlp.py 256"
bash tools/gpu_ncu1.sh mlp64 fused_mlp_tc_kernel 1 "tools/run_mlp.py 64"
bash tools/gpu_ncu1.sh mlp16 fused_mlp_tc_kernel 1 "tools/run_mlp.py 16"
