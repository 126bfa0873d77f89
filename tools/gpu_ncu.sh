#!/bin/bash
# Full ncu captures (one launch each) of the decoder kernels; exported to CSV on the box so that only small files
# travel back (gpurun_out is capped at 64 MiB).
set -u
mkdir -p gpurun_out/ncu
bash tools/gpu_ncu1.sh vattn_bwd_oh vattn_bwd_oh_kernel 1 tools/run_decoder_bwd.py
bash tools/gpu_ncu1.sh dw_tc_vattn dw_tc_kernel 8 tools/run_decoder_bwd.py
bash tools/gpu_ncu1.sh dw_tc_tail dw_tc_kernel 1 tools/run_decoder_bwd.py
bash tools/gpu_ncu1.sh vattn_fwd_oh vattn_fwd_oh_kernel 0 tools/run_decoder_fwd.py
bash tools/gpu_ncu1.sh tail_bwd_tc resnet_tail_bwd_tc_kernel 1 tools/run_decoder_bwd.py
bash tools/gpu_ncu1.sh tail_fwd_tc resnet_tail_tc_kernel 0 tools/run_decoder_fwd.py
du -sh gpurun_out
