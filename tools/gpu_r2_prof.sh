#!/bin/bash
# round-2 profiling pass: FPS launch variants, per-launch list of one (graph-replayed) training step, full ncu captures of the
# decoder kernels. Only CSV exports travel back.
set -u
mkdir -p gpurun_out/ncu
for v in 0 1 2 3 4 5; do NSDP_FPS_VARIANT=$v timeout 120 python tools/microbench_fps.py 2>&1 | tail -1; done
WARM=6 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_r2_one_step.csv python tools/one_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/launches_r2_one_step.csv 40 > gpurun_out/launches_r2_summary.txt 2>&1; head -30 gpurun_out/launches_r2_summary.txt
bash tools/gpu_ncu1.sh vattn_bwd_oh vattn_bwd_oh_kernel 1 tools/run_decoder_bwd.py
bash tools/gpu_ncu1.sh dw_tc_vattn dw_tc_kernel 4 tools/run_decoder_bwd.py
bash tools/gpu_ncu1.sh vattn_fwd_oh vattn_fwd_oh_kernel 0 tools/run_decoder_fwd.py
bash tools/gpu_ncu1.sh tail_bwd_tc resnet_tail_bwd_tc_kernel 1 tools/run_decoder_bwd.py
# DRAM bytes of EVERY launch of one decoder-attention backward op + tail backward (chain segments and reductions)
REPS=2 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/ncu/decoder_bwd_dram.csv python tools/run_decoder_bwd.py > gpurun_out/ncu/decoder_bwd_dram.log 2>&1
echo "dram list rc=$?"
du -sh gpurun_out
