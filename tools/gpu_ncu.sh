#!/bin/bash
# Full ncu captures (one launch each) of the decoder kernels; exported to CSV on the box so that only small files
# travel back (gpurun_out is capped at 64 MiB).
set -u
mkdir -p gpurun_out/ncu
cap() {  # name regex skip
  REPS=2 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:"$2" -s "$3" -c 1 -f -o gpurun_out/ncu/$1 python tools/run_decoder_bwd.py > gpurun_out/ncu/$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/ncu/$1.ncu-rep --page raw --csv > gpurun_out/ncu/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$1.ncu-rep --page source --csv > gpurun_out/ncu/$1.source.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$1.ncu-rep --page details --csv > gpurun_out/ncu/$1.details.csv 2>/dev/null
  ls -la gpurun_out/ncu/$1.ncu-rep
  [ $(stat -c %s gpurun_out/ncu/$1.ncu-rep) -gt 12000000 ] && rm -f gpurun_out/ncu/$1.ncu-rep
}
cap vattn_bwd_tc vattn_bwd_tc_kernel 1
cap dw_tc_vattn dw_tc_kernel 4
cap vattn_fwd_tc vattn_fwd_tc_kernel 0
cap tail_bwd_tc resnet_tail_bwd_tc_kernel 0
du -sh gpurun_out
