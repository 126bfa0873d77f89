#!/bin/bash
# usage: gpu_ncu1.sh <name> <kernel regex> <skip> <script>
set -u
mkdir -p gpurun_out/ncu
REPS=2 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"$2" -s "$3" -c 1 -f -o gpurun_out/ncu/$1 python $4 > gpurun_out/ncu/$1.log 2>&1
echo "$1 rc=$?"
ncu -i gpurun_out/ncu/$1.ncu-rep --page raw --csv > gpurun_out/ncu/$1.raw.csv 2>/dev/null
ncu -i gpurun_out/ncu/$1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/ncu/$1.cs.csv 2>/dev/null
ncu -i gpurun_out/ncu/$1.ncu-rep --page details --csv > gpurun_out/ncu/$1.details.csv 2>/dev/null
rm -f gpurun_out/ncu/$1.ncu-rep
