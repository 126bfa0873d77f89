#!/bin/bash
# staging segment length sweep: short segments keep the staged tiles L2-resident (fp16: 294 KB per tile, 148 tiles = 43.5 MB)
set -u
mkdir -p gpurun_out
for seg in 148 296 444 592 1184 4144; do
NSDP_VATTN_SEG=$seg NSDP_TAIL_SEG=$seg timeout 600 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_seg$seg.json 2> gpurun_out/bench_seg.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_seg$seg.json').read().strip().splitlines()[-1])
k = d['roofline']['kernel_ms_per_step']
print('seg', $seg, 'step', round(d['ms_per_step'], 3), 'vbwd', k['vattn_bwd_D200_K7_M50000'], 'tailbwd', k['resnet_tail_bwd'])
PY
done
for seg in 148 296; do
NSDP_VATTN_SEG=$seg NSDP_TAIL_SEG=$seg REPS=2 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/dram_seg$seg.csv python tools/run_decoder_bwd.py > /dev/null 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open('gpurun_out/dram_seg$seg.csv')))
for i, r in enumerate(rows):
    if "Kernel Name" in r: hdr, start = r, i; break
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = collections.Counter()
for r in rows[start+1:]:
    if len(r) <= vi or 'dram' not in r[mi]: continue
    try: v = float(r[vi].replace(",", "")) * mult.get(r[ui], 1)
    except ValueError: continue
    n = r[ki]
    key = 'chain' if 'vattn_bwd_oh' in n else ('dw' if 'dw_tc' in n else ('tail' if 'tail_bwd' in n else 'other'))
    tot[key + ' ' + r[mi].split('_')[3].split('.')[0]] += v / 1e9
print('seg', $seg, dict((k, round(v, 2)) for k, v in sorted(tot.items())))
PY
done
