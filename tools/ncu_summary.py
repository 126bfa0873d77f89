"""Condenses `ncu --page raw --csv` dumps (tools/gpu_ncu1.sh) into one markdown table: duration, DRAM traffic, tensor-pipe
and memory utilisation, registers, IPC, top stall lines."""
import csv, sys, os, subprocess
names = sys.argv[1:]
want = [("gpu__time_duration.sum", "duration"), ("sm__cycles_elapsed.avg.per_second", "SM clock"), ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "IPC"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 LSU data pipe %"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem->tensor operand pipe %")]
cols = {}
for n in names:
    rows = list(csv.reader(open(n)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for i, h in enumerate(hdr):
        d[h] = (vals[i], units[i])
    d["__kernel"] = d.get("Kernel Name", ("", ""))[0]
    cols[os.path.basename(n).replace(".raw.csv", "")] = d
print("| metric | " + " | ".join(cols) + " |")
print("|---|" + "---|" * len(cols))
for key, label in want:
    cells = []
    for d in cols.values():
        v, u = d.get(key, ("-", ""))
        try: v = f"{float(v):.4g}"
        except ValueError: pass
        cells.append(f"{v} {u}".strip())
    print(f"| {label} | " + " | ".join(cells) + " |")
