"""Turns the ncu exports of tools/gpu_r2_prof.sh (under gpurun_out/) into the tracked round-2 profile files:
    profiles/ncu_r2_summary.md        metrics table of the full captures, per-kernel DRAM bytes of one decoder backward, hot lines
    profiles/ncu_r2_traffic.json      dram__bytes_read.sum + dram__bytes_write.sum of the dominant op (read by bench.py: roofline.traffic)
    profiles/launches_r2_one_step.csv / launches_r2_summary.txt   the per-launch list of one graph-replayed training step
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def run(*a):
    return subprocess.run([sys.executable, *a], capture_output=True, text=True, cwd=ROOT).stdout


def dram_by_kernel(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i
            break
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0}
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    cnt = collections.Counter()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        agg[name][r[mi]] += v * mult.get(r[ui], 1)
        if r[mi] == "gpu__time_duration.sum":
            cnt[name] += 1
    return agg, cnt


def main():
    names = ["vattn_bwd_oh", "dw_tc_vattn", "vattn_fwd_oh", "tail_bwd_tc"]
    raw = [os.path.join(G, "ncu", n + ".raw.csv") for n in names]
    out = ["# ncu captures, round 2 (B200, `ncu --set full --clock-control none`, bench workload: 8 shapes x 4096 surface points x 50 000 queries)",
           "",
           "Produced by `tools/gpu_r2_prof.sh` (captures) and `tools/make_profiles_r2.py` (this file). Times under ncu are cold-cache,",
           "serialised and at ~1.75 GHz; the bench line's CUDA-event numbers are the ones to quote. Staging format: fp16 (the default).",
           "",
           run("tools/ncu_summary.py", *raw)]
    agg, cnt = dram_by_kernel(os.path.join(G, "ncu", "decoder_bwd_dram.csv"))
    out += ["", "## DRAM bytes of EVERY launch of one decoder backward (cross-attention + tail; `--metrics dram__bytes_read.sum,dram__bytes_write.sum`)", "",
            "| kernel | launches | time (ms, under ncu) | DRAM read (GB) | DRAM write (GB) |", "|---|---|---|---|---|"]
    top = sorted(agg.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0))[:8]
    for n, d in top:
        out.append(f"| `{n[-70:]}` | {cnt[n]} | {d.get('gpu__time_duration.sum', 0):.3f} | {d.get('dram__bytes_read.sum', 0) / 1e9:.3f} | {d.get('dram__bytes_write.sum', 0) / 1e9:.3f} |")
    find = lambda key: next((d for n, d in agg.items() if key in n), {})
    chain, dw, tail = find("vattn_bwd_oh_kernel"), find("dw_tc_kernel"), find("resnet_tail_bwd_tc_kernel")
    b = lambda d: d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    # the reduction launches serve both ops: split its bytes by what each chain kernel staged
    share = chain.get("dram__bytes_write.sum", 0) / max(chain.get("dram__bytes_write.sum", 0) + tail.get("dram__bytes_write.sum", 0), 1)
    vattn_bytes = b(chain) + share * b(dw)
    tail_bytes = b(tail) + (1 - share) * b(dw)
    out += ["", f"Decoder-attention backward op (chain segments + its share of the reduction launches, split by staged bytes): "
                f"**{vattn_bytes / 1e9:.1f} GB** (round 1, bf16 hi + lo staging: 34.8 GB); tail backward: {tail_bytes / 1e9:.1f} GB (round 1: ~12 GB).",
            "Algorithmic compulsory traffic of the attention op is ~20 MB; the rest is the price of keeping the 3 x (2 x 208) + 2 x 208 TMEM",
            "columns of weight / table gradient accumulators out of the chain kernel (DESIGN.md section 4)."]
    with open(os.path.join(P, "ncu_r2_traffic.json"), "w") as f:
        json.dump({"op": "nsdp_vattn_bwd_f32 (decoder cross-attention, D=200, 7+1 rows/query, 8 x 50 000 queries)",
                   "traffic_bytes": vattn_bytes, "chain_read": chain.get("dram__bytes_read.sum", 0),
                   "chain_write": chain.get("dram__bytes_write.sum", 0), "reduction_bytes_share": share * b(dw),
                   "tail_bwd_bytes": tail_bytes, "staging": "fp16",
                   "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every launch of tools/run_decoder_bwd.py "
                             "(tools/gpu_r2_prof.sh -> gpurun_out/ncu/decoder_bwd_dram.csv)"}, f, indent=1)
    for n in names:
        cs = os.path.join(G, "ncu", n + ".cs.csv")
        if os.path.exists(cs):
            lines = run("tools/ncu_lines.py", cs).splitlines()[:16]
            out += ["", f"## {n}: source lines by stall samples", "", "```"] + lines + ["```"]
    with open(os.path.join(P, "ncu_r2_summary.md"), "w") as f:
        f.write("\n".join(out) + "\n")
    for src, dst in (("launches_r2_one_step.csv", "launches_r2_one_step.csv"), ("launches_r2_summary.txt", "launches_r2_summary.txt")):
        if os.path.exists(os.path.join(G, src)):
            shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))
    print(open(os.path.join(P, "ncu_r2_traffic.json")).read())


if __name__ == "__main__":
    main()
