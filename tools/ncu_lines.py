"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump: instructions executed,
stall samples and the dominant stall reasons (top lines first)."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None; hdr = None; rows = []
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0] and r[2] == "-":
        try:
            samples = int(r[6]); inst = int(r[7])
        except ValueError:
            continue
        stalls = {}
        for i in range(32, 49):
            try: v = int(r[i])
            except ValueError: v = 0
            if v: stalls[hdr[i].replace("stall_", "")] = v
        rows.append((samples, inst, cur, r[0], r[1].strip()[:90], stalls))
tot_s = sum(r[0] for r in rows); tot_i = sum(r[1] for r in rows)
print(f"total samples {tot_s}, total warp instructions {tot_i}")
for key, name in ((0, "samples"), (1, "instructions")):
    print(f"--- top by {name}")
    for s, i, f, ln, src, st in sorted(rows, key=lambda x: -x[key])[:top]:
        st3 = ", ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{100*s/tot_s:5.1f}%s {100*i/tot_i:5.1f}%i {f}:{ln:>4} {src}  [{st3}]")
