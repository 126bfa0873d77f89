"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = 0; agg = collections.Counter(); cnt = collections.Counter()
for r in rows[start + 1:]:
    if len(r) <= vi: continue
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    u = r[ui]
    v = v / 1e6 if u == "ns" else (v / 1e3 if u.startswith("us") else v)
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("nsdp::", "")[:70]
    agg[name] += v; cnt[name] += 1; tot += v
print(f"total {tot:.3f} ms, {sum(cnt.values())} launches")
for n, v in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{v:8.3f} ms {100*v/tot:5.1f}% {cnt[n]:4d}  {n}")
