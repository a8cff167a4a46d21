#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv` output: top stall locations (SASS) with reasons.
Usage: ncu -i rep --page source --csv --kernel-id ::regex:NAME:1 > src.csv ; python tools/ncu_top_stalls.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[col["# Samples"]].isdigit()]
tot = sum(int(r[col["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for r in data:
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h and r[col[h]].isdigit():
            agg[h] = agg.get(h, 0) + int(r[col[h]])
print("by reason:", ", ".join(f"{k[6:]}={100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:n]:
    s = int(r[col["# Samples"]])
    reasons = {h[6:]: int(r[col[h]]) for h in hdr if h.startswith("stall_") and "Not Issued" not in h and r[col[h]].isdigit() and int(r[col[h]]) > 0}
    best = sorted(reasons.items(), key=lambda kv: -kv[1])[:2]
    print(f"{s:6d} {100 * s / tot:5.1f}% {r[col['Source']].strip()[:72]:72s} x{r[col['Instructions Executed']]:>9s} {best}")
