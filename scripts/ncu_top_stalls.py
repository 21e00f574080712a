#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv` output.
    python scripts/ncu_top_stalls.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# several kernels may be concatenated: split on "Kernel Name" rows
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr): body.append(rows[j])
            j += 1
        H = {h: k for k, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[H["# Samples"]] or 0) for r in body)
        agg = {s: sum(int(r[H[s]] or 0) for r in body) for s in stalls}
        print("==", name[:100], "samples", tot, "instructions", len(body))
        print("   ", ", ".join(f"{k[6:]}={v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 // max(tot, 1) >= 1))
        top = sorted(range(len(body)), key=lambda k: -int(body[k][H["# Samples"]] or 0))[:n]
        for k in sorted(top):
            r = body[k]
            why = sorted(((int(r[H[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
            print(f"  {k:5d} {int(r[H['# Samples']]):6d} {r[H['Instructions Executed']]:>9} {r[H['Source']][:90]:90s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
        i = j
    else:
        i += 1
