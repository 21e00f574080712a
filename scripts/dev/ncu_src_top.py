"""Top stalled SASS instructions of one kernel from an ncu report (source page), grouped in address order:
    python scripts/dev/ncu_src_top.py gpurun_out/x.ncu-rep [n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = [(r[isrc], int(r[isamp] or 0), int(r[iex] or 0)) for r in rows[2:] if len(r) > iex]
tot = sum(d[1] for d in data)
print("total samples", tot, "instructions", len(data))
top = sorted(enumerate(data), key=lambda t: -t[1][1])[:n]
for i, (s, k, e) in sorted(top):
    print(i, s.strip()[:100], k, f"{100 * k / tot:.1f}%", e)
# cumulative by window of 200 instructions
print("-- samples per 250-instruction window")
for w in range(0, len(data), 250):
    k = sum(d[1] for d in data[w:w + 250])
    if k * 50 > tot:
        print(w, f"{100 * k / tot:.1f}%")
