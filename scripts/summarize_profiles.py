#!/usr/bin/env python
"""Turn the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/:
    python scripts/summarize_profiles.py <round tag, e.g. r01>
  gpurun_out/<tag>_launches.csv   (ncu --metrics gpu__time_duration.sum --csv)  -> profiles/<tag>_launches.csv + _summary.md
  gpurun_out/<tag>_*.ncu-rep      (ncu --set full)                              -> profiles/<tag>_ncu_<name>.md
  and profiles/top_kernel_traffic.json (DRAM bytes per launch of the fused attention kernels, read by bench.py)."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("matcha::<unnamed>::", "").replace("unnamed>::", "").replace("matcha::", "")


def launches(tag):
    src = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    if not os.path.exists(src):
        return
    lines = [ln for ln in open(src) if ln.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w").write("".join(lines))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    out = [f"# {tag} - ncu launch list (bench.py, fused + row-chain kernels on)\n",
           "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline [--no-pairs]`",
           "(400 launches after the first 200 = about six training steps; cold-cache, serialised: compare SHARES, not absolutes)\n",
           "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k[:90]}` | {n} | {us:.1f} | {100 * us / tot:.1f}% |")
    open(os.path.join(ROOT, "profiles", f"{tag}_launches_summary.md"), "w").write("\n".join(out) + "\n")


def ncu_raw(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def ncu_reports(tag):
    traffic = {}
    for fn in sorted(os.listdir(os.path.join(ROOT, "gpurun_out"))):
        if not (fn.startswith(tag + "_") and fn.endswith(".ncu-rep")):
            continue
        recs, units = ncu_raw(os.path.join(ROOT, "gpurun_out", fn))
        name = fn[len(tag) + 1:-8]
        out = [f"# {tag} - `ncu --set full --clock-control none` of `{name}` (1x B200), values per launch\n"]
        for d in recs:
            kn = short(d["Kernel Name"])
            out += [f"## `{kn}`\n", "| metric | value | unit |", "|---|---:|---|"]
            for k in KEYS:
                if k in d:
                    out.append(f"| {k} | {d[k]} | {units.get(k, '')} |")
            out.append("")

            def mb(key):
                v, u = float(d[key].replace(",", "")), units.get(key, "")
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
            if "attn_fused_bwd" in kn:
                traffic["fused_attn_bwd"] = {"dram_read_bytes": mb("dram__bytes_read.sum"), "dram_write_bytes": mb("dram__bytes_write.sum"),
                                             "tokens": 81920, "source": f"profiles/{tag}_ncu_{name}.md"}
            if "attn_fused_fwd" in kn:
                traffic["fused_attn_fwd"] = {"dram_read_bytes": mb("dram__bytes_read.sum"), "dram_write_bytes": mb("dram__bytes_write.sum"),
                                             "tokens": 81920, "source": f"profiles/{tag}_ncu_{name}.md"}
        open(os.path.join(ROOT, "profiles", f"{tag}_ncu_{name}.md"), "w").write("\n".join(out) + "\n")
    if traffic:
        path = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")
        if os.path.exists(path):
            old = json.load(open(path))
            old.update(traffic)
            traffic = old
        json.dump(traffic, open(os.path.join(ROOT, "profiles", "top_kernel_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    launches(tag)
    ncu_reports(tag)
