#!/usr/bin/env python
"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/summarize_profiles.py <round-tag> <launches.csv> <prof.ncu-rep> <workload>"""
import collections
import csv
import json
import os
import subprocess
import sys

tag, launches, rep, workload = sys.argv[1:5]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel share of the step ----------------------------------------------------
rows = list(csv.reader(open(launches)))
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.Counter()
cnt = collections.Counter()
for r in rows:
    if len(r) != len(hdr) or r is hdr or r[ki] == "Kernel Name":
        continue
    name = r[ki].split("(")[0].replace("void ", "")
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[ui], 1.0)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
lines = [f"# ncu launch list summary ({tag}, workload {workload})", "",
         "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare shares, "
         "not absolutes).", "", "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
for name, v in tot.most_common():
    lines.append(f"| `{name}` | {cnt[name]} | {v:.1f} | {100 * v / total:.1f} % |")
open(os.path.join(out_dir, f"{tag}_launches_summary.md"), "w").write("\n".join(lines) + "\n")

# ---- full capture of the dominant kernel -------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]
out = [f"# ncu --set full: dominant kernel ({tag}, workload {workload})", ""]
traffic = []
for r in rr[2:]:
    d = dict(zip(h, r))
    u = dict(zip(h, units))
    out.append("```")
    for k in want:
        if k in d:
            out.append(f"{k} = {d[k]} {u.get(k, '')}")
    out.append("```")
    try:
        conv = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = float(d["dram__bytes_read.sum"]) * conv[u["dram__bytes_read.sum"]]
        wr = float(d["dram__bytes_write.sum"]) * conv[u["dram__bytes_write.sum"]]
        traffic.append(rd + wr)
    except Exception:
        pass
open(os.path.join(out_dir, f"{tag}_integrate_full.md"), "w").write("\n".join(out) + "\n")
if traffic:
    p = os.path.join(out_dir, "integrate_traffic.json")
    cur = json.load(open(p)) if os.path.exists(p) else {}
    cur[workload] = sum(traffic) / len(traffic)
    cur[workload + "_note"] = f"mean dram read+write bytes per k_integrate launch over {len(traffic)} captured launches ({tag})"
    json.dump(cur, open(p, "w"), indent=1)
print(open(os.path.join(out_dir, f"{tag}_launches_summary.md")).read())
