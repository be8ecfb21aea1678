#!/usr/bin/env python
"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  launches : python tools/summarize_profiles.py launches <tag> <launches.csv> <workload note>
  full     : python tools/summarize_profiles.py full <tag> <prof.ncu-rep> <title> [--figure name:workload:kernel-regex ...]

`full` writes profiles/<tag>.md with the key metrics of every captured launch and, per --figure, stores a derived
number in profiles/<name>.json under <workload> (bench.py reads integrate_traffic / integrate_issue / mc_traffic):
  *_traffic : dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the matching launches)
  *_issue   : smsp__issue_active.avg.pct_of_peak_sustained_active / 100
  *_fma     : sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active / 100 (FP32 pipe busy cycles)
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
CONV = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    name = name.replace("void ", "")
    m = re.match(r"([A-Za-z_0-9]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def launches(tag, path, note):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        if len(r) != len(hdr) or r[ki] == "Kernel Name":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[ui], 1.0)
        tot[short(r[ki])] += v
        cnt[short(r[ki])] += 1
    total = sum(tot.values())
    lines = [f"# ncu launch list summary ({tag})", "", note, "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare shares, "
             "not absolutes).", "", "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for name, v in tot.most_common():
        lines.append(f"| `{name}` | {cnt[name]} | {v:.1f} | {100 * v / total:.2f} % |")
    open(os.path.join(OUT, f"{tag}.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def full(tag, rep, title, figures):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    out = [f"# ncu --set full: {title} ({tag})", "",
           "`ncu --set full --clock-control none --import-source on` under gpurun on one B200; the report itself stays in "
           "gpurun_out/ (scratch), these are its key metrics per captured launch.", ""]
    recs = []
    for r in rr[2:]:
        d, u = dict(zip(h, r)), dict(zip(h, units))
        recs.append((d, u))
        out.append(f"## `{short(d['Kernel Name'])}`")
        out.append("```")
        for k in WANT:
            if k in d:
                out.append(f"{k} = {d[k]} {u.get(k, '')}")
        out.append("```")
    open(os.path.join(OUT, f"{tag}.md"), "w").write("\n".join(out) + "\n")
    for fig in figures:
        name, workload, rx = fig.split(":", 2)
        vals = []
        for d, u in recs:
            if not re.search(rx, d["Kernel Name"]):
                continue
            if name.endswith("traffic"):
                vals.append(float(d["dram__bytes_read.sum"]) * CONV[u["dram__bytes_read.sum"]] +
                            float(d["dram__bytes_write.sum"]) * CONV[u["dram__bytes_write.sum"]])
            elif name.endswith("issue"):
                vals.append(float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]) / 100.0)
            elif name.endswith("fma"):
                vals.append(float(d["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]) / 100.0)
        if not vals:
            print(f"no launch matches {rx} for {name}")
            continue
        p = os.path.join(OUT, name + ".json")
        cur = json.load(open(p)) if os.path.exists(p) else {}
        cur[workload] = sum(vals) if name.startswith("mc_") else sum(vals) / len(vals)
        cur[workload + "_note"] = (f"{'sum' if name.startswith('mc_') else 'mean'} over {len(vals)} captured launch(es) matching "
                                   f"/{rx}/ in profiles/{tag}.md")
        json.dump(cur, open(p, "w"), indent=1)
        print(name, workload, cur[workload])


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        figs = [a.split("=", 1)[1] if a.startswith("--figure=") else a for a in sys.argv[5:] if a != "--figure"]
        full(sys.argv[2], sys.argv[3], sys.argv[4], figs)
