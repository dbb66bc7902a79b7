#!/usr/bin/env python3
"""Turn gpurun_out/*.csv / *.ncu-rep into the small tracked summaries under profiles/ (run in the build container).

    python tools/summarize_ncu.py launches gpurun_out/launches_X.csv profiles/NAME.csv "title"
    python tools/summarize_ncu.py raw gpurun_out/prof_X.ncu-rep profiles/NAME.csv
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = ("gpu__time_duration", "dram__bytes", "gpu__dram_throughput", "sm__pipe_tensor", "sm__warps_active", "launch__registers",
        "launch__occupancy_limit", "sm__throughput", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_tensor", "smsp__issue_active",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit", "lts__t_sector_hit", "smsp__average_warps_issue_stalled", "launch__shared_mem",
        "launch__grid_size", "launch__block_size", "l1tex__t_sectors_pipe_lsu_mem_global_op_st", "l1tex__t_requests_pipe_lsu_mem_global_op_st")


def launches(src, dst, title):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        ns = float(row["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(row["Metric Unit"], 1)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        tot += ns
    with open(dst, "w") as f:
        f.write(f"# {title}\n# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("kernel,launches,total_us,avg_us,share\n")
        for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{c},{ns / 1e3:.1f},{ns / 1e3 / c:.1f},{ns / tot:.4f}\n")
        f.write(f"TOTAL,{sum(c for c, _ in agg.values())},{tot / 1e3:.1f},,1.0\n")


def raw(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    keep = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name") or h.startswith(KEEP)]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in keep])
        w.writerow([units[i] for i in keep])
        for row in r[2:]:
            w.writerow([row[i][:60] for i in keep])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        raw(sys.argv[2], sys.argv[3])
