#!/usr/bin/env python
"""Turn gpurun_out/<tag>_launches.csv and <tag>_prof.ncu-rep into small text summaries under profiles/."""
import csv
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1]
out_dir = ROOT / "profiles"
out_dir.mkdir(exist_ok=True)
src = ROOT / "gpurun_out"

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]

launches = src / f"{tag}_launches.csv"
if launches.exists():
    rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(list)
    for r in rows[1:]:
        agg[r[ki].split("(")[0].split("::")[-1][:48]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(out_dir / f"{tag}_launches_summary.txt", "w") as fh:
        fh.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 2 --warmup 3 ({tag}); cold-cache, serialised: compare SHARES\n")
        fh.write(f"{'kernel':50s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            fh.write(f"{k:50s} {len(v):8d} {sum(v)/1e3:12.1f} {sum(v)/len(v)/1e3:10.2f} {sum(v)/tot*100:6.1f}%\n")
    print("wrote", out_dir / f"{tag}_launches_summary.txt")

rep = src / f"{tag}_prof.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    hdr, units = rows[0], rows[1]
    with open(out_dir / f"{tag}_ncu_full_summary.txt", "w") as fh:
        fh.write(f"# ncu --set full --clock-control none --import-source on ({tag}); per launch\n")
        for r in rows[2:]:
            fh.write(r[hdr.index("Kernel Name")][:120] + "\n")
            for k in KEYS:
                if k in hdr:
                    fh.write(f"    {k:95s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}\n")
    print("wrote", out_dir / f"{tag}_ncu_full_summary.txt")
for name in (f"{tag}_bench.json", f"{tag}_bench_reference.json"):
    if (src / name).exists():
        (out_dir / name).write_text((src / name).read_text())
