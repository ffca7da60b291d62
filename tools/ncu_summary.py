#!/usr/bin/env python3
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/r01_launches_v3.csv  > profiles/r01_launches_v3.txt
  python tools/ncu_summary.py full     gpurun_out/r01_accumulate_v3.ncu-rep > profiles/r01_accumulate_v3.txt

`launches`: per-kernel launch count, total and mean device time and share of the captured window (the
            per-launch times are cold-cache and serialised: compare shares, not absolutes).
`full`    : the metrics DESIGN.md / bench.py quote for the dominant kernel from one `ncu --set full` capture.
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_lsu.sum",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "local_load", "local_store", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
]


def launches(path):
    rows = []
    with open(path) as f:
        text = f.read()
    start = text.find('"ID"')
    rd = csv.DictReader(io.StringIO(text[start:]))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
        unit = r["Metric Unit"]
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((name, v, r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for name, ms, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows)} launches, {total:.3f} ms of device time in the captured window")
    print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'mean ms':>9s} {'share':>7s}  grid block")
    for name, (cnt, ms, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:60]:60s} {cnt:8d} {ms:10.3f} {ms / cnt:9.4f} {100 * ms / total:6.1f}%  {grid} {block}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"# {path}")
        for i, h in enumerate(hdr):
            if h in ("Kernel Name", "Grid Size", "Block Size") or any(h == k or h.startswith(k) for k in KEEP):
                if r[i] not in ("", "n/a"):
                    print(f"{h:90s} {units[i]:14s} {r[i]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
