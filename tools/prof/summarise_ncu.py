#!/usr/bin/env python
"""Turn ncu reports / launch lists under gpurun_out/ into the small text summaries kept in profiles/.

    python tools/prof/summarise_ncu.py launches gpurun_out/launches.csv  > profiles/rN_launches.md
    python tools/prof/summarise_ncu.py full gpurun_out/prof_x.ncu-rep    > profiles/rN_x.md
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def full(path):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## {name}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {m} | {r[i]} | {units[i]} |")
        stalls = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
                  if h.startswith(STALL_PREFIX) and h.endswith("_per_issue_active.ratio") and r[i]]
        if stalls:
            print("\nTop warp stall reasons (warps stalled per issue-active cycle):\n")
            for v, h in sorted(stalls, reverse=True)[:6]:
                print(f"* {h[len(STALL_PREFIX):-len('_per_issue_active.ratio')]}: {v:.2f}")
        rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", "")) if "dram__bytes_read.sum" in hdr else 0
        wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", "")) if "dram__bytes_write.sum" in hdr else 0
        ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        print(f"\nDRAM traffic: read {rd} {ur} + write {wr} {uw}\n")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
    agg = defaultdict(list)
    for r in rows:
        agg[r[4].split("(")[0]].append(float(r[14].replace(",", "")) * (1e-3 if r[13] == "ns" else 1.0 if r[13] in ("us", "usecond") else 1e3))
    total = sum(sum(v) for v in agg.values())
    print("| kernel | launches | avg us | total us | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {k} | {len(v)} | {sum(v) / len(v):.1f} | {sum(v):.1f} | {100 * sum(v) / total:.1f} % |")


if __name__ == "__main__":
    {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2])
