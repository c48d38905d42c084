#!/usr/bin/env python
"""Turn ncu reports / launch lists under gpurun_out/ into the small text summaries kept in profiles/.

    python tools/prof/summarise_ncu.py launches gpurun_out/launches.csv  > profiles/rN_launches.md
    python tools/prof/summarise_ncu.py full gpurun_out/prof_x.ncu-rep    > profiles/rN_x.md
    python tools/prof/summarise_ncu.py issue gpurun_out/prof_x.ncu-rep <workload> [profiles/issue.json profiles/traffic.json [frames]]
        integer-issue roofline of every captured kernel (merged into the two JSON files bench.py reads); `frames` = frames
        the captured launches decoded, so that bench.py can scale the DRAM traffic to the batch it runs
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


def issue(path, workload, issue_json=None, traffic_json=None, frames=None):
    """Issue-slot accounting per kernel name (launches of one name are summed): warp instructions executed
    against the issue slots of the launch (4 sub-partitions x SMs x elapsed cycles), pipe shares, occupancy,
    and the share of issue-active cycles lost to instruction fetch."""
    import json
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    num = lambda r, m: float(r[hdr.index(m)].replace(",", "")) if m in hdr and r[hdr.index(m)] not in ("", "n/a") else 0.0
    unit = lambda m: rows[1][hdr.index(m)] if m in hdr else ""
    to_bytes = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("alacgpu::", "").split("<")[0]
        a = agg.setdefault(name, {"launches": 0, "inst": 0.0, "slots": 0.0, "ms": 0.0, "dram": 0.0, "w": []})
        cyc = num(r, "sm__cycles_elapsed.max")
        sms = num(r, "launch__sm_count") or 148.0
        dur = num(r, "gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit("gpu__time_duration.sum"), 1.0)
        a["launches"] += 1
        a["inst"] += num(r, "smsp__inst_executed.sum")
        a["slots"] += 4.0 * sms * cyc
        a["ms"] += dur
        a["dram"] += num(r, "dram__bytes_read.sum") * to_bytes.get(unit("dram__bytes_read.sum"), 1.0) + \
            num(r, "dram__bytes_write.sum") * to_bytes.get(unit("dram__bytes_write.sum"), 1.0)
        a["w"].append((dur, {
            "alu_pipe_pct": num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": num(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "stall_no_instruction": num(r, STALL_PREFIX + "no_instruction_per_issue_active.ratio"),
            "registers": num(r, "launch__registers_per_thread"),
        }))
    res, traf = {}, {}
    for name, a in agg.items():
        tot = sum(d for d, _ in a["w"]) or 1.0
        blk = {k: sum(d * m[k] for d, m in a["w"]) / tot for k in a["w"][0][1]}
        blk.update({"launches_captured": a["launches"], "inst_executed": a["inst"], "issue_slots": a["slots"],
                    "issue_frac": a["inst"] / a["slots"] if a["slots"] else None, "capture_ms": a["ms"],
                    "source": path.split("/")[-1]})
        if frames:
            blk["frames_in_capture"] = int(frames)
        res[name] = blk
        traf[name] = {"bytes": a["dram"], "frames": int(frames)} if frames else a["dram"]
        print(name, json.dumps(blk))
    for fn, val in ((issue_json, res), (traffic_json, traf)):
        if fn:
            try:
                cur = json.load(open(fn))
            except Exception:
                cur = {}
            cur.setdefault(workload, {}).update(val)
            json.dump(cur, open(fn, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "issue":
        issue(*sys.argv[2:])
    else:
        {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2])
