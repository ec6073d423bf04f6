#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep, read here on the CPU box with `ncu -i`) into JSON lines:
    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<round>_<kernel>_ncu_full.json
and a launch list (ncu --metrics gpu__time_duration.sum --csv log) into a markdown table:
    python profiles/summarize_ncu.py --launches gpurun_out/launches.csv "<command>" > profiles/<round>_launch_list.md
"""
import collections
import csv
import io
import json
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__cluster_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_membar_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
        "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TSCALE = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = f"{r[i]} {units[i]}".strip()
        try:
            rd, ru = d["dram__bytes_read.sum"].split()
            wr, wu = d["dram__bytes_write.sum"].split()
            t, tu = d["gpu__time_duration.sum"].split()
            traffic = float(rd) * SCALE[ru] + float(wr) * SCALE[wu]
            d["derived.dram_traffic_bytes_per_launch"] = traffic
            d["derived.dram_GBps_under_profiler"] = traffic / (float(t) * TSCALE[tu]) / 1e9
        except Exception:
            pass
        print(json.dumps(d))


def launches(path, cmd):
    txt = [l for l in open(path) if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(txt):
        v = float(r["Metric Value"]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
        k = r["Kernel Name"][:100]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list: `{cmd}`\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    else:
        full(sys.argv[1])
