#!/usr/bin/env python
"""Per-kernel table (launches, average duration, DRAM bytes per launch, DRAM GB/s) from an ncu CSV log taken with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum (cold-cache, serialised launches):
    python profiles/launch_table.py gpurun_out/launches_cfg4.csv "<command>" > profiles/<round>_launch_list_<cfg>.md"""
import collections
import csv
import sys

UNIT = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(path, cmd):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ix = {k: i for i, k in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[1:]:
        try:
            val = float(r[ix["Metric Value"]].replace(",", "")) * UNIT.get(r[ix["Metric Unit"]], 1.0)
        except ValueError:
            continue
        k = per.setdefault((r[ix["ID"]], r[ix["Kernel Name"]]), {})
        k[r[ix["Metric Name"]]] = val
    agg = collections.OrderedDict()
    for (_, name), m in per.items():
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values()) or 1.0
    print(f"# ncu launch list: `{cmd}`\n")
    print("`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` "
          "(cold-cache, serialised: compare SHARES, not absolutes)\n")
    print("| kernel | launches | avg us | share | DRAM MB / launch | DRAM GB/s |")
    print("|---|---|---|---|---|---|")
    for name, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        short = name.replace("speckv::<unnamed>::", "").replace("void ", "")[:70]
        print(f"| `{short}` | {n} | {t / n * 1e6:.1f} | {100 * t / tot:.1f}% | {b / n / 1e6:.2f} | {b / t / 1e9 if t else 0:.0f} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
