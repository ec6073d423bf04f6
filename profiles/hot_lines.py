#!/usr/bin/env python
"""Top source lines of a kernel by executed warp instructions, from an ncu report captured with
--import-source on (kernels built with -lineinfo):  python profiles/hot_lines.py rep.ncu-rep [N]"""
import csv
import io
import subprocess
import sys


def main(path, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[2] == "-":      # per-source-line summary row
            try:
                n = int(r[hdr.index("Instructions Executed")])
                smp = int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            lines.append((n, smp, cur_file, r[0], r[1].strip()[:100]))
    tot = sum(l[0] for l in lines) or 1
    tots = sum(l[1] for l in lines) or 1
    print(f"total warp instructions {tot}, samples {tots}")
    for n, smp, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{100 * n / tot:5.1f}% inst {100 * smp / tots:5.1f}% smp  {f}:{ln}: {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
