#!/usr/bin/env python
"""SASS opcode mix of a profiled kernel (executed warp instructions per opcode):
python profiles/opmix.py rep.ncu-rep [elements_per_launch]"""
import collections
import csv
import io
import subprocess
import sys


def main(path, elems=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    isrc, iinst = hdr.index("Source"), hdr.index("Instructions Executed")
    ops = collections.Counter()
    for r in rows[2:]:
        if len(r) <= iinst or not r[iinst].isdigit():
            continue
        toks = r[isrc].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ops[op.split(".")[0]] += int(r[iinst])
    tot = sum(ops.values())
    print(f"total warp instructions {tot}" + (f" = {tot * 32 / elems:.1f} per element" if elems else ""))
    for k, v in ops.most_common(30):
        print(f"{k:10s} {v:12d} {100 * v / tot:5.1f}%" + (f" {v * 32 / elems:6.2f}/elem" if elems else ""))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
