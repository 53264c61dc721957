#!/usr/bin/env python
"""Aggregate an ncu report's warp-stall samples by source function / line (needs -lineinfo and --import-source on).

    python tools/ncu_by_function.py gpurun_out/prof.ncu-rep k_connect_finish [top_lines]
"""
import bisect
import csv
import io
import os
import re
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "clive2_b200", "csrc")
files = {n: open(os.path.join(csrc, n)).read().split("\n") for n in ("clv_device.cuh", "clive_b200.cu", "clv_math.h")}


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


hdr, data = None, []
for r in rows:
    if not r:
        continue
    if r[0] == "Line No":
        hdr = r
        iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        continue
    if hdr and len(r) > iT and r[0].isdigit():
        data.append((int(r[0]), r[1], I(r[iS]), I(r[iI]), I(r[iT])))


def which(line, text):
    for n, ls in files.items():
        if line - 1 < len(ls) and ls[line - 1].strip() == text.strip():
            return n
    return "?"


funcs = [(i, m.group(1)) for i, l in enumerate(files["clv_device.cuh"], 1)
         for m in [re.match(r"^CLV_HD\s+[\w:<>\s\*&]+?\b(\w+)\(", l)] if m]
starts = [f[0] for f in funcs]
tot = sum(x[2] for x in data) or 1
agg, lines = {}, []
for line, text, smp, inst, ti in data:
    f = which(line, text)
    name = funcs[bisect.bisect_right(starts, line) - 1][1] if f == "clv_device.cuh" and starts and line >= starts[0] else f
    a = agg.setdefault(name, [0, 0, 0])
    a[0] += smp; a[1] += inst; a[2] += ti
    lines.append((smp, f, line, text, inst, ti))
print(f"{kernel}: {tot} samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"  {k:28s} samples {v[0] / tot:.3f}  warp-inst {v[1]:>12d}  lanes {v[2] / max(v[1], 1):5.1f}")
print()
for smp, f, line, text, inst, ti in sorted(lines, reverse=True)[:top]:
    print(f"  {smp / tot:.3f} {f}:{line} lanes {ti / max(inst, 1):4.1f} | {text.strip()[:105]}")
