#!/usr/bin/env python
"""profiles/dram_traffic.json (the `roofline.traffic` source of bench.py) from ncu CSV logs of ONE fused pass per config:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file <cfg>.csv \
        -k regex:'k_walk|k_fused|k_connect' -s <launches of the warm-up passes> -c <launches of one pass> \
        python tools/profile_pass.py --config <cfg> --passes 3          (CLV_PASS_SLICES=1)
    python tools/dram_traffic_from_csv.py <commit> c4=gpurun_out/x_c4.csv c2=... > profiles/dram_traffic.json

Per kernel NAME the bytes of all its launches in the pass are summed (the wavefront walk kernels run once per round).
"""
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {"_note": f"dram__bytes_read.sum + dram__bytes_write.sum per PASS and kernel name (all launches of the kernel in one "
                f"fused pass summed), ncu --metrics pass of tools/profile_pass.py with CLV_PASS_SLICES=1, code at commit {sys.argv[1]}"}
for arg in sys.argv[2:]:
    cfg, path = arg.split("=", 1)
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    iK, iM, iU, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    acc, ms = {}, {}
    for r in rows:
        if r is hdr or len(r) <= iV or not r[iV].replace(".", "").replace(",", "").isdigit():
            continue
        name = re.sub(r"<.*", "", r[iK].replace("void ", "")).split("(")[0].strip()
        v = float(r[iV].replace(",", ""))
        if r[iM].startswith("dram__bytes"):
            acc[name] = acc.get(name, 0.0) + v * UNIT.get(r[iU], 1.0)
        elif r[iM].startswith("gpu__time_duration"):
            ms[name] = ms.get(name, 0.0) + v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iU], 1e-6)
    out[cfg] = {k: float(v) for k, v in acc.items()}
    out[cfg]["_ncu_ms"] = {k: round(v, 4) for k, v in ms.items()}
json.dump(out, sys.stdout, indent=1)
print()
