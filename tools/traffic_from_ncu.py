#!/usr/bin/env python
"""Turn an ncu CSV (`--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`) of
`tools/prof_encoder.py B passes` into profiles/r1_traffic.json: DRAM bytes of the encoder kernels per pass.

Usage: python tools/traffic_from_ncu.py launches.csv passes out.json"""
import csv
import json
import re
import sys

src, passes, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0,
         "usecond": 1e3, "msecond": 1e6}
per = {}
ids = set()
for r in rows:
    name = r[ki]
    if not re.search(r"segment_kernel|subblock_kernel", name):
        continue
    m = re.search(r"(\w+_kernel)<([^>]*)>", name)
    key = f"{m.group(1)}<{m.group(2)}>" if m else name[:40]
    d = per.setdefault(key, {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "time_ns": 0.0})
    v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    if r[mi] == "dram__bytes_read.sum":
        d["dram_read"] += v
    elif r[mi] == "dram__bytes_write.sum":
        d["dram_write"] += v
    elif r[mi] == "gpu__time_duration.sum":
        d["time_ns"] += v
        d["launches"] += 1
tot_r = sum(d["dram_read"] for d in per.values()) / passes
tot_w = sum(d["dram_write"] for d in per.values()) / passes
n = sum(d["launches"] for d in per.values()) / passes
res = {
    "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum on tools/prof_encoder.py "
              f"({passes} encoder passes, B=256 x 5 s, f16x3, 2 sub-batch streams); per-pass sums over the encoder kernels",
    "launches_per_step": n,
    "dram_read_bytes_per_step": tot_r,
    "dram_write_bytes_per_step": tot_w,
    "traffic_bytes_per_step": tot_r + tot_w,
    "sum_kernel_ms_under_ncu_per_step": sum(d["time_ns"] for d in per.values()) / passes / 1e6,
    "per_kernel": {k: {kk: (vv / passes) for kk, vv in d.items()} for k, d in per.items()},
}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
