#!/usr/bin/env python
"""Summarise an ncu source-page CSV (`ncu -i rep --page source --csv --print-source sass`) of one of the encoder
kernels by CODE REGION: consecutive SASS ranges are classified by the instructions they contain (depthwise FFMA2
loops, TMEM loads = epilogue, UTCHMMA = MMA issuer, UTMALDG = TMA producers ...) and the warp-stall samples of each
range are added up per stall reason.  Usage: python tools/ncu_regions.py file.csv [--top N]"""
import csv
import sys
from collections import Counter, defaultdict

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = rows[2:]


def num(r, name):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return 0.0


# region markers: walk the SASS; a region ends at each BRA / EXIT-free boundary?  Simpler: fixed-size windows of 64
# instructions labelled by their dominant opcode class.
def opclass(src):
    op = src.split()[0] if src.split() else ""
    if op.startswith("@"):
        op = src.split()[1]
    return op.split(".")[0]


W = 48
out = []
for i in range(0, len(data), W):
    chunk = data[i:i + W]
    ops = Counter(opclass(r[col["Source"]]) for r in chunk)
    samples = sum(num(r, "# Samples") for r in chunk)
    inst = sum(num(r, "Instructions Executed") for r in chunk)
    st = {s: sum(num(r, s) for r in chunk) for s in stalls}
    out.append((i, ops, samples, inst, st))
tot = sum(o[2] for o in out)
print(f"total samples {tot:.0f}")
print("first-idx  samples  share   inst-exec   dominant ops                      top stalls")
for i, ops, samples, inst, st in out:
    if samples < 0.004 * tot:
        continue
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    print(f"{i:6d} {samples:9.0f} {100 * samples / tot:6.2f}% {inst:11.0f}   {' '.join(f'{k}:{v}' for k, v in ops.most_common(4)):34s} "
          + " ".join(f"{k[6:]}={100 * v / max(samples, 1):.0f}%" for k, v in top))
