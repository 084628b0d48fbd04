#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` output: stall-reason totals and the hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = None
tot = {}
data = []
kern = None
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1]; continue
    if r and r[0] == "Address":
        hdr = r; idx = {h: i for i, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    def gi(h):
        try: return int(r[idx[h]] or 0)
        except ValueError: return 0
    n = gi("# Samples")
    st = {h: gi(h) for h in stalls}
    for h, v in st.items():
        tot[h] = tot.get(h, 0) + v
    data.append((n, kern, r[idx["Source"]][:100], {h: v for h, v in st.items() if v}))
S = sum(tot.values()) or 1
print("total stall samples", S)
for h, v in sorted(tot.items(), key=lambda x: -x[1]):
    if v: print(f"  {h:26s} {v:8d} {100*v/S:5.1f}%")
print()
for n, k, src, st in sorted(data, key=lambda x: -x[0])[:topn]:
    print(n, src.strip(), st)
