#!/usr/bin/env python
"""Decode the scheduling control fields (stall count, write/read scoreboard, wait mask) of sm_100 SASS from
`cuobjdump -sass` and print a window of a kernel's instructions.  Used to see which scoreboards the shared-memory
loads of the depthwise loop share.  Usage: python tools/sass_ctrl.py lib.so kernel_substr nth_ffma2 [count]"""
import re
import subprocess
import sys

lib, sub, nth = sys.argv[1], sys.argv[2], int(sys.argv[3])
count = int(sys.argv[4]) if len(sys.argv) > 4 else 60
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
ins, on = [], False
i = 0
while i < len(txt):
    ln = txt[i]
    if "Function :" in ln:
        on = sub in ln
    if on:
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", ln)
        if m and i + 1 < len(txt):
            m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", txt[i + 1])
            if m2:
                c = (int(m2.group(1), 16) >> 41) & 0x1FFFFF
                ins.append((m.group(1), m.group(2).strip(), c & 0xF, (c >> 5) & 7, (c >> 8) & 7, (c >> 11) & 0x3F))
                i += 2
                continue
    i += 1
n = 0
for k, x in enumerate(ins):
    if x[1].startswith("FFMA2"):
        n += 1
        if n == nth:
            for y in ins[k:k + count]:
                print(y[0], f"st={y[2]:2d} wr={y[3]} rd={y[4]} wait={y[5]:06b}", y[1][:72])
            break
