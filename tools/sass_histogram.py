#!/usr/bin/env python
"""SASS opcode histogram of the encoder kernels in the built library (cuobjdump -sass): the evidence that the product
kernels are tcgen05 / TMEM / TMA code (B200_PROFILING.md "What proves a Blackwell-native kernel").
Usage: python tools/sass_histogram.py [lib.so] > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "viet-asr_b200", "libvasr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCATOMSWS", "SYNCS", "FFMA2", "FFMA", "FMNMX", "LDS", "STS", "F2FP",
       "HADD2", "FADD", "ATOMG", "ATOMS", "LDGSTS", "MEMBAR", "FENCE", "UCGABAR_ARV", "BAR", "STL", "LDL", "HMMA"]
fn = None
hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        hist[fn] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)([.\w]*)", line)
    if m and fn:
        hist[fn][m.group(1)] += 1
        if m.group(1) in ("UTCHMMA", "UTMALDG", "UTCBAR", "UTMASTG") and m.group(2):
            hist[fn][m.group(1) + m.group(2)] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: instructions per kernel (static counts), selected opcodes")
for fn, h in hist.items():
    if not any(k in fn for k in ("segment", "subblock", "stft_mel", "normalize", "beam_kernel", "ctc_collapse", "pw_gemm", "decoder_kernel", "resample")):
        continue
    tot = sum(v for k, v in h.items() if "." not in k)
    cells = [f"{k}={h[k]}" for k in KEY if h.get(k)]
    mods = [f"{k}={v}" for k, v in sorted(h.items()) if "." in k]
    print(f"\n{fn[:110]}\n  total {tot}: " + " ".join(cells) + ("\n  variants: " + " ".join(mods) if mods else ""))
