#!/usr/bin/env python
"""Per-layer algorithmic work and roofline floors of the encoder, regenerated from the model definition
(SURVEY.md section 8d formulae: dw = 2 Cin K T_out, pw = 2 Cin Cout T_out, res = 2 C_blockin Cout T,
bytes = 4 (Cin T_in + Cout T_out) per sub-block + 4 C_blockin T per residual).  Floors per batch on one B200 from
MEASURED_PEAKS.json (HBM copy GB/s, sustained bf16 TFLOP/s; f16x3 issues three MMAs per MAC) and the measured FFMA2
peak (125 FMA lanes per clock per SM, tools/ubench/fma_rate).
Usage: python tools/layer_table.py [model] [batch] [seconds] > profiles/r1_layer_table.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "viet-asr_b200"))
import configs  # noqa: E402  (plain module import: no CUDA library needed)

model = sys.argv[1] if len(sys.argv) > 1 else "quartznet15x5"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 5.0
md = configs.MODELS[model]()
peaks = {"hbm_gbs": 6552.0, "bf16_tflops_sustained": 1384.4, "sm_max_mhz": 1965.0}
try:
    peaks.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
except Exception:
    pass
hbm = peaks["hbm_gbs"] * 1e9
tc = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1384.4)) * 1e12
fp32 = 125.0 * 2 * 148 * peaks["sm_max_mhz"] * 1e6

T = 1 + int(secs * 16000) // 160
cin = 64
rows = []
for bi, blk in enumerate(md["JasperEncoder"]["jasper"]):
    k, s, d = blk["kernel"][0], blk["stride"][0], blk["dilation"][0]
    pad = (d * k) // 2 - 1 if d > 1 else k // 2
    cout = blk["filters"]
    bcin, bT, c = cin, T, cin
    for r in range(blk["repeat"]):
        To = (T + 2 * pad - d * (k - 1) - 1) // s + 1
        last = r == blk["repeat"] - 1
        res = blk["residual"] and last
        by = 4 * (c * T + cout * To) + (4 * bcin * bT if res else 0)
        dw = 2 * c * k * To if blk.get("separable", False) else 0
        pw = 2 * c * cout * To + (2 * bcin * cout * To if res else 0)
        rows.append((bi, r, c, cout, k if blk.get("separable", False) else 1, s, d, To, res, by, dw, pw))
        c, T = cout, To
    cin = cout

print(f"# Encoder layers of {model}, {B} x {secs:g} s clips (T_f = {1 + int(secs * 16000) // 160}): algorithmic work and floors on one B200\n")
print(f"Peaks: HBM {hbm / 1e9:.0f} GB/s (measured copy), tensor {tc / 1e12:.0f} TFLOP/s (measured sustained bf16; f16x3 = 3 MMAs per MAC), "
      f"FP32 {fp32 / 1e12:.1f} TFLOP/s (measured FFMA2 rate x {peaks['sm_max_mhz']:.0f} MHz).\n")
print("| block.sub | Cin -> Cout | K (s, d) | T_out | res | MB / utt | dw MFLOP | 1x1 MFLOP | HBM floor us | f16x3 tensor floor us | FP32 (dw) floor us | bound |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
tot = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0]
for (bi, r, c, co, k, s, d, To, res, by, dw, pw) in rows:
    th, tt, tf = B * by / hbm * 1e6, 3 * B * pw / tc * 1e6, B * dw / fp32 * 1e6
    bound = max((th, "HBM"), (tt, "tensor"), (tf, "FP32"))[1]
    print(f"| {bi}.{r} | {c} -> {co} | {k} ({s}, {d}) | {To} | {'yes' if res else ''} | {by / 1e6:.2f} | {dw / 1e6:.1f} | {pw / 1e6:.1f} | "
          f"{th:.1f} | {tt:.1f} | {tf:.1f} | {bound} |")
    for i, v in enumerate((by, dw, pw, th, tt, tf)):
        tot[i] += v
print(f"| **sum** | | | | | {tot[0] / 1e6:.1f} | {tot[1] / 1e6:.0f} | {tot[2] / 1e6:.0f} | {tot[3]:.0f} | {tot[4]:.0f} | {tot[5]:.0f} | "
      f"sum of per-layer maxima {sum(max(B * by / hbm, 3 * B * pw / tc, B * dw / fp32) for (_, _, _, _, _, _, _, _, _, by, dw, pw) in rows) * 1e6:.0f} us |")
