#!/usr/bin/env python
"""CPU numerics study (build container, needs weights/ and tests/golden/): which fp16-split scheme of the 1x1
convolutions keeps the path inside north_star's parity bar (log-probs within 1e-3 rel, greedy ids bit-exact on the
real-speech fixtures)?  Emulates the tensor-core arithmetic in torch: operands rounded to fp16 (hi) plus an fp16
residual (lo), products accumulated in fp32; BatchNorm stays a separate fp32 op here (the CUDA path folds its scale into
the weights before the split, which only moves where the rounding happens).  Schemes:
  x1   hi*hi                      (1 MMA per MAC)
  xw   hi*hi + lo_x*hi_w          (2: activation rounding corrected)
  wx   hi*hi + hi_x*lo_w          (2: weight rounding corrected)
  x3   hi*hi + lo_x*hi_w + hi_x*lo_w   (3: what the shipped kernel does)
Usage: python tools/studies/split_precision_study.py   -> prints a table (recorded in profiles/)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import quartznet_oracle as O  # noqa: E402

torch.set_num_threads(8)
SCHEME = {"name": "fp32"}
_orig = O.masked_conv1d


def _split(t):
    hi = t.half().float()
    lo = (t - hi).half().float()
    return hi, lo


def emu_conv(x, lens, weight, stride=1, padding=0, dilation=1, groups=1):
    if SCHEME["name"] == "fp32" or groups != 1 or weight.shape[2] != 1:
        return _orig(x, lens, weight, stride, padding, dilation, groups)
    lens_i = lens.to(dtype=torch.long)
    max_len = x.size(2)
    mask = torch.arange(max_len).expand(len(lens_i), max_len) >= lens_i.unsqueeze(1)
    x = x.masked_fill(mask.unsqueeze(1), 0)
    # per-layer power-of-two pre-scale like csrc/encoder_tc.cu (largest |w| into [2^13, 2^14))
    mx = weight.abs().max().item()
    sc = 2.0 ** (14 - int(np.floor(np.log2(mx))) - 1) if mx > 0 else 1.0
    w_hi, w_lo = _split(weight * sc)
    x_hi, x_lo = _split(x)
    out = F.conv1d(x_hi, w_hi)
    s = SCHEME["name"]
    if s in ("xw", "x3"):
        out = out + F.conv1d(x_lo, w_hi)
    if s in ("wx", "x3"):
        out = out + F.conv1d(x_hi, w_lo)
    out = out / sc
    new_lens = (lens_i + 2 * padding - dilation * (weight.shape[2] - 1) - 1) / stride + 1
    return out, new_lens


def main():
    from conftest import load_golden, load_weights, pcm_to_wave, MODEL_OF
    import viet_asr_b200  # noqa: F401  (configs)
    from viet_asr_b200 import configs
    O.masked_conv1d = emu_conv
    rows = []
    for tag, kinds in (("vi12x1", ("real_batch", "real_single")), ("en15x5", ("real_batch",))):
        md = configs.MODELS[MODEL_OF[tag]]()
        enc_sd, dec_sd = load_weights(tag)
        for kind in kinds:
            g = load_golden(f"{tag}_{kind}")
            wave, length = pcm_to_wave(g["pcm16"]), torch.from_numpy(g["lens"])
            ref_logp = torch.from_numpy(g["logits"]).log_softmax(-1)
            ref_ids = torch.from_numpy(g["ids"])
            top2 = ref_logp.topk(2, -1).values
            margin = (top2[..., 0] - top2[..., 1]).min().item()
            for s in ("fp32", "x1", "xw", "wx", "x3"):
                SCHEME["name"] = s
                r = O.full_path(enc_sd, dec_sd, md["JasperEncoder"]["jasper"], wave, length)
                rel = ((r["logp"] - ref_logp).norm() / ref_logp.norm()).item()
                mism = int((r["ids"] != ref_ids).sum())
                maxabs = (r["logp"] - ref_logp).abs().max().item()
                rows.append((f"{tag}_{kind}", s, rel, maxabs, mism, ref_ids.numel(), margin))
                print(f"{tag}_{kind:12s} {s:5s} rel-L2 {rel:.2e}  max|dlogp| {maxabs:.2e}  id mismatches {mism}/{ref_ids.numel()}  "
                      f"(min top-2 margin of the reference {margin:.3f})", flush=True)
    return rows


if __name__ == "__main__":
    main()
