#!/usr/bin/env python
"""Encoder-only driver for profiling (ncu / VASR_TC_PROF): QuartzNet15x5, B x 5 s synthetic clips, N encoder passes.

Launch order of one pass with 2 sub-batch streams: block 0 (x2), 256-channel segment (x2), 512-channel segment (x2),
K=87 layer (x2), final 1x1 (x2)  ->  `ncu -k regex:segment_kernel -s 6 -c 1` captures the 512-channel segment of the
second pass.  Usage: python tools/prof_encoder.py [B] [passes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import viet_asr_b200 as V  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = 80000
md = V.configs.MODELS["quartznet15x5"]() if "quartznet15x5" in V.configs.MODELS else list(V.configs.MODELS.values())[-1]()
V.NeuralModuleFactory(placement=V.DeviceType.GPU)
eng = V.VietASR(model_definition=md, gemm_mode=os.environ.get("VASR_GEMM_MODE", "f16x3"))
wdir = os.path.join(ROOT, "weights", "en15x5")
if os.path.exists(os.path.join(wdir, "JasperEncoder.pt")):
    eng.encoder.restore_from(os.path.join(wdir, "JasperEncoder.pt"))
    eng.decoder.restore_from(os.path.join(wdir, "JasperDecoderForCTC.pt"))
g = torch.Generator().manual_seed(1234)
wave = (0.1 * torch.randn(B, L, generator=g)).clamp_(-1.0, 1.0).cuda()
length = torch.full((B,), L, dtype=torch.int64).cuda()
feat, seq = eng.preprocessor.forward_channels_last(wave, length)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
ev[0].record()
for i in range(N):
    enc, enc_len = eng.encoder.forward_channels_last(feat, seq)
    ev[i + 1].record()
torch.cuda.synchronize()
print("encoder ms per pass:", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(N)])
