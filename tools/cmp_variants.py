#!/usr/bin/env python
"""Run the 15x5 encoder under several environment settings (one subprocess each, the library reads its switches
once) on the same seeded input and compare every variant's output with the first one.

Usage: python tools/cmp_variants.py B seconds "NAME=ENV1=a;ENV2=b" "NAME2=..."      (first variant = reference)
Prints per variant: encoder ms per pass (back to back), max |diff| and rel-L2 vs the reference variant."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(B, secs, out):
    sys.path.insert(0, ROOT)
    import torch
    import viet_asr_b200 as V
    L = int(16000 * secs)
    md = V.configs.MODELS["quartznet15x5"]()
    V.NeuralModuleFactory(placement=V.DeviceType.GPU)
    eng = V.VietASR(model_definition=md, gemm_mode=os.environ.get("VASR_GEMM_MODE", "f16x3"))
    wdir = os.path.join(ROOT, "weights", "en15x5")
    eng.encoder.restore_from(os.path.join(wdir, "JasperEncoder.pt"))
    eng.decoder.restore_from(os.path.join(wdir, "JasperDecoderForCTC.pt"))
    g = torch.Generator().manual_seed(1234)
    wave = (0.1 * torch.randn(B, L, generator=g)).clamp_(-1.0, 1.0).cuda()
    length = torch.full((B,), L, dtype=torch.int64)
    if os.environ.get("CMP_RAGGED"):
        for i in range(B):
            length[i] = L - (i * 977) % (L // 2)
    length = length.cuda()
    feat, seq = eng.preprocessor.forward_channels_last(wave, length)
    torch.cuda.synchronize()
    N = 4
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
    ev[0].record()
    for i in range(N):
        enc, enc_len = eng.encoder.forward_channels_last(feat, seq)
        ev[i + 1].record()
    torch.cuda.synchronize()
    print("  encoder ms per pass:", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(N)], flush=True)
    torch.save(enc.float().cpu(), out)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(int(sys.argv[2]), float(sys.argv[3]), sys.argv[4])
        sys.exit(0)
    import torch
    B, secs = int(sys.argv[1]), float(sys.argv[2])
    ref = None
    for spec in sys.argv[3:]:
        name, _, envs = spec.partition("=")
        env = dict(os.environ)
        for kv in filter(None, envs.split(";")):
            k, _, v = kv.partition("=")
            env[k] = v
        out = f"/tmp/cmp_{name}.pt"
        if os.path.exists(out):
            os.remove(out)
        print(f"[{name}] {envs}", flush=True)
        try:
            r = subprocess.run([sys.executable, __file__, "--child", str(B), str(secs), out], env=env, timeout=100,
                               stderr=subprocess.STDOUT, stdout=subprocess.PIPE, text=True)
            txt = r.stdout
            rc = r.returncode
        except subprocess.TimeoutExpired as e:
            txt = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            rc = "TIMEOUT"
        for line in txt.splitlines():
            if "encoder ms" in line or "TCSEG" in line or "TCPROF" in line or "rror" in line or "ignored" in line:
                print("  " + line.strip())
        print(f"  rc={rc}", flush=True)
        if not os.path.exists(out):
            print("  no output")
            continue
        t = torch.load(out)
        if ref is None:
            ref = t
            print(f"  reference: shape {tuple(t.shape)} |x|max {t.abs().max().item():.4g}")
        else:
            d = (t - ref)
            print(f"  vs reference: max|diff| {d.abs().max().item():.3e} rel-L2 {(d.norm() / ref.norm()).item():.3e} "
                  f"nan {int(torch.isnan(t).sum())} equal {bool((t == ref).all())}")
