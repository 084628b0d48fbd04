#!/usr/bin/env python
"""On-device bring-up aid for the tcgen05 path: runs truncated models in fp32 (CUDA-core) mode and in the
tensor-core modes and reports where they diverge.  Each case runs in its own subprocess with a timeout so a
hung kernel cannot take the rest of the session down.  Usage: python tools/debug_tc.py [case ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (blocks builder, B, L)
    "gemm_64_256": lambda blk: [blk(256, 1, 1, residual=False, separable=False)],
    "gemm_64_512": lambda blk: [blk(512, 1, 1, residual=False, separable=False)],
    "dw11_256": lambda blk: [blk(256, 1, 11, residual=False)],
    "block0_s2": lambda blk: [blk(256, 1, 33, stride=2, residual=False)],
    "block0_res": lambda blk: [blk(256, 1, 33, stride=2, residual=False), blk(256, 1, 33)],
    "rep2_res512": lambda blk: [blk(256, 1, 33, stride=2, residual=False), blk(512, 2, 51)],
    "dil2": lambda blk: [blk(256, 1, 33, stride=2, residual=False), blk(512, 1, 87, dilation=2, residual=False)],
    "final1024": lambda blk: [blk(256, 1, 33, stride=2, residual=False), blk(1024, 1, 1, residual=False, separable=False)],
}


def run_case(name, mode):
    import numpy as np
    import torch
    import viet_asr_b200 as V
    from oracle import quartznet_oracle as O
    jasper = CASES[name](V.configs._blk)
    nlab = 28
    enc_sd, dec_sd = O.random_state_dicts(jasper, 64, nlab, seed=3)
    md = {"AudioToMelSpectrogramPreprocessor": dict(V.configs.PREPROCESSOR_DEFAULT),
          "JasperEncoder": {"activation": "relu", "conv_mask": True, "jasper": jasper}, "labels": V.configs.EN_LABELS}
    g = torch.Generator().manual_seed(5)
    B, Lw = 2, 30000
    wave = 0.1 * torch.randn(B, Lw, generator=g)
    length = torch.tensor([Lw, 21111])
    wave[1, 21111:] = 0
    ref = O.full_path(enc_sd, dec_sd, jasper, wave, length)["enc"]       # [B, C, T]
    out = {}
    for m in ("fp32", mode):
        eng = V.VietASR(model_definition=md, gemm_mode=m)
        eng.load_state_dicts(enc_sd, dec_sd)
        r = eng.forward_device(wave.cuda(), length.cuda(), want_log_probs=False)
        torch.cuda.synchronize()
        out[m] = r["enc"].cpu().transpose(1, 2)
    res = {"case": name, "mode": mode, "shape": list(ref.shape)}
    for m in out:
        d = (out[m] - ref).abs()
        res[f"{m}_maxabs"] = d.max().item()
        res[f"{m}_rel"] = ((out[m] - ref).norm() / ref.norm()).item()
    if res[f"{mode}_rel"] > 1e-3:
        d = (out[mode] - ref).abs()
        bad = (d > 1e-3 * ref.abs().max()).nonzero()
        res["n_bad"] = int(bad.shape[0])
        res["first_bad"] = bad[:8].tolist()
        res["bad_channels"] = sorted(set(bad[:, 1].tolist()))[:40]
        res["bad_times"] = sorted(set(bad[:, 2].tolist()))[:40]
        b0 = bad[0].tolist()
        res["sample"] = {"got": out[mode][b0[0], b0[1], b0[2]].item(), "want": ref[b0[0], b0[1], b0[2]].item()}
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"dbg_{name}_{mode}.npz"), got=out[mode].numpy(), want=ref.numpy())
    print("DBG " + json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--one":
        run_case(sys.argv[2], sys.argv[3])
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for mode in ("f16x3", "f16x1"):
        for n in names:
            try:
                p = subprocess.run([sys.executable, __file__, "--one", n, mode], capture_output=True, text=True, timeout=45)
                lines = [l for l in p.stdout.splitlines() if l.startswith("DBG ")]
                print(lines[-1] if lines else f"DBG {{\"case\": \"{n}\", \"mode\": \"{mode}\", \"rc\": {p.returncode}, \"err\": {json.dumps(p.stderr[-600:])}}}")
            except subprocess.TimeoutExpired:
                print(f"DBG {{\"case\": \"{n}\", \"mode\": \"{mode}\", \"timeout\": true}}")
            sys.stdout.flush()
