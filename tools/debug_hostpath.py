import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import viet_asr_b200 as V
md = V.configs.quartznet12x1_vi()
from oracle import quartznet_oracle as O
enc_sd, dec_sd = O.random_state_dicts(md["JasperEncoder"]["jasper"], 64, len(md["labels"]), seed=7)
eng = V.VietASR(model_definition=md, gemm_mode="f16x3")
eng.load_state_dicts(enc_sd, dec_sd)
B, L = int(os.environ.get("B", "256")), 80000
g = torch.Generator().manual_seed(1)
wave = (0.1 * torch.randn(B, L, generator=g)).clamp_(-1, 1); length = torch.full((B,), L, dtype=torch.int64)
r = eng.forward_device(wave.cuda(), length.cuda())
torch.cuda.synchronize()
for it in range(3):
    ids_h, len_h = eng.transcribe_host_ids(wave.pin_memory(), length.pin_memory())
    same_rows = (ids_h == r["out_ids"].cpu()).all(dim=1)
    bad = (~same_rows).nonzero().flatten().tolist()
    print(f"iter {it}: {len(bad)} utterances differ; first {bad[:10]} last {bad[-5:]}; len equal {torch.equal(len_h, r['out_len'].cpu())}")

import numpy as np
d = os.environ.get("VASR_HOST_DUMP")
if d:
    feat = np.fromfile(d + "/feat.bin", dtype=np.float32).reshape(r["feat"].shape)
    enc = np.fromfile(d + "/enc.bin", dtype=np.float32).reshape(r["enc"].shape)
    seq = np.fromfile(d + "/seq.bin", dtype=np.int64)
    lens = np.fromfile(d + "/lens.bin", dtype=np.int32).reshape(-1, B)
    fd = np.abs(feat - r["feat"].cpu().numpy()).reshape(B, -1).max(1)
    ed = np.abs(enc - r["enc"].cpu().numpy()).reshape(B, -1).max(1)
    print("feat rows differing:", np.nonzero(fd > 0)[0][:20].tolist(), "count", int((fd > 0).sum()))
    print("enc rows differing:", np.nonzero(ed > 0)[0][:20].tolist(), "count", int((ed > 0).sum()), "max", float(ed.max()))
    print("seq ok:", bool((seq == r["seq"].cpu().numpy()).all()), "lens rows:", lens[:, [0, 63, 64, 127, 128, 255]].tolist())
    bad = np.nonzero(ed > 0)[0]
    if len(bad):
        b = int(bad[0]); e0 = enc[b]; e1 = r["enc"][b].cpu().numpy()
        tt = np.nonzero(np.abs(e0 - e1).max(1) > 0)[0]
        print("utt", b, "bad time rows:", tt[:10].tolist(), "...", tt[-5:].tolist(), "n", len(tt), "of", e0.shape[0])
        cc = np.nonzero(np.abs(e0 - e1).max(0) > 0)[0]
        print("bad channels n", len(cc), cc[:10].tolist())
