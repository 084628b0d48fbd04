#!/usr/bin/env python
"""Round-2 golden fixtures (runs ONLY in the build container: needs /root/reference).  Same method as
make_golden.py - the reference's own `nemo/collections/asr/parts/jasper.py`, imported unmodified by file path, with
the shipped checkpoints - for two more cases:

  tests/golden/vi12x1_real_all.npz   all 7 native-16 kHz sample WAVs + test1.wav (8 kHz, brought to 16 kHz with
                                      scipy.signal.resample_poly so that both sides consume the same samples: the
                                      reference's librosa resampler is not installable here), each
                                      transcribed ALONE like infer.py:167-171 does (B = 1): ids, logits, texts
  tests/golden/en15x5_real_b48.npz   BASELINE configs[2] shape: 48 five-second clips cut from the same speech
                                      (recipe stored, not the audio: clip i = wav[src[i]] rolled by off[i] and tiled
                                      to 80 000 samples), run through the reference as batches: greedy ids of every
                                      clip + logits of the first four - the pin for the 128-row CTA-pair kernel,
                                      which small batches never reach.
Also copies the 4-gram KenLM binary next to the other language models (weights/lm, git-ignored).
"""
import os
import shutil
import sys
import wave

import numpy as np
import torch
import torch.nn as nn
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import quartznet_oracle as O  # noqa: E402
from oracle.make_golden import RefEncoder, load_ref_parts  # noqa: E402

torch.set_num_threads(8)
WAVS16 = ["510_cbsk___file_goc_510201920_3.wav", "510_cbsk___file_goc_510201920_7.wav", "6.wav", "91.wav",
          "V1 09 11 12H00 THOI SU 2019_143.wav", "V1 1 11 12H00 THOI SU 2019_5.wav", "V1 31 10 12h00 THOI SU 2019_2_235.wav"]


def read_wav(path):
    w = wave.open(path)
    assert w.getnchannels() == 1 and w.getsampwidth() == 2
    return np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16), w.getframerate()


def ref_decode(ref_enc, dec_sd, pcm16, lens):
    """Reference encoder + the decoder conv / log_softmax / argmax (jasper.py:249-254, greedy_ctc_decoder.py:33-36)."""
    wave_f = torch.from_numpy(pcm16.astype(np.float32) / 32768.0)
    with torch.no_grad():
        feats, seq = O.filterbank_features(wave_f, torch.from_numpy(lens))
        out, _ = ref_enc(feats, seq)
        conv = nn.Conv1d(dec_sd["decoder_layers.0.weight"].shape[1], dec_sd["decoder_layers.0.weight"].shape[0], 1)
        conv.load_state_dict({"weight": dec_sd["decoder_layers.0.weight"], "bias": dec_sd["decoder_layers.0.bias"]})
        logits = conv(out).transpose(1, 2).contiguous()
        ids = torch.log_softmax(logits, dim=-1).argmax(dim=-1)
    return logits.numpy(), ids.numpy()


def clip_from(pcm, off, n):
    x = np.roll(pcm, -off)
    reps = -(-n // len(x))
    return np.tile(x, reps)[:n]


def main():
    parts = load_ref_parts()
    samples = os.path.join(REF, "audio_samples")
    listing = sorted(os.listdir(samples))
    wavs16 = [w for w in listing if w.endswith(".wav") and read_wav(os.path.join(samples, w))[1] == 16000]
    assert sorted(wavs16) == sorted(WAVS16), wavs16
    dst = os.path.join(ROOT, "weights", "lm")
    os.makedirs(dst, exist_ok=True)
    for name in ("4-gram-lm.binary",):
        src = os.path.join(REF, "models/language_model", name)
        if os.path.exists(src):
            shutil.copyfile(src, os.path.join(dst, name))
            print("copied", name, os.path.getsize(src), "bytes")

    # ---------------- vi 12x1: every sample alone
    cfg = yaml.safe_load(open(os.path.join(REF, "configs/quartznet12x1_vi.yaml"), encoding="utf-8"))
    jasper, labels = cfg["JasperEncoder"]["jasper"], cfg["labels"]
    enc_sd = torch.load(os.path.join(REF, "models/acoustic_model/vietnamese/JasperEncoder-STEP-289936.pt"), map_location="cpu")
    dec_sd = torch.load(os.path.join(REF, "models/acoustic_model/vietnamese/JasperDecoderForCTC-STEP-289936.pt"), map_location="cpu")
    ref_enc = RefEncoder(parts, jasper, 64).eval()
    print("vi12x1 load_state_dict:", ref_enc.load_state_dict(enc_sd))
    pcms, names = [], []
    for w in WAVS16:
        pcms.append(read_wav(os.path.join(samples, w))[0]); names.append(w)
    from scipy.signal import resample_poly
    p8, sr8 = read_wav(os.path.join(samples, "test1.wav"))
    assert sr8 == 8000
    up = resample_poly(p8.astype(np.float64), 2, 1)
    pcms.append(np.clip(np.rint(up), -32768, 32767).astype(np.int16)); names.append("test1.wav@16k(resample_poly)")
    lmax = max(len(p) for p in pcms)
    pcm = np.zeros((len(pcms), lmax), np.int16)
    lens = np.array([len(p) for p in pcms], np.int64)
    tmax = max(O.utterance_frames(jasper, lens))
    ids_all = np.full((len(pcms), tmax), -1, np.int16)
    logits_all = np.zeros((len(pcms), tmax, len(labels) + 1), np.float32)
    frames, texts = [], []
    for i, p in enumerate(pcms):
        pcm[i, : len(p)] = p
        lg, ids = ref_decode(ref_enc, dec_sd, p[None], lens[i:i + 1])
        t = ids.shape[1]
        assert t == O.utterance_frames(jasper, [len(p)])[0]
        ids_all[i, :t] = ids[0]; logits_all[i, :t] = lg[0]; frames.append(t)
        texts.append(O.ids_to_text(O.ctc_collapse(ids, len(labels)), labels)[0])
        print(f"  {names[i]:45s} L={len(p):6d} T_e={t:3d} -> {texts[-1]!r}")
    out = os.path.join(ROOT, "tests/golden/vi12x1_real_all.npz")
    np.savez_compressed(out, pcm16=pcm, lens=lens, frames=np.array(frames, np.int32), ids=ids_all, logits=logits_all,
                        texts=np.array(texts), names=np.array(names))
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")

    # ---------------- en 15x5: 48 five-second clips (benchmark shape), recipe only
    cfg = yaml.safe_load(open(os.path.join(REF, "configs/quartznet15x5.yaml"), encoding="utf-8"))
    jasper, labels = cfg["JasperEncoder"]["jasper"], cfg["labels"]
    enc_sd = torch.load(os.path.join(REF, "models/acoustic_model/english/JasperEncoder-STEP-247400.pt"), map_location="cpu")
    dec_sd = torch.load(os.path.join(REF, "models/acoustic_model/english/JasperDecoderForCTC-STEP-247400.pt"), map_location="cpu")
    ref_enc = RefEncoder(parts, jasper, 64).eval()
    print("en15x5 load_state_dict:", ref_enc.load_state_dict(enc_sd))
    NB, L = 48, 80000
    src = np.array([i % 7 for i in range(NB)], np.int32)
    off = np.array([(i // 7) * 9973 + 1234 * (i % 5) for i in range(NB)], np.int32)
    clips = np.stack([clip_from(pcms[src[i]], int(off[i]), L) for i in range(NB)])
    lens48 = np.full((NB,), L, np.int64)
    ids48, logits4, margins = [], None, []
    for s in range(0, NB, 8):
        lg, ids = ref_decode(ref_enc, dec_sd, clips[s:s + 8], lens48[s:s + 8])
        ids48.append(ids)
        t2 = torch.log_softmax(torch.from_numpy(lg), -1).topk(2, -1).values
        margins.append((t2[..., 0] - t2[..., 1]).numpy())          # the reference's own top-2 margin of every frame
        if s == 0:
            logits4 = lg[:4]
        print(f"  clips {s}..{s + 7} done")
    ids48 = np.concatenate(ids48)
    top = np.sort(torch.log_softmax(torch.from_numpy(logits4), -1).numpy(), axis=-1)
    out = os.path.join(ROOT, "tests/golden/en15x5_real_b48.npz")
    np.savez_compressed(out, src=src, off=off, L=np.int64(L), ids=ids48.astype(np.int8), logits4=logits4,
                        margin=np.concatenate(margins).astype(np.float16),
                        min_margin4=np.float32((top[..., -1] - top[..., -2]).min()),
                        texts=np.array(O.ids_to_text(O.ctc_collapse(ids48, len(labels)), labels)))
    print("wrote", out, os.path.getsize(out) // 1024, "KiB; first texts:", O.ids_to_text(O.ctc_collapse(ids48[:3], len(labels)), labels))


if __name__ == "__main__":
    main()
