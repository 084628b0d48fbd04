#!/usr/bin/env python
"""Pin the oracle against the reference and write the golden fixtures.

Runs ONLY in the build container (needs /root/reference).  It

 1. imports the reference's own arithmetic file
    ``nemo/collections/asr/parts/jasper.py`` UNMODIFIED by file path (it only
    needs torch), builds the JasperBlock stack exactly as
    ``nemo/collections/asr/jasper.py:153-194`` does and loads the shipped
    checkpoints (all keys must match);
 2. runs it side by side with ``oracle/quartznet_oracle.py`` on real 16 kHz
    WAVs (ragged batch -> exercises masking) and on a seeded random model,
    asserting agreement (max abs diff printed);
 3. checks the mel-basis restatement against torchaudio's slaney filterbank;
 4. writes small fixtures to ``tests/golden/*.npz`` (inputs as int16 PCM,
    reference outputs as float32) and copies the shipped checkpoints into
    ``weights/`` (git-ignored, travels to the GPU box with the snapshot).

The reference ships no tests / golden vectors; these
fixtures - outputs of the reference's own code executed here - are the pin.
"""
import importlib.util
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

torch.set_num_threads(8)


def load_ref_parts():
    spec = importlib.util.spec_from_file_location(
        "ref_parts_jasper", os.path.join(REF, "nemo/collections/asr/parts/jasper.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class RefEncoder(nn.Module):
    """nn.Sequential(JasperBlock...) under attribute ``encoder`` - the exact module
    tree of JasperEncoder (jasper.py:153-194) without the NeuralModule base."""

    def __init__(self, parts, jasper, feat_in):
        super().__init__()
        act = parts.jasper_activations["relu"]()
        layers = []
        for lcfg in jasper:
            layers.append(parts.JasperBlock(
                feat_in, lcfg["filters"], repeat=lcfg["repeat"], kernel_size=lcfg["kernel"],
                stride=lcfg["stride"], dilation=lcfg["dilation"], dropout=lcfg["dropout"],
                residual=lcfg["residual"], groups=lcfg.get("groups", 1),
                separable=lcfg.get("separable", False), heads=lcfg.get("heads", -1),
                residual_mode="add", normalization="batch", norm_groups=-1, activation=act,
                residual_panes=[], conv_mask=True, se=lcfg.get("se", False),
                se_reduction_ratio=lcfg.get("se_reduction_ratio", 16),
                kernel_size_factor=lcfg.get("kernel_size_factor", 1.0)))
            feat_in = lcfg["filters"]
        self.encoder = nn.Sequential(*layers)

    def forward(self, x, length):
        s, length = self.encoder(([x], length))
        return s[-1], length


def read_wav16(path):
    w = wave.open(path)
    assert w.getframerate() == 16000 and w.getnchannels() == 1 and w.getsampwidth() == 2
    pcm = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
    return pcm


def batch_from_pcm(pcms):
    lmax = max(len(p) for p in pcms)
    x = np.zeros((len(pcms), lmax), dtype=np.int16)
    for i, p in enumerate(pcms):
        x[i, : len(p)] = p
    lens = np.array([len(p) for p in pcms], dtype=np.int64)
    return x, lens


def run_case(name, ref_enc, enc_sd, dec_sd, jasper, pcm16, lens, out_path, labels):
    wave_f = torch.from_numpy(pcm16.astype(np.float32) / 32768.0)  # soundfile/librosa float convention
    length = torch.from_numpy(lens)
    with torch.no_grad():
        feats, seq, logmel = O.filterbank_features(wave_f, length, return_pre_norm=True)
        ref_out, ref_len = ref_enc(feats, seq)
        taps = []
        my_out, my_len = O.encoder_forward(enc_sd, jasper, feats, seq, taps)
        d_enc = (ref_out - my_out).abs().max().item()
        assert torch.equal(ref_len.float(), my_len.float()), (ref_len, my_len)
        # decoder = one Conv1d + log_softmax (jasper.py:249,254)
        conv = nn.Conv1d(dec_sd["decoder_layers.0.weight"].shape[1], dec_sd["decoder_layers.0.weight"].shape[0], 1)
        conv.load_state_dict({"weight": dec_sd["decoder_layers.0.weight"], "bias": dec_sd["decoder_layers.0.bias"]})
        ref_logits = conv(ref_out).transpose(1, 2)
        ref_logp = torch.log_softmax(ref_logits, dim=-1)
        ref_ids = ref_logp.argmax(dim=-1)
        my_logp = O.decoder_forward(dec_sd, my_out)
        d_lp = (ref_logp - my_logp).abs().max().item()
        assert torch.equal(ref_ids, O.greedy_argmax(my_logp))
    texts = O.ids_to_text(O.ctc_collapse(ref_ids.numpy(), len(labels)), labels)
    print(f"[{name}] B={len(lens)} T_f={feats.shape[2]} T_e={ref_out.shape[2]} enc_len={ref_len.tolist()} "
          f"|ref-oracle| enc {d_enc:.3e} logp {d_lp:.3e}")
    for t in texts:
        print("    ->", repr(t))
    assert d_enc < 2e-4 and d_lp < 2e-4
    top2 = ref_logp.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1]).min().item()
    np.savez_compressed(
        out_path, pcm16=pcm16, lens=lens, feats=feats.numpy(), seq=seq.numpy(),
        blk0_sub=taps[0][:, ::16, :].contiguous().numpy(),  # every 16th channel of block 0's output
        enc_sub=ref_out[:, ::32, :].contiguous().numpy(),   # every 32nd channel of the encoder output
        enc_len=ref_len.numpy().astype(np.float32),
        logits=ref_logits.contiguous().numpy(), ids=ref_ids.numpy().astype(np.int64),
        min_margin=np.float32(margin), texts=np.array(texts))
    print(f"    wrote {out_path} ({os.path.getsize(out_path)/1024:.0f} KiB), min top-2 margin {margin:.4f}")


def copy_language_models():
    """The shipped KenLM binaries are data, not source: copy them next to the checkpoints (git-ignored, they travel to
    the GPU box with the snapshot) so the LM parity tests can run on real models there."""
    dst = os.path.join(ROOT, "weights", "lm")
    os.makedirs(dst, exist_ok=True)
    for name in ("3-gram-lm.binary", "5-gram-lm.binary"):
        shutil.copyfile(os.path.join(REF, "models/language_model", name), os.path.join(dst, name))
        print("copied", name, os.path.getsize(os.path.join(dst, name)), "bytes")


def write_wer_cases():
    """Accuracy metric: run the reference's own `word_error_rate`
    (nemo/collections/asr/metrics.py:30-63, imported unmodified by file path - it only needs torch) on seeded
    hypothesis / reference pairs built from the shipped corpus and write inputs + outputs to
    tests/golden/wer_cases.json."""
    import json
    import random
    spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(REF, "nemo/collections/asr/metrics.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = random.Random(20260925)
    corpus = os.path.join(REF, "models/language_model/alltext.txt")
    lines = []
    if os.path.exists(corpus):
        with open(corpus, encoding="utf-8") as f:
            for i, ln in enumerate(f):
                if i >= 4000:
                    break
                ln = ln.strip()
                if 3 <= len(ln.split()) <= 24:
                    lines.append(ln)
    if not lines:
        lines = ["xin chào các bạn", "hôm nay trời đẹp quá", "một hai ba bốn năm sáu bảy"]
    rng.shuffle(lines)

    def corrupt(s):
        w = s.split()
        out = []
        for x in w:
            r = rng.random()
            if r < 0.12:
                continue                                   # deletion
            if r < 0.24:
                out.append(rng.choice(lines).split()[0])   # substitution
                continue
            if r < 0.32:
                out.append(x[:-1] if len(x) > 1 else x + "a")   # character edit
                continue
            out.append(x)
            if r > 0.93:
                out.append(rng.choice(lines).split()[-1])  # insertion
        return " ".join(out)

    cases = []
    for n in (1, 2, 5, 16):
        refs = [lines[rng.randrange(len(lines))] for _ in range(n)]
        hyps = [corrupt(r) for r in refs]
        cases.append({"hyp": hyps, "ref": refs})
    cases.append({"hyp": ["", "a b"], "ref": ["a b c", ""]})
    cases.append({"hyp": ["a  b   c"], "ref": ["a b c"]})
    cases.append({"hyp": [""], "ref": [""]})                  # -> inf
    cases.append({"hyp": ["same words here"], "ref": ["same words here"]})
    for c in cases:
        for cer in (False, True):
            v = mod.word_error_rate(c["hyp"], c["ref"], use_cer=cer)
            c["cer" if cer else "wer"] = "inf" if v == float("inf") else v
    path = os.path.join(ROOT, "tests/golden/wer_cases.json")
    with open(path, "w", encoding="utf-8") as f:
        json.dump({"source": "nemo/collections/asr/metrics.py word_error_rate run by oracle/make_golden.py", "cases": cases},
                  f, ensure_ascii=False, indent=1)
    print("wrote", path, len(cases), "cases")


def main():
    if "--wer-only" in sys.argv:
        write_wer_cases()
        return
    copy_language_models()
    if "--lm-only" in sys.argv:
        return
    write_wer_cases()
    parts = load_ref_parts()
    os.makedirs(os.path.join(ROOT, "tests/golden"), exist_ok=True)
    os.makedirs(os.path.join(ROOT, "weights"), exist_ok=True)

    # 3. mel basis vs torchaudio slaney
    import torchaudio
    fb = O.slaney_mel_filterbank(16000, 512, 64, 0.0, 8000.0)
    ta = torchaudio.functional.melscale_fbanks(257, 0.0, 8000.0, 64, 16000, norm="slaney", mel_scale="slaney").T.numpy()
    rel = np.abs(fb - ta).max() / np.abs(ta).max()
    print(f"mel basis vs torchaudio slaney: max rel diff {rel:.2e}; nnz/row {np.count_nonzero(fb,axis=1).min()}..{np.count_nonzero(fb,axis=1).max()}")
    assert rel < 1e-5

    samples = os.path.join(REF, "audio_samples")
    models = {
        "vi12x1": ("quartznet12x1_vi", "configs/quartznet12x1_vi.yaml",
                   "models/acoustic_model/vietnamese/JasperEncoder-STEP-289936.pt",
                   "models/acoustic_model/vietnamese/JasperDecoderForCTC-STEP-289936.pt"),
        "en15x5": ("quartznet15x5", "configs/quartznet15x5.yaml",
                   "models/acoustic_model/english/JasperEncoder-STEP-247400.pt",
                   "models/acoustic_model/english/JasperDecoderForCTC-STEP-247400.pt"),
    }
    wavs_a = ["91.wav", "V1 09 11 12H00 THOI SU 2019_143.wav", "V1 1 11 12H00 THOI SU 2019_5.wav"]
    wavs_b = ["V1 1 11 12H00 THOI SU 2019_5.wav", "V1 09 11 12H00 THOI SU 2019_143.wav"]
    for tag, (cfgname, yml, encp, decp) in models.items():
        cfg = yaml.safe_load(open(os.path.join(REF, yml), encoding="utf-8"))
        jasper = cfg["JasperEncoder"]["jasper"]
        labels = cfg["labels"]
        # the oracle's restated block list must equal the YAML
        mine, nlab = O.quartznet_cfg(cfgname)
        assert nlab == len(labels)
        for a, b in zip(mine, jasper):
            for key in ("filters", "repeat", "kernel", "stride", "dilation", "residual"):
                assert a[key] == b[key], (key, a, b)
            assert a.get("separable", False) == b.get("separable", False)
        assert len(mine) == len(jasper)
        enc_sd = torch.load(os.path.join(REF, encp), map_location="cpu")
        dec_sd = torch.load(os.path.join(REF, decp), map_location="cpu")
        ref_enc = RefEncoder(parts, jasper, 64).eval()
        msg = ref_enc.load_state_dict(enc_sd)
        print(tag, "reference load_state_dict:", msg)
        wl = wavs_a if tag == "vi12x1" else wavs_b
        pcm, lens = batch_from_pcm([read_wav16(os.path.join(samples, w)) for w in wl])
        run_case(f"{tag}_real_batch", ref_enc, enc_sd, dec_sd, jasper, pcm, lens,
                 os.path.join(ROOT, f"tests/golden/{tag}_real_batch.npz"), labels)
        if tag == "vi12x1":
            pcm1, lens1 = batch_from_pcm([read_wav16(os.path.join(samples, "V1 1 11 12H00 THOI SU 2019_5.wav"))])
            run_case(f"{tag}_real_single", ref_enc, enc_sd, dec_sd, jasper, pcm1, lens1,
                     os.path.join(ROOT, f"tests/golden/{tag}_real_single.npz"), labels)
        # copy checkpoints (git-ignored) so the GPU box can run the real-weight parity tests
        dst = os.path.join(ROOT, "weights", tag)
        os.makedirs(dst, exist_ok=True)
        shutil.copyfile(os.path.join(REF, encp), os.path.join(dst, "JasperEncoder.pt"))
        shutil.copyfile(os.path.join(REF, decp), os.path.join(dst, "JasperDecoderForCTC.pt"))
        with open(os.path.join(dst, "labels.yaml"), "w", encoding="utf-8") as f:
            yaml.safe_dump({"labels": labels}, f, allow_unicode=True)

        # seeded random model of the same architecture: pinned to the reference code
        # without needing the checkpoints on the box
        r_enc, r_dec = O.random_state_dicts(jasper, 64, len(labels), seed=20260925)
        ref_r = RefEncoder(parts, jasper, 64).eval()
        print(tag, "random load_state_dict:", ref_r.load_state_dict(r_enc))
        pcm2, lens2 = batch_from_pcm([read_wav16(os.path.join(samples, w))[:24000 + 3333 * i] for i, w in enumerate(wavs_b)])
        run_case(f"{tag}_rand", ref_r, r_enc, r_dec, jasper, pcm2, lens2,
                 os.path.join(ROOT, f"tests/golden/{tag}_rand.npz"), labels)


if __name__ == "__main__":
    main()
