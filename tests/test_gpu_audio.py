"""GPU: audio ingest kernels (csrc/audio.cu) through the C ABI against oracle/resample_oracle.py, and the batched
length-aware data path in front of the model (`VietASR.transcribe_signals` / `transcribe_files`).

librosa / resampy are absent from the image: the resampler's parity with the packages is UNPINNED (oracle header);
what is asserted here is CUDA kernel == scalar restatement, and that the bucketed batch pipeline gives exactly the
transcripts of the same batches sent through `transcribe_batch` with each utterance resampled on its own."""
import os
import wave

import numpy as np
import pytest
import torch

from conftest import load_golden, model_and_weights
from oracle import resample_oracle as R

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import viet_asr_b200 as V
    return V


def _signal(n, sr, seed):
    g = np.random.default_rng(seed)
    t = np.arange(n) / sr
    x = 0.4 * np.sin(2 * np.pi * 220.0 * t) + 0.2 * np.sin(2 * np.pi * 1730.0 * t + 0.3) + 0.05 * g.standard_normal(n)
    return np.clip(x, -1, 1).astype(np.float32)


def test_pcm16_to_float_bit_exact():
    V = _cuda()
    g = np.random.default_rng(0)
    pcm = g.integers(-32768, 32768, size=(3, 5001)).astype(np.int16)
    pcm[0, :3] = [-32768, 32767, 0]
    lens = torch.tensor([5001, 17, 4000], dtype=torch.int64)
    rs = V.Resampler()
    y, n = rs(torch.from_numpy(pcm).cuda(), lens.cuda(), 16000, 16000)
    want = R.pcm16_to_float(pcm)
    for b in range(3):
        want[b, int(lens[b]):] = 0
    assert y.dtype == torch.float32 and torch.equal(n.cpu(), lens)
    assert np.array_equal(y.cpu().numpy(), want)
    xf = torch.from_numpy(want).cuda()
    y2, _ = rs(xf, lens.cuda(), 16000, 16000)
    assert y2.data_ptr() == xf.data_ptr()                  # float input at the model rate passes through untouched


@pytest.mark.parametrize("sr_in,sr_out", [(8000, 16000), (44100, 16000), (22050, 16000), (48000, 16000), (16000, 8000),
                                          (11025, 16000)])
def test_resample_matches_oracle(sr_in, sr_out):
    V = _cuda()
    lens = [2999, 1, 1777, 640]
    L = max(lens)
    x = np.zeros((len(lens), L), np.float32)
    for b, n in enumerate(lens):
        x[b, :n] = _signal(n, sr_in, b)
    rs = V.Resampler()
    y, n_out = rs(torch.from_numpy(x).cuda(), torch.tensor(lens).cuda(), sr_in, sr_out)
    assert y.shape == (len(lens), V.Resampler.out_len(L, sr_in, sr_out))
    y = y.cpu().numpy()
    for b, n in enumerate(lens):
        want = R.librosa_resample(x[b, :n], sr_in, sr_out)
        assert int(n_out[b]) == want.shape[0] == int(np.ceil(n * sr_out / sr_in))
        np.testing.assert_allclose(y[b, : want.shape[0]], want, rtol=0, atol=2e-5)
        assert not y[b, want.shape[0]:].any()              # zero beyond the resampled length


def test_resample_pcm16_input_equals_float_input():
    V = _cuda()
    g = np.random.default_rng(5)
    pcm = g.integers(-20000, 20000, size=(2, 4096)).astype(np.int16)
    lens = torch.tensor([4096, 3001]).cuda()
    rs = V.Resampler()
    a, na = rs(torch.from_numpy(pcm).cuda(), lens, 8000, 16000)
    xf = torch.from_numpy(R.pcm16_to_float(pcm)).cuda()
    b, nb = rs(xf, lens, 8000, 16000)
    assert torch.equal(na, nb) and torch.equal(a, b)


def test_resample_argument_errors():
    V = _cuda()
    rs = V.Resampler()
    with pytest.raises(RuntimeError):
        rs(torch.zeros(1, 10), torch.tensor([10]), 8000, 16000)                     # CPU tensors
    with pytest.raises(ValueError):
        rs(torch.zeros(1, 10, dtype=torch.float64).cuda(), torch.tensor([10]).cuda(), 8000, 16000)
    with pytest.raises(ValueError):
        rs(torch.zeros(10).cuda(), torch.tensor([10]).cuda(), 8000, 16000)
    with pytest.raises((RuntimeError, ValueError)):
        rs(torch.zeros(1, 10).cuda(), torch.tensor([10]).cuda(), 16000000, 16000)   # ratio below the table resolution


def _write_wav16(path, pcm, sr):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes(pcm.astype("<i2").tobytes())


def test_transcribe_files_equals_per_batch_route(tmp_path):
    """Mixed 8 kHz / 16 kHz files of different lengths through the bucketed batch pipeline.  Expected transcripts: the
    same length-bucketed batches (a zero-padded batch is NOT equivalent to single utterances in the reference either:
    the STFT reflect-pads the padded row, features.py:181-188), each utterance resampled on its own, sent through
    `transcribe_batch` (the host route).  Resampler-vs-oracle parity is asserted above; here bit-identical inputs
    make the transcripts comparable exactly."""
    V = _cuda()
    md, enc_sd, dec_sd = model_and_weights("vi12x1", "rand")
    eng = V.VietASR(model_definition=md, gemm_mode="fp32", decoder="greedy")
    eng.load_state_dicts(enc_sd, dec_sd)
    g = load_golden("vi12x1_rand")
    rs = V.Resampler()
    paths, sigs, srs = [], [], []
    for i, (row, n) in enumerate(zip(g["pcm16"], g["lens"])):
        pcm16k = row[: int(n)]
        for sr, pcm in ((16000, pcm16k), (8000, pcm16k[::2].copy()), (8000, pcm16k[: int(n) // 2][::2].copy())):
            p = tmp_path / f"a{i}_{sr}_{len(pcm)}.wav"
            _write_wav16(p, pcm, sr)
            paths.append(str(p)); sigs.append(pcm); srs.append(sr)
    want = [None] * len(paths)
    for sr in sorted(set(srs)):
        idx = [i for i in range(len(paths)) if srs[i] == sr]
        for batch in V.plan_batches([len(sigs[i]) for i in idx], 3, int(2560.0 * sr)):
            ids = [idx[j] for j in batch]
            res = []
            for i in ids:
                y, n = rs(torch.from_numpy(sigs[i])[None].cuda(), torch.tensor([len(sigs[i])]).cuda(), sr, 16000)
                res.append(y[0, : int(n[0])].cpu().numpy())
            for i, t in zip(ids, eng.transcribe_batch(res, decoder="greedy")):
                want[i] = t
    got = eng.transcribe_files(paths, batch_size=3)
    assert got == want and all(isinstance(t, str) for t in got)
    # the CLI's 10 s cap (infer.py:201-203): longer clips are skipped (None), input order is kept
    plong = tmp_path / "long.wav"
    _write_wav16(plong, np.zeros(8000 * 11, np.int16), 8000)
    got2 = eng.transcribe_files([paths[0], str(plong)], max_duration=10.0, batch_size=1)
    assert got2[1] is None and isinstance(got2[0], str)


def test_oracle_resampled_input_gives_same_log_probs():
    """8 kHz real speech: log-probs from the device-resampled waveform vs from the oracle-resampled waveform agree to
    the path's logit tolerance (the resampler's fp32 accumulation is far below it)."""
    V = _cuda()
    md, enc_sd, dec_sd = model_and_weights("vi12x1", "rand")
    eng = V.VietASR(model_definition=md, gemm_mode="fp32", decoder="greedy")
    eng.load_state_dicts(enc_sd, dec_sd)
    g = load_golden("vi12x1_rand")
    pcm8 = g["pcm16"][0, : int(g["lens"][0])][::2].copy()
    y, n = V.Resampler()(torch.from_numpy(pcm8)[None].cuda(), torch.tensor([len(pcm8)]).cuda(), 8000, 16000)
    want = torch.from_numpy(R.librosa_resample(R.pcm16_to_float(pcm8), 8000, 16000))[None]
    assert int(n[0]) == want.shape[1]
    a = eng.forward_device(y, n, want_log_probs=True)["log_probs"]
    b = eng.forward_device(want.cuda(), n, want_log_probs=True)["log_probs"]
    assert ((a - b).norm() / b.norm()).item() < 1e-3


def test_wer_over_a_manifest(tmp_path):
    """WER / CER over a NeMo-format manifest next to the throughput.  The manifest holds the reference's
    sample utterances (written out as WAV files) with the REFERENCE's own greedy transcripts as text
    (tests/golden/vi12x1_real_all.npz, oracle/make_golden_r2.py): the B200 path, run one utterance per batch like
    infer.py, must score WER = CER = 0 against them; batched with padding it may differ in the last frame only."""
    import json
    V = _cuda()
    from conftest import have_weights, load_weights
    if not have_weights("vi12x1"):
        pytest.skip("weights/vi12x1 not present")
    md = V.configs.quartznet12x1_vi()
    eng = V.VietASR(model_definition=md, gemm_mode="f16x3", decoder="greedy")
    eng.load_state_dicts(*load_weights("vi12x1"))
    g = load_golden("vi12x1_real_all")
    man = tmp_path / "manifest.json"
    with open(man, "w", encoding="utf-8") as f:
        for i, n in enumerate(g["lens"]):
            p = tmp_path / f"utt{i}.wav"
            _write_wav16(p, g["pcm16"][i, : int(n)], 16000)
            f.write(json.dumps({"audio_filepath": p.name, "duration": int(n) / 16000.0, "text": str(g["texts"][i])}, ensure_ascii=False) + "\n")
    alone = V.evaluate_manifest(eng, str(man), batch_size=1)
    assert alone["utterances"] == 8 and alone["wer"] == 0.0 and alone["cer"] == 0.0, alone
    batched = V.evaluate_manifest(eng, str(man), batch_size=8)
    assert batched["wer"] < 0.03 and batched["audio_s_per_s"] > 0
    with open(man, "a", encoding="utf-8") as f:
        f.write(json.dumps({"audio_filepath": "x.wav", "text": "a"}) + "\n")
    with pytest.raises(ValueError, match="without proper duration key"):
        list(V.read_manifest(str(man)))
