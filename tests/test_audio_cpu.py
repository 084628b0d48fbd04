"""CPU: audio ingest host logic (WAV decode, collation, length-bucketed batching), the resampling oracle's sanity
checks, and the WER/CER metric against the reference's own `word_error_rate` outputs (tests/golden/wer_cases.json,
written by oracle/make_golden.py from nemo/collections/asr/metrics.py)."""
import json
import os
import wave

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import resample_oracle as R


# ----------------------------------------------------------------------------- WER / CER
def _wer_cases():
    return json.load(open(os.path.join(GOLDEN, "wer_cases.json"), encoding="utf-8"))["cases"]


def test_word_error_rate_matches_reference_golden():
    import viet_asr_b200 as V
    cases = _wer_cases()
    assert len(cases) >= 8
    for c in cases:
        for key, cer in (("wer", False), ("cer", True)):
            got = V.word_error_rate(c["hyp"], c["ref"], use_cer=cer)
            want = float("inf") if c[key] == "inf" else c[key]
            assert got == want, (c, key, got, want)


def test_word_error_rate_errors_like_reference():
    import viet_asr_b200 as V
    with pytest.raises(ValueError, match="same number of elements"):
        V.word_error_rate(["a"], ["a", "b"])
    assert V.word_error_rate([], []) == float("inf")


def test_levenshtein_against_scalar_dp():
    """the row-vectorised recurrence equals the textbook DP (metrics.py:7-27 restated as the checker)."""
    from viet_asr_b200.metrics import levenshtein

    def dp(a, b):
        cur = list(range(len(a) + 1))
        for i in range(1, len(b) + 1):
            prev, cur = cur, [i] + [0] * len(a)
            for j in range(1, len(a) + 1):
                cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (a[j - 1] != b[i - 1]))
        return cur[len(a)]
    g = np.random.default_rng(3)
    for _ in range(200):
        a = g.integers(0, 4, size=int(g.integers(0, 14))).tolist()
        b = g.integers(0, 4, size=int(g.integers(0, 14))).tolist()
        assert levenshtein(a, b) == dp(a, b), (a, b)


# ----------------------------------------------------------------------------- host-side batching
def test_plan_batches_properties():
    import viet_asr_b200 as V
    g = np.random.default_rng(0)
    lens = g.integers(1000, 160000, size=777).tolist()
    batches = V.plan_batches(lens, max_batch=64, max_padded_samples=64 * 80000)
    flat = [i for b in batches for i in b]
    assert sorted(flat) == list(range(777))                      # every utterance exactly once
    for b in batches:
        assert 1 <= len(b) <= 64
        mx = max(lens[i] for i in b)
        assert len(b) * mx <= 64 * 80000 or len(b) == 1
    maxes = [max(lens[i] for i in b) for b in batches]
    assert maxes == sorted(maxes, reverse=True)                  # longest first
    # padding waste of the bucketed plan is far below that of arrival-order batching
    waste = sum(len(b) * max(lens[i] for i in b) - sum(lens[i] for i in b) for b in batches) / sum(lens)
    assert waste < 0.08, waste
    assert V.plan_batches([5 * 10 ** 9], 4, 1000) == [[0]]       # over-long utterance gets its own batch
    assert V.plan_batches([], 4, 1000) == []
    with pytest.raises(ValueError):
        V.plan_batches([1], 0, 10)


def test_collate_layout_like_seq_collate_fn():
    import viet_asr_b200 as V
    a = np.arange(5, dtype=np.int16)
    b = np.arange(9, dtype=np.int16) - 4
    w, n = V.collate([a, b], pin=False)
    assert w.dtype == torch.int16 and w.shape == (2, 9) and n.tolist() == [5, 9]
    assert w[0, :5].tolist() == a.tolist() and w[0, 5:].abs().sum() == 0 and w[1].tolist() == b.tolist()
    wf, nf = V.collate([a.astype(np.float32), b.astype(np.float64)], pin=False)
    assert wf.dtype == torch.float32 and nf.dtype == torch.int64
    with pytest.raises(ValueError):
        V.collate([a, b.astype(np.float32)], pin=False)
    with pytest.raises(ValueError):
        V.collate([], pin=False)
    with pytest.raises(ValueError):
        V.collate([np.zeros(0, np.float32)], pin=False)


def _write_wav(path, data, sr, width, nch=1):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(nch); w.setsampwidth(width); w.setframerate(sr)
        w.writeframes(data)


def test_read_wav_formats(tmp_path):
    import viet_asr_b200 as V
    g = np.random.default_rng(1)
    pcm = g.integers(-32768, 32767, size=4000).astype("<i2")
    _write_wav(tmp_path / "m16.wav", pcm.tobytes(), 8000, 2)
    a, sr = V.read_wav(str(tmp_path / "m16.wav"))
    assert sr == 8000 and a.dtype == np.int16 and np.array_equal(a, pcm)      # int16 stays int16 (converted on the device)
    st = g.integers(-32768, 32767, size=(4000, 2)).astype("<i2")
    _write_wav(tmp_path / "s16.wav", st.tobytes(), 44100, 2, nch=2)
    a, sr = V.read_wav(str(tmp_path / "s16.wav"))
    assert sr == 44100 and a.dtype == np.float32
    np.testing.assert_allclose(a, (st.astype(np.float32) / 32768.0).mean(axis=1), rtol=0, atol=1e-7)
    u8 = g.integers(0, 255, size=1000).astype(np.uint8)
    _write_wav(tmp_path / "m8.wav", u8.tobytes(), 16000, 1)
    a, _ = V.read_wav(str(tmp_path / "m8.wav"))
    np.testing.assert_allclose(a, (u8.astype(np.float32) - 128) / 128, atol=1e-7)
    i32 = g.integers(-2 ** 31, 2 ** 31 - 1, size=1000).astype("<i4")
    _write_wav(tmp_path / "m32.wav", i32.tobytes(), 16000, 4)
    a, _ = V.read_wav(str(tmp_path / "m32.wav"))
    np.testing.assert_allclose(a, i32.astype(np.float32) / 2 ** 31, atol=1e-7)
    v = g.integers(-2 ** 23, 2 ** 23 - 1, size=500).astype(np.int32)
    raw = np.stack([v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF], axis=1).astype(np.uint8)
    _write_wav(tmp_path / "m24.wav", raw.tobytes(), 16000, 3)
    a, _ = V.read_wav(str(tmp_path / "m24.wav"))
    np.testing.assert_allclose(a, v.astype(np.float32) / 2 ** 23, atol=1e-7)


def test_audio_batch_layer_ports_and_iteration():
    import viet_asr_b200 as V
    V.NeuralModuleFactory(placement=V.DeviceType.CPU)
    dl = V.AudioBatchLayer(sample_rate=16000)
    assert set(dl.output_ports) == {"audio_signal", "a_sig_length"}
    assert list(iter(dl)) == []
    dl.set_signals([np.ones(7, np.float32), np.ones(3, np.float32)])
    (sig, n), = list(dl)
    assert sig.shape == (2, 7) and n.tolist() == [7, 3] and sig[1, 3:].sum() == 0
    dl.set_signal(np.ones((1, 5), np.float32))          # infer.py:39-43 reshapes to 1-D
    (sig, n), = list(dl)
    assert sig.shape == (1, 5) and n.tolist() == [5]


def test_resampler_fails_loudly_without_cuda():
    import viet_asr_b200 as V
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        rs = V.Resampler()          # creating the device filter table needs a GPU
        rs(torch.zeros(1, 10), torch.tensor([10]), 8000, 16000)


# ----------------------------------------------------------------------------- resampling oracle sanity
def test_host_filter_table_equals_oracle_design():
    from viet_asr_b200.audio import kaiser_best_window
    win, nt = kaiser_best_window()
    ref, nt2 = R.sinc_window(**R.KAISER_BEST)
    assert nt == nt2 == 512 and win.shape == ref.shape == (64 * 512 + 1,)
    np.testing.assert_allclose(win, ref.astype(np.float32), rtol=0, atol=1e-7)
    assert abs(float(ref[0]) - R.KAISER_BEST["rolloff"]) < 1e-12 and abs(float(ref[-1])) < 1e-6


@pytest.mark.parametrize("sr_in,sr_out", [(8000, 16000), (44100, 16000), (22050, 16000), (16000, 8000)])
def test_resample_oracle_is_a_sane_low_pass(sr_in, sr_out):
    """Not a pin of librosa (absent here - parity UNPINNED): the restated windowed-sinc interpolator must agree with
    scipy's polyphase resampler on a band-limited signal away from the edges, and have librosa's output length."""
    from scipy.signal import resample_poly
    from math import gcd
    n = 3000
    t = np.arange(n) / sr_in
    band = 0.4 * min(sr_in, sr_out)
    x = (0.5 * np.sin(2 * np.pi * 0.11 * band * t) + 0.3 * np.sin(2 * np.pi * 0.53 * band * t + 1.0)
         + 0.1 * np.sin(2 * np.pi * 0.83 * band * t + 2.0)).astype(np.float32)
    y = R.librosa_resample(x, sr_in, sr_out)
    assert y.dtype == np.float32 and y.shape[0] == int(np.ceil(n * sr_out / sr_in))
    g = gcd(sr_in, sr_out)
    z = resample_poly(x.astype(np.float64), sr_out // g, sr_in // g)
    m = min(len(y), len(z))
    edge = int(80 * max(1.0, sr_out / sr_in))
    err = np.abs(y[edge:m - edge] - z[edge:m - edge]).max()
    assert err < 5e-3, err


def test_resample_oracle_identity_and_dc():
    x = np.random.default_rng(0).standard_normal(100).astype(np.float32)
    assert R.librosa_resample(x, 16000, 16000) is x or np.array_equal(R.librosa_resample(x, 16000, 16000), x)
    y = R.librosa_resample(np.ones(2000, np.float32), 8000, 16000)
    assert y.shape == (4000,) and np.abs(y[300:-300] - 1.0).max() < 2e-3      # unity DC gain away from the edges
    np.testing.assert_array_equal(R.pcm16_to_float(np.array([-32768, 0, 16384], np.int16)),
                                  np.array([-1.0, 0.0, 0.5], np.float32))


def test_resample_oracle_table_interpolation_vs_direct_windowed_sinc():
    """The restated algorithm interpolates a 512-entries-per-zero-crossing table linearly; evaluating the same
    Kaiser-windowed sinc directly at every tap position must give the same samples to ~1e-6 (8 kHz -> 16 kHz, where
    the table stride is exact), i.e. the table / index arithmetic of the restatement is self-consistent."""
    from scipy.special import i0
    kb = R.KAISER_BEST
    n_in = 600
    x = np.random.default_rng(4).standard_normal(n_in)
    y = R.resampy_resample(x.astype(np.float64), 8000, 16000)

    def h(tau):                                  # continuous low-pass: rolloff * sinc(rolloff * tau) * kaiser(tau / num_zeros)
        tau = np.abs(tau)
        w = np.where(tau <= kb["num_zeros"], i0(kb["beta"] * np.sqrt(np.clip(1 - (tau / kb["num_zeros"]) ** 2, 0, 1))) / i0(kb["beta"]), 0.0)
        return kb["rolloff"] * np.sinc(kb["rolloff"] * tau) * w

    n = np.arange(n_in)
    for t in (0, 1, 7, 200, 201, 555, 1198, 1199):
        tau = t * 0.5 - n                        # distance (input samples) between the output instant and every input
        # resampy's wings stop at the table end: |tau| < num_zeros (left) / <= num_zeros (right); edges are covered by w = 0
        direct = float(np.sum(h(tau) * x))
        assert abs(direct - y[t]) < 2e-6 * max(1.0, abs(direct)), (t, direct, y[t])
