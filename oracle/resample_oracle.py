"""CPU restatement of `librosa.load(path, sr=16000)`'s resampling step - TEST INFRASTRUCTURE ONLY.

Reference call sites: infer.py:200, app.py:66,82 (`librosa.load(..., sr=16000)`).  librosa (unpinned in
requirements.txt) delegates to resampy's "kaiser_best" filter; neither package is vendored under /root/reference
nor installed in this image, so this file restates the PUBLISHED algorithm (resampy 0.2.x: filters.sinc_window,
core.resample, interpn.resample_f; librosa 0.8 core.audio.resample) and parity with the packages themselves is
UNPINNED.  What the tests pin: the CUDA kernel == this restatement, and this restatement
against scipy's polyphase resampler on band-limited signals (sanity of the low-pass design, loose tolerance).
"""
import numpy as np

# resampy's "kaiser_best" design parameters (filters.py / data/kaiser_best.npz)
KAISER_BEST = dict(num_zeros=64, precision=9, beta=14.769656459379492, rolloff=0.9475937167399596)


def sinc_window(num_zeros=64, precision=9, beta=14.769656459379492, rolloff=0.9475937167399596):
    """resampy.filters.sinc_window with a Kaiser taper: right half of the low-pass, num_zeros * 2**precision + 1
    samples; returns (interp_win float64, num_table)."""
    from scipy.signal.windows import kaiser
    num_bits = 2 ** precision
    n = num_bits * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = kaiser(2 * n + 1, beta)[n:]
    return taper * sinc_win, num_bits


def resampy_resample(x, sr_orig, sr_new):
    """resampy.core.resample(x, sr_orig, sr_new, filter='kaiser_best') for a 1-D signal (interpn.resample_f restated
    with the per-sample loops vectorised over the taps)."""
    x = np.asarray(x)
    ratio = float(sr_new) / sr_orig
    n_out = int(x.shape[0] * ratio)
    interp_win, num_table = sinc_window(**KAISER_BEST)
    if ratio < 1:
        interp_win = interp_win * ratio
    interp_delta = np.zeros_like(interp_win)
    interp_delta[:-1] = np.diff(interp_win)
    scale = min(1.0, ratio)
    time_increment = 1.0 / ratio
    index_step = int(scale * num_table)
    nwin = interp_win.shape[0]
    n_orig = x.shape[0]
    y = np.zeros(n_out, dtype=np.float64)
    # resampy accumulates time_register += time_increment (sequential float64 sum)
    tr = np.concatenate([[0.0], np.cumsum(np.full(max(n_out - 1, 0), time_increment))]) if n_out else np.zeros(0)
    xd = x.astype(np.float64)
    for t in range(n_out):
        time_register = tr[t]
        n = int(time_register)
        frac = scale * (time_register - n)
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        i_max = min(n + 1, (nwin - offset) // index_step)
        idx = offset + np.arange(i_max) * index_step
        w = interp_win[idx] + eta * interp_delta[idx]
        y[t] += np.dot(w, xd[n - np.arange(i_max)])
        frac = scale - frac
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        k_max = min(n_orig - n - 1, (nwin - offset) // index_step)
        idx = offset + np.arange(k_max) * index_step
        w = interp_win[idx] + eta * interp_delta[idx]
        y[t] += np.dot(w, xd[n + 1 + np.arange(k_max)])
    return y


def librosa_resample(y, orig_sr, target_sr):
    """librosa.core.resample(y, orig_sr, target_sr, res_type='kaiser_best', fix=True, scale=False): resampy, then
    fix_length to ceil(n * ratio)."""
    y = np.asarray(y, dtype=np.float32)
    if orig_sr == target_sr:
        return y
    ratio = float(target_sr) / orig_sr
    n_samples = int(np.ceil(y.shape[-1] * ratio))
    y_hat = resampy_resample(y, orig_sr, target_sr)
    if len(y_hat) < n_samples:
        y_hat = np.pad(y_hat, (0, n_samples - len(y_hat)))
    return np.ascontiguousarray(y_hat[:n_samples], dtype=np.float32)


def pcm16_to_float(pcm):
    """soundfile / AudioSegment._convert_samples_to_float32 (parts/segment.py:61-74): int16 -> float32 / 2**15."""
    return np.asarray(pcm, dtype=np.int16).astype(np.float32) * np.float32(1.0 / 32768.0)
