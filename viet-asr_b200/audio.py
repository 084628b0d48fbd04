"""Audio ingest in front of the hot path: WAV decode, int16 -> float32 and 8 kHz -> 16 kHz (any rate) resampling on the
GPU, zero-pad collation and length-bucketed batching.

Reference behaviour mirrored here:
  * `librosa.load(path, sr=16000)` in the CLI and the web app (infer.py:200, app.py:66,82): float32 mono, resampled
    with resampy's "kaiser_best" windowed sinc -> `Resampler` (csrc/audio.cu `resample_kernel`; parity UNPINNED, see
    oracle/resample_oracle.py);
  * clips longer than 10 s are skipped by the CLI (infer.py:201-203) -> `max_duration`;
  * `AudioDataLayer` (infer.py:16-54) feeds ONE utterance per call -> `AudioBatchLayer` feeds a zero-padded batch in the
    `seq_collate_fn` layout (nemo/collections/asr/parts/dataset.py:14-53): `[B, Tmax]` float32 + `[B]` int64.
There is no CPU fallback for the arithmetic: decoding the container format (stdlib `wave`) is host work, sample
conversion and resampling run on the device.
"""
from __future__ import annotations

import ctypes as C
import wave
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .nm import AudioSignal, DataLayerNM, LengthsType, NeuralType

KAISER_BEST = dict(num_zeros=64, precision=9, beta=14.769656459379492, rolloff=0.9475937167399596)


def kaiser_best_window(num_zeros: int = 64, precision: int = 9, beta: float = 14.769656459379492,
                       rolloff: float = 0.9475937167399596) -> Tuple[np.ndarray, int]:
    """Right half of the Kaiser-windowed sinc low-pass of resampy's "kaiser_best" filter:
    num_zeros * 2**precision + 1 samples, 2**precision table entries per zero crossing."""
    num_table = 2 ** precision
    n = num_table * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = np.kaiser(2 * n + 1, beta)[n:]
    return (taper * sinc_win).astype(np.float32), num_table


class Resampler:
    """Device-side `librosa.resample(..., res_type="kaiser_best")` over a zero-padded batch."""

    def __init__(self):
        self._lib = _lib.load()
        win, num_table = kaiser_best_window()
        h = C.c_void_p()
        _lib.check(self._lib.vasr_resampler_create(win.ctypes.data, int(win.shape[0]), int(num_table), C.byref(h)))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.vasr_resampler_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @staticmethod
    def out_len(n_in: int, sr_in: int, sr_out: int) -> int:
        return int(_lib.load().vasr_resample_out_len(int(n_in), int(sr_in), int(sr_out)))

    @torch.no_grad()
    def __call__(self, x: torch.Tensor, length: torch.Tensor, sr_in: int, sr_out: int):
        """x [B, L_in] float32 or int16 (PCM) CUDA tensor, length [B] int64 CUDA -> (y [B, L_out] float32, len_out [B] int64)."""
        if not (x.is_cuda and length.is_cuda):
            raise RuntimeError("Resampler: tensors must be CUDA tensors (there is no CPU path)")
        if x.dtype not in (torch.float32, torch.int16):
            raise ValueError(f"Resampler: float32 or int16 input expected, got {x.dtype}")
        if x.dim() != 2 or length.dim() != 1 or length.shape[0] != x.shape[0]:
            raise ValueError("Resampler: x must be [B, L] and length [B]")
        x = x.contiguous()
        length = length.to(torch.int64).contiguous()
        B, L_in = x.shape
        stream = torch.cuda.current_stream().cuda_stream
        if sr_in == sr_out:
            if x.dtype == torch.float32:
                return x, length
            y = torch.empty((B, L_in), dtype=torch.float32, device=x.device)
            _lib.check(self._lib.vasr_pcm16_to_float(x.data_ptr(), length.data_ptr(), B, L_in, y.data_ptr(), stream))
            return y, length
        L_out = self.out_len(L_in, sr_in, sr_out)
        y = torch.empty((B, L_out), dtype=torch.float32, device=x.device)
        len_out = torch.empty((B,), dtype=torch.int64, device=x.device)
        _lib.check(self._lib.vasr_resample(self._h, x.data_ptr(), 1 if x.dtype == torch.int16 else 0, length.data_ptr(),
                                           B, L_in, int(sr_in), int(sr_out), y.data_ptr(), len_out.data_ptr(), L_out, stream))
        return y, len_out


def read_wav(path: str) -> Tuple[np.ndarray, int]:
    """PCM WAV -> (mono samples, sample rate).  16-bit mono files (all of the reference's audio_samples/) come back
    as int16 so that the sample conversion happens on the device; other layouts are converted to float32 on the host
    (soundfile's scaling, channel mean = librosa.to_mono)."""
    with wave.open(path, "rb") as w:
        sr, nch, width, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        a = np.frombuffer(raw, dtype="<i2")
        if nch == 1:
            return a.copy(), sr
        return (a.reshape(-1, nch).astype(np.float32) / 32768.0).mean(axis=1).astype(np.float32), sr
    if width == 1:
        a = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif width == 4:
        a = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
        a = v.astype(np.float32) / 8388608.0
    else:
        raise ValueError(f"{path}: unsupported sample width {width}")
    if nch > 1:
        a = a.reshape(-1, nch).mean(axis=1)
    return a.astype(np.float32), sr


def collate(signals: Sequence[np.ndarray], pin: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Zero-pad 1-D signals of one dtype (int16 or float32) to the longest: the `seq_collate_fn` audio layout
    (parts/dataset.py:14-53) -> (`[B, Tmax]`, `[B]` int64), in pinned host memory by default."""
    if len(signals) == 0:
        raise ValueError("collate: empty batch")
    dt = np.asarray(signals[0]).dtype
    tdt = torch.int16 if dt == np.int16 else torch.float32
    lens = torch.tensor([int(np.asarray(s).shape[0]) for s in signals], dtype=torch.int64)
    if int(lens.min()) <= 0:
        raise ValueError("collate: empty signal in batch")
    L = int(lens.max())
    w = torch.zeros((len(signals), L), dtype=tdt)
    for i, s in enumerate(signals):
        a = np.asarray(s)
        if a.ndim != 1:
            raise ValueError("collate: signals must be 1-D (mono)")
        if (a.dtype == np.int16) != (tdt == torch.int16):
            raise ValueError("collate: mixed int16 / float signals in one batch")
        w[i, : a.shape[0]] = torch.from_numpy(np.ascontiguousarray(a if tdt == torch.int16 else a.astype(np.float32)))
    if pin and torch.cuda.is_available():
        w, lens = w.pin_memory(), lens.pin_memory()
    return w, lens


def plan_batches(lengths: Sequence[int], max_batch: int = 256, max_padded_samples: int = 256 * 160000) -> List[List[int]]:
    """Length-bucketed batching: indices sorted by length (longest first, stable), packed greedily so that a batch
    holds at most `max_batch` utterances and `B * Tmax <= max_padded_samples` padded samples.  Every index appears
    exactly once; a single utterance longer than the cap gets a batch of its own."""
    if max_batch < 1 or max_padded_samples < 1:
        raise ValueError("plan_batches: caps must be positive")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    batches: List[List[int]] = []
    cur: List[int] = []
    cur_max = 0
    for i in order:
        n = int(lengths[i])
        new_max = max(cur_max, n)
        if cur and (len(cur) + 1 > max_batch or (len(cur) + 1) * new_max > max_padded_samples):
            batches.append(cur)
            cur, cur_max = [], 0
            new_max = n
        cur.append(i)
        cur_max = new_max
    if cur:
        batches.append(cur)
    return batches


class AudioBatchLayer(DataLayerNM):
    """Batched analogue of the reference's in-memory `AudioDataLayer` (infer.py:16-54): `set_signals([...])` then one
    iteration yields `(audio_signal [B, Tmax] float32, a_sig_length [B] int64)`."""

    @property
    def output_ports(self):
        return {
            "audio_signal": NeuralType(("B", "T"), AudioSignal(freq=self._sample_rate)),
            "a_sig_length": NeuralType(tuple("B"), LengthsType()),
        }

    def __init__(self, sample_rate: int):
        super().__init__()
        self._sample_rate = sample_rate
        self.output = False
        self.signal = None
        self.signal_shape = None

    def __iter__(self):
        return self

    def __next__(self):
        if not self.output:
            raise StopIteration
        self.output = False
        return self.signal, self.signal_shape

    def set_signal(self, signal):
        """infer.py:39-43 (single utterance)."""
        self.set_signals([np.reshape(signal, [-1])])

    def set_signals(self, signals: Sequence[np.ndarray]):
        w, lens = collate([np.asarray(s, dtype=np.float32) for s in signals], pin=False)
        self.signal, self.signal_shape = w, lens
        self.output = True

    def __len__(self):
        return 1

    @property
    def dataset(self):
        return None

    @property
    def data_iterator(self):
        return self
