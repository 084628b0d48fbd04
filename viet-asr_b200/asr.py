"""`nemo.collections.asr`-compatible neural modules backed by libvasr_b200.so.

Same class names, constructor kwargs, port names and `forward` kwargs as the
reference (nemo/collections/asr/__init__.py:15-47), so `infer.py`'s wiring
(infer.py:99-160) works unchanged.  All arithmetic runs in hand-written CUDA
through the C ABI (`_lib.py`); tensors are PyTorch-owned device buffers.

Layout note: the library computes in channels-last [B, T, C].  The [B, C, T]
tensors the reference's ports carry are returned as transposed *views* of that
memory (same shape/values, no copy) and recognised again on the way in.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .nm import (AcousticEncodedRepresentation, AudioSignal, LengthsType, LogprobsType,
                 MelSpectrogramType, NeuralType, NonTrainableNM, PredictionsType, SpectrogramType,
                 TrainableNM)


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (vasr_b200 has no CPU path), got device {t.device}")


def _as_channels_last(x: torch.Tensor) -> torch.Tensor:
    """[B, C, T] (any strides) -> contiguous [B, T, C] fp32 storage; free if `x` is our own view."""
    xt = x.transpose(1, 2)
    if xt.dtype == torch.float32 and xt.is_contiguous():
        return xt
    return xt.to(torch.float32).contiguous()


# --------------------------------------------------------------------------- mel basis
def slaney_mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """What `librosa.filters.mel(sr, n_fft, n_mels=, fmin=, fmax=)` returns with librosa's defaults
    (htk=False, norm='slaney', float32) - called at parts/features.py:199-202.  librosa is not a
    dependency of this package; the basis is a constructor-time constant computed on the host."""
    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        lin = f / (200.0 / 3)
        log = 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) / (math.log(6.4) / 27.0)
        return np.where(f >= 1000.0, log, lin)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((math.log(6.4) / 27.0) * (m - 15.0)), (200.0 / 3) * m)

    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


# --------------------------------------------------------------------------- preprocessor
_WINDOWS = {"hann": torch.hann_window, "hamming": torch.hamming_window, "blackman": torch.blackman_window,
            "bartlett": torch.bartlett_window}   # features.py:171-178 ('none' dereferences a None window in the reference)


class AudioToMelSpectrogramPreprocessor(NonTrainableNM):
    """audio_preprocessing.py:212-383 (wrapper) over FilterbankFeatures (parts/features.py:113-301).

    Built configuration = what the shipped configs use on the inference path: window 'hann' (also 'hamming',
    'blackman', 'bartlett'), normalize 'per_feature', log with the 'add' guard, mag_power 2, frame_splicing 1,
    stft_conv False (vi config) or True (quartznet15x5.yaml: the torch_stft convolution STFT = the same frames with
    scipy's periodic window; parity with that un-vendored package is unpinned), dither 0 (infer.py:89 forces 0; a
    non-zero dither is rejected rather than silently ignored)."""

    @property
    def input_ports(self):
        return {"input_signal": NeuralType(("B", "T"), AudioSignal(freq=self._sample_rate)),
                "length": NeuralType(tuple("B"), LengthsType())}

    @property
    def output_ports(self):
        return {"processed_signal": NeuralType(("B", "D", "T"), MelSpectrogramType()),
                "processed_length": NeuralType(tuple("B"), LengthsType())}

    def __init__(self, sample_rate=16000, window_size=0.02, window_stride=0.01, n_window_size=None,
                 n_window_stride=None, window="hann", normalize="per_feature", n_fft=None, preemph=0.97,
                 features=64, lowfreq=0, highfreq=None, log=True, log_zero_guard_type="add",
                 log_zero_guard_value=2 ** -24, dither=1e-5, pad_to=16, frame_splicing=1, stft_conv=False,
                 pad_value=0, mag_power=2.0):
        self._sample_rate = sample_rate
        if window_size and n_window_size:
            raise ValueError(f"{self} received both window_size and n_window_size. Only one should be specified.")
        if window_stride and n_window_stride:
            raise ValueError(f"{self} received both window_stride and n_window_stride. Only one should be specified.")
        if window_size:
            n_window_size = int(window_size * self._sample_rate)
        if window_stride:
            n_window_stride = int(window_stride * self._sample_rate)
        super().__init__()
        if (n_window_size is None or n_window_stride is None or not isinstance(n_window_size, int)
                or not isinstance(n_window_stride, int) or n_window_size <= 0 or n_window_stride <= 0):
            raise ValueError(f"{self} got an invalid value for either n_window_size or n_window_stride. "
                             f"Both must be positive ints.")  # parts/features.py:137-148
        if log_zero_guard_type not in ("add", "clamp"):
            raise ValueError(f"{self} received {log_zero_guard_type} for the log_zero_guard_type parameter. "
                             f"It must be either 'add' or 'clamp'.")  # parts/features.py:216-221
        unsupported = []
        if window not in _WINDOWS: unsupported.append(f"window={window!r}")
        if normalize != "per_feature": unsupported.append(f"normalize={normalize!r}")
        if not log or log_zero_guard_type != "add": unsupported.append("log/log_zero_guard_type")
        if isinstance(log_zero_guard_value, str): unsupported.append(f"log_zero_guard_value={log_zero_guard_value!r}")
        if mag_power != 2.0: unsupported.append(f"mag_power={mag_power}")
        if frame_splicing != 1: unsupported.append(f"frame_splicing={frame_splicing}")
        if pad_value != 0: unsupported.append(f"pad_value={pad_value}")
        if dither and dither > 0: unsupported.append(f"dither={dither} (inference path uses 0, infer.py:89)")
        if pad_to == "max": unsupported.append("pad_to='max'")
        if preemph is None: unsupported.append("preemph=None")
        if unsupported:
            raise ValueError("vasr_b200 AudioToMelSpectrogramPreprocessor: not built for " + ", ".join(unsupported))
        self.win_length, self.hop_length = n_window_size, n_window_stride
        self.n_fft = n_fft or 2 ** math.ceil(math.log2(n_window_size))
        self.nfilt = features
        self.pad_to = int(pad_to)
        highfreq = highfreq or sample_rate / 2
        # constructor-time constants, computed exactly like the reference does
        # features.py:171-180: symmetric torch window; `stft_conv: true` (features.py:156-167, quartznet15x5.yaml:26)
        # goes through torch_stft, whose basis is scaled by scipy's PERIODIC window (fftbins=True) - the same frames,
        # reflect padding and power spectrum otherwise, so it is the same kernel with another window table
        self.stft_conv = bool(stft_conv)
        self._window = _WINDOWS[window](n_window_size, periodic=bool(stft_conv)).to(torch.float)
        self._fb = slaney_mel_filterbank(sample_rate, self.n_fft, features, lowfreq, highfreq)  # :199-205
        if self.n_fft != 512:
            raise ValueError(f"vasr_b200 AudioToMelSpectrogramPreprocessor: only n_fft=512 is built (got {self.n_fft})")
        self._lib = _lib.load()
        self._cfg = _lib.FrontendCfg(n_window_size, n_window_stride, self.n_fft, features, float(preemph),
                                     float(log_zero_guard_value), self.pad_to)
        self._h_ = None   # device-side handle, created on first use (needs a CUDA context)

    @property
    def _h(self):
        if self._h_ is None:
            h = C.c_void_p()
            win = np.ascontiguousarray(self._window.numpy())
            fb = np.ascontiguousarray(self._fb)
            _lib.check(self._lib.vasr_frontend_create(C.byref(self._cfg), win.ctypes.data, fb.ctypes.data, C.byref(h)))
            self._h_ = h
        return self._h_

    def __del__(self):
        h = getattr(self, "_h_", None)
        if h is not None and h.value:
            self._lib.vasr_frontend_destroy(h)
            self._h_ = None

    @property
    def filter_banks(self):
        return torch.from_numpy(self._fb).unsqueeze(0)

    def get_seq_len(self, length):
        return torch.ceil(length.float() / self.hop_length).to(dtype=torch.long)

    def num_frames(self, L: int) -> int:
        """T_f = 1 + L // hop (torch.stft, center=True), padded to a multiple of pad_to (features.py:292-300)."""
        t = 1 + int(L) // self.hop_length
        if self.pad_to > 0 and t % self.pad_to:
            t += self.pad_to - t % self.pad_to
        return t

    def set_padding(self, per_utterance: bool):
        """False (default): the STFT reflects at the end of the padded batch row, the reference's semantics for a [B, L]
        tensor.  True: every utterance is reflected at its own length - what it sees when the reference transcribes it
        alone - so batched features equal the single-utterance ones for every length (vasr_frontend_set_padding)."""
        _lib.check(self._lib.vasr_frontend_set_padding(self._h, 1 if per_utterance else 0))
        self.pad_per_utterance = bool(per_utterance)

    def forward_channels_last(self, input_signal: torch.Tensor, length: torch.Tensor):
        _require_cuda(input_signal, "AudioToMelSpectrogramPreprocessor")
        x = input_signal.to(torch.float32).contiguous()
        ln = length.to(device=x.device, dtype=torch.int64).contiguous()
        if x.dim() != 2 or ln.dim() != 1 or ln.shape[0] != x.shape[0]:
            raise ValueError(f"input_signal must be [B, T] and length [B]; got {tuple(x.shape)}, {tuple(ln.shape)}")
        B, L = x.shape
        T = self.num_frames(L)
        feat = torch.empty((B, T, self.nfilt), dtype=torch.float32, device=x.device)
        seq = torch.empty((B,), dtype=torch.int64, device=x.device)
        _lib.check(self._lib.vasr_frontend_forward(self._h, x.data_ptr(), ln.data_ptr(), B, L,
                                                   feat.data_ptr(), seq.data_ptr(), _stream_ptr()))
        return feat, seq

    def forward(self, input_signal, length):
        feat, seq = self.forward_channels_last(input_signal, length)
        return feat.transpose(1, 2), seq

    def get_features(self, input_signal, length):
        return self.forward(input_signal, length)[0]


# --------------------------------------------------------------------------- acoustic model handle
class _ModelHandle:
    """One vasr_model shared by JasperEncoder and JasperDecoderForCTC instances of one pipeline.
    The C handle owns encoder and decoder weights; either module can be restored first."""

    def __init__(self, jasper: Sequence[dict], feat_in: int, num_classes_with_blank: int):
        lib = _lib.load()
        blocks = (_lib.BlockCfg * len(jasper))()
        for i, c in enumerate(jasper):
            k, s, d = c["kernel"], c["stride"], c["dilation"]
            k = k[0] if isinstance(k, (list, tuple)) else k
            s = s[0] if isinstance(s, (list, tuple)) else s
            d = d[0] if isinstance(d, (list, tuple)) else d
            blocks[i] = _lib.BlockCfg(int(c["filters"]), int(c["repeat"]), int(k), int(s), int(d),
                                      int(bool(c["residual"])), int(bool(c.get("separable", False))))
        h = C.c_void_p()
        _lib.check(lib.vasr_model_create(blocks, len(jasper), int(feat_in), int(num_classes_with_blank), C.byref(h)))
        self.h, self.lib = h, lib
        self.have_enc = self.have_dec = False
        self.finalized_mode: Optional[int] = None
        self.gemm_mode = _lib.GEMM_MODES["fp32"]

    def __del__(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.vasr_model_destroy(self.h)
            self.h = None

    def load(self, sd: Dict[str, torch.Tensor]):
        for name, t in sd.items():
            if name.endswith("num_batches_tracked"):
                continue
            a = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
            dims = (C.c_int64 * max(a.dim(), 1))(*a.shape)
            _lib.check(self.lib.vasr_model_load_tensor(self.h, name.encode(), a.data_ptr(), dims, a.dim(), 0))
        self.finalized_mode = None

    def ensure_final(self):
        if self.finalized_mode != self.gemm_mode:
            _lib.check(self.lib.vasr_model_finalize(self.h, self.gemm_mode))
            self.finalized_mode = self.gemm_mode


class _ConvHolder(nn.Module):
    """Parameter container with the reference's key layout (`<idx>.conv.weight`), never called."""

    def __init__(self, cin, cout, k, groups=1):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, k, groups=groups, bias=False)


class _BlockHolder(nn.Module):
    def __init__(self, cin, cfg):
        super().__init__()
        cout, rep = int(cfg["filters"]), int(cfg["repeat"])
        k = cfg["kernel"][0] if isinstance(cfg["kernel"], (list, tuple)) else cfg["kernel"]
        sep = bool(cfg.get("separable", False))
        mods: List[nn.Module] = []
        c = cin
        for r in range(rep):
            if sep:
                mods += [_ConvHolder(c, c, k, groups=c), _ConvHolder(c, cout, 1)]
            else:
                mods += [_ConvHolder(c, cout, k)]
            mods += [nn.BatchNorm1d(cout, eps=1e-3, momentum=0.1)]
            if r != rep - 1:
                mods += [nn.Identity(), nn.Identity()]      # activation + dropout slots (parts/jasper.py:236)
            c = cout
        self.mconv = nn.ModuleList(mods)
        if cfg["residual"]:
            self.res = nn.ModuleList([nn.ModuleList([_ConvHolder(cin, cout, 1),
                                                     nn.BatchNorm1d(cout, eps=1e-3, momentum=0.1)])])
        else:
            self.res = None


class JasperEncoder(TrainableNM):
    """nemo/collections/asr/jasper.py:17-204.  Options that no shipped config uses (groups>1, heads,
    SE, dense residual, non-batch norms, residual_mode='max', hardtanh/selu) raise ValueError."""

    @property
    def input_ports(self):
        return {"audio_signal": NeuralType(("B", "D", "T"), SpectrogramType()),
                "length": NeuralType(tuple("B"), LengthsType())}

    @property
    def output_ports(self):
        return {"outputs": NeuralType(("B", "D", "T"), AcousticEncodedRepresentation()),
                "encoded_lengths": NeuralType(tuple("B"), LengthsType())}

    def __init__(self, jasper, activation, feat_in, normalization_mode="batch", residual_mode="add",
                 norm_groups=-1, conv_mask=True, frame_splicing=1, init_mode="xavier_uniform",
                 gemm_mode: str = "fp32"):
        super().__init__()
        bad = []
        if activation != "relu": bad.append(f"activation={activation!r}")
        if normalization_mode != "batch": bad.append(f"normalization_mode={normalization_mode!r}")
        if residual_mode != "add": bad.append(f"residual_mode={residual_mode!r}")
        if not conv_mask: bad.append("conv_mask=False")
        if frame_splicing != 1: bad.append(f"frame_splicing={frame_splicing}")
        for i, c in enumerate(jasper):
            for key, dflt in (("groups", 1), ("heads", -1), ("se", False), ("residual_dense", False),
                              ("kernel_size_factor", 1.0), ("tied", False)):
                if c.get(key, dflt) != dflt:
                    bad.append(f"jasper[{i}].{key}={c[key]!r}")
        if bad:
            raise ValueError("vasr_b200 JasperEncoder: not built for " + ", ".join(bad))
        if gemm_mode not in _lib.GEMM_MODES:
            raise ValueError(f"gemm_mode must be one of {sorted(_lib.GEMM_MODES)}, got {gemm_mode!r}")
        self._jasper = [dict(c) for c in jasper]
        self._feat_in = feat_in * frame_splicing
        self._out_ch = int(jasper[-1]["filters"])
        blocks, cin = [], self._feat_in
        for c in jasper:
            blocks.append(_BlockHolder(cin, c))
            cin = int(c["filters"])
        self.encoder = nn.Sequential(*blocks)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p, gain=1.0)       # init_weights(mode='xavier_uniform'), parts/jasper.py:27-52
        self._gemm_mode = gemm_mode
        self._model: Optional[_ModelHandle] = None
        self.__dict__["_decoder"] = None        # Optional[JasperDecoderForCTC], see attach_decoder
        self._dirty = True
        self._ws: Optional[torch.Tensor] = None
        if self._device.type != "cuda" or torch.cuda.is_available():
            self.to(self._device)      # jasper.py:196,251 (without a GPU only the symbolic graph can be built)

    # -- C handle; `attach_decoder` makes it a whole-model handle shared with the decoder module
    def _handle(self) -> _ModelHandle:
        if self._model is None:
            nc = self._decoder._num_classes if self._decoder is not None else 0
            self._model = _ModelHandle(self._jasper, self._feat_in, nc)
        return self._model

    def attach_decoder(self, decoder: "JasperDecoderForCTC"):
        """Share one C handle (encoder + decoder head) - needed by the fused whole-path host call."""
        if decoder._feat_in != self._out_ch:
            raise ValueError(f"decoder feat_in {decoder._feat_in} != encoder output channels {self._out_ch}")
        # plain attribute slots: registering either module as a child of the other would make the
        # nn.Module tree cyclic
        self.__dict__["_decoder"] = decoder
        decoder.__dict__["_shared_from"] = self
        self._model = None
        self._dirty = True
        decoder._dirty = True

    def load_state_dict(self, state_dict, strict=True):
        res = super().load_state_dict(state_dict, strict=strict)
        self._dirty = True
        return res

    def set_gemm_mode(self, mode: str):
        if mode not in _lib.GEMM_MODES:
            raise ValueError(f"gemm_mode must be one of {sorted(_lib.GEMM_MODES)}, got {mode!r}")
        self._gemm_mode = mode

    def _sync_weights(self):
        h = self._handle()
        if self._dirty:
            h.load(self.state_dict())
            self._dirty = False
        if self._decoder is not None and self._decoder._dirty:
            h.load(self._decoder.state_dict())
            self._decoder._dirty = False
        h.gemm_mode = _lib.GEMM_MODES[self._gemm_mode]
        h.ensure_final()
        return h

    def out_frames(self, T_f: int) -> int:
        return int(self._handle().lib.vasr_model_out_frames(self._handle().h, int(T_f)))

    def out_frames_of(self, feat_frames: torch.Tensor) -> torch.Tensor:
        """Per-utterance T_e for per-utterance feature-frame counts [B] (int tensor, any device): the conv arithmetic
        of `MaskedConv1d.get_seq_len` (parts/jasper.py:108-111) on integers.  T' = (T + c) // s + 1 with
        c = 2p - d(k-1) - 1; 'same'-padded stride-1 layers (c = -1) are the identity and are skipped, so QuartzNet
        costs one integer op per strided block instead of one per sub-block."""
        steps = self.__dict__.get("_frame_steps")
        if steps is None:
            steps = []
            first = lambda v: int(v[0] if isinstance(v, (list, tuple)) else v)
            for c in self._jasper:
                k, s_, d = first(c["kernel"]), first(c.get("stride", 1)), first(c.get("dilation", 1))
                pad = (d * k) // 2 - 1 if d > 1 else k // 2
                cc = 2 * pad - d * (k - 1) - 1
                if s_ == 1 and cc == -1:
                    continue
                steps += [(cc, s_)] * int(c["repeat"])
            self.__dict__["_frame_steps"] = steps
        t = feat_frames.to(torch.int64)
        for cc, s_ in steps:
            t = torch.div(t + cc, s_, rounding_mode="floor") + 1
        return t.clamp(min=0).to(torch.int32)

    def check_range(self, B: int):
        """Raise RuntimeError if the last forward on this module overflowed the fp16 operand range of the tensor-core
        modes (synchronises the current stream; see vasr_encoder_check in include/vasr_b200.h)."""
        if self._ws is None or self._model is None:
            return
        h = self._model
        _lib.check(h.lib.vasr_encoder_check(h.h, self._ws.data_ptr(), self._ws.numel(), int(B), _stream_ptr()))

    def forward_channels_last(self, feat: torch.Tensor, length: torch.Tensor):
        """feat [B, T_f, feat_in] contiguous fp32, length [B] -> enc [B, T_e, C], enc_len [B] f32."""
        _require_cuda(feat, "JasperEncoder")
        h = self._sync_weights()
        B, T_f, F = feat.shape
        if F != self._feat_in:
            raise ValueError(f"JasperEncoder: expected {self._feat_in} input features, got {F}")
        ln = length.to(device=feat.device, dtype=torch.int64).contiguous()
        T_e = h.lib.vasr_model_out_frames(h.h, T_f)
        enc = torch.empty((B, T_e, self._out_ch), dtype=torch.float32, device=feat.device)
        enc_len = torch.empty((B,), dtype=torch.float32, device=feat.device)
        need = int(h.lib.vasr_encoder_workspace_bytes(h.h, B, T_f))
        if self._ws is None or self._ws.numel() < need or self._ws.device != feat.device:
            self._ws = torch.empty((need,), dtype=torch.uint8, device=feat.device)
        _lib.check(h.lib.vasr_encoder_forward(h.h, feat.data_ptr(), ln.data_ptr(), B, T_f, enc.data_ptr(),
                                              enc_len.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                              _stream_ptr()))
        return enc, enc_len

    def forward(self, audio_signal, length=None):
        feat = _as_channels_last(audio_signal)
        if length is None:
            length = torch.full((feat.shape[0],), feat.shape[1], dtype=torch.int64, device=feat.device)
            return self.forward_channels_last(feat, length)[0].transpose(1, 2)
        enc, enc_len = self.forward_channels_last(feat, length)
        return enc.transpose(1, 2), enc_len


class JasperDecoderForCTC(TrainableNM):
    """nemo/collections/asr/jasper.py:207-254: Conv1d(feat_in, V+1, 1) -> transpose -> log_softmax."""

    @property
    def input_ports(self):
        return {"encoder_output": NeuralType(("B", "D", "T"), AcousticEncodedRepresentation())}

    @property
    def output_ports(self):
        return {"output": NeuralType(("B", "T", "D"), LogprobsType())}

    def __init__(self, feat_in, num_classes, init_mode="xavier_uniform"):
        super().__init__()
        self._feat_in = feat_in
        self._num_classes = num_classes + 1          # + blank (jasper.py:246-247)
        self.decoder_layers = nn.Sequential(nn.Conv1d(feat_in, self._num_classes, kernel_size=1, bias=True))
        nn.init.xavier_uniform_(self.decoder_layers[0].weight, gain=1.0)
        self._own: Optional[_ModelHandle] = None      # decoder-only C handle (no encoder blocks)
        self.__dict__["_shared_from"] = None     # Optional[JasperEncoder], see JasperEncoder.attach_decoder
        self._dirty = True
        if self._device.type != "cuda" or torch.cuda.is_available():
            self.to(self._device)      # jasper.py:196,251 (without a GPU only the symbolic graph can be built)

    def load_state_dict(self, state_dict, strict=True):
        res = super().load_state_dict(state_dict, strict=strict)
        self._dirty = True
        return res

    def _sync(self):
        if self._shared_from is not None:
            return self._shared_from._sync_weights()
        if self._own is None:
            self._own = _ModelHandle([], self._feat_in, self._num_classes)
        h = self._own
        if self._dirty:
            h.load(self.state_dict())
            self._dirty = False
        h.ensure_final()
        return h

    def forward_channels_last(self, enc: torch.Tensor, want_log_probs: bool = True):
        """enc [B, T_e, feat_in] -> (log_probs [B, T_e, V+1] or None, ids [B, T_e] i64)."""
        _require_cuda(enc, "JasperDecoderForCTC")
        h = self._sync()
        B, T_e, Cc = enc.shape
        if Cc != self._feat_in:
            raise ValueError(f"JasperDecoderForCTC: expected {self._feat_in} channels, got {Cc}")
        logp = torch.empty((B, T_e, self._num_classes), dtype=torch.float32, device=enc.device) if want_log_probs else None
        ids = torch.empty((B, T_e), dtype=torch.int64, device=enc.device)
        _lib.check(h.lib.vasr_decoder_forward(h.h, enc.data_ptr(), B, T_e,
                                              logp.data_ptr() if logp is not None else None,
                                              ids.data_ptr(), _stream_ptr()))
        return logp, ids

    def forward(self, encoder_output):
        enc = _as_channels_last(encoder_output)
        logp, ids = self.forward_channels_last(enc, True)
        logp._vasr_greedy_ids = ids       # lets GreedyCTCDecoder reuse the fused argmax
        return logp


class GreedyCTCDecoder(TrainableNM):
    """nemo/collections/asr/greedy_ctc_decoder.py:9-36."""

    @property
    def input_ports(self):
        return {"log_probs": NeuralType(("B", "T", "D"), LogprobsType())}

    @property
    def output_ports(self):
        return {"predictions": NeuralType(("B", "T"), PredictionsType())}

    def __init__(self):
        super().__init__()

    def forward(self, log_probs):
        ids = getattr(log_probs, "_vasr_greedy_ids", None)
        if ids is not None:
            return ids
        _require_cuda(log_probs, "GreedyCTCDecoder")
        lp = log_probs.to(torch.float32).contiguous()
        B, T, V = lp.shape
        out = torch.empty((B, T), dtype=torch.int64, device=lp.device)
        lib = _lib.load()
        _lib.check(lib.vasr_greedy_argmax(lp.data_ptr(), B * T, V, out.data_ptr(), _stream_ptr()))
        return out


# --------------------------------------------------------------------------- beam search
class NGramLM:
    """Device copy of a KenLM binary for the beam search (what pyctcdecode gets from `kenlm.Model(lm_path)`,
    beam_search_decoder.py:82-87).  The file is decoded on the host (kenlm_binary.KenlmModel) and uploaded once by
    `vasr_lm_create`; `vocab` (the acoustic model's labels) fixes the spelling -> word-id table."""

    def __init__(self, lm_path: str, vocab: Sequence[str]):
        from .kenlm_binary import KenlmModel
        if not torch.cuda.is_available():
            raise RuntimeError("NGramLM needs a CUDA device (vasr_b200 has no CPU path)")
        lib = _lib.load()
        m = KenlmModel(lm_path)

        def hash_labels(ids):
            arr = (C.c_int32 * len(ids))(*ids)
            return int(lib.vasr_lm_hash_labels(arr, len(ids)))
        table = m.vocabulary_table(list(vocab), hash_labels)
        keep = []                                           # host arrays must outlive the call

        def ptr(a, dtype):
            a = np.ascontiguousarray(a, dtype=dtype)
            keep.append(a)
            return a.ctypes.data
        arrs = _lib.LmArrays()
        arrs.order, arrs.vocab, arrs.bos, arrs.eos = m.order, m.counts[0], m.bos, m.eos
        arrs.counts = ptr(np.asarray(m.counts), np.uint64)
        arrs.uni_prob, arrs.uni_backoff = ptr(m.uni_prob, np.float32), ptr(m.uni_backoff, np.float32)
        arrs.uni_next = ptr(m.uni_next, np.uint32)
        for k in range(m.order - 2):
            arrs.mid_word[k] = ptr(m.mid_word[k], np.int32)
            arrs.mid_prob[k] = ptr(m.mid_prob[k], np.float32)
            arrs.mid_backoff[k] = ptr(m.mid_backoff[k], np.float32)
            arrs.mid_next[k] = ptr(m.mid_next[k], np.uint32)
        arrs.long_word, arrs.long_prob = ptr(m.long_word, np.int32), ptr(m.long_prob, np.float32)
        arrs.vocab_keys, arrs.vocab_vals = ptr(table["keys"], np.uint64), ptr(table["vals"], np.int32)
        arrs.vocab_slots = len(table["keys"])
        h = C.c_void_p()
        _lib.check(lib.vasr_lm_create(C.byref(arrs), C.byref(h)))
        self._h = h
        self.order, self.counts, self.path = m.order, list(m.counts), lm_path
        self.words, self.word2id, self.bos, self.eos = m.words, m.word2id, m.bos, m.eos

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            _lib.load().vasr_lm_destroy(h)
            self._h = None

    def score_batch(self, ctx: torch.Tensor, nctx: torch.Tensor, word: torch.Tensor) -> torch.Tensor:
        """log10 P(word | ctx) on the device: ctx [N, 4] i32 (oldest -> newest), nctx [N] i32, word [N] i32 -> [N] f64."""
        _require_cuda(ctx, "NGramLM.score_batch")
        ctx = ctx.to(torch.int32).contiguous(); nctx = nctx.to(torch.int32).contiguous(); word = word.to(torch.int32).contiguous()
        if ctx.dim() != 2 or ctx.shape[1] != 4:
            raise ValueError("ctx must be [N, 4]")
        out = torch.empty((ctx.shape[0],), dtype=torch.float64, device=ctx.device)
        _lib.check(_lib.load().vasr_lm_score_batch(self._h, ctx.data_ptr(), nctx.data_ptr(), word.data_ptr(),
                                                    out.data_ptr(), ctx.shape[0], _stream_ptr()))
        return out


def _frames_arg(frames, B: int, device):
    """Optional per-utterance frame counts -> (tensor kept alive, device pointer or None)."""
    if frames is None:
        return None, None
    f = torch.as_tensor(frames).to(device=device, dtype=torch.int32).contiguous()
    if f.shape != (B,):
        raise ValueError(f"frames must have shape [{B}], got {tuple(f.shape)}")
    return f, f.data_ptr()


def ctc_beam_search(log_probs: torch.Tensor, vocab: Sequence[str], beam_width: int,
                    token_min_logp: float = -5.0, beam_prune_logp: float = -10.0, lm: Optional[NGramLM] = None,
                    alpha: float = 0.5, beta: float = 1.5, unk_score_offset: float = -10.0, frames=None):
    """Device prefix beam search, optionally with n-gram LM fusion: log_probs [B, T, V+1] ->
    (ids [B, T] i32, len [B] i32, score [B] f32).  `frames` [B]: utterance b is searched over its first frames[b]
    frames only (a zero-padded batch; None = all T frames, the reference's single-utterance behaviour)."""
    _require_cuda(log_probs, "ctc_beam_search")
    lp = log_probs.to(torch.float32).contiguous()
    B, T, V1 = lp.shape
    fr, fr_ptr = _frames_arg(frames, B, lp.device)
    if V1 != len(vocab) + 1:
        raise ValueError(f"log_probs has {V1} classes but the vocabulary has {len(vocab)} labels (+1 blank)")
    lib = _lib.load()
    ids = torch.empty((B, T), dtype=torch.int32, device=lp.device)
    n = torch.empty((B,), dtype=torch.int32, device=lp.device)
    sc = torch.empty((B,), dtype=torch.float32, device=lp.device)
    space_id = list(vocab).index(" ") if " " in vocab else -1
    if lm is None:
        ws = torch.empty((int(lib.vasr_ctc_beam_workspace_bytes(B, T)),), dtype=torch.uint8, device=lp.device)
        _lib.check(lib.vasr_ctc_beam_search(lp.data_ptr(), fr_ptr, B, T, V1, len(vocab), space_id, int(beam_width),
                                            float(token_min_logp), float(beam_prune_logp), ws.data_ptr(), ws.numel(),
                                            ids.data_ptr(), n.data_ptr(), sc.data_ptr(), _stream_ptr()))
    else:
        nbytes = int(lib.vasr_ctc_beam_lm_workspace_bytes(B, T, int(beam_width)))
        ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=lp.device)
        _lib.check(lib.vasr_ctc_beam_search_lm(lp.data_ptr(), fr_ptr, B, T, V1, len(vocab), space_id, int(beam_width),
                                               float(token_min_logp), float(beam_prune_logp), lm._h, float(alpha),
                                               float(beta), float(unk_score_offset), ws.data_ptr(), ws.numel(),
                                               ids.data_ptr(), n.data_ptr(), sc.data_ptr(), _stream_ptr()))
    return ids, n, sc


class BeamSearchDecoderWithLM(NonTrainableNM):
    """nemo/collections/asr/beam_search_decoder.py:14-102: pyctcdecode's prefix beam search, here as a batched CUDA
    kernel (the reference is CPU-only and asserts batch size 1, :96).  `lm_path=None`: no language model (the mode
    infer.py:118-130 falls back to when kenlm is not importable).  `lm_path=<KenLM binary>`: n-gram shallow fusion
    with `alpha`, `beta` like `build_ctcdecoder(vocab, kenlm_model_path=lm_path, alpha, beta)` (:82-87); the file is
    decoded by kenlm_binary.py and searched on the GPU.  `log_probs_length` is ignored like in the reference
    (:95-101).  pyctcdecode and kenlm are un-vendored third-party packages: parity is against the restatements in
    oracle/beam_oracle.py and oracle/kenlm_oracle.py (unpinned against the packages)."""

    @property
    def input_ports(self):
        return {"log_probs": NeuralType(("B", "T", "D"), LogprobsType()),
                "log_probs_length": NeuralType(tuple("B"), LengthsType())}

    @property
    def output_ports(self):
        return {"predictions": NeuralType(("B", "T"), PredictionsType())}

    def __init__(self, lm_path, vocab, beam_width, alpha, beta, num_cpus, cutoff_prob=1.0, cutoff_top_n=40,
                 input_tensor=True):
        super().__init__()
        if self._factory.world_size > 1:
            raise ValueError("BeamSearchDecoderWithLM does not run in distributed mode")   # beam_search_decoder.py:79-80
        if not 1 <= int(beam_width) <= 128:
            raise ValueError(f"beam_width must be in [1, 128], got {beam_width}")
        self.vocab = list(vocab)
        self.beam_width = int(beam_width)
        self.alpha, self.beta = float(alpha), float(beta)
        self.num_cpus, self.cutoff_prob, self.cutoff_top_n, self.input_tensor = num_cpus, cutoff_prob, cutoff_top_n, input_tensor
        self.lm_path = lm_path
        self.lm = NGramLM(lm_path, self.vocab) if lm_path is not None else None

    def decode_batch(self, log_probs, frames=None) -> List[str]:
        """`frames` [B]: frames of every utterance in a zero-padded batch (each is decoded as if it were alone, which
        is the only way the reference ever runs this decoder, :96); None = all frames of the tensor."""
        ids, n, _ = ctc_beam_search(log_probs, self.vocab, self.beam_width, lm=self.lm, alpha=self.alpha, beta=self.beta,
                                    frames=frames)
        return [" ".join(t.split()) for t in ids_to_text(ids, n, self.vocab)]

    def forward(self, log_probs, log_probs_length=None):
        texts = self.decode_batch(log_probs)
        return texts[0] if len(texts) == 1 else texts


# --------------------------------------------------------------------------- helpers.py
def ctc_collapse(predictions: torch.Tensor, blank: int, frames=None):
    """Device CTC collapse -> (ids [B, T] int32 padded with -1, lengths [B] int32).  `frames` [B]: see
    `ctc_beam_search` (None = all T frames, helpers.py:26-30)."""
    _require_cuda(predictions, "ctc_collapse")
    p = predictions.to(torch.int64).contiguous()
    B, T = p.shape
    fr, fr_ptr = _frames_arg(frames, B, p.device)
    out = torch.empty((B, T), dtype=torch.int32, device=p.device)
    n = torch.empty((B,), dtype=torch.int32, device=p.device)
    lib = _lib.load()
    _lib.check(lib.vasr_ctc_collapse(p.data_ptr(), fr_ptr, B, T, int(blank), out.data_ptr(), n.data_ptr(), _stream_ptr()))
    return out, n


def ids_to_text(out_ids, out_len, labels: Sequence[str]) -> List[str]:
    out_ids = out_ids.cpu().numpy() if isinstance(out_ids, torch.Tensor) else np.asarray(out_ids)
    out_len = out_len.cpu().numpy() if isinstance(out_len, torch.Tensor) else np.asarray(out_len)
    return ["".join(labels[c] for c in row[:n]) for row, n in zip(out_ids, out_len)]


def post_process_predictions(predictions: List[torch.Tensor], labels: Sequence[str]) -> List[str]:
    """nemo/collections/asr/helpers.py:207-208 -> __ctc_decoder_predictions_tensor (:7-33), with the
    per-frame Python loop replaced by the device collapse kernel.  blank id = len(labels)."""
    hyps: List[str] = []
    for p in predictions:
        if not p.is_cuda:
            p = p.cuda()
        ids, n = ctc_collapse(p, len(labels))
        hyps += ids_to_text(ids, n, labels)
    return hyps
