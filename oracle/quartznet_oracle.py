"""CPU oracle for the VietASR CTC inference hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product path
(``viet-asr_b200/``) never imports anything from ``oracle/`` and fails
loudly when its CUDA library is missing.

It is a plain torch-CPU (fp32) restatement of the reference's arithmetic,
function by function, each citing the reference file:line it follows
(paths relative to the reference checkout):

  * ``slaney_mel_filterbank``  - librosa.filters.mel(htk=False, norm='slaney'),
    called at nemo/collections/asr/parts/features.py:199-202 (librosa itself
    is a third-party dependency that is absent from the reference tree and
    from this image; restated from its published algorithm and pinned
    against torchaudio.functional.melscale_fbanks in tests).
  * ``filterbank_features``    - FilterbankFeatures.forward,
    nemo/collections/asr/parts/features.py:245-301 (+ ctor :113-236).
  * ``normalize_batch``        - nemo/collections/asr/parts/features.py:17-30.
  * ``masked_conv1d``          - MaskedConv1d.forward / get_seq_len,
    nemo/collections/asr/parts/jasper.py:108-132.
  * ``jasper_block`` / ``encoder_forward`` - JasperBlock.forward
    (parts/jasper.py:408-448, structure :175-288, :329-400) and
    JasperEncoder.forward (nemo/collections/asr/jasper.py:198-204).
  * ``decoder_forward``        - JasperDecoderForCTC.forward, jasper.py:249-254.
  * ``greedy_argmax``          - GreedyCTCDecoder.forward, greedy_ctc_decoder.py:33-36.
  * ``ctc_collapse`` / ``ids_to_text`` - __ctc_decoder_predictions_tensor,
    nemo/collections/asr/helpers.py:7-33.

Parity pinning: ``oracle/make_golden.py`` runs this restatement side by side
with the reference's own ``parts/jasper.py`` (imported unmodified by file
path, in the build container where /root/reference exists) on the shipped
checkpoints and WAVs, asserts agreement, and writes the fixtures under
``tests/golden/``.  The reference ships no tests or golden vectors of its
own, so those generated fixtures are the pin.
Beam search + KenLM (pyctcdecode) is NOT pinned: see oracle/beam_oracle.py.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

CONSTANT = 1e-5  # features.py:14


# --------------------------------------------------------------------------
# front end
# --------------------------------------------------------------------------
def _hz_to_mel_slaney(f: np.ndarray) -> np.ndarray:
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3.0
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_part = min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, log_part, mels)


def _mel_to_hz_slaney(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3.0
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def slaney_mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with librosa's defaults
    (htk=False, norm='slaney', float32 result) -> [n_mels, 1 + n_fft//2]."""
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    mel_f = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


def get_seq_len(length: torch.Tensor, hop: int) -> torch.Tensor:
    # features.py:238-239
    return torch.ceil(length.float() / hop).to(dtype=torch.long)


def normalize_batch(x: torch.Tensor, seq_len: torch.Tensor) -> torch.Tensor:
    """features.py:17-30, normalize_type == 'per_feature'."""
    x_mean = torch.zeros((seq_len.shape[0], x.shape[1]), dtype=x.dtype)
    x_std = torch.zeros((seq_len.shape[0], x.shape[1]), dtype=x.dtype)
    for i in range(x.shape[0]):
        x_mean[i, :] = x[i, :, : seq_len[i]].mean(dim=1)
        x_std[i, :] = x[i, :, : seq_len[i]].std(dim=1)  # unbiased
    x_std += CONSTANT
    return (x - x_mean.unsqueeze(2)) / x_std.unsqueeze(2)


def conv_stft_magnitude(x: torch.Tensor, n_fft: int, hop: int, win_length: int, window: str) -> torch.Tensor:
    """`torch_stft.STFT(n_fft, hop, win_length, window).transform(x)[0]` - what `stft_conv: true` selects
    (features.py:156-167; quartznet15x5.yaml:26).  torch_stft is an un-vendored third-party package that is absent
    from the image: restated from its published algorithm, parity with the package UNPINNED.  Algorithm: Fourier basis
    rows [cos; -sin] of the first n_fft/2+1 bins, multiplied by `scipy.signal.get_window(window, win_length,
    fftbins=True)` (a PERIODIC window, unlike the symmetric one of the torch path) zero-padded to n_fft around the
    centre; the signal is reflect-padded by n_fft/2 and correlated with the basis at stride `hop` (conv1d);
    magnitude = sqrt(re^2 + im^2)."""
    from scipy.signal import get_window
    import numpy as np
    cutoff = n_fft // 2 + 1
    basis = np.fft.fft(np.eye(n_fft))
    basis = np.vstack([np.real(basis[:cutoff, :]), np.imag(basis[:cutoff, :])])
    win = get_window(window, win_length, fftbins=True)
    lpad = (n_fft - win_length) // 2
    win = np.pad(win, (lpad, n_fft - win_length - lpad))
    fwd = torch.from_numpy((basis * win[None, :]).astype(np.float32))[:, None, :]
    xp = F.pad(x.unsqueeze(1).unsqueeze(1), (n_fft // 2, n_fft // 2, 0, 0), mode="reflect").squeeze(1)
    ft = F.conv1d(xp, fwd, stride=hop, padding=0)
    re, im = ft[:, :cutoff, :], ft[:, cutoff:, :]
    return torch.sqrt(re ** 2 + im ** 2)


def filterbank_features(
    x: torch.Tensor,
    length: torch.Tensor,
    sample_rate: int = 16000,
    n_window_size: int = 320,
    n_window_stride: int = 160,
    n_fft: int = 512,
    nfilt: int = 64,
    preemph: float = 0.97,
    log_zero_guard_value: float = 2 ** -24,
    pad_to: int = 0,
    return_pre_norm: bool = False,
    window: str = "hann",
    stft_conv: bool = False,
):
    """FilterbankFeatures.forward (features.py:245-301) with the settings that
    apply on the inference path: dither=0 and pad_to=0 (infer.py:89-90),
    window='hann', normalize='per_feature', log 'add' guard, mag_power=2,
    stft_conv=False, frame_splicing=1 (ctor defaults
    audio_preprocessing.py:314-337; configs/quartznet12x1_vi.yaml:8-18).

    Forced deviation: torch>=2 requires return_complex=True (features.py:181-188
    uses the legacy real view); re^2+im^2 is computed identically.
    x: [B, L] float32, length: [B] int64 -> ([B, nfilt, T_f], [B] int64)
    """
    x = x.to(torch.float32)
    seq_len = get_seq_len(length, n_window_stride)  # :247
    # :255 preemphasis over the whole padded row
    x = torch.cat((x[:, 0].unsqueeze(1), x[:, 1:] - preemph * x[:, :-1]), dim=1)
    if stft_conv:
        power = conv_stft_magnitude(x, n_fft, n_window_stride, n_window_size, window).pow(2.0)  # :156-167, :260-261
    else:
        window_fn = {"hann": torch.hann_window, "hamming": torch.hamming_window, "blackman": torch.blackman_window,
                     "bartlett": torch.bartlett_window, "none": None}.get(window, None)  # :171-178
        win = window_fn(n_window_size, periodic=False).to(torch.float) if window_fn else None  # :179-180
        spec = torch.stft(
            x, n_fft=n_fft, hop_length=n_window_stride, win_length=n_window_size,
            center=True, window=win, return_complex=True,
        )  # :181-188, pad_mode='reflect' default
        power = torch.view_as_real(spec).pow(2.0).sum(-1)  # :260-263
    fb = torch.from_numpy(
        slaney_mel_filterbank(sample_rate, n_fft, nfilt, 0.0, sample_rate / 2)
    ).unsqueeze(0)  # :199-205
    mel = torch.matmul(fb, power)  # :266
    logmel = torch.log(mel + log_zero_guard_value)  # :269-271
    out = normalize_batch(logmel, seq_len)  # :284
    max_len = out.size(-1)
    mask = torch.arange(max_len).expand(out.size(0), max_len) >= seq_len.unsqueeze(1)
    out = out.masked_fill(mask.unsqueeze(1), 0.0)  # :287-290
    if pad_to > 0:  # :292-300
        pad_amt = out.size(-1) % pad_to
        if pad_amt != 0:
            out = F.pad(out, (0, pad_to - pad_amt), value=0.0)
    if return_pre_norm:
        return out, seq_len, logmel
    return out, seq_len


# --------------------------------------------------------------------------
# encoder
# --------------------------------------------------------------------------
def get_same_padding(kernel_size: int, stride: int, dilation: int) -> int:
    # parts/jasper.py:60-65
    if stride > 1 and dilation > 1:
        raise ValueError("Only stride OR dilation may be greater than 1")
    if dilation > 1:
        return (dilation * kernel_size) // 2 - 1
    return kernel_size // 2


def masked_conv1d(x, lens, weight, stride=1, padding=0, dilation=1, groups=1):
    """MaskedConv1d.forward (parts/jasper.py:113-132) with use_mask=True."""
    lens_i = lens.to(dtype=torch.long)  # :115
    max_len = x.size(2)
    mask = torch.arange(max_len).expand(len(lens_i), max_len) >= lens_i.unsqueeze(1)
    x = x.masked_fill(mask.unsqueeze(1), 0)  # :116-118
    k = weight.shape[2]
    new_lens = (lens_i + 2 * padding - dilation * (k - 1) - 1) / stride + 1  # :108-111 (true division)
    out = F.conv1d(x, weight, None, stride=stride, padding=padding, dilation=dilation, groups=groups)
    return out, new_lens


def _bn_eval(x, sd, prefix, eps=1e-3):
    # nn.BatchNorm1d(C, eps=1e-3) in eval mode, parts/jasper.py:392
    return F.batch_norm(
        x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
        sd[prefix + ".weight"], sd[prefix + ".bias"], training=False, eps=eps,
    )


def jasper_block(sd: Dict[str, torch.Tensor], b: int, cfg: dict, x: torch.Tensor, lens: torch.Tensor,
                 taps: Optional[list] = None):
    """One JasperBlock (parts/jasper.py:408-448) in eval mode; dropout p=0 is a no-op.

    State-dict keys follow the shipped checkpoints:
    sub-block r: mconv.{5r} dw, mconv.{5r+1} pw, mconv.{5r+2} BN (separable) or
    mconv.{4r} conv, mconv.{4r+1} BN (non separable); residual res.0.0 / res.0.1.
    """
    repeat = int(cfg["repeat"])
    k = int(cfg["kernel"][0]); stride = int(cfg["stride"][0]); dil = int(cfg["dilation"][0])
    separable = bool(cfg.get("separable", False))
    residual = bool(cfg["residual"])
    pad = get_same_padding(k, stride, dil)
    lens_orig = lens
    xin = x
    out = x
    per = 5 if separable else 4  # conv(s) + BN + act + dropout slots per sub-block (:219-257)
    for r in range(repeat):
        base = f"encoder.{b}.mconv.{per * r}"
        if separable:
            w_dw = sd[f"encoder.{b}.mconv.{per * r}.conv.weight"]
            w_pw = sd[f"encoder.{b}.mconv.{per * r + 1}.conv.weight"]
            out, lens = masked_conv1d(out, lens, w_dw, stride, pad, dil, groups=w_dw.shape[0])
            out, lens = masked_conv1d(out, lens, w_pw, 1, 0, 1, 1)
            out = _bn_eval(out, sd, f"encoder.{b}.mconv.{per * r + 2}")
        else:
            w = sd[base + ".conv.weight"]
            out, lens = masked_conv1d(out, lens, w, stride, pad, dil, 1)
            out = _bn_eval(out, sd, f"encoder.{b}.mconv.{per * r + 1}")
        if r != repeat - 1:
            out = F.relu(out)
            if taps is not None:
                taps.append(out)
    if residual:
        w_r = sd[f"encoder.{b}.res.0.0.conv.weight"]
        res, _ = masked_conv1d(xin, lens_orig, w_r, 1, 0, 1, 1)  # :430-434 masks with block-input lens
        res = _bn_eval(res, sd, f"encoder.{b}.res.0.1")
        out = out + res  # :439
    out = F.relu(out)  # :444
    if taps is not None:
        taps.append(out)
    return out, lens


def encoder_forward(sd, jasper_cfg: Sequence[dict], feats: torch.Tensor, lens: torch.Tensor,
                    taps: Optional[list] = None):
    """JasperEncoder.forward (jasper.py:198-204): [B,64,T_f],[B] i64 -> [B,1024,T_e],[B] f32."""
    x = feats
    for b, cfg in enumerate(jasper_cfg):
        x, lens = jasper_block(sd, b, cfg, x, lens, taps)
    return x, lens


def decoder_forward(sd, enc: torch.Tensor) -> torch.Tensor:
    """JasperDecoderForCTC.forward (jasper.py:249-254): -> log-probs [B, T_e, V+1]."""
    y = F.conv1d(enc, sd["decoder_layers.0.weight"], sd["decoder_layers.0.bias"])
    return F.log_softmax(y.transpose(1, 2), dim=-1)


def decoder_logits(sd, enc: torch.Tensor) -> torch.Tensor:
    """Pre-softmax logits [B, T_e, V+1] (for the 1e-3 rel parity bound)."""
    y = F.conv1d(enc, sd["decoder_layers.0.weight"], sd["decoder_layers.0.bias"])
    return y.transpose(1, 2).contiguous()


def greedy_argmax(log_probs: torch.Tensor) -> torch.Tensor:
    # greedy_ctc_decoder.py:35 ; ties -> lowest index (torch.argmax)
    return log_probs.argmax(dim=-1, keepdim=False)


def utterance_frames(jasper_cfg: Sequence[dict], length, hop: int = 160) -> List[int]:
    """Encoder frames each utterance has when the reference runs it ALONE (infer.py:167-171: one utterance per call):
    T_f = 1 + L // hop (features.py:245-301, center=True, pad_to=0), then `(T + 2p - d(k-1) - 1) // s + 1` per
    sub-block (parts/jasper.py:108-111).  A zero-padded batch is decoded utterance by utterance over these frames."""
    out = []
    for L in np.asarray(length).reshape(-1).tolist():
        t = int(L) // hop + 1
        for c in jasper_cfg:
            k = c["kernel"][0] if isinstance(c["kernel"], (list, tuple)) else c["kernel"]
            st = c["stride"][0] if isinstance(c.get("stride", 1), (list, tuple)) else c.get("stride", 1)
            d = c["dilation"][0] if isinstance(c.get("dilation", 1), (list, tuple)) else c.get("dilation", 1)
            for _ in range(int(c["repeat"])):
                t = (t + 2 * get_same_padding(k, st, d) - d * (k - 1) - 1) // st + 1
        out.append(max(t, 0))
    return out


def filterbank_features_each_alone(x: torch.Tensor, length: torch.Tensor, **kw):
    """Features of a zero-padded batch computed the way the reference's inference path computes them: one utterance per
    call (infer.py:167-171), i.e. `filterbank_features` on x[b, :length[b]] alone, then zero-padded to the batch's frame
    count (frames beyond an utterance's own are masked zeros in the batched tensor anyway, features.py:287-290).  This is
    what `vasr_frontend_set_padding(fe, 1)` reproduces on the device."""
    T = x.shape[1] // kw.get("n_window_stride", 160) + 1
    feats, seqs = [], []
    for b in range(x.shape[0]):
        n = int(length[b])
        f, s = filterbank_features(x[b:b + 1, :n], length[b:b + 1], **kw)
        feats.append(F.pad(f, (0, T - f.shape[2])))
        seqs.append(s)
    return torch.cat(feats), torch.cat(seqs)


def ctc_collapse(ids: np.ndarray, blank: int, frames: Optional[Sequence[int]] = None) -> List[List[int]]:
    """helpers.py:20-32: iterate ALL frames of the utterance (no truncation at the encoded length).  `frames`: rows of a
    zero-padded batch are cut to the frames the utterance has on its own (`utterance_frames`) first."""
    out = []
    for i, row in enumerate(np.asarray(ids)):
        if frames is not None:
            row = row[: int(frames[i])]
        dec = []
        prev = blank
        for p in row:
            p = int(p)
            if (p != prev or prev == blank) and p != blank:
                dec.append(p)
            prev = p
        out.append(dec)
    return out


def ids_to_text(collapsed: List[List[int]], labels: Sequence[str]) -> List[str]:
    return ["".join(labels[c] for c in row) for row in collapsed]


# --------------------------------------------------------------------------
# whole path + helpers shared by tests / bench
# --------------------------------------------------------------------------
def full_path(enc_sd, dec_sd, jasper_cfg, wave: torch.Tensor, length: torch.Tensor):
    """wave [B,L] f32, length [B] i64 -> dict of every intermediate on the path."""
    with torch.no_grad():
        feats, seq = filterbank_features(wave, length)
        enc, enc_len = encoder_forward(enc_sd, jasper_cfg, feats, seq)
        logits = decoder_logits(dec_sd, enc)
        logp = F.log_softmax(logits, dim=-1)
        ids = greedy_argmax(logp)
    return {"feats": feats, "seq": seq, "enc": enc, "enc_len": enc_len,
            "logits": logits, "logp": logp, "ids": ids}


def quartznet_cfg(name: str) -> Tuple[List[dict], int]:
    """Block lists restating configs/quartznet12x1_vi.yaml:25-162 and
    configs/quartznet15x5.yaml:33-197 -> (jasper list, number of labels)."""
    def blk(filters, repeat, kernel, stride=1, dilation=1, residual=True, separable=True):
        return {"filters": filters, "repeat": repeat, "kernel": [kernel], "stride": [stride],
                "dilation": [dilation], "dropout": 0.0, "residual": residual, "separable": separable}
    if name == "quartznet12x1_vi":
        ks = [33] * 3 + [39] * 3 + [51] * 3 + [63] * 3 + [75]
        fs = [256] * 6 + [512] * 7
        blocks = [blk(256, 1, 33, stride=2, residual=False)]
        blocks += [blk(f, 1, k) for f, k in zip(fs, ks)]
        blocks += [blk(1024, 1, 1, residual=False, separable=False)]
        return blocks, 90
    if name == "quartznet15x5":
        ks = [33] * 3 + [39] * 3 + [51] * 3 + [63] * 3 + [75] * 3
        fs = [256] * 6 + [512] * 9
        blocks = [blk(256, 1, 33, stride=2, residual=False)]
        blocks += [blk(f, 5, k) for f, k in zip(fs, ks)]
        blocks += [blk(512, 1, 87, dilation=2, residual=False)]
        blocks += [blk(1024, 1, 1, residual=False, separable=False)]
        return blocks, 28
    raise KeyError(name)


def random_state_dicts(jasper_cfg: Sequence[dict], feat_in: int, num_classes: int, seed: int):
    """Seeded random weights with the checkpoint key layout (used when the shipped
    checkpoints are not on the box).  BN statistics are non-trivial so that a
    wrong fold shows up."""
    g = torch.Generator().manual_seed(seed)
    enc: Dict[str, torch.Tensor] = {}

    def bn(prefix, c):
        enc[prefix + ".weight"] = 0.5 + torch.rand(c, generator=g)
        enc[prefix + ".bias"] = 0.2 * torch.randn(c, generator=g)
        enc[prefix + ".running_mean"] = 0.2 * torch.randn(c, generator=g)
        enc[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=g)
        enc[prefix + ".num_batches_tracked"] = torch.tensor(1)

    cin = feat_in
    for b, cfg in enumerate(jasper_cfg):
        cout = int(cfg["filters"]); k = int(cfg["kernel"][0]); rep = int(cfg["repeat"])
        sep = bool(cfg.get("separable", False))
        per = 5 if sep else 4
        c = cin
        for r in range(rep):
            if sep:
                enc[f"encoder.{b}.mconv.{per*r}.conv.weight"] = torch.randn(c, 1, k, generator=g) / math.sqrt(k)
                enc[f"encoder.{b}.mconv.{per*r+1}.conv.weight"] = torch.randn(cout, c, 1, generator=g) / math.sqrt(c)
                bn(f"encoder.{b}.mconv.{per*r+2}", cout)
            else:
                enc[f"encoder.{b}.mconv.{per*r}.conv.weight"] = torch.randn(cout, c, k, generator=g) / math.sqrt(c * k)
                bn(f"encoder.{b}.mconv.{per*r+1}", cout)
            c = cout
        if cfg["residual"]:
            enc[f"encoder.{b}.res.0.0.conv.weight"] = torch.randn(cout, cin, 1, generator=g) / math.sqrt(cin)
            bn(f"encoder.{b}.res.0.1", cout)
        cin = cout
    dec = {
        "decoder_layers.0.weight": torch.randn(num_classes + 1, cin, 1, generator=g) / math.sqrt(cin),
        "decoder_layers.0.bias": 0.1 * torch.randn(num_classes + 1, generator=g),
    }
    return enc, dec
