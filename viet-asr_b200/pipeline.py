"""Batched greedy-CTC inference engine: the data-parallel hot path end to end.

`VietASR` keeps the reference class's constructor/`transcribe` contract
(infer.py:57-171) but runs the greedy wiring (infer.py:113) on the GPU and
accepts batches.  Two execution routes share the same kernels:

  * `transcribe_batch_device` - module by module through the neural-module API
    (device tensors in, device tensors out); used by parity tests and the
    device-resident throughput measurement;
  * `transcribe_host` - one C-ABI call `vasr_transcribe_host` with HOST buffers
    (H2D of the waveforms and D2H of the collapsed ids inside the call); the
    end-to-end number of bench.py.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib, asr, audio, configs
from .nm import DeviceType, NeuralModuleFactory


class GraphedGreedyPath:
    """The fixed-shape greedy path (front end -> encoder -> decoder -> collapse) for batches of exactly [B, L], captured
    ONCE as a CUDA graph and replayed per call.  The reference re-sorts the module DAG and
    re-applies `.eval()` on every `infer` call (nemo/backends/pytorch/actions.py:1437, 414); here the call chain is
    resolved at capture time - static device buffers, cached tensor maps and descriptor tables, the library's kernels
    recorded in launch order - so a call is: H2D of the waveform, one graph launch, D2H of the collapsed ids."""

    def __init__(self, eng: "VietASR", B: int, L: int):
        import ctypes
        self.eng, self.B, self.L = eng, int(B), int(L)
        dev = torch.device("cuda", torch.cuda.current_device())
        pre, enc, dec = eng.preprocessor, eng.encoder, eng.decoder
        h = enc._sync_weights()
        lib = h.lib
        T_f = pre.num_frames(self.L)
        T_e = enc.out_frames(T_f)
        f32, i64, i32 = torch.float32, torch.int64, torch.int32
        self.wave = torch.zeros((B, L), dtype=f32, device=dev)
        self.length = torch.full((B,), L, dtype=i64, device=dev)
        self.feat = torch.empty((B, T_f, pre.nfilt), dtype=f32, device=dev)
        self.seq = torch.empty((B,), dtype=i64, device=dev)
        self.enc = torch.empty((B, T_e, enc._out_ch), dtype=f32, device=dev)
        self.enc_len = torch.empty((B,), dtype=f32, device=dev)
        self.ids = torch.empty((B, T_e), dtype=i64, device=dev)
        self.frames = torch.empty((B,), dtype=i32, device=dev)
        self.out_ids = torch.empty((B, T_e), dtype=i32, device=dev)
        self.out_len = torch.empty((B,), dtype=i32, device=dev)
        self.ws = torch.empty((int(lib.vasr_encoder_workspace_bytes(h.h, B, T_f)),), dtype=torch.uint8, device=dev)
        self.host_ids = torch.empty((B, T_e), dtype=i32).pin_memory()
        self.host_len = torch.empty((B,), dtype=i32).pin_memory()
        blank = len(eng.labels)

        def chain():
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(lib.vasr_frontend_forward(pre._h, self.wave.data_ptr(), self.length.data_ptr(), B, L,
                                                 self.feat.data_ptr(), self.seq.data_ptr(), st))
            _lib.check(lib.vasr_encoder_forward(h.h, self.feat.data_ptr(), self.seq.data_ptr(), B, T_f, self.enc.data_ptr(),
                                                self.enc_len.data_ptr(), self.ws.data_ptr(), self.ws.numel(), st))
            _lib.check(lib.vasr_decoder_forward(h.h, self.enc.data_ptr(), B, T_e, None, self.ids.data_ptr(), st))
            self.frames.copy_(eng.utterance_frames(self.length))
            _lib.check(lib.vasr_ctc_collapse(self.ids.data_ptr(), self.frames.data_ptr(), B, T_e, blank,
                                             self.out_ids.data_ptr(), self.out_len.data_ptr(), st))

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up: every cache the chain needs exists before capture
            chain(); chain()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            chain()
        self._h, self._lib_ = h, lib

    def __call__(self, wave_host: torch.Tensor, length_host: torch.Tensor):
        """wave_host [B, L] f32, length_host [B] i64 (pinned host tensors) -> (ids [B, T_e] i32, len [B] i32) pinned."""
        if tuple(wave_host.shape) != (self.B, self.L):
            raise ValueError(f"this graph was captured for [{self.B}, {self.L}] waveforms, got {tuple(wave_host.shape)}")
        self.wave.copy_(wave_host, non_blocking=True)
        self.length.copy_(length_host, non_blocking=True)
        self.graph.replay()
        self.host_ids.copy_(self.out_ids, non_blocking=True)
        self.host_len.copy_(self.out_len, non_blocking=True)
        _lib.check(self._lib_.vasr_encoder_check(self._h.h, self.ws.data_ptr(), self.ws.numel(), self.B,
                                                 torch.cuda.current_stream().cuda_stream))       # synchronises
        return self.host_ids, self.host_len


class VietASR:
    def __init__(self, config_file: Optional[str] = None, encoder_checkpoint: Optional[str] = None,
                 decoder_checkpoint: Optional[str] = None, device: str = "gpu", lm_path: Optional[str] = None,
                 beam_width: int = 20, lm_alpha: float = 0.5, lm_beta: float = 1.5, *,
                 model_definition: Optional[Dict] = None, gemm_mode: str = "f16x3", decoder: str = "beam",
                 batch_invariant: bool = False):
        """`batch_invariant=True`: zero-padded batches are processed so that every utterance gets exactly the result it
        gets alone (the only way the reference ever runs, infer.py:167-171): the STFT reflects at each utterance's own
        length (`set_padding`) in addition to the per-utterance decoding frames that are always on.  Default False =
        the reference's tensor semantics for a [B, L] batch (features.py:181-188)."""
        if device != "gpu" or not torch.cuda.is_available():
            raise RuntimeError("vasr_b200.VietASR runs on a CUDA device only (device='gpu'); there is no CPU path")
        if lm_path is not None and not os.path.exists(lm_path):
            raise FileNotFoundError(f"language model {lm_path!r} does not exist")
        if decoder not in ("beam", "greedy"):
            raise ValueError(f"decoder must be 'beam' or 'greedy', got {decoder!r}")
        self.decoder_kind = decoder
        if model_definition is None:
            if config_file is None:
                raise ValueError("either config_file or model_definition is required")
            model_definition = configs.load_model_definition(config_file)
        md = model_definition
        pre = dict(md["AudioToMelSpectrogramPreprocessor"])
        pre["dither"] = 0; pre["pad_to"] = 0                      # infer.py:89-90
        if NeuralModuleFactory.get_default_factory() is None:
            NeuralModuleFactory(placement=DeviceType.GPU)
        self.labels: List[str] = list(md["labels"])
        self.sample_rate = pre["sample_rate"]
        self.preprocessor = asr.AudioToMelSpectrogramPreprocessor(**pre)
        self.batch_invariant = bool(batch_invariant)
        if self.batch_invariant:
            self.preprocessor.set_padding(True)
        self.encoder = asr.JasperEncoder(feat_in=pre["features"], gemm_mode=gemm_mode, **md["JasperEncoder"])
        self.decoder = asr.JasperDecoderForCTC(feat_in=md["JasperEncoder"]["jasper"][-1]["filters"],
                                               num_classes=len(self.labels))
        self.greedy = asr.GreedyCTCDecoder()
        self.beam = asr.BeamSearchDecoderWithLM(lm_path=lm_path, vocab=self.labels, beam_width=beam_width,
                                                alpha=lm_alpha, beta=lm_beta, num_cpus=1)
        self.encoder.attach_decoder(self.decoder)
        if encoder_checkpoint:
            self.encoder.restore_from(encoder_checkpoint)
        if decoder_checkpoint:
            self.decoder.restore_from(decoder_checkpoint)

    # ---- weights
    def load_state_dicts(self, enc_sd, dec_sd):
        self.encoder.load_state_dict(enc_sd)
        self.decoder.load_state_dict(dec_sd)

    def set_gemm_mode(self, mode: str):
        self.encoder.set_gemm_mode(mode)

    # ---- device route (module by module)
    def utterance_frames(self, length: torch.Tensor) -> torch.Tensor:
        """Encoder frames every utterance of a zero-padded batch would have on its own: T_f = 1 + L // hop
        (features.py:245-301 with pad_to = 0), then the encoder's stride arithmetic.  The decoders run each utterance
        over exactly these frames, so a transcript does not depend on what else is in the batch."""
        return self.encoder.out_frames_of(torch.div(length.to(torch.int64), self.preprocessor.hop_length, rounding_mode="floor") + 1)

    @torch.no_grad()
    def forward_device(self, wave: torch.Tensor, length: torch.Tensor, want_log_probs: bool = False):
        feat, seq = self.preprocessor.forward_channels_last(wave, length)
        enc, enc_len = self.encoder.forward_channels_last(feat, seq)
        logp, ids = self.decoder.forward_channels_last(enc, want_log_probs)
        frames = self.utterance_frames(length)
        out_ids, out_len = asr.ctc_collapse(ids, len(self.labels), frames=frames)
        return {"feat": feat, "seq": seq, "enc": enc, "enc_len": enc_len, "log_probs": logp, "ids": ids,
                "out_ids": out_ids, "out_len": out_len, "frames": frames}

    def transcribe_batch_device(self, wave: torch.Tensor, length: torch.Tensor) -> List[str]:
        r = self.forward_device(wave, length)
        texts = asr.ids_to_text(r["out_ids"], r["out_len"], self.labels)      # (device -> host: synchronises)
        self.encoder.check_range(wave.shape[0])
        return texts

    def capture_graph(self, B: int, L: int) -> GraphedGreedyPath:
        """Capture the greedy path for batches of exactly [B, L] samples as a CUDA graph (cached per shape)."""
        cache = self.__dict__.setdefault("_graphs", {})
        key = (int(B), int(L), self.encoder._gemm_mode)
        if key not in cache:
            cache[key] = GraphedGreedyPath(self, B, L)
        return cache[key]

    # ---- host route (one C-ABI call, host buffers)
    def out_frames(self, L: int) -> int:
        return self.encoder.out_frames(self.preprocessor.num_frames(L))

    def transcribe_host_ids(self, wave_host: torch.Tensor, length_host: torch.Tensor,
                            out_ids: Optional[torch.Tensor] = None, out_len: Optional[torch.Tensor] = None):
        """wave_host [B, L] f32 / length_host [B] i64 CPU tensors (pinned for async copies)."""
        if wave_host.is_cuda or length_host.is_cuda:
            raise ValueError("transcribe_host_ids takes host tensors")
        w = wave_host.to(torch.float32).contiguous()
        ln = length_host.to(torch.int64).contiguous()
        B, L = w.shape
        T_e = self.out_frames(L)
        if out_ids is None:
            out_ids = torch.empty((B, T_e), dtype=torch.int32).pin_memory()
        if out_len is None:
            out_len = torch.empty((B,), dtype=torch.int32).pin_memory()
        h = self.encoder._sync_weights()
        _lib.check(h.lib.vasr_transcribe_host(self.preprocessor._h, h.h, w.data_ptr(), ln.data_ptr(), B, L,
                                              out_ids.data_ptr(), out_len.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream))
        return out_ids, out_len

    def transcribe_host_to_device(self, wave_host: torch.Tensor, length_host: torch.Tensor,
                                  out_ids: Optional[torch.Tensor] = None, out_len: Optional[torch.Tensor] = None):
        """Host waveforms in (pinned for asynchronous copies), collapsed ids left on the DEVICE, nothing synchronised:
        the copy / compute pipeline of `transcribe_host_ids` for callers that hand the result to a collective
        (`dist.gather_results`).  `check_range()` afterwards reports an fp16 overflow of the tensor-core modes."""
        if wave_host.is_cuda or length_host.is_cuda:
            raise ValueError("transcribe_host_to_device takes host tensors")
        w = wave_host.to(torch.float32).contiguous()
        ln = length_host.to(torch.int64).contiguous()
        B, L = w.shape
        T_e = self.out_frames(L)
        dev = torch.device("cuda", torch.cuda.current_device())
        if out_ids is None:
            out_ids = torch.empty((B, T_e), dtype=torch.int32, device=dev)
        if out_len is None:
            out_len = torch.empty((B,), dtype=torch.int32, device=dev)
        h = self.encoder._sync_weights()
        _lib.check(h.lib.vasr_transcribe_host_to_device(self.preprocessor._h, h.h, w.data_ptr(), ln.data_ptr(), B, L,
                                                        out_ids.data_ptr(), out_len.data_ptr(),
                                                        torch.cuda.current_stream().cuda_stream))
        return out_ids, out_len

    def check_range(self):
        h = self.encoder._sync_weights()
        _lib.check(h.lib.vasr_transcribe_check(h.h, torch.cuda.current_stream().cuda_stream))

    @torch.no_grad()
    def beam_batch_device(self, wave: torch.Tensor, length: torch.Tensor) -> List[str]:
        """Device tensors -> transcripts through the beam-search decoder (with the n-gram LM when the engine was
        built with `lm_path`), batched."""
        feat, seq = self.preprocessor.forward_channels_last(wave, length)
        enc, _ = self.encoder.forward_channels_last(feat, seq)
        logp, _ = self.decoder.forward_channels_last(enc, True)
        texts = self.beam.decode_batch(logp, frames=self.utterance_frames(length))
        self.encoder.check_range(wave.shape[0])
        return texts

    def transcribe_batch(self, signals: Sequence[np.ndarray], decoder: Optional[str] = None) -> List[str]:
        """List of 1-D float waveforms (16 kHz) -> transcripts; zero-pads to the longest
        (the `seq_collate_fn` convention, parts/dataset.py:14-53).  `decoder`: 'greedy' or 'beam'
        (default: the engine's, i.e. beam search - fused with the KenLM model when `lm_path` was given - like the
        reference's `transcribe`)."""
        kind = decoder or self.decoder_kind
        lens = torch.tensor([len(s) for s in signals], dtype=torch.int64)
        L = int(lens.max())
        w = torch.zeros((len(signals), L), dtype=torch.float32)
        for i, s in enumerate(signals):
            w[i, : len(s)] = torch.as_tensor(np.asarray(s), dtype=torch.float32)
        if kind == "greedy":
            ids, n = self.transcribe_host_ids(w.pin_memory(), lens.pin_memory())
            return asr.ids_to_text(ids, n, self.labels)
        return self.beam_batch_device(w.cuda(non_blocking=True), lens.cuda(non_blocking=True))

    # ---- batched, length-aware data path in front of the kernels
    @torch.no_grad()
    def transcribe_signals(self, signals: Sequence[np.ndarray], sample_rates, decoder: Optional[str] = None,
                           batch_size: int = 256, max_padded_seconds: float = 2560.0,
                           max_duration: Optional[float] = None) -> List[Optional[str]]:
        """Mono signals (int16 PCM or float32) at their native sample rates -> transcripts, in input order.

        What `infer.py:196-206` / `app.py:58-91` do one file at a time - `librosa.load(path, sr=16000)` then
        `transcribe` - as a batched device pipeline: utterances are bucketed by length (per sample rate), zero-padded
        (`seq_collate_fn` layout), copied once, converted / resampled to the model rate on the GPU and decoded.
        `max_duration` mirrors the CLI's skip of clips longer than 10 s (infer.py:201-203): those come back as None."""
        kind = decoder or self.decoder_kind
        if isinstance(sample_rates, int):
            sample_rates = [sample_rates] * len(signals)
        if len(sample_rates) != len(signals):
            raise ValueError("transcribe_signals: one sample rate per signal")
        out: List[Optional[str]] = [None] * len(signals)
        if not hasattr(self, "_resampler"):
            self._resampler = audio.Resampler()
        by_sr: Dict[int, List[int]] = {}
        for i, (s, sr) in enumerate(zip(signals, sample_rates)):
            n = int(np.asarray(s).shape[0])
            if n == 0 or (max_duration is not None and n / float(sr) > max_duration):
                continue
            by_sr.setdefault(int(sr), []).append(i)
        for sr, idxs in by_sr.items():
            groups = {}
            for i in idxs:                                   # int16 and float signals are collated separately
                groups.setdefault(np.asarray(signals[i]).dtype == np.int16, []).append(i)
            for _, gi in groups.items():
                lens = [int(np.asarray(signals[i]).shape[0]) for i in gi]
                for batch in audio.plan_batches(lens, batch_size, int(max_padded_seconds * sr)):
                    ids = [gi[j] for j in batch]
                    w, ln = audio.collate([signals[i] for i in ids])
                    w, ln = w.cuda(non_blocking=True), ln.cuda(non_blocking=True)
                    w, ln = self._resampler(w, ln, sr, self.sample_rate)
                    if kind == "greedy":
                        texts = self.transcribe_batch_device(w, ln)
                    else:
                        texts = self.beam_batch_device(w, ln)
                    for i, t in zip(ids, texts):
                        out[i] = t
        return out

    def transcribe_files(self, paths: Sequence[str], **kwargs) -> List[Optional[str]]:
        """WAV files -> transcripts (see `transcribe_signals`)."""
        sigs, srs = [], []
        for p in paths:
            a, sr = audio.read_wav(p)
            sigs.append(a); srs.append(sr)
        return self.transcribe_signals(sigs, srs, **kwargs)

    def transcribe(self, audio_signal: np.ndarray) -> str:
        """infer.py:167-171: one utterance -> text (beam search, LM-fused when `lm_path` was given; greedy when the
        engine was built with decoder='greedy')."""
        return self.transcribe_batch([np.reshape(audio_signal, [-1])])[0]
