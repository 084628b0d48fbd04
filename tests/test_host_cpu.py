"""CPU: the C-ABI library loads and exports every declared symbol; the neural-module mirror builds and
type-checks the infer.py graph; configuration errors match the reference's behaviour."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT

import viet_asr_b200 as V
from viet_asr_b200 import _lib, nm


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "vasr_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vasr_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"libvasr_b200.so does not export {s}"
    assert sorted(_lib.PROTOTYPES) == syms, "python prototypes out of sync with include/vasr_b200.h"
    assert _lib.load().vasr_abi_version() == 2


def test_model_create_validation_no_gpu_needed():
    lib = _lib.load()
    h = ctypes.c_void_p()
    blocks = (_lib.BlockCfg * 1)(_lib.BlockCfg(256, 1, 33, 2, 2, 0, 1))
    rc = lib.vasr_model_create(blocks, 1, 64, 29, ctypes.byref(h))
    assert rc == _lib.VASR_EINVAL
    assert b"Only stride OR dilation may be greater than 1" in lib.vasr_last_error()   # parts/jasper.py:61-62
    with pytest.raises(ValueError):
        _lib.check(rc)
    blocks = (_lib.BlockCfg * 1)(_lib.BlockCfg(256, 1, 32, 1, 1, 0, 1))
    assert lib.vasr_model_create(blocks, 1, 64, 29, ctypes.byref(h)) == _lib.VASR_EINVAL
    blocks = (_lib.BlockCfg * 1)(_lib.BlockCfg(256, 1, 33, 1, 1, 0, 1))
    assert lib.vasr_model_create(blocks, 1, 64, 29, ctypes.byref(h)) == 0
    assert lib.vasr_model_out_frames(h, 501) == 501
    lib.vasr_model_destroy(h)


def test_out_frames_matches_reference_length_formula():
    lib = _lib.load()
    for name, Tf, Te in (("quartznet12x1_vi", 1001, 501), ("quartznet15x5", 501, 251), ("quartznet15x5", 437, 219)):
        md = V.configs.MODELS[name]()
        m = V.asr._ModelHandle(md["JasperEncoder"]["jasper"], 64, len(md["labels"]) + 1)
        assert lib.vasr_model_out_frames(m.h, Tf) == Te


def test_configs_match_oracle_restatement():
    from oracle import quartznet_oracle as O
    for name in V.configs.MODELS:
        md = V.configs.MODELS[name]()
        blocks, nlab = O.quartznet_cfg(name)
        assert md["JasperEncoder"]["jasper"] == blocks
        assert len(md["labels"]) == nlab


def test_preprocessor_ctor_errors_like_reference():
    nm.NeuralModuleFactory(placement=nm.DeviceType.GPU)
    with pytest.raises(ValueError, match="received both window_size and n_window_size"):
        V.AudioToMelSpectrogramPreprocessor(window_size=0.02, n_window_size=320)
    with pytest.raises(ValueError, match="log_zero_guard_type"):
        V.AudioToMelSpectrogramPreprocessor(dither=0, log_zero_guard_type="bogus")
    with pytest.raises(ValueError, match="dither"):
        V.AudioToMelSpectrogramPreprocessor()          # default dither 1e-5 is not the inference path
    with pytest.raises(ValueError, match="window"):
        V.AudioToMelSpectrogramPreprocessor(dither=0, window="none")
    pc = V.AudioToMelSpectrogramPreprocessor(dither=0, stft_conv=True)      # quartznet15x5.yaml:26
    from scipy.signal import get_window
    np.testing.assert_allclose(pc._window.numpy(), get_window("hann", 320, fftbins=True), atol=1e-6)   # periodic
    ps = V.AudioToMelSpectrogramPreprocessor(dither=0, window="hamming")
    np.testing.assert_allclose(ps._window.numpy(), torch.hamming_window(320, periodic=False).numpy(), atol=0)
    p = V.AudioToMelSpectrogramPreprocessor(dither=0, pad_to=0, n_fft=512)
    assert p.num_frames(80000) == 501 and p.num_frames(69813) == 437
    p16 = V.AudioToMelSpectrogramPreprocessor(dither=0, pad_to=16, n_fft=512)
    assert p16.num_frames(80000) == 512
    assert p.get_seq_len(torch.tensor([80000, 69813])).tolist() == [500, 437]


def test_modules_fail_loudly_without_cuda():
    nm.NeuralModuleFactory(placement=nm.DeviceType.GPU)
    p = V.AudioToMelSpectrogramPreprocessor(dither=0, pad_to=0, n_fft=512)
    with pytest.raises(RuntimeError, match="no CPU path"):
        p.forward(input_signal=torch.zeros(1, 1000), length=torch.tensor([1000]))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA device only"):
            V.VietASR(model_definition=V.configs.quartznet12x1_vi())


def test_symbolic_graph_like_infer_py():
    """infer.py:99-160: data layer -> preprocessor -> encoder -> decoder -> greedy, with port type checks."""
    nm.NeuralModuleFactory(placement=nm.DeviceType.GPU)
    md = V.configs.quartznet12x1_vi()

    class AudioDataLayer(nm.DataLayerNM):
        @property
        def output_ports(self):
            return {"audio_signal": nm.NeuralType(("B", "T"), nm.AudioSignal(freq=16000)),
                    "a_sig_length": nm.NeuralType(tuple("B"), nm.LengthsType())}

        def __len__(self): return 1
        @property
        def dataset(self): return None
        @property
        def data_iterator(self): return iter(())

    dl = AudioDataLayer()
    pre = V.AudioToMelSpectrogramPreprocessor(**md["AudioToMelSpectrogramPreprocessor"])
    enc = V.JasperEncoder(feat_in=64, **md["JasperEncoder"])
    dec = V.JasperDecoderForCTC(feat_in=1024, num_classes=len(md["labels"]))
    greedy = V.GreedyCTCDecoder()
    sig, sig_len = dl()
    feat, feat_len = pre(input_signal=sig, length=sig_len)
    encoded, encoded_len = enc(audio_signal=feat, length=feat_len)
    logp = dec(encoder_output=encoded)
    pred = greedy(log_probs=logp)
    assert isinstance(pred, nm.NmTensor) and pred.producer is greedy
    with pytest.raises(nm.NeuralPortNameMismatchError):
        enc(audio=feat, length=feat_len)
    with pytest.raises(nm.NeuralPortNmTensorMismatchError):
        enc(audio_signal=logp, length=feat_len)            # log-probs into a spectrogram port
    with pytest.raises(nm.NeuralPortNmTensorMismatchError):
        dec(encoder_output=feat)
    # state-dict key layout of the shipped checkpoints
    keys = set(enc.state_dict().keys())
    assert "encoder.0.mconv.0.conv.weight" in keys and "encoder.0.mconv.2.running_var" in keys
    assert "encoder.1.res.0.0.conv.weight" in keys and "encoder.14.mconv.1.running_mean" in keys
    assert enc.state_dict()["encoder.0.mconv.0.conv.weight"].shape == (64, 1, 33)
    assert set(dec.state_dict().keys()) == {"decoder_layers.0.weight", "decoder_layers.0.bias"}
    md15 = V.configs.quartznet15x5()
    enc15 = V.JasperEncoder(feat_in=64, **md15["JasperEncoder"])
    k15 = set(enc15.state_dict().keys())
    assert "encoder.1.mconv.20.conv.weight" in k15 and "encoder.1.mconv.22.weight" in k15
    assert "encoder.17.mconv.0.conv.weight" in k15 and enc15.state_dict()["encoder.17.mconv.0.conv.weight"].shape == (1024, 512, 1)


def test_encoder_rejects_unbuilt_options():
    nm.NeuralModuleFactory(placement=nm.DeviceType.GPU)
    md = V.configs.quartznet12x1_vi()["JasperEncoder"]
    with pytest.raises(ValueError, match="activation"):
        V.JasperEncoder(feat_in=64, jasper=md["jasper"], activation="hardtanh")
    bad = [dict(md["jasper"][0], se=True)] + md["jasper"][1:]
    with pytest.raises(ValueError, match="se="):
        V.JasperEncoder(feat_in=64, jasper=bad, activation="relu")


def test_shipped_checkpoint_keys_load_when_present():
    from conftest import have_weights, load_weights
    if not have_weights("vi12x1"):
        pytest.skip("weights/vi12x1 not present")
    nm.NeuralModuleFactory(placement=nm.DeviceType.GPU)
    md = V.configs.quartznet12x1_vi()
    enc = V.JasperEncoder(feat_in=64, **md["JasperEncoder"])
    dec = V.JasperDecoderForCTC(feat_in=1024, num_classes=len(md["labels"]))
    e, d = load_weights("vi12x1")
    assert str(enc.load_state_dict(e)) == "<All keys matched successfully>"
    assert str(dec.load_state_dict(d)) == "<All keys matched successfully>"
