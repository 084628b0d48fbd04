"""Drop-in boundary against the reference's OWN base classes.  Build-container only: needs
/root/reference (skipped on the GPU box).  A subprocess imports the vendored NeMo 0.10 core unmodified
(`nemo.core`, `nemo.backends.pytorch.nm`; the two third-party imports that are absent here - wget, ruamel.yaml - are
stubbed and the `np.int/np.float/np.str` aliases NumPy 2 removed are restored), re-bases the classes of
`viet_asr_b200.asr` on `nemo.backends.pytorch.nm.TrainableNM / NonTrainableNM` and the reference's neural types
(VASR_NM_BACKEND=nemo, the one-import swap of INTEGRATION.md section 1) and runs the symbolic wiring of infer.py:96-160
under the reference's `nemo.core.NeuralModuleFactory`: constructor kwargs splatted from the model definition, port
names, neural-type checks, `restore_from` key layout.  No GPU, no compute."""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT, have_weights

REF = "/root/reference"

SCRIPT = textwrap.dedent(r'''
    import os, sys, types, warnings
    warnings.simplefilter("ignore")
    import numpy as np
    for n, t in (("int", int), ("float", float), ("str", str), ("bool", bool), ("object", object)):
        if not hasattr(np, n):
            setattr(np, n, t)
    def stub(name, **attrs):
        m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m; return m
    stub("wget")
    import yaml as _y
    class YAML:
        def __init__(self, typ=None): pass
        def load(self, f): return _y.safe_load(f)
        def dump(self, d, f): return _y.safe_dump(d, f)
    stub("ruamel").yaml = stub("ruamel.yaml", YAML=YAML)
    sys.path.insert(0, REF); sys.path.insert(0, ROOT)
    os.environ["VASR_NM_BACKEND"] = "nemo"
    import torch
    import nemo
    from nemo.backends.pytorch.nm import DataLayerNM, NonTrainableNM, TrainableNM
    from nemo.core.neural_types import AudioSignal, LengthsType, NeuralType, NmTensor, NeuralPortNmTensorMismatchError
    import viet_asr_b200 as V
    from viet_asr_b200 import asr
    # the drop-in classes now derive from the reference's base classes, not from this repo's mirror
    assert issubclass(asr.JasperEncoder, TrainableNM) and issubclass(asr.JasperDecoderForCTC, TrainableNM)
    assert issubclass(asr.GreedyCTCDecoder, TrainableNM)
    assert issubclass(asr.AudioToMelSpectrogramPreprocessor, NonTrainableNM) and issubclass(asr.BeamSearchDecoderWithLM, NonTrainableNM)
    assert asr.JasperEncoder.__mro__[1].__module__ == "nemo.backends.pytorch.nm"

    with open(os.path.join(REF, "configs/quartznet12x1_vi.yaml"), encoding="utf-8") as f:
        model_definition = _y.safe_load(f)                                     # infer.py:85-90
    model_definition["AudioToMelSpectrogramPreprocessor"]["dither"] = 0
    model_definition["AudioToMelSpectrogramPreprocessor"]["pad_to"] = 0
    neural_factory = nemo.core.NeuralModuleFactory(placement=nemo.core.DeviceType.CPU)   # infer.py:96 (no GPU here)

    class AudioDataLayer(DataLayerNM):                                         # infer.py:16-54, ports only
        @property
        def output_ports(self):
            return {"audio_signal": NeuralType(("B", "T"), AudioSignal(freq=self._sample_rate)),
                    "a_sig_length": NeuralType(tuple("B"), LengthsType())}
        def __init__(self, sample_rate):
            super().__init__()
            self._sample_rate = sample_rate
        def __len__(self): return 1
        @property
        def dataset(self): return None
        @property
        def data_iterator(self): return iter(())

    data_layer = AudioDataLayer(sample_rate=model_definition["AudioToMelSpectrogramPreprocessor"]["sample_rate"])
    data_preprocessor = asr.AudioToMelSpectrogramPreprocessor(**model_definition["AudioToMelSpectrogramPreprocessor"])
    jasper_encoder = asr.JasperEncoder(feat_in=model_definition["AudioToMelSpectrogramPreprocessor"]["features"],
                                       **model_definition["JasperEncoder"])
    jasper_decoder = asr.JasperDecoderForCTC(feat_in=model_definition["JasperEncoder"]["jasper"][-1]["filters"],
                                             num_classes=len(model_definition["labels"]))
    greedy_decoder = asr.GreedyCTCDecoder()
    beamsearch_decoder = asr.BeamSearchDecoderWithLM(vocab=model_definition["labels"], beam_width=20, alpha=0.5, beta=1.5,
                                                     lm_path=None, num_cpus=max(1, os.cpu_count()))
    if WEIGHTS:                                                                # infer.py:143-144
        jasper_encoder.restore_from(os.path.join(ROOT, "weights/vi12x1/JasperEncoder.pt"))
        jasper_decoder.restore_from(os.path.join(ROOT, "weights/vi12x1/JasperDecoderForCTC.pt"))
    audio_signal, audio_signal_len = data_layer()                              # infer.py:147-160
    processed_signal, processed_signal_len = data_preprocessor(input_signal=audio_signal, length=audio_signal_len)
    encoded, encoded_len = jasper_encoder(audio_signal=processed_signal, length=processed_signal_len)
    log_probs = jasper_decoder(encoder_output=encoded)
    beam_predictions = beamsearch_decoder(log_probs=log_probs, log_probs_length=encoded_len)
    greedy_predictions = greedy_decoder(log_probs=log_probs)
    for t in (processed_signal, encoded, log_probs, beam_predictions, greedy_predictions):
        assert isinstance(t, NmTensor), type(t)
    assert beam_predictions.producer is beamsearch_decoder and log_probs.producer is jasper_decoder
    try:
        jasper_decoder(encoder_output=processed_signal)                        # a spectrogram into the encoded port
        raise SystemExit("type check did not fire")
    except NeuralPortNmTensorMismatchError:
        pass
    # what PtActions does before executing (actions.py:414-415, 428): eval(), then module(force_pt=True, **tensors) -
    # on CPU tensors the drop-in module must refuse loudly (no CPU path), i.e. the call reaches forward()
    jasper_decoder.eval()
    try:
        jasper_decoder(force_pt=True, encoder_output=torch.zeros(1, 1024, 4))
        raise SystemExit("CPU tensors were accepted")
    except RuntimeError as e:
        assert "CUDA" in str(e) or "cuda" in str(e), e
    print("BOUNDARY-OK", nemo.__version__, type(neural_factory).__module__)
''')


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "nemo")), reason="needs the reference checkout (build container only)")
def test_drop_in_modules_on_the_reference_base_classes():
    code = f"REF = {REF!r}\nROOT = {ROOT!r}\nWEIGHTS = {have_weights('vi12x1')!r}\n" + SCRIPT
    env = dict(os.environ)
    env.pop("VASR_NM_BACKEND", None)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd="/tmp")
    assert r.returncode == 0 and "BOUNDARY-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
