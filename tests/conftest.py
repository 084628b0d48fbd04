import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
WEIGHTS = os.path.join(ROOT, "weights")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def have_weights(tag):
    return os.path.exists(os.path.join(WEIGHTS, tag, "JasperEncoder.pt"))


def load_weights(tag):
    enc = torch.load(os.path.join(WEIGHTS, tag, "JasperEncoder.pt"), map_location="cpu")
    dec = torch.load(os.path.join(WEIGHTS, tag, "JasperDecoderForCTC.pt"), map_location="cpu")
    return enc, dec


MODEL_OF = {"vi12x1": "quartznet12x1_vi", "en15x5": "quartznet15x5"}
RAND_SEED = 20260925  # oracle/make_golden.py


def model_and_weights(tag, kind):
    """(model_definition, enc_sd, dec_sd) for golden case `<tag>_<kind>`; skips if real weights are absent."""
    import viet_asr_b200 as V
    from oracle import quartznet_oracle as O
    md = V.configs.MODELS[MODEL_OF[tag]]()
    jasper = md["JasperEncoder"]["jasper"]
    if kind == "rand":
        enc, dec = O.random_state_dicts(jasper, 64, len(md["labels"]), seed=RAND_SEED)
    else:
        if not have_weights(tag):
            pytest.skip(f"weights/{tag} not present (run oracle/make_golden.py where /root/reference exists)")
        enc, dec = load_weights(tag)
    return md, enc, dec


def pcm_to_wave(pcm16):
    return torch.from_numpy(pcm16.astype(np.float32) / 32768.0)
