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


# ----------------------------------------------------------------------------- language-model fixtures
def tiny_lm_path(order):
    """tests/golden/tiny_lm_{3,5}gram.binary: small KenLM-format files written by oracle/kenlm_writer.py."""
    return os.path.join(GOLDEN, f"tiny_lm_{order}gram.binary")


def shipped_lm_path(name="3-gram-lm.binary"):
    """The reference's own KenLM binary, copied to weights/lm by oracle/make_golden.py; skips when absent."""
    p = os.path.join(WEIGHTS, "lm", name)
    if not os.path.exists(p):
        pytest.skip(f"weights/lm/{name} not present (run oracle/make_golden.py --lm-only where /root/reference exists)")
    return p


def spelled_posteriors(sentences, labels, seed, peak=5.0, noise=1.5, confusions=()):
    """Synthetic CTC log-posteriors [B, T, V+1] that spell `sentences` (blank = last class) with random repeats,
    blanks and logit noise, so a beam search has real alternatives to weigh.  `confusions`: (utterance, char
    position, other char) -> that frame group gets a second peak almost as high as the spelled one."""
    g = np.random.default_rng(seed)
    lab = {c: i for i, c in enumerate(labels)}
    blank = len(labels)
    seqs = []
    for u, sent in enumerate(sentences):
        frames = []
        prev = None
        for pos, ch in enumerate(sent):
            c = lab[ch]
            if prev == c:
                frames.append((blank, None))
            alt = [o for (uu, pp, o) in confusions if uu == u and pp == pos]
            for _ in range(int(g.integers(1, 3))):
                frames.append((c, lab[alt[0]] if alt else None))
            for _ in range(int(g.integers(0, 2))):
                frames.append((blank, None))
            prev = c
        seqs.append(frames)
    T = max(len(f) for f in seqs) + 2
    x = noise * g.standard_normal((len(sentences), T, blank + 1)).astype(np.float32)
    for u, frames in enumerate(seqs):
        for t in range(T):
            c, alt = frames[t] if t < len(frames) else (blank, None)
            x[u, t, c] += peak
            if alt is not None:
                x[u, t, alt] += peak - 0.3
    return torch.from_numpy(x).log_softmax(-1)
