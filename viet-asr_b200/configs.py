"""Model definitions of the two shipped QuartzNet variants and a YAML loader.

The block lists restate configs/quartznet12x1_vi.yaml:25-162 and
configs/quartznet15x5.yaml:33-197 of the reference; `load_model_definition`
reads a user-supplied YAML with the same sections `infer.py:85-111` consumes
(and maps the legacy `AudioPreprocessing` section of quartznet15x5.yaml onto
AudioToMelSpectrogramPreprocessor kwargs).
"""
from __future__ import annotations

from typing import Dict, List

VI_LABELS = list(" abcdeghiklmnopqrstuvxyàáâãèéêìíòóôõùúýăđĩũơưạảấầẩẫậắằẳẵặẹẻẽếềểễệỉịọỏốồổỗộớờởỡợụủứừửữựỳỵỷỹ")
EN_LABELS = list(" abcdefghijklmnopqrstuvwxyz'")

PREPROCESSOR_DEFAULT = dict(sample_rate=16000, window_size=0.02, window_stride=0.01, window="hann",
                            normalize="per_feature", n_fft=512, features=64, dither=0, pad_to=0, stft_conv=False)


def _blk(filters, repeat, kernel, stride=1, dilation=1, residual=True, separable=True) -> dict:
    return {"filters": filters, "repeat": repeat, "kernel": [kernel], "stride": [stride], "dilation": [dilation],
            "dropout": 0.0, "residual": residual, "separable": separable}


def quartznet12x1_vi() -> Dict:
    ks = [33] * 3 + [39] * 3 + [51] * 3 + [63] * 3 + [75]
    fs = [256] * 6 + [512] * 7
    blocks = [_blk(256, 1, 33, stride=2, residual=False)]
    blocks += [_blk(f, 1, k) for f, k in zip(fs, ks)]
    blocks += [_blk(1024, 1, 1, residual=False, separable=False)]
    return {"AudioToMelSpectrogramPreprocessor": dict(PREPROCESSOR_DEFAULT),
            "JasperEncoder": {"activation": "relu", "conv_mask": True, "jasper": blocks},
            "labels": list(VI_LABELS)}


def quartznet15x5() -> Dict:
    ks = [33] * 3 + [39] * 3 + [51] * 3 + [63] * 3 + [75] * 3
    fs = [256] * 6 + [512] * 9
    blocks = [_blk(256, 1, 33, stride=2, residual=False)]
    blocks += [_blk(f, 5, k) for f, k in zip(fs, ks)]
    blocks += [_blk(512, 1, 87, dilation=2, residual=False)]
    blocks += [_blk(1024, 1, 1, residual=False, separable=False)]
    return {"AudioToMelSpectrogramPreprocessor": dict(PREPROCESSOR_DEFAULT),
            "JasperEncoder": {"activation": "relu", "conv_mask": True, "jasper": blocks},
            "labels": list(EN_LABELS)}


MODELS = {"quartznet12x1_vi": quartznet12x1_vi, "quartznet15x5": quartznet15x5}


def load_model_definition(path: str) -> Dict:
    """YAML -> the dict `infer.py` builds its modules from (infer.py:85-111)."""
    import yaml
    with open(path, encoding="utf-8") as f:
        d = yaml.safe_load(f)
    if "AudioToMelSpectrogramPreprocessor" not in d and "AudioPreprocessing" in d:
        pre = dict(d["AudioPreprocessing"])
        pre.pop("feat_type", None)
        pre.setdefault("sample_rate", d.get("sample_rate", 16000))
        d["AudioToMelSpectrogramPreprocessor"] = pre
    pre = d["AudioToMelSpectrogramPreprocessor"]
    pre["dither"] = 0      # infer.py:89
    pre["pad_to"] = 0      # infer.py:90
    return d
