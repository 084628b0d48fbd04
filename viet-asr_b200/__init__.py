"""viet-asr_b200: B200-native CTC inference hot path of dangvansam/viet-asr.

Import as `viet_asr_b200` (the repo root ships a tiny module of that name that
loads this directory, whose on-disk name `viet-asr_b200` is not an identifier).
"""
from . import _lib, nm, asr, configs, audio, metrics          # noqa: F401
from .asr import (AudioToMelSpectrogramPreprocessor, JasperEncoder, JasperDecoderForCTC,  # noqa: F401
                  GreedyCTCDecoder, BeamSearchDecoderWithLM, post_process_predictions, ctc_collapse,
                  ctc_beam_search, ids_to_text, NGramLM)
from .nm import (NeuralModuleFactory, DeviceType, NeuralType, NmTensor, DataLayerNM,       # noqa: F401
                 TrainableNM, NonTrainableNM, AudioSignal, LengthsType)
from .pipeline import VietASR, GraphedGreedyPath   # noqa: F401
from .audio import Resampler, AudioBatchLayer, read_wav, collate, plan_batches   # noqa: F401
from .metrics import word_error_rate, read_manifest, evaluate_manifest   # noqa: F401

__all__ = ["AudioToMelSpectrogramPreprocessor", "JasperEncoder", "JasperDecoderForCTC", "GreedyCTCDecoder", "BeamSearchDecoderWithLM", "ctc_beam_search", "NGramLM",
           "post_process_predictions", "ctc_collapse", "ids_to_text", "NeuralModuleFactory", "DeviceType",
           "NeuralType", "NmTensor", "DataLayerNM", "TrainableNM", "NonTrainableNM", "VietASR", "configs",
           "Resampler", "AudioBatchLayer", "read_wav", "collate", "plan_batches", "word_error_rate", "read_manifest",
           "evaluate_manifest", "GraphedGreedyPath"]
