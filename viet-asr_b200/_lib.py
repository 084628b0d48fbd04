"""ctypes binding of libvasr_b200.so (the C ABI declared in include/vasr_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails
this module raises.  VASR_EINVAL maps to ValueError (the reference raises
ValueError for bad configuration, e.g. parts/features.py:145,217 and
parts/jasper.py:61-62), every other status to RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VASR_B200_LIB: developer override to time an experimental build of the same library (tools/); never a fallback
LIB_PATH = os.environ.get("VASR_B200_LIB") or os.path.join(_HERE, "libvasr_b200.so")

VASR_OK, VASR_EINVAL, VASR_ECUDA, VASR_ESTATE, VASR_ENOMEM, VASR_ERANGE = 0, -1, -2, -3, -4, -5
GEMM_FP32_SIMT, GEMM_F16X3, GEMM_F16X1 = 0, 1, 2
GEMM_MODES = {"fp32": GEMM_FP32_SIMT, "f16x3": GEMM_F16X3, "f16x1": GEMM_F16X1}


class BlockCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("filters", "repeat", "kernel", "stride", "dilation", "residual", "separable")]


class FrontendCfg(C.Structure):
    _fields_ = [("n_window_size", C.c_int32), ("n_window_stride", C.c_int32), ("n_fft", C.c_int32),
                ("nfilt", C.c_int32), ("preemph", C.c_float), ("log_zero_guard", C.c_float),
                ("pad_to", C.c_int32)]


class LmArrays(C.Structure):
    """vasr_lm_arrays (include/vasr_b200.h): host pointers to the decoded KenLM trie."""
    _fields_ = [("order", C.c_int32), ("vocab", C.c_int32), ("bos", C.c_int32), ("eos", C.c_int32),
                ("counts", C.c_void_p),
                ("uni_prob", C.c_void_p), ("uni_backoff", C.c_void_p), ("uni_next", C.c_void_p),
                ("mid_word", C.c_void_p * 3), ("mid_prob", C.c_void_p * 3), ("mid_backoff", C.c_void_p * 3),
                ("mid_next", C.c_void_p * 3),
                ("long_word", C.c_void_p), ("long_prob", C.c_void_p),
                ("vocab_keys", C.c_void_p), ("vocab_vals", C.c_void_p), ("vocab_slots", C.c_int32)]


# every symbol include/vasr_b200.h declares: (restype, argtypes)
_vp, _i, _i64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
PROTOTYPES = {
    "vasr_abi_version": (_i, []),
    "vasr_last_error": (C.c_char_p, []),
    "vasr_launch_count": (_i64, []),
    "vasr_frontend_create": (_i, [C.POINTER(FrontendCfg), _vp, _vp, C.POINTER(_vp)]),
    "vasr_frontend_destroy": (None, [_vp]),
    "vasr_frontend_set_padding": (_i, [_vp, _i]),
    "vasr_frontend_num_frames": (_i, [_vp, _i64]),
    "vasr_frontend_forward": (_i, [_vp, _vp, _vp, _i, _i64, _vp, _vp, _vp]),
    "vasr_model_create": (_i, [C.POINTER(BlockCfg), _i, _i, _i, C.POINTER(_vp)]),
    "vasr_model_destroy": (None, [_vp]),
    "vasr_model_load_tensor": (_i, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i, _i]),
    "vasr_model_finalize": (_i, [_vp, _i]),
    "vasr_model_gemm_mode": (_i, [_vp]),
    "vasr_model_out_frames": (_i, [_vp, _i]),
    "vasr_model_out_channels": (_i, [_vp]),
    "vasr_model_num_classes": (_i, [_vp]),
    "vasr_encoder_workspace_bytes": (_sz, [_vp, _i, _i]),
    "vasr_encoder_forward": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "vasr_encoder_check": (_i, [_vp, _vp, _sz, _i, _vp]),
    "vasr_decoder_forward": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "vasr_greedy_argmax": (_i, [_vp, _i, _i, _vp, _vp]),
    "vasr_ctc_collapse": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "vasr_ctc_beam_workspace_bytes": (_sz, [_i, _i]),
    "vasr_ctc_beam_search": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, C.c_float, C.c_float, _vp, _sz, _vp, _vp, _vp, _vp]),
    "vasr_lm_create": (_i, [C.POINTER(LmArrays), C.POINTER(_vp)]),
    "vasr_lm_destroy": (None, [_vp]),
    "vasr_lm_order": (_i, [_vp]),
    "vasr_lm_hash_labels": (C.c_uint64, [_vp, _i]),
    "vasr_lm_score_batch": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "vasr_ctc_beam_lm_workspace_bytes": (_sz, [_i, _i, _i]),
    "vasr_ctc_beam_search_lm": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, C.c_float, C.c_float, _vp, C.c_double, C.c_double,
                                     C.c_double, _vp, _sz, _vp, _vp, _vp, _vp]),
    "vasr_resampler_create": (_i, [_vp, _i, _i, C.POINTER(_vp)]),
    "vasr_resampler_destroy": (None, [_vp]),
    "vasr_resample_out_len": (_i64, [_i64, _i, _i]),
    "vasr_pcm16_to_float": (_i, [_vp, _vp, _i, _i64, _vp, _vp]),
    "vasr_resample": (_i, [_vp, _vp, _i, _vp, _i, _i64, _i, _i, _vp, _vp, _i64, _vp]),
    "vasr_transcribe_host": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _vp, _vp, _vp]),
    "vasr_transcribe_host_to_device": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _vp, _vp, _vp]),
    "vasr_transcribe_check": (_i, [_vp, _vp]),
}

_lib = None


def load():
    """Load (once) and return the ctypes library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C viet-asr_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.vasr_abi_version() != 2:
        raise RuntimeError("libvasr_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc == VASR_OK:
        return
    msg = load().vasr_last_error().decode("utf-8", "replace")
    if rc == VASR_EINVAL:
        raise ValueError(msg)
    raise RuntimeError(f"vasr_b200 error {rc}: {msg}")


def launch_count() -> int:
    return int(load().vasr_launch_count())
