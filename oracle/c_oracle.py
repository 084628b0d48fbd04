"""ctypes view of oracle/c/libctc_oracle.so (plain-C restatement of greedy argmax + CTC collapse) - TEST
INFRASTRUCTURE ONLY.  Built by `make -C oracle/c` (also run by __graft_entry__.build())."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def load():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "c", "libctc_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", os.path.join(_HERE, "c")], check=True, capture_output=True)
        _LIB = C.CDLL(path)
    return _LIB


def greedy_argmax(logp: np.ndarray) -> np.ndarray:
    logp = np.ascontiguousarray(logp, dtype=np.float32)
    B, T, V = logp.shape
    ids = np.empty((B, T), dtype=np.int64)
    load().oracle_greedy_argmax(logp.ctypes.data_as(C.c_void_p), B, T, V, ids.ctypes.data_as(C.c_void_p))
    return ids


def ctc_collapse(ids: np.ndarray, blank: int):
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    B, T = ids.shape
    out = np.empty((B, T), dtype=np.int32)
    n = np.empty((B,), dtype=np.int32)
    load().oracle_ctc_collapse(ids.ctypes.data_as(C.c_void_p), B, T, int(blank), out.ctypes.data_as(C.c_void_p),
                               n.ctypes.data_as(C.c_void_p))
    return out, n
