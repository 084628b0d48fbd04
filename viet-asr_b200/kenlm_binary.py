"""Loader for the KenLM binaries the reference ships (models/language_model/*-gram-lm.binary) - product side.

The reference hands `lm_path` to pyctcdecode, which opens it with the `kenlm` python module
(nemo/collections/asr/beam_search_decoder.py:82-87; README.md:43-45).  Here the file is decoded once on the host into
flat arrays (word ids, de-quantised log10 prob / back-off, child ranges per order) that `vasr_lm_create` uploads to
HBM; the beam-search kernel walks them directly (csrc/beam.cu).  Only the format the shipped files use is accepted:
format version 5, model type QUANT_ARRAY_TRIE, ArrayBhiksha with 0 chopped bits, vocabulary strings included.
File layout: see the header of oracle/kenlm_oracle.py (an independent scalar reader used by the tests).
"""
from __future__ import annotations

import struct
from typing import Dict, List, Sequence

import numpy as np

_MAGIC = b"mmap lm http://kheafield.com/code format version 5\n\x00"


def _unpack(raw: np.ndarray, base: int, total_bits: int, n: int) -> np.ndarray:
    """n little-endian bit-packed records of total_bits (<= 57) starting at byte `base` -> uint64 each."""
    bit = np.arange(n, dtype=np.uint64) * np.uint64(total_bits)
    byte = (bit >> np.uint64(3)).astype(np.int64) + base
    v = np.zeros(n, dtype=np.uint64)
    for j in range(8):
        v |= raw[byte + j].astype(np.uint64) << np.uint64(8 * j)
    return (v >> (bit & np.uint64(7))) & np.uint64((1 << total_bits) - 1)


class KenlmModel:
    """Flat, de-quantised view of a KenLM QUANT_ARRAY_TRIE binary.

    order, counts[order]; words (list[str], id order), bos, eos;
    uni_prob/uni_backoff f32[V], uni_next u32[V+1];
    mid_word[k] i32[n_k], mid_prob[k]/mid_backoff[k] f32[n_k], mid_next[k] u32[n_k+1]   (k = 0 -> bigrams);
    long_word i32[n], long_prob f32[n].
    """

    def __init__(self, path: str):
        with open(path, "rb") as f:
            blob = f.read()
        if blob[: len(_MAGIC)] != _MAGIC:
            raise ValueError(f"{path}: not a KenLM binary (format version 5)")
        if struct.unpack_from("<fffII", blob, 56) != (0.0, 1.0, -0.5, 1, 0xFFFFFFFF) or struct.unpack_from("<Q", blob, 80)[0] != 1:
            raise ValueError(f"{path}: KenLM sanity header mismatch")
        order = blob[88]
        (model_type,) = struct.unpack_from("<I", blob, 96)
        if model_type != 5:
            raise ValueError(f"{path}: KenLM model type {model_type} is not supported (only QUANT_ARRAY_TRIE = 5)")
        if blob[100] == 0:
            raise ValueError(f"{path}: the binary was built without vocabulary strings")
        if order < 2 or order > 5:
            raise ValueError(f"{path}: n-gram order {order} outside the supported range [2, 5]")
        counts = list(struct.unpack_from(f"<{order}Q", blob, 108))
        raw = np.frombuffer(blob + b"\x00" * 8, dtype=np.uint8)
        off = (108 + 8 * order + 7) // 8 * 8
        if struct.unpack_from("<Q", blob, off)[0] + 1 != counts[0]:
            raise ValueError(f"{path}: vocabulary size does not match the unigram count")
        off += 8 + 8 * counts[0]
        if blob[off] != 2:
            raise ValueError(f"{path}: unknown quantiser version {blob[off]}")
        pbits, bbits = blob[off + 1], blob[off + 2]
        off += 8
        mid_ptab, mid_btab = [], []
        for _ in range(order - 2):
            mid_ptab.append(np.frombuffer(blob, dtype="<f4", count=1 << pbits, offset=off)); off += 4 << pbits
            mid_btab.append(np.frombuffer(blob, dtype="<f4", count=1 << bbits, offset=off)); off += 4 << bbits
        long_ptab = np.frombuffer(blob, dtype="<f4", count=1 << pbits, offset=off); off += 4 << pbits

        V = counts[0]
        uni = np.frombuffer(blob, dtype=np.dtype([("p", "<f4"), ("b", "<f4"), ("n", "<u8")]), count=V + 1, offset=off)
        off += 16 * (V + 2)
        self.uni_prob = (-np.abs(uni["p"][:V])).astype(np.float32)        # the sign bit is a flag, not a sign
        self.uni_backoff = np.ascontiguousarray(uni["b"][:V], dtype=np.float32)
        self.uni_next = uni["n"].astype(np.uint32)

        wbits = V.bit_length()
        self.mid_word: List[np.ndarray] = []
        self.mid_prob: List[np.ndarray] = []
        self.mid_backoff: List[np.ndarray] = []
        self.mid_next: List[np.ndarray] = []
        for k in range(order - 2):
            n, max_next = counts[k + 1], counts[k + 2]
            if blob[off] != 0:
                raise ValueError(f"{path}: unknown pointer-compression version")
            if blob[off + 1] != 0:
                raise ValueError(f"{path}: compressed trie pointers (-a {blob[off + 1]}) are not supported; the shipped models use 0")
            nbits = max_next.bit_length()
            off += 8 * 2 + 7
            total = wbits + bbits + pbits + nbits
            if total > 57:
                raise ValueError(f"{path}: {total}-bit trie records are not supported")
            v = _unpack(raw, off, total, n + 1)
            self.mid_word.append((v[:n] & np.uint64((1 << wbits) - 1)).astype(np.int32))
            self.mid_backoff.append(mid_btab[k][((v[:n] >> np.uint64(wbits)) & np.uint64((1 << bbits) - 1)).astype(np.int64)].astype(np.float32))
            self.mid_prob.append(mid_ptab[k][((v[:n] >> np.uint64(wbits + bbits)) & np.uint64((1 << pbits) - 1)).astype(np.int64)].astype(np.float32))
            self.mid_next.append((v >> np.uint64(wbits + bbits + pbits)).astype(np.uint32))
            off += ((1 + n) * total + 7) // 8 + 8
        n = counts[-1]
        total = wbits + pbits
        v = _unpack(raw, off, total, n)
        self.long_word = (v & np.uint64((1 << wbits) - 1)).astype(np.int32)
        self.long_prob = long_ptab[(v >> np.uint64(wbits)).astype(np.int64)].astype(np.float32)
        off += ((1 + n) * total + 7) // 8 + 8

        strings = blob[off:].split(b"\x00")
        if strings[-1] != b"" or len(strings) - 1 != V:
            raise ValueError(f"{path}: section sizes do not add up (vocabulary strings expected at byte {off})")
        self.words = [w.decode("utf-8") for w in strings[:-1]]
        self.word2id: Dict[str, int] = {w: i for i, w in enumerate(self.words)}
        if "<s>" not in self.word2id or "</s>" not in self.word2id or self.words[0] != "<unk>":
            raise ValueError(f"{path}: <unk>/<s>/</s> missing from the vocabulary")
        self.bos, self.eos = self.word2id["<s>"], self.word2id["</s>"]
        self.order, self.counts, self.path = order, counts, path
        self._check()

    def _check(self) -> None:
        """Structural invariants of a reverse trie (cheap, vectorised): child ranges monotone and complete."""
        nxt = [self.uni_next] + self.mid_next
        for k, a in enumerate(nxt):
            if a[0] != 0 or a[-1] != self.counts[k + 1] or np.any(np.diff(a.astype(np.int64)) < 0):
                raise ValueError(f"{self.path}: child ranges of order {k + 1} are not monotone (file layout not understood)")
        words = self.mid_word + [self.long_word]
        for k, (a, w) in enumerate(zip(nxt, words)):
            d = np.diff(w.astype(np.int64))
            starts = np.zeros(len(w), dtype=bool)
            starts[a[:-1][a[:-1] < len(w)]] = True          # first child of every node
            if np.any((d <= 0) & ~starts[1:]):
                raise ValueError(f"{self.path}: word ids of order {k + 2} are not ascending inside a node")

    def vocabulary_table(self, labels: Sequence[str], hash_fn) -> Dict[str, np.ndarray]:
        """Open-addressing table {hash(label-id sequence of a word) -> word id} for the words the acoustic model can
        spell (single-character labels); hash_fn(list[int]) must be the kernel's rolling hash."""
        lab = {c: i for i, c in enumerate(labels)}
        if any(len(c) != 1 for c in labels):
            raise NotImplementedError("LM fusion needs single-character labels")
        items = []
        for wid, w in enumerate(self.words):
            if w and all(ch in lab for ch in w) and " " not in w:
                items.append((hash_fn([lab[ch] for ch in w]), wid))
        size = 1
        while size < 2 * max(len(items), 1):
            size <<= 1
        keys = np.zeros(size, dtype=np.uint64)
        vals = np.full(size, -1, dtype=np.int32)
        for h, wid in items:
            if h == 0:
                raise ValueError("vocabulary hash collision with the empty-slot marker")
            p = h & (size - 1)
            while keys[p] != 0:
                if int(keys[p]) == h:
                    raise ValueError(f"64-bit hash collision between LM vocabulary words ({self.words[wid]!r})")
                p = (p + 1) & (size - 1)
            keys[p] = h
            vals[p] = wid
        return {"keys": keys, "vals": vals}
