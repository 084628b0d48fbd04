"""Writer for small KenLM "QUANT_ARRAY_TRIE" binaries - TEST INFRASTRUCTURE ONLY.

Produces tests/golden/tiny_lm_*.binary so that the LM parity tests (host parser, device trie walk, LM-fused beam
search) run without the reference's model files, which cannot travel to the GPU box.
The byte layout is the one documented in oracle/kenlm_oracle.py, which was derived from and validated on the
shipped models/language_model/{3,4,5}-gram-lm.binary (section sizes add up to the file size, back-off
distributions normalise).  `murmur64a` is the vocabulary hash of lm/vocab.cc (util/murmur_hash.cc MurmurHash64A,
seed 0): tests/test_oracle_cpu.py checks it against the hashes stored in the shipped files when they are present.

Run `python oracle/kenlm_writer.py` to regenerate the fixtures (deterministic).
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Sequence, Tuple

import numpy as np

MAGIC = b"mmap lm http://kheafield.com/code format version 5\n\x00"
_M = 0xC6A4A7935BD1E995
_MASK = (1 << 64) - 1


def murmur64a(data: bytes, seed: int = 0) -> int:
    h = (seed ^ (len(data) * _M)) & _MASK
    nblocks = len(data) // 8
    for i in range(nblocks):
        k = int.from_bytes(data[8 * i: 8 * i + 8], "little")
        k = (k * _M) & _MASK
        k ^= k >> 47
        k = (k * _M) & _MASK
        h ^= k
        h = (h * _M) & _MASK
    tail = data[8 * nblocks:]
    if tail:
        h ^= int.from_bytes(tail, "little")
        h = (h * _M) & _MASK
    h ^= h >> 47
    h = (h * _M) & _MASK
    h ^= h >> 47
    return h


def _pack(records: Sequence[int], total_bits: int) -> bytes:
    acc = 0
    for i, r in enumerate(records):
        acc |= r << (i * total_bits)
    return acc.to_bytes((len(records) * total_bits + 7) // 8, "little")


def write_quant_array_trie(out_path: str, words: Sequence[str], ngrams: Dict[Tuple[str, ...], Tuple[float, float]],
                           order: int, prob_bits: int = 8, backoff_bits: int = 7, seed: int = 0) -> None:
    """words: vocabulary without <unk> (must hold <s> and </s>); ngrams: {(w1..wn): (log10 prob, log10 backoff)} for
    n = 1..order, closed under taking suffixes (the reverse trie needs every suffix as a node).  Probabilities of
    order >= 2 are snapped to randomly drawn quantisation bins (unigrams are stored as floats, like KenLM does)."""
    rng = np.random.default_rng(seed)
    vocab = sorted(set(words), key=lambda w: murmur64a(w.encode("utf-8")))
    hashes = [murmur64a(w.encode("utf-8")) for w in vocab]
    vocab = ["<unk>"] + vocab
    wid = {w: i for i, w in enumerate(vocab)}
    V = len(vocab)
    # reverse trie: children[path] with path = (w_n, w_{n-1}, ...) as ids
    levels: List[Dict[Tuple[int, ...], Tuple[float, float]]] = [dict() for _ in range(order)]
    for gram, pb in ngrams.items():
        ids = tuple(wid[w] for w in gram)
        levels[len(ids) - 1][ids[::-1]] = pb
    for n in range(1, order):
        for path in levels[n]:
            if path[:-1] not in levels[n - 1]:
                raise ValueError(f"n-gram set is not suffix-closed: {path}")
    for w in range(V):
        levels[0].setdefault((w,), (-9.0, 0.0))
    # node order per level: parents in their order, children ascending by word id
    order_nodes: List[List[Tuple[int, ...]]] = [[(w,) for w in range(V)]]
    child_begin: List[List[int]] = []
    for n in range(1, order):
        kids: Dict[Tuple[int, ...], List[int]] = {}
        for path in levels[n]:
            kids.setdefault(path[:-1], []).append(path[-1])
        nodes, begins = [], []
        for parent in order_nodes[n - 1]:
            begins.append(len(nodes))
            for w in sorted(kids.get(parent, [])):
                nodes.append(parent + (w,))
        begins.append(len(nodes))
        order_nodes.append(nodes)
        child_begin.append(begins)
    counts = [len(x) for x in order_nodes]

    # quantisation tables (ascending bins; back-off bins 0/1 are KenLM's reserved -0.0 / 0.0)
    def table(nbins, lo, hi, reserved):
        t = np.sort(rng.uniform(lo, hi, nbins).astype(np.float32))
        if reserved:
            t[0], t[1] = np.float32(-0.0), np.float32(0.0)
        return t
    mid_ptab = [table(1 << prob_bits, -6.0, -0.05, False) for _ in range(order - 2)]
    mid_btab = [table(1 << backoff_bits, -2.0, -0.01, True) for _ in range(order - 2)]
    long_ptab = table(1 << prob_bits, -6.0, -0.05, False)

    def nearest(tab, x, first):
        return int(first + np.argmin(np.abs(tab[first:] - np.float32(x))))

    out = bytearray()
    out += MAGIC.ljust(56, b"\x00")
    out += struct.pack("<fffII", 0.0, 1.0, -0.5, 1, 0xFFFFFFFF) + b"\x00" * 4 + struct.pack("<Q", 1)
    out += struct.pack("<B3xfIB3xI", order, 1.5, 5, 1, 1)
    out += struct.pack(f"<{order}Q", *counts)
    out += b"\x00" * ((-len(out)) % 8)
    out += struct.pack("<Q", V - 1) + struct.pack(f"<{V - 1}Q", *hashes) + struct.pack("<Q", 0)
    out += bytes([2, prob_bits, backoff_bits, 0, 0, 0, 0, 0])
    for k in range(order - 2):
        out += mid_ptab[k].tobytes() + mid_btab[k].tobytes()
    out += long_ptab.tobytes()
    ends = child_begin[0] if order > 1 else [0] * (V + 1)
    for w in range(V):
        p, b = levels[0][(w,)]
        out += struct.pack("<ffQ", p, b, ends[w])
    out += struct.pack("<ffQ", 0.0, 0.0, ends[V]) + struct.pack("<ffQ", 0.0, 0.0, 0)
    wbits = V.bit_length()
    for k in range(order - 2):                               # level index k + 1
        nodes = order_nodes[k + 1]
        nbits = counts[k + 2].bit_length()
        total = wbits + backoff_bits + prob_bits + nbits
        recs = []
        for i, path in enumerate(nodes):
            p, b = levels[k + 1][path]
            qb = nearest(mid_btab[k], b, 1) if b != 0.0 else 1
            qp = nearest(mid_ptab[k], p, 0)
            recs.append(path[-1] | (qb << wbits) | (qp << (wbits + backoff_bits)) |
                        (child_begin[k + 1][i] << (wbits + backoff_bits + prob_bits)))
        recs.append(child_begin[k + 1][len(nodes)] << (wbits + backoff_bits + prob_bits))
        out += b"\x00" * 8 + struct.pack("<Q", 0) + b"\x00" * 7
        out += _pack(recs, total) + b"\x00" * 8
    nodes = order_nodes[order - 1]
    total = wbits + prob_bits
    recs = [path[-1] | (nearest(long_ptab, levels[order - 1][path][0], 0) << wbits) for path in nodes] + [0]
    out += _pack(recs, total) + b"\x00" * 8
    out += b"".join(w.encode("utf-8") + b"\x00" for w in vocab)
    with open(out_path, "wb") as f:
        f.write(bytes(out))


TINY_WORDS = ["<s>", "</s>", "the", "cat", "sat", "on", "mat", "a", "dog", "hi", "there", "hat", "he", "she", "it's",
              "that", "at", "an", "and", "ant", "then", "them", "this", "is", "his", "mad", "had", "ham", "hot", "to",
              "too", "tot", "dot", "do", "so", "sad", "sit", "set", "sea", "see", "tea", "ten", "tan", "can", "cap"]


def tiny_ngrams(order: int, seed: int):
    """A random suffix-closed n-gram set over TINY_WORDS (probabilities are arbitrary, not normalised)."""
    rng = np.random.default_rng(seed)
    real = [w for w in TINY_WORDS if w != "<s>"]
    grams: Dict[Tuple[str, ...], Tuple[float, float]] = {}
    for w in TINY_WORDS + ["<unk>"]:
        grams[(w,)] = (0.0 if w == "<s>" else float(-rng.uniform(1.0, 3.5)), float(-rng.uniform(0.05, 1.0)))
    n_top = {2: 400, 3: 700, 4: 500, 5: 300}
    for n in range(2, order + 1):
        for _ in range(n_top[n]):
            ctx = [TINY_WORDS[int(rng.integers(len(TINY_WORDS)))] for _ in range(n - 1)]
            ctx = [w if w != "</s>" else "the" for w in ctx]
            ctx = [ctx[0]] + [w if w != "<s>" else "cat" for w in ctx[1:]]          # <s> only sentence-initial
            gram = tuple(ctx + [real[int(rng.integers(len(real)))]])
            for s in range(len(gram)):                        # suffix closure
                g = gram[s:]
                if g not in grams:
                    grams[g] = (float(-rng.uniform(0.1, 4.0)), 0.0 if len(g) == order else float(-rng.uniform(0.02, 1.5)))
    return grams


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(os.path.dirname(here), "tests", "golden")
    for order, seed in ((3, 11), (5, 12)):
        path = os.path.join(out, f"tiny_lm_{order}gram.binary")
        write_quant_array_trie(path, TINY_WORDS, tiny_ngrams(order, seed), order, seed=seed)
        print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
