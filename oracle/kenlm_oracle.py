"""CPU oracle for n-gram queries on the KenLM binaries the reference ships
(models/language_model/{3,4,5}-gram-lm.binary).

TEST INFRASTRUCTURE ONLY (see oracle/quartznet_oracle.py for the import rules).

PARITY UNPINNED.  The reference scores words through the third-party packages `kenlm`
(README.md:43-45, installed from github master, unpinned) and `pyctcdecode`
(requirements.txt:16; call site nemo/collections/asr/beam_search_decoder.py:82-87).  Neither is in
/root/reference nor installed here, so this file restates, from the published KenLM source
(lm/binary_format.cc, lm/vocab.cc, lm/quantize.cc, lm/bhiksha.cc, lm/trie.cc, lm/search_trie.cc,
lm/model.cc), how a "QUANT_ARRAY_TRIE" (model type 5) file is laid out and queried.  What pins it
instead of a golden vector: the files themselves.  A wrong bit layout cannot satisfy all of
  * section sizes derived from the header counts add up to the file size exactly,
  * word ids are strictly ascending inside every trie node, child ranges are monotone and end at
    the next order's count,
  * sum_w P(w | context) == 1 for observed contexts (a back-off model is normalised),
  * the training text (models/language_model/alltext.txt) gets a low perplexity,
which tests/test_oracle_cpu.py checks when the files are present.

Layout restated (little endian):
  [0,88)    Sanity: 52-byte magic "mmap lm http://kheafield.com/code format version 5\\n\\0" padded
            to 56, floats 0/1/-0.5, uint32 1, uint32 max, uint64 1
  [88,108)  FixedWidthParameters: u8 order, f32 probing_multiplier, u32 model_type (5),
            u8 has_vocabulary, u32 search_version
  [108,..)  u64 counts[order]; padded to 8
  vocab     SortedVocabulary: u64 n (= counts[0]-1, <unk> has no hash), u64 hash[counts[0]] sorted
            (MurmurHash64A of the word; word id = rank + 1, <unk> = 0)
  quant     u8 version(2), u8 prob_bits, u8 backoff_bits, pad to 8; then for every middle order
            f32 prob_bins[2^prob_bits], f32 backoff_bins[2^backoff_bits]; for the longest order
            f32 prob_bins[2^prob_bits]
  unigram   (counts[0]+2) x {f32 prob, f32 backoff, u64 next}
  middle k  (orders 2..n-1) ArrayBhiksha offsets (8-byte header {u8 version, u8 chop_bits config},
            u64 offsets[array_count], +7), then bit-packed entries of
            word_bits | backoff_bits | prob_bits | inline_next_bits, (count+1) entries, +8 bytes
  longest   bit-packed word_bits | prob_bits, (count+1) entries, +8 bytes
  strings   the vocabulary as NUL-terminated strings in word-id order (has_vocabulary)
The trie is keyed in reverse: the path w_n -> w_{n-1} -> ... reaches the node of n-gram
(w_1 .. w_n); its prob is log10 P(w_n | w_1..w_{n-1}), its backoff that of the n-gram as a context.
"""
from __future__ import annotations

import math
import struct
from typing import Dict, List, Sequence, Tuple

MAGIC = b"mmap lm http://kheafield.com/code format version 5\n\x00"


def required_bits(max_value: int) -> int:
    """util::RequiredBits: bits needed to store values in [0, max_value]."""
    return max_value.bit_length() if max_value else 0


class KenlmBinary:
    """Scalar, read-on-demand view of a QUANT_ARRAY_TRIE file."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.d = f.read()
        d = self.d
        if d[: len(MAGIC)] != MAGIC:
            raise ValueError("not a KenLM binary (format version 5)")
        zero_f, one_f, mhalf_f, one_w, max_w = struct.unpack_from("<fffII", d, 56)
        (one_q,) = struct.unpack_from("<Q", d, 80)
        if (zero_f, one_f, mhalf_f, one_w, max_w, one_q) != (0.0, 1.0, -0.5, 1, 0xFFFFFFFF, 1):
            raise ValueError("KenLM sanity header mismatch (endianness / ABI)")
        self.order = d[88]
        (self.model_type,) = struct.unpack_from("<I", d, 96)
        self.has_vocabulary = d[100] != 0
        if self.model_type != 5:
            raise ValueError(f"only QUANT_ARRAY_TRIE (5) is restated, file has model type {self.model_type}")
        self.counts = list(struct.unpack_from(f"<{self.order}Q", d, 108))
        off = (108 + 8 * self.order + 7) // 8 * 8
        (n_hash,) = struct.unpack_from("<Q", d, off)
        if n_hash + 1 != self.counts[0]:
            raise ValueError("vocabulary size does not match the unigram count")
        off += 8 + 8 * self.counts[0]
        # ---- quantisation tables
        if d[off] != 2:
            raise ValueError("unknown SeparatelyQuantize version")
        self.prob_bits, self.backoff_bits = d[off + 1], d[off + 2]
        off += 8
        self.mid_prob: List[Tuple[float, ...]] = []
        self.mid_backoff: List[Tuple[float, ...]] = []
        for _ in range(self.order - 2):
            self.mid_prob.append(struct.unpack_from(f"<{1 << self.prob_bits}f", d, off)); off += 4 << self.prob_bits
            self.mid_backoff.append(struct.unpack_from(f"<{1 << self.backoff_bits}f", d, off)); off += 4 << self.backoff_bits
        self.long_prob = struct.unpack_from(f"<{1 << self.prob_bits}f", d, off); off += 4 << self.prob_bits
        # ---- unigrams
        self.uni_off = off
        off += 16 * (self.counts[0] + 2)
        # ---- middles
        self.word_bits = required_bits(self.counts[0])
        self.word_mask = (1 << self.word_bits) - 1
        self.mid_base: List[int] = []
        self.mid_total_bits: List[int] = []
        self.mid_next_bits: List[int] = []
        for k in range(self.order - 2):                      # k = 0 -> bigrams
            entries, max_next = self.counts[k + 1], self.counts[k + 2]
            if d[off] != 0:
                raise ValueError("unknown ArrayBhiksha version")
            if d[off + 1] != 0:
                raise ValueError("ArrayBhiksha pointer compression (chop bits > 0) is not restated; the shipped files use 0")
            next_bits = required_bits(max_next)              # chop == 0: the whole pointer is stored inline
            array_count = (max_next >> next_bits) + 1        # == 1
            off += 8 * (1 + array_count) + 7
            total = self.word_bits + self.backoff_bits + self.prob_bits + next_bits
            self.mid_base.append(off); self.mid_total_bits.append(total); self.mid_next_bits.append(next_bits)
            off += ((1 + entries) * total + 7) // 8 + 8
        # ---- longest
        self.long_base = off
        self.long_total_bits = self.word_bits + self.prob_bits
        off += ((1 + self.counts[-1]) * self.long_total_bits + 7) // 8 + 8
        self.strings_off = off
        if not self.has_vocabulary:
            raise ValueError("file carries no vocabulary strings")
        words = d[off:].split(b"\x00")
        if words[-1] != b"" or len(words) - 1 != self.counts[0]:
            raise ValueError(f"section sizes do not add up: strings at {off}, {len(words) - 1} words, file {len(d)} bytes")
        self.words = [w.decode("utf-8") for w in words[:-1]]
        self.word2id: Dict[str, int] = {w: i for i, w in enumerate(self.words)}
        self.bos, self.eos = self.word2id["<s>"], self.word2id["</s>"]

    # ------------------------------------------------------------------ raw reads (util/bit_packing.hh ReadInt57)
    def _bits(self, base: int, bit_off: int, nbits: int) -> int:
        byte = base + (bit_off >> 3)
        return (int.from_bytes(self.d[byte: byte + 8], "little") >> (bit_off & 7)) & ((1 << nbits) - 1)

    def unigram(self, w: int) -> Tuple[float, float, int, int]:
        prob, backoff, nxt = struct.unpack_from("<ffQ", self.d, self.uni_off + 16 * w)
        (end,) = struct.unpack_from("<Q", self.d, self.uni_off + 16 * (w + 1) + 8)
        return -abs(prob), backoff, nxt, end                 # the sign bit of prob is a flag, not a sign

    def middle_word(self, k: int, i: int) -> int:
        return self._bits(self.mid_base[k], i * self.mid_total_bits[k], self.word_bits)

    def middle(self, k: int, i: int) -> Tuple[float, float, int, int]:
        base, tb = self.mid_base[k], self.mid_total_bits[k]
        at = i * tb + self.word_bits
        q_b = self._bits(base, at, self.backoff_bits)
        q_p = self._bits(base, at + self.backoff_bits, self.prob_bits)
        at += self.backoff_bits + self.prob_bits
        nb = self.mid_next_bits[k]
        return self.mid_prob[k][q_p], self.mid_backoff[k][q_b], self._bits(base, at, nb), self._bits(base, at + tb, nb)

    def longest_word(self, i: int) -> int:
        return self._bits(self.long_base, i * self.long_total_bits, self.word_bits)

    def longest(self, i: int) -> float:
        return self.long_prob[self._bits(self.long_base, i * self.long_total_bits + self.word_bits, self.prob_bits)]

    def _find(self, word_at, lo: int, hi: int, w: int) -> int:
        """index of word w in the sorted node range [lo, hi), or -1 (lm/trie.cc FindBitPacked; any search works)."""
        while lo < hi:
            mid = (lo + hi) // 2
            x = word_at(mid)
            if x < w:
                lo = mid + 1
            elif x > w:
                hi = mid
            else:
                return mid
        return -1

    # ------------------------------------------------------------------ queries
    def walk(self, rev_words: Sequence[int]) -> List[Tuple[float, float]]:
        """(prob, backoff) of the n-grams ending the reversed word path, for every length that exists."""
        out: List[Tuple[float, float]] = []
        if not rev_words:
            return out
        p, b, lo, hi = self.unigram(rev_words[0])
        out.append((p, b))
        for depth, w in enumerate(rev_words[1: self.order], start=1):
            if depth < self.order - 1:
                k = depth - 1
                i = self._find(lambda j: self.middle_word(k, j), lo, hi, w)
                if i < 0:
                    break
                p, b, lo, hi = self.middle(k, i)
                out.append((p, b))
            else:
                i = self._find(self.longest_word, lo, hi, w)
                if i < 0:
                    break
                out.append((self.longest(i), 0.0))
        return out

    def score(self, context: Sequence[int], w: int) -> float:
        """log10 P(w | context) with back-off (lm/model.cc GenericModel::FullScore); context oldest -> newest."""
        ctx = list(context)[-(self.order - 1):] if self.order > 1 else []
        found = self.walk([w] + ctx[::-1])
        prob = found[-1][0]
        if len(found) <= len(ctx):                           # charge the back-off of the context n-grams not matched
            cfound = self.walk(ctx[::-1])
            for j in range(len(found), len(cfound) + 1):     # context suffix of length j
                if j >= 1:
                    prob += cfound[j - 1][1]
        return prob

    def index(self, word: str) -> int:
        return self.word2id.get(word, 0)

    def score_words(self, words: Sequence[str], bos: bool = True, eos: bool = True) -> float:
        ctx = [self.bos] if bos else []
        total = 0.0
        for w in words:
            i = self.index(w)
            total += self.score(ctx, i)
            ctx = (ctx + [i])[-(self.order - 1):]
        if eos:
            total += self.score(ctx, self.eos)
        return total


def perplexity(lm: KenlmBinary, lines: Sequence[str]) -> float:
    total, n = 0.0, 0
    for ln in lines:
        ws = ln.split()
        total += lm.score_words(ws)
        n += len(ws) + 1
    return math.pow(10.0, -total / max(n, 1))
