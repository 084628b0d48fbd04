"""CPU oracle for CTC prefix beam search WITHOUT a language model.

TEST INFRASTRUCTURE ONLY (see oracle/quartznet_oracle.py for the import rules).

PARITY UNPINNED.  The reference delegates beam search to the third-party package
`pyctcdecode` (requirements.txt:16, unpinned; call sites
nemo/collections/asr/beam_search_decoder.py:12, 82-87, 98-101).  Its source is not
in /root/reference and the package is not installed in this image, and the
reference ships no transcripts or golden vectors for it.  This file restates the
published algorithm of pyctcdecode's `BeamSearchDecoderCTC._decode_logits`
(version ~0.5, defaults beam_prune_logp=-10, token_min_logp=-5, no hotwords, no
LM) from memory of the public package; the CUDA kernel is tested against THIS
restatement, not against pyctcdecode itself.

Behaviour restated (lm_path=None, what infer.py:118-130 falls back to when kenlm
is missing):
  * input: `exp(log_probs[0])` rows sum to 1 -> treated as probabilities ->
    log(clip(p, 1e-15, 1));
  * per frame: candidate symbols = {c : logp[c] >= token_min_logp} U {argmax};
  * a beam is (text, next_word, word_part, last_char, logit_score); for symbol c:
      blank or c == last_char -> same text, last_char = c
      c == ' '               -> next_word = word_part, word_part = ''
      else                   -> word_part += c
    symbols are visited in ascending vocabulary index (blank = last index),
    beams in their current order;
  * beams with equal (text + next_word, word_part, last_char) merge by log-sum-exp,
    keeping first-seen order; next_word is folded into text;
  * prune: keep score >= best + beam_prune_logp, then the beam_width best
    (stable sort, ties keep order);
  * end: fold word_part into text, merge equal texts, best text with whitespace
    normalised.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

MIN_TOKEN_CLIP_P = 1e-15
MAX_CANDIDATES = 16   # candidate symbols per frame in the CUDA kernel (csrc/beam.cu MC)
NEG_INF = float("-inf")


def _logaddexp(a: float, b: float) -> float:
    if a == NEG_INF:
        return b
    if b == NEG_INF:
        return a
    m = max(a, b)
    return m + math.log(math.exp(a - m) + math.exp(b - m))


def _merge_tokens(a: str, b: str) -> str:
    if not b:
        return a
    return b if not a else a + " " + b


def beam_search_no_lm(log_probs: np.ndarray, labels: Sequence[str], beam_width: int,
                      token_min_logp: float = -5.0, beam_prune_logp: float = -10.0) -> Tuple[str, float]:
    """log_probs [T, V+1] (blank = last class) -> (best text, its log score)."""
    # the reference feeds exp(log_probs); pyctcdecode takes log(clip(p, 1e-15, 1)) of that - i.e. the float32
    # log-probs themselves, floored at ln(1e-15) (the exp/log round trip's last-ulp noise is not modelled)
    lp32 = np.minimum(np.maximum(np.asarray(log_probs, dtype=np.float32), np.float32(np.log(np.float32(MIN_TOKEN_CLIP_P)))), np.float32(0))
    lp = lp32.astype(np.float64)
    vocab = list(labels) + [""]
    beams: List[Tuple[str, str, str, object, float]] = [("", "", "", None, 0.0)]
    for t in range(lp.shape[0]):
        col = lp[t]
        amax = int(col.argmax())
        cand = sorted(set(np.where(col >= token_min_logp)[0].tolist()) | {amax})
        if len(cand) > MAX_CANDIDATES:   # kernel limit (never reached by a peaked CTC posterior): keep the most probable
            keep = sorted((c for c in cand if c != amax), key=lambda c: -col[c])[: MAX_CANDIDATES - 1]
            cand = sorted(set(keep) | {amax})
        new = []
        for c in cand:
            p = float(col[c])
            ch = vocab[c]
            for text, next_word, word_part, last_char, score in beams:
                if ch == "" or last_char == ch:
                    new.append((text, next_word, word_part, ch, score + p))
                elif ch == " ":
                    new.append((text, word_part, "", ch, score + p))
                else:
                    new.append((text, next_word, word_part + ch, ch, score + p))
        merged = {}
        for text, next_word, word_part, last_char, score in new:
            key = (_merge_tokens(text, next_word), word_part, last_char)   # pyctcdecode's _merge_beams hash
            merged[key] = _logaddexp(merged[key], score) if key in merged else score
        scored = [(k[0], "", k[1], k[2], v) for k, v in merged.items()]
        best = max(b[4] for b in scored)
        scored = [b for b in scored if b[4] >= best + beam_prune_logp]
        scored.sort(key=lambda b: -b[4])          # stable: ties keep first-seen order
        beams = scored[:beam_width]
    final = {}
    for text, _next_word, word_part, _last, score in beams:
        full = _merge_tokens(text, word_part)
        final[full] = _logaddexp(final[full], score) if full in final else score
    best_text, best_score = max(final.items(), key=lambda kv: kv[1])
    return " ".join(best_text.split()), best_score


def beam_search_batch(log_probs: np.ndarray, labels: Sequence[str], beam_width: int) -> List[str]:
    return [beam_search_no_lm(lp, labels, beam_width)[0] for lp in log_probs]
