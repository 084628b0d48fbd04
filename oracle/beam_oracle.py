"""CPU oracle for CTC prefix beam search: without a language model (`beam_search_no_lm`) and with KenLM
shallow fusion (`beam_search_lm`, see the second half of this file).

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


# ======================================================================================================================
# Beam search WITH KenLM shallow fusion - what `BeamSearchDecoderWithLM(lm_path=<binary>)` runs
# (beam_search_decoder.py:82-87 -> pyctcdecode.build_ctcdecoder(vocab, kenlm_model_path, alpha, beta);
#  infer.py:184-191: 3-gram-lm.binary, beam_width 100, alpha 0.5, beta 1.5).
#
# PARITY UNPINNED, like the no-LM search above: restated from the public pyctcdecode package
# (decoder.py `_decode_logits`, `_get_lm_beams`; language_model.py `LanguageModel.score`,
# `score_partial_token`), defaults unk_score_offset=-10, score_boundary=True, AVG_TOKEN_LEN=6,
# LOG_BASE_CHANGE_FACTOR=1/log10(e), no hotwords, prune_history=False.  Behaviour restated:
#   * a KenLM *binary* gives pyctcdecode no unigram list ("No known unigrams provided"), so its character trie is
#     None and EVERY non-empty partial word is charged unk_score_offset (x len/6 when longer than 6 characters);
#     an out-of-vocabulary word is detected with `word not in kenlm_model` and charged unk_score_offset (log10 units,
#     before alpha);
#   * per frame, after the merge, every candidate beam (text, next_word, word_part) gets
#       lm(text + next_word) + partial(word_part);  lm(text') is cached BY TEXT: the first time a text is seen its
#       score is   lm(text) + alpha * ln(10) * log10 P(next_word | state(text)) [+ unk] + beta   and is never recomputed;
#   * candidates are pruned at best(combined) + beam_prune_logp and the beam_width best by combined score kept
#     (stable); beams carry only the acoustic score;
#   * at the end word_part becomes next_word and EVERY final text is scored with is_last_word=True
#     (+ log10 P(</s> | state)): pyctcdecode >= 0.5 keys its cache by (text, is_eos), so a text that was already
#     cached without </s> during the search is scored again (older versions keyed by text only).  A final
#     text with no pending word gets the </s> term alone here (the package would additionally score an empty word:
#     not restated - unpinned either way);
#   * the LM state of a text is the KenLM state after <s> w1 .. wn, i.e. its last order-1 words.
LOG10_TO_LN = 1.0 / math.log10(math.e)
AVG_TOKEN_LEN = 6


def _partial_token_score(word_part: str, unk_score_offset: float) -> float:
    if not word_part:
        return 0.0
    s = unk_score_offset                               # char trie is None -> is_oov = 1
    if len(word_part) > AVG_TOKEN_LEN:
        s = s * len(word_part) / AVG_TOKEN_LEN
    return s


def beam_search_lm(log_probs: np.ndarray, labels: Sequence[str], beam_width: int, lm, alpha: float = 0.5,
                   beta: float = 1.5, unk_score_offset: float = -10.0, token_min_logp: float = -5.0,
                   beam_prune_logp: float = -10.0, return_all: bool = False):
    """log_probs [T, V+1] (blank = last class), lm = oracle.kenlm_oracle.KenlmBinary -> (best text, combined score)."""
    lp32 = np.minimum(np.maximum(np.asarray(log_probs, dtype=np.float32), np.float32(np.log(np.float32(MIN_TOKEN_CLIP_P)))), np.float32(0))
    lp = lp32.astype(np.float64)
    vocab = list(labels) + [""]
    ctx_len = lm.order - 1
    # text -> (lm score of the text, KenLM context ids)
    cache = {"": (0.0, [lm.bos])}

    eos_cache = {}                                     # pyctcdecode >= 0.5 keys its cache by (text, is_eos)

    def lm_of(text: str, next_word: str, is_eos: bool) -> float:
        new_text = _merge_tokens(text, next_word)
        if is_eos:
            # end of the utterance: EVERY final text is scored as the end of the sentence, also one that was already
            # committed (and cached without </s>) during the search
            if new_text not in eos_cache:
                prev, ctx = cache[text]
                if next_word:
                    wid = lm.index(next_word)
                    raw = lm.score(ctx, wid)
                    if next_word not in lm.word2id:
                        raw += unk_score_offset
                    nctx = (ctx + [wid])[-ctx_len:] if ctx_len > 0 else []
                    raw += lm.score(nctx, lm.eos)
                    eos_cache[new_text] = prev + alpha * raw * LOG10_TO_LN + beta
                else:                                   # no pending word: the </s> term alone, no word bonus
                    eos_cache[new_text] = prev + alpha * lm.score(ctx, lm.eos) * LOG10_TO_LN
            return eos_cache[new_text]
        if new_text not in cache:
            prev, ctx = cache[text]
            wid = lm.index(next_word)
            raw = lm.score(ctx, wid)
            if next_word not in lm.word2id:
                raw += unk_score_offset
            nctx = (ctx + [wid])[-ctx_len:] if ctx_len > 0 else []
            cache[new_text] = (prev + alpha * raw * LOG10_TO_LN + beta, nctx)
        return cache[new_text][0]

    def score_beams(merged, is_eos):
        out = []
        for (text, next_word, word_part, last_char), score in merged:
            lm_score = lm_of(text, next_word, is_eos) + _partial_token_score(word_part, unk_score_offset)
            out.append((_merge_tokens(text, next_word), "", word_part, last_char, score, score + lm_score))
        best = max(b[5] for b in out)
        out = [b for b in out if b[5] >= best + beam_prune_logp]
        out.sort(key=lambda b: -b[5])                   # stable
        return out[:beam_width]

    beams: List[Tuple[str, str, str, object, float]] = [("", "", "", None, 0.0)]
    for t in range(lp.shape[0]):
        col = lp[t]
        amax = int(col.argmax())
        cand = sorted(set(np.where(col >= token_min_logp)[0].tolist()) | {amax})
        if len(cand) > MAX_CANDIDATES:
            keep = sorted((c for c in cand if c != amax), key=lambda c: -col[c])[: MAX_CANDIDATES - 1]
            cand = sorted(set(keep) | {amax})
        merged = {}
        for c in cand:
            p = float(col[c])
            ch = vocab[c]
            for text, next_word, word_part, last_char, score in beams:
                if ch == "" or last_char == ch:
                    nb = (text, next_word, word_part, ch)
                elif ch == " ":
                    nb = (text, word_part, "", ch)
                else:
                    nb = (text, next_word, word_part + ch, ch)
                key = (_merge_tokens(nb[0], nb[1]), nb[2], nb[3])
                if key in merged:
                    merged[key] = (merged[key][0], _logaddexp(merged[key][1], score + p))
                else:
                    merged[key] = (nb, score + p)
        scored = score_beams(list(merged.values()), False)
        beams = [b[:5] for b in scored]
    final = {}
    for text, _nw, word_part, _last, score in beams:
        key = _merge_tokens(text, word_part)
        if key in final:
            final[key] = (final[key][0], _logaddexp(final[key][1], score))
        else:
            final[key] = ((text, word_part, "", None), score)
    scored = score_beams(list(final.values()), True)
    if return_all:
        return [(" ".join(b[0].split()), b[4], b[5]) for b in scored]
    return " ".join(scored[0][0].split()), scored[0][5]
