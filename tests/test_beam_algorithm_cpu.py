"""CPU model of the per-frame selection logic of csrc/beam.cu (phases 3-6), checked against the sort-based definition.

The CUDA kernel merges equal candidate states through a hash table (first-seen member = lowest insertion index, members
reduced in insertion order), prunes, radix-selects the key of the beam_width-th best state when many states are alive
and ranks the survivors by counting.  The definition it must reproduce - what oracle/beam_oracle.py and pyctcdecode do -
is: stable sort by (key, insertion index), run-wise log-sum-exp in that order, prune, stable sort by score descending,
keep the first beam_width.  This test restates BOTH in numpy/Python (no GPU) and compares them on random candidate sets
with duplicates, exact score ties and wide frames; the kernel itself is compared with the oracle in the GPU tests.
"""
import math
import struct

import numpy as np
import pytest

BW_MAX = 128


def _logaddexp(a, b):
    m = max(a, b)
    return m + math.log(1.0 + math.exp(min(a, b) - m))          # csrc/beam.cu logaddexp_d


def _dkey(x):
    u = struct.unpack("<Q", struct.pack("<d", x))[0]
    return (~u) & 0xFFFFFFFFFFFFFFFF if u & 0x8000000000000000 else u | 0x8000000000000000


def by_sorting(keys, scores, prune, beam_width):
    """The definition: (rep insertion index, merged score) of the new beams in rank order."""
    order = sorted(range(len(keys)), key=lambda i: (keys[i], i))
    merged = []                                                  # (rep, score)
    p = 0
    while p < len(order):
        q = p
        acc = scores[order[p]]
        while q + 1 < len(order) and keys[order[q + 1]] == keys[order[p]]:
            q += 1
            acc = _logaddexp(acc, scores[order[q]])
        merged.append((order[p], acc))
        p = q + 1
    best = max(s for _, s in merged)
    alive = [(rep, s) for rep, s in merged if s >= best + prune]
    alive.sort(key=lambda t: (-t[1], t[0]))                      # score descending, first-seen ascending
    return alive[:beam_width]


def by_hashing(keys, scores, prune, beam_width, rng):
    """What the kernel does, with the thread interleaving replaced by a random processing order."""
    n = len(keys)
    npow = 1
    while npow < n:
        npow <<= 1
    mask = 2 * npow - 1
    EMPTY = 0xFFFF
    table = [EMPTY] * (2 * npow)
    head = [EMPTY] * (2 * npow)
    nxt = [EMPTY] * n
    slot = [0] * n
    for idx in rng.permutation(n):                               # phase 3: any arrival order must give the same result
        idx = int(idx)
        p = (keys[idx] >> 20) & mask
        while True:
            e = table[p]
            if e == EMPTY:
                table[p] = idx
                break
            if keys[e] == keys[idx]:
                table[p] = min(table[p], idx)
                break
            p = (p + 1) & mask
        slot[idx] = p
        nxt[idx], head[p] = head[p], idx
    merged = {}
    for idx in range(n):                                         # phase 4: first-seen members reduce in insertion order
        p = slot[idx]
        if table[p] != idx:
            continue
        acc, last = scores[idx], idx
        while True:
            pick = None
            e = head[p]
            while e != EMPTY:
                if e > last and (pick is None or e < pick):
                    pick = e
                e = nxt[e]
            if pick is None:
                break
            acc = _logaddexp(acc, scores[pick])
            last = pick
        merged[idx] = acc
    best = max(merged.values())
    thr = best + prune
    alive = [(_dkey(s), rep) for rep, s in merged.items() if s >= thr]
    alive = [alive[int(i)] for i in rng.permutation(len(alive))]   # atomicAdd order is arbitrary
    A = len(alive)
    surv = alive
    if A > 2 * BW_MAX:                                           # phase 6: radix select
        kbest, kthr = _dkey(best), _dkey(thr)
        diff = kbest ^ kthr
        hb = diff.bit_length() - 1
        shift = -8 if hb < 0 else (hb // 8) * 8
        pmask = 0 if shift >= 56 else (0xFFFFFFFFFFFFFFFF << (shift + 8)) & 0xFFFFFFFFFFFFFFFF
        prefix = kbest & pmask
        need = beam_width
        while shift >= 0:
            hist = [0] * 256
            for k, _ in alive:
                if (k & pmask) == prefix:
                    hist[(k >> shift) & 255] += 1
            acc = 0
            for d in range(255, -1, -1):
                if need <= acc + hist[d]:
                    digit, need, cnt = d, need - acc, hist[d]
                    break
                acc += hist[d]
            prefix |= digit << shift
            pmask |= 0xFF << shift
            if (beam_width - need) + cnt <= beam_width + 32:
                break
            shift -= 8
        surv = [(k, rep) for k, rep in alive if (k & pmask) >= prefix]
        assert len(surv) >= min(beam_width, A)
    out = {}
    for ka, ia in surv:                                          # rank by counting
        k = sum(1 for ko, io in surv if ko > ka or (ko == ka and io < ia))
        if k < beam_width:
            assert k not in out
            out[k] = (ia, merged[ia])
    return [out[k] for k in range(len(out))]


def _frame(rng, n, n_states, tie_fraction, spread):
    state = rng.integers(0, n_states, size=n)
    salt = rng.integers(1, 2**63, size=n_states, dtype=np.uint64)
    keys = [int(salt[s]) for s in state]
    scores = (-spread * rng.random(n) - 20.0).tolist()
    if tie_fraction > 0:                                         # exact float64 ties (flat posteriors produce them)
        pool = [-20.0 - 0.25 * j for j in range(6)]
        for i in range(n):
            if rng.random() < tie_fraction:
                scores[i] = pool[int(rng.integers(0, len(pool)))]
    return keys, scores


@pytest.mark.parametrize("n,n_states,ties,spread", [
    (1, 1, 0.0, 1.0), (7, 3, 0.0, 5.0), (300, 200, 0.0, 12.0), (300, 40, 0.5, 12.0),
    (2048, 1500, 0.0, 8.0),          # wide frame: radix select over ~1500 alive states
    (2048, 1500, 0.0, 0.01),         # alive keys differ only in low mantissa bits
    (2048, 2048, 1.0, 8.0),          # hundreds of exact ties around the beam_width-th score
    (2048, 600, 0.3, 30.0),          # pruning removes most states
])
@pytest.mark.parametrize("beam_width", [1, 20, 128])
def test_hash_merge_and_counting_rank_equal_the_sort_based_definition(n, n_states, ties, spread, beam_width):
    rng = np.random.default_rng(1000 * n + 10 * beam_width + int(100 * ties))
    for _ in range(3):
        keys, scores = _frame(rng, n, n_states, ties, spread)
        want = by_sorting(keys, scores, -10.0, beam_width)
        got = by_hashing(keys, scores, -10.0, beam_width, rng)
        assert got == want                                       # same representatives, bit-equal float64 scores, same order
