// CTC prefix beam search, batched on the GPU (one CTA per utterance), without a language model or with an n-gram
// LM (KenLM trie uploaded by vasr_lm_create) fused into the search.
// Replaces BeamSearchDecoderWithLM.forward (nemo/collections/asr/beam_search_decoder.py:95-102), i.e. pyctcdecode's
// BeamSearchDecoderCTC.decode: with lm_path=None the mode infer.py:118-130 falls back to, with lm_path the default
// of infer.py:184-191 (3-gram-lm.binary, beam 100, alpha 0.5, beta 1.5).
// pyctcdecode and kenlm are un-vendored third-party packages: the algorithms are restated in oracle/beam_oracle.py
// and oracle/kenlm_oracle.py (PARITY UNPINNED against the packages themselves) and this kernel is tested against
// those restatements.
//
// LM fusion (template parameter LM): every beam also carries the LM score of its committed text, the KenLM context
// (last order-1 word ids), and a rolling hash + length of the partial word.  When ' ' is a candidate symbol the
// partial word of every beam is looked up in the vocabulary table and scored once per frame (commit_*); candidates
// are ranked by acoustic + lm(text) + partial-word penalty, beams keep the acoustic score.  At the end of the utterance
// every final text is scored once more as the end of the sentence (pyctcdecode >= 0.5 keys its LM cache by
// (text, is_eos)): a pending partial word is committed with the </s> term, a text without one gets the </s> term alone.
//
// A beam is a CTC state (prefix, last_char).  Prefixes are identified by a 64-bit rolling hash of their symbol
// sequence (merging = equal hash + equal last_char); the text is recovered at the end by back-tracing per-frame
// (parent beam, appended symbol) records.  Per frame: candidate symbols {logp >= token_min_logp} U {argmax} in
// ascending index, expansion in (symbol, beam) order, merge by log-sum-exp in first-seen order (hash table in shared
// memory, members reduced in insertion order), prune at best + beam_prune_logp, keep the beam_width best (ties keep
// first-seen order; rank by counting).  Round 1 did the merge and the ranking with two bitonic sorts per frame
// (~100 block barriers, 49 us per frame); this version has 6 barriers per frame.
#include "common.cuh"
#include "kernels.cuh"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

namespace vasr {
namespace beam {

constexpr int BW_MAX = 128;             // beam width limit
constexpr int MC = 16;                  // candidate symbols per frame (blank included)
constexpr int NC_MAX = BW_MAX * MC;     // expansion limit per frame
#ifndef VASR_BEAM_THREADS
#define VASR_BEAM_THREADS 256
#endif
// 256 and 512 threads measure within 3 % of each other (512 is ahead on wide frames - 128 beams x 16 symbols -, 256 on
// the narrow frames of real speech: fewer warps per barrier); 2 CTAs per SM either way (shared memory)
constexpr int THREADS = VASR_BEAM_THREADS;
constexpr unsigned long long H0 = 0x9E3779B97F4A7C15ull;
constexpr int SYM_NONE = 255, KEY_BLANK = 254, KEY_NONE = 255;

__host__ __device__ __forceinline__ unsigned long long mix(unsigned long long h, int c)
{
    h ^= (unsigned long long)(c + 1) * 0xD6E8FEB86659FD93ull;
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 33;
    h *= 0xC4CEB9FE1A85EC53ull;
    h ^= h >> 29;
    return h;
}
__device__ __forceinline__ double logaddexp_d(double a, double b)
{
    // m + log(exp(a - m) + exp(b - m)) as pyctcdecode's _log_sum_exp writes it; one of the two terms is exp(0) = 1 exactly
    const double m = fmax(a, b);
    return m + log(1.0 + exp(fmin(a, b) - m));
}
// order-preserving map double -> u64 (ascending)
__device__ __forceinline__ unsigned long long dkey(double x)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

// ---------------------------------------------------------------------------------------------------------------------
// n-gram LM in HBM: reverse trie (path w_n -> w_{n-1} -> ...), de-quantised log10 prob / back-off per node
// (viet-asr_b200/kenlm_binary.py decodes the KenLM file; lm/model.cc GenericModel::FullScore is what lm_score restates)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int MAX_ORDER = 5;
struct DeviceLM {
    int order, vocab, bos, eos;
    const float* uni_prob; const float* uni_backoff; const uint32_t* uni_next;
    const int32_t* mid_word[MAX_ORDER - 2]; const float* mid_prob[MAX_ORDER - 2]; const float* mid_backoff[MAX_ORDER - 2];
    const uint32_t* mid_next[MAX_ORDER - 2];
    const int32_t* long_word; const float* long_prob;
    const unsigned long long* vkeys; const int32_t* vvals; int vmask;
};

__device__ __forceinline__ int find_word(const int32_t* __restrict__ words, uint32_t lo, uint32_t hi, int w)
{
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        const int x = __ldg(words + mid);
        if (x < w) lo = mid + 1;
        else if (x > w) hi = mid;
        else return (int)mid;
    }
    return -1;
}
// rev[0..n) = w_n, w_{n-1}, ...: (prob, backoff) of every n-gram suffix that exists; returns how many do (>= 1)
__device__ int lm_walk(const DeviceLM& lm, const int* rev, int n, float* probs, float* backoffs)
{
    const int w0 = rev[0];
    probs[0] = __ldg(lm.uni_prob + w0); backoffs[0] = __ldg(lm.uni_backoff + w0);
    uint32_t lo = __ldg(lm.uni_next + w0), hi = __ldg(lm.uni_next + w0 + 1);
    int L = 1;
    for (int depth = 1; depth < n && depth < lm.order; ++depth) {
        const int w = rev[depth];
        if (depth < lm.order - 1) {
            const int k = depth - 1;
            const int i = find_word(lm.mid_word[k], lo, hi, w);
            if (i < 0) break;
            probs[depth] = __ldg(lm.mid_prob[k] + i); backoffs[depth] = __ldg(lm.mid_backoff[k] + i);
            lo = __ldg(lm.mid_next[k] + i); hi = __ldg(lm.mid_next[k] + i + 1);
        } else {
            const int i = find_word(lm.long_word, lo, hi, w);
            if (i < 0) break;
            probs[depth] = __ldg(lm.long_prob + i); backoffs[depth] = 0.f;
        }
        ++L;
    }
    return L;
}
// log10 P(w | ctx) with back-off; ctx oldest -> newest, nctx <= order - 1
__device__ double lm_score(const DeviceLM& lm, const int* ctx, int nctx, int w)
{
    int rev[MAX_ORDER];
    float p[MAX_ORDER], bo[MAX_ORDER];
    rev[0] = w;
    for (int j = 0; j < nctx; ++j) rev[1 + j] = ctx[nctx - 1 - j];
    const int L = lm_walk(lm, rev, nctx + 1, p, bo);
    double prob = (double)p[L - 1];
    if (L <= nctx) {                                   // charge the back-off of the context n-grams that did not match
        const int CL = lm_walk(lm, rev + 1, nctx, p, bo);
        for (int j = L; j <= CL; ++j) prob = __dadd_rn(prob, (double)bo[j - 1]);
    }
    return prob;
}
__device__ __forceinline__ int vocab_lookup(const DeviceLM& lm, unsigned long long h)
{
    int p = (int)(h & (unsigned long long)lm.vmask);
    while (true) {
        const unsigned long long k = __ldg(lm.vkeys + p);
        if (k == h) return __ldg(lm.vvals + p);
        if (k == 0ull) return -1;
        p = (p + 1) & lm.vmask;
    }
}
struct LmParams {
    double alpha, beta, unk, log10_to_ln;
};
// lm(text + word) from lm(text): (prev + alpha * raw * ln10) + beta, raw in log10 units (pyctcdecode LanguageModel.score)
__device__ __forceinline__ double lm_accumulate(const LmParams& q, double prev, double raw)
{
    return __dadd_rn(__dadd_rn(prev, __dmul_rn(__dmul_rn(q.alpha, raw), q.log10_to_ln)), q.beta);
}
// </s> on a text that has no pending word: alpha * ln P(</s> | context), no word-insertion bonus (no word is added)
__device__ __forceinline__ double lm_accumulate_eos(const LmParams& q, double prev, double raw_eos)
{
    return __dadd_rn(prev, __dmul_rn(__dmul_rn(q.alpha, raw_eos), q.log10_to_ln));
}
__device__ __forceinline__ double partial_penalty(const LmParams& q, int wl)      // score_partial_token, char trie absent
{
    if (wl == 0) return 0.0;
    return wl > 6 ? __ddiv_rn(__dmul_rn(q.unk, (double)wl), 6.0) : q.unk;
}

// the state of the (at most beam_width) beams between two frames; two sets alternate (frame t reads set t & 1 and
// writes the other), so a frame needs no copy-back pass
struct BeamSet {
    unsigned long long hash[BW_MAX], fhash[BW_MAX];       // fhash: hash of the text without a trailing space
    double score[BW_MAX];
    unsigned char lastsym[BW_MAX], lastkey[BW_MAX];
    // LM state per beam (unused without LM)
    double lmscore[BW_MAX];
    unsigned long long wph[BW_MAX];
    int wplen[BW_MAX];
    int ctx[BW_MAX][MAX_ORDER - 1];
    int nctx[BW_MAX];
};
constexpr unsigned EMPTY = 0xFFFFu;      // > every insertion index (< NC_MAX = 2048)
constexpr int TABLE_MAX = 2 * NC_MAX;    // merge table: open addressing at a load factor <= 1/2
struct Smem {
    BeamSet bs[2];
    double commit_lm[BW_MAX];
    int commit_ctx[BW_MAX][MAX_ORDER - 1];
    int commit_nctx[BW_MAX];
    int has_space;
    __align__(16) float lp[128];
    int cand[MC];
    float candp[MC];
    int ncand, nbeam, nalive, nsurv, nmerge;
    int wcnt[THREADS / 32];
    int sel_digit, sel_need, sel_cnt;    // radix select: digit of the bucket holding the beam_width-th key, what is still needed from it, its size
    unsigned int hist[256];
    double wbest[THREADS / 32];
    // by insertion index (candidate = symbol-major (cand, beam) pair)
    union alignas(16) {
        unsigned long long ckey[NC_MAX]; // merge key of the candidate's state (phases 2-3)
        double ucomb[NC_MAX];            // ranking score of a merged state, at its first-seen candidate (phases 4-5)
        unsigned long long surv_key[NC_MAX];    // phase 6, many states alive: the ones the radix select kept
    };
    double cscore[NC_MAX];               // acoustic score; after the merge the first-seen candidate holds the merged score
    unsigned int cmeta[NC_MAX];          // src beam | appended sym << 8 | lastsym << 16 | lastkey << 24
    union {
        unsigned short cslot[NC_MAX];    // table slot of the candidate's state (phases 3-4)
        unsigned short surv_idx[NC_MAX]; // phase 6: first-seen candidate of the states the radix select kept
    };
    unsigned short next[NC_MAX];         // next candidate in the slot's member list (EMPTY = end)
    unsigned short mlist[NC_MAX];        // first-seen members of the states that have more than one member
    union alignas(16) {
        unsigned int table[TABLE_MAX];   // slot -> a member of the state stored there, in the end its FIRST-SEEN member
        unsigned long long alive_key[NC_MAX];   // after the merge: order-preserving keys of the states that survive pruning
    };
    union alignas(16) {
        unsigned int head[TABLE_MAX];    // slot -> most recently pushed member
        unsigned short alive_idx[NC_MAX];       // after the merge: first-seen candidate of every surviving state
    };
};

// developer builds (make dev): cycles per phase of CTA 0, printed by the launcher when VASR_BEAM_PROF is set
#ifdef VASR_DEV
__device__ unsigned long long g_beam_prof[16];
#define BPROF_DECL() long long _bt = clock64(); unsigned long long _bacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define BPROF(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long _n = clock64(); _bacc[i] += (unsigned long long)(_n - _bt); _bt = _n; } } while (0)
#define BPROF_FLUSH() do { if (blockIdx.x == 0 && threadIdx.x == 0) for (int _i = 0; _i < 16; ++_i) g_beam_prof[_i] += _bacc[_i]; } while (0)
#else
#define BPROF_DECL() do {} while (0)
#define BPROF(i) do {} while (0)
#define BPROF_FLUSH() do {} while (0)
#endif

// One CTA per utterance, T_e sequential frames.  A frame is six block-wide phases (no sort):
//   1. warp 0 clips the frame's log-probs and picks the candidate symbols; the other warps clear the merge table
//   2. expansion of (symbol, beam) pairs into candidate states (+ LM: the word every beam would commit on ' ')
//   3. merge: every candidate inserts its 64-bit state key into an open-addressing table in shared memory; the slot
//      keeps the lowest insertion index (atomicMin) and a linked list of its members
//   4. the first-seen member of every state log-sum-exps the members IN INSERTION ORDER (what a sort by
//      (key, insertion index) followed by a run-wise reduction gives - the order matters for bit-equal float64 sums)
//   5. states within beam_prune_logp of the best are collected; each finds its rank by counting the states that
//      precede it in (score descending, first-seen ascending) order - O(A^2 / 256) compares on shared-memory
//      broadcasts, A ~ 100-400 - which is the position a stable sort would give it
//   6. ranks < beam_width become the beams of the next frame (other beam set) and write their back-pointers
template <bool LM>
__global__ void __launch_bounds__(THREADS, 2)
beam_kernel(const float* __restrict__ logp, const int* __restrict__ frames, int T, int V1, int blank, int space_id, int beam_width,
            float tok_min, float prune, unsigned char* __restrict__ bp_parent, unsigned char* __restrict__ bp_sym,
            int* __restrict__ out_ids, int* __restrict__ out_len, float* __restrict__ out_score,
            const DeviceLM lm, const LmParams q)
{
    extern __shared__ __align__(16) unsigned char raw[];
    Smem& s = *reinterpret_cast<Smem*>(raw);
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* lpb = logp + (size_t)b * T * V1;
    unsigned char* bpp = bp_parent + (size_t)b * T * BW_MAX;
    unsigned char* bps = bp_sym + (size_t)b * T * BW_MAX;
    const float clip_lo = logf(1e-15f);
    const int nkeep = LM ? lm.order - 1 : 0;               // KenLM context length
    // frames of THIS utterance: a zero-padded batch must decode every utterance over the frames the reference sees
    // when it runs that utterance alone (beam_search_decoder.py:96 asserts batch size 1), not over the padding
    const int Tb = frames ? min(max(frames[b], 0), T) : T;

    if (tid == 0) {
        BeamSet& z = s.bs[0];
        z.hash[0] = H0; z.fhash[0] = H0; z.score[0] = 0.0; z.lastsym[0] = SYM_NONE; z.lastkey[0] = KEY_NONE; s.nbeam = 1;
        if (LM) {                                           // start state: <s> (score_boundary=True)
            z.lmscore[0] = 0.0; z.wph[0] = H0; z.wplen[0] = 0;
            z.nctx[0] = nkeep > 0 ? 1 : 0; z.ctx[0][0] = lm.bos;
        }
    }
    __syncthreads();

    // the word a beam would commit if ' ' came next: id lookup, LM score, new context (one thread per beam)
    auto commit_word = [&](const BeamSet& c, int i, bool eos) {
        const int wid_lm = vocab_lookup(lm, c.wph[i]);
        const int w = wid_lm < 0 ? 0 : wid_lm;             // <unk> = 0
        double r = lm_score(lm, c.ctx[i], c.nctx[i], w);
        if (wid_lm < 0) r = __dadd_rn(r, q.unk);            // `word not in kenlm_model`
        int nc = 0;
        if (nkeep > 0) {
            const int have = c.nctx[i];
            const int drop = have + 1 > nkeep ? have + 1 - nkeep : 0;
            for (int j = drop; j < have; ++j) s.commit_ctx[i][nc++] = c.ctx[i][j];
            s.commit_ctx[i][nc++] = w;
        }
        s.commit_nctx[i] = nc;
        if (eos) r = __dadd_rn(r, lm_score(lm, s.commit_ctx[i], nc, lm.eos));
        s.commit_lm[i] = lm_accumulate(q, c.lmscore[i], r);
    };

    BPROF_DECL();
    if (tid >= V1 && tid < 128) s.lp[tid] = -INFINITY;                     // padding of the symbol table (never rewritten)
    float lp_next = (tid < V1 && Tb > 0) ? lpb[tid] : 0.f;                 // log-probs of the next frame, one symbol per thread
    for (int t = 0; t < Tb; ++t) {
        BPROF(15);
        const BeamSet& cur = s.bs[t & 1];
        BeamSet& nxt = s.bs[(t & 1) ^ 1];
        // ---- 1. frame log-probs (clipped like log(clip(p, 1e-15, 1))) and candidate symbols --------------------
        // thread v owns symbol v.  The argmax (ties -> lowest index, np.argmax) is always a candidate; the others need
        // logp >= token_min_logp, and if more than MC symbols qualify only the argmax and the MC - 1 most probable
        // others are kept, ties to the lower index (documented kernel limit; a peaked CTC posterior never reaches it).
        // Every thread ranks its own symbol against the V1 others on shared-memory broadcasts.
        if (tid < V1) s.lp[tid] = fminf(fmaxf(lp_next, clip_lo), 0.f);
        if (tid < V1 && t + 1 < Tb) lp_next = lpb[(size_t)(t + 1) * V1 + tid];      // in flight during this frame
        if (tid == 0) { s.nalive = 0; s.nsurv = 0; s.nmerge = 0; s.has_space = 0; }
        {
            // clear the part of the merge table this frame can touch (2 x the next power of two of nbeam * MC
            // bounds 2 x Np below)
            int cap = 1; while (cap < s.nbeam * MC) cap <<= 1;
            cap *= 2;
            const uint4 e4 = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
            uint4* t4 = reinterpret_cast<uint4*>(s.table);
            uint4* h4 = reinterpret_cast<uint4*>(s.head);
            for (int i = tid; i < cap / 4; i += THREADS) { t4[i] = e4; h4[i] = e4; }
        }
        __syncthreads();
        BPROF(0);
        // symbols with logp >= token_min_logp; when there are between 1 and MC of them they ARE the candidates (the
        // argmax is one of them) and nothing has to be ranked - the usual case on speech
        float x = -INFINITY;
        if (tid < V1) x = s.lp[tid];
        const bool qual = tid < V1 && x >= tok_min;
        const int Q = __syncthreads_count(qual ? 1 : 0);
        bool sel = qual;
        if (Q < 1 || Q > MC) {
            sel = false;
            if (tid < V1 && (qual || Q < 1)) {
                int bf[4] = {0, 0, 0, 0}, bq[4] = {0, 0, 0, 0};   // symbols that precede this one in (logp descending, index ascending) order
                const float4* lp4 = reinterpret_cast<const float4*>(s.lp);       // entries [V1, 128) are -inf: they precede nothing
                for (int u4 = 0; u4 < (V1 + 3) / 4; ++u4) {
                    const float4 y4 = lp4[u4];
                    const float y[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool prec = (y[j] > x) || (y[j] == x && u4 * 4 + j < tid);
                        bf[j] += prec ? 1 : 0;
                        bq[j] += (prec && y[j] >= tok_min) ? 1 : 0;
                    }
                }
                const int before = bf[0] + bf[1] + bf[2] + bf[3], before_q = bq[0] + bq[1] + bq[2] + bq[3];
                // before == 0: the argmax.  Otherwise before_q counts the argmax too (it qualifies whenever anything does)
                sel = before == 0 || (qual && before_q <= MC - 1);
            }
        }
        BPROF(1);
        const unsigned selbal = __ballot_sync(0xffffffffu, sel);
        if (lane == 0) s.wcnt[wid] = __popc(selbal);
        if (sel && tid == space_id) s.has_space = 1;
        __syncthreads();
        BPROF(2);
        if (sel) {                                 // candidates in ascending index
            int pos = __popc(selbal & ((1u << lane) - 1u));
            for (int w = 0; w < wid; ++w) pos += s.wcnt[w];
            s.cand[pos] = tid; s.candp[pos] = x;
        }
        if (tid == 0) { int c = 0; for (int w = 0; w < THREADS / 32; ++w) c += s.wcnt[w]; s.ncand = c; }
        __syncthreads();
        BPROF(3);
        const int n = s.nbeam, m = s.ncand, N = n * m;
        int Np = 1; while (Np < N) Np <<= 1;
        const unsigned mask = (unsigned)(2 * Np - 1);

        // ---- 2. LM: score the word every beam would commit on ' ' (once per beam and frame) --------------------
        if (LM && s.has_space && tid < n && cur.wplen[tid] > 0) commit_word(cur, tid, false);
        BPROF(4);

        // ---- 2. expansion, insertion index = cand * n + beam (symbol-major like the reference loop) ------------
        for (int idx = tid; idx < N; idx += THREADS) {
            const int j = idx / n, i = idx - j * n;
            const int c = s.cand[j];
            const unsigned long long h = cur.hash[i];
            const int ls = cur.lastsym[i], lk = cur.lastkey[i];
            unsigned long long nh = h; int nls = ls, nlk, app = SYM_NONE;
            if (c == blank) nlk = KEY_BLANK;
            else if (lk == c) nlk = c;                                     // repeat of the last emitted symbol
            else if (c == space_id) {
                if (h == H0 || ls == space_id) { nls = space_id; nlk = space_id; }   // leading / repeated space: no new word
                else { nh = mix(h, c); nls = c; nlk = c; app = c; }
            } else { nh = mix(h, c); nls = c; nlk = c; app = c; }
            s.ckey[idx] = mix(nh, nlk);
            s.cscore[idx] = cur.score[i] + (double)s.candp[j];
            s.cmeta[idx] = (unsigned)i | ((unsigned)app << 8) | ((unsigned)nls << 16) | ((unsigned)nlk << 24);
        }
        __syncthreads();
        BPROF(5);

        // ---- 3. merge equal states: insert into the table; the slot ends up holding the first-seen member -------
        // (a plain read first: members that join an existing state - and every probe past an occupied slot - cost no
        // atomic; the shared-memory atomics are what bounds this phase.  Issuing a thread's candidates as a batch of
        // unconditional atomicCAS was measured 40 % slower, profiles/r2_experiments.md)
        for (int idx = tid; idx < N; idx += THREADS) {
            const unsigned long long key = s.ckey[idx];
            unsigned p = (unsigned)(key >> 20) & mask;
            while (true) {
                unsigned e = *reinterpret_cast<volatile unsigned*>(&s.table[p]);
                if (e == EMPTY) {
                    e = atomicCAS(&s.table[p], EMPTY, (unsigned)idx);
                    if (e == EMPTY) break;                                 // claimed an empty slot
                }
                if (s.ckey[e] == key) { atomicMin(&s.table[p], (unsigned)idx); break; }   // any member carries the key
                p = (p + 1) & mask;
            }
            s.cslot[idx] = (unsigned short)p;
            s.next[idx] = (unsigned short)atomicExch(&s.head[p], (unsigned)idx);
        }
        __syncthreads();
        BPROF(6);

        // ---- 4. first-seen members: log-sum-exp of the members in insertion order, ranking score ----------------
        // ranking score: acoustic, plus lm(text) and the partial-word penalty of the state when an LM is fused
        auto finish = [&](int idx, double acc) -> double {
            double comb = acc;
            if (LM) {
                const unsigned meta = s.cmeta[idx];
                const int src = meta & 0xff, app = (meta >> 8) & 0xff;
                double lmv; int wl;
                if (app == space_id) { lmv = s.commit_lm[src]; wl = 0; }
                else { lmv = cur.lmscore[src]; wl = cur.wplen[src] + (app != SYM_NONE ? 1 : 0); }
                comb = __dadd_rn(acc, __dadd_rn(lmv, partial_penalty(q, wl)));
            }
            s.ucomb[idx] = comb;
            return comb;
        };
        unsigned own = 0;                                                  // bit k: candidate tid + k * THREADS is a first-seen member
        double best = -INFINITY;
        for (int idx = tid, k = 0; idx < N; idx += THREADS, ++k) {
            const unsigned p = s.cslot[idx];
            if (s.table[p] != (unsigned)idx) continue;
            own |= 1u << k;
            if (s.head[p] == (unsigned)idx && s.next[idx] == EMPTY) best = fmax(best, finish(idx, s.cscore[idx]));   // the only member
            else s.mlist[atomicAdd(&s.nmerge, 1)] = (unsigned short)idx;    // several members: the float64 log-sum-exp chains are
        }                                                                   // spread evenly over the threads below
        __syncthreads();
        const int M = s.nmerge;
        for (int w = tid; w < M; w += THREADS) {
            const int idx = s.mlist[w];
            const unsigned p = s.cslot[idx];
            double acc = s.cscore[idx];
            int last = idx;
            while (true) {                                                 // next member in ascending insertion index
                int pick = 0x7fffffff;
                for (unsigned e = s.head[p]; e != EMPTY; e = s.next[e])
                    if ((int)e > last && (int)e < pick) pick = (int)e;
                if (pick == 0x7fffffff) break;
                acc = logaddexp_d(acc, s.cscore[pick]);
                last = pick;
            }
            s.cscore[idx] = acc;
            best = fmax(best, finish(idx, acc));
        }
        for (int o = 16; o >= 1; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (lane == 0) s.wbest[wid] = best;
        __syncthreads();
        best = s.wbest[0];
        for (int w = 1; w < THREADS / 32; ++w) best = fmax(best, s.wbest[w]);
        BPROF(7);

        // ---- 5. prune at best + beam_prune_logp -----------------------------------------------------------------
        const double thr = best + (double)prune;
        for (int idx = tid, k = 0; idx < N; idx += THREADS, ++k) {
            if (!((own >> k) & 1u)) continue;
            const double c = s.ucomb[idx];
            if (c >= thr) {
                const int a = atomicAdd(&s.nalive, 1);
                s.alive_key[a] = dkey(c);
                s.alive_idx[a] = (unsigned short)idx;
            }
        }
        __syncthreads();
        BPROF(8);

        // ---- 6. rank by (score descending, first-seen ascending); the beam_width best become the new beams ------
        const int A = s.nalive;
        const bool many = A > 2 * BW_MAX;
        if (many) {
            // many states alive (diffuse posteriors): counting ranks among all of them is O(A^2).  Radix-select the
            // key of the beam_width-th best state first, 8 bits at a time from the highest bit in which the alive
            // keys can differ (they all lie between the keys of thr and best), and rank only the states at or above
            // it.  The select stops as soon as the bucket that holds the beam_width-th key plus everything above it
            // is at most beam_width + 32 states; states with equal scores stay together, the counting below orders them.
            const unsigned long long kbest = dkey(best), kthr = dkey(thr);
            const int hb = 63 - __clzll((long long)(kbest ^ kthr));         // -1: all alive keys are equal
            int shift = hb < 0 ? -8 : (hb / 8) * 8;
            unsigned long long pmask = shift >= 56 ? 0ull : (~0ull << (shift + 8));
            unsigned long long prefix = kbest & pmask;
            static_assert(THREADS >= 256, "one histogram bin per thread");
            if (tid == 0) s.sel_need = beam_width;
            for (; shift >= 0; shift -= 8) {
                if (tid < 256) s.hist[tid] = 0u;
                __syncthreads();
                for (int a = tid; a < A; a += THREADS) {
                    const unsigned long long ka = s.alive_key[a];
                    if ((ka & pmask) == prefix) atomicAdd(&s.hist[(unsigned)(ka >> shift) & 255u], 1u);
                }
                __syncthreads();
                if (wid == 0) {
                    unsigned c = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) c += s.hist[lane * 8 + j];
                    unsigned above = c;                                      // inclusive suffix sum over the lanes
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned v = __shfl_down_sync(0xffffffffu, above, o);
                        if (lane + o < 32) above += v;
                    }
                    above -= c;                                              // states in higher buckets than this lane's
                    const unsigned need = (unsigned)s.sel_need;
                    __syncwarp();
                    if (above < need && need <= above + c) {                 // exactly one lane
                        unsigned acc = above;
                        for (int j = 7; j >= 0; --j) {
                            const unsigned hcnt = s.hist[lane * 8 + j];
                            if (need <= acc + hcnt) { s.sel_digit = lane * 8 + j; s.sel_need = (int)(need - acc); s.sel_cnt = (int)hcnt; break; }
                            acc += hcnt;
                        }
                    }
                }
                __syncthreads();
                prefix |= (unsigned long long)s.sel_digit << shift;
                pmask |= 0xFFull << shift;
                if ((beam_width - s.sel_need) + s.sel_cnt <= beam_width + 32) break;
            }
            for (int a = tid; a < A; a += THREADS) {
                const unsigned long long ka = s.alive_key[a];
                if ((ka & pmask) >= prefix) {
                    const int u = atomicAdd(&s.nsurv, 1);
                    s.surv_key[u] = ka;
                    s.surv_idx[u] = s.alive_idx[a];
                }
            }
            __syncthreads();
        }
        BPROF(9);
        auto rank_and_fill = [&](const unsigned long long* rkey, const unsigned short* ridx, const int R) {
        for (int a = tid; a < R; a += THREADS) {
            const unsigned long long ka = rkey[a];
            const unsigned ia = ridx[a];
            // states ahead of this one: larger key, or equal key and seen earlier.  Equal float64 scores are rare, so the
            // main loop only counts larger and equal keys (two keys per 16-byte load, independent counters)
            int g0 = 0, g1 = 0, e0 = 0, e1 = 0;
            const ulonglong2* rk2 = reinterpret_cast<const ulonglong2*>(rkey);
            for (int o = 0; o < R / 2; ++o) {
                const ulonglong2 kk = rk2[o];
                g0 += kk.x > ka ? 1 : 0; e0 += kk.x == ka ? 1 : 0;
                g1 += kk.y > ka ? 1 : 0; e1 += kk.y == ka ? 1 : 0;
            }
            if (R & 1) { const unsigned long long ko = rkey[R - 1]; g0 += ko > ka ? 1 : 0; e0 += ko == ka ? 1 : 0; }
            int k = g0 + g1;
            if (e0 + e1 > 1)
                for (int o = 0; o < R; ++o) k += (rkey[o] == ka && ridx[o] < ia) ? 1 : 0;
            if (k >= beam_width) continue;
            const unsigned meta = s.cmeta[ia];
            const int src = meta & 0xff, app = (meta >> 8) & 0xff;
            const unsigned long long nh = (app == SYM_NONE) ? cur.hash[src] : mix(cur.hash[src], app);
            nxt.hash[k] = nh;
            // text without a trailing space: unchanged by a committed ' ' (the source cannot end in one)
            nxt.fhash[k] = (app == SYM_NONE) ? cur.fhash[src] : (app == space_id ? cur.hash[src] : nh);
            nxt.score[k] = s.cscore[ia];
            nxt.lastsym[k] = (unsigned char)((meta >> 16) & 0xff);
            nxt.lastkey[k] = (unsigned char)((meta >> 24) & 0xff);
            if (LM) {
                if (app == space_id) {
                    nxt.lmscore[k] = s.commit_lm[src]; nxt.wph[k] = H0; nxt.wplen[k] = 0;
                    nxt.nctx[k] = s.commit_nctx[src];
                    for (int j = 0; j < MAX_ORDER - 1; ++j) nxt.ctx[k][j] = s.commit_ctx[src][j];
                } else {
                    nxt.lmscore[k] = cur.lmscore[src];
                    nxt.wph[k] = (app == SYM_NONE) ? cur.wph[src] : mix(cur.wph[src], app);
                    nxt.wplen[k] = cur.wplen[src] + (app != SYM_NONE ? 1 : 0);
                    nxt.nctx[k] = cur.nctx[src];
                    for (int j = 0; j < MAX_ORDER - 1; ++j) nxt.ctx[k][j] = cur.ctx[src][j];
                }
            }
            bpp[(size_t)t * BW_MAX + k] = (unsigned char)src;
            bps[(size_t)t * BW_MAX + k] = (unsigned char)app;
        }
        };
        if (many) rank_and_fill(s.surv_key, s.surv_idx, s.nsurv);
        else rank_and_fill(s.alive_key, s.alive_idx, A);
        BPROF(10);
        if (tid == 0) s.nbeam = A < beam_width ? A : beam_width;
        __syncthreads();
        BPROF(11);
    }
    BPROF_FLUSH();
    const BeamSet& fin = s.bs[Tb & 1];

    // ---- end of utterance: the partial word joins the text; states with equal text merge (first-seen order);
    //      with an LM every final text is scored with the </s> term, the best text wins; back-trace ---------------
    if (LM && tid < s.nbeam) {
        // every final text is scored as the END of the sentence: pyctcdecode >= 0.5 keys its LM cache by
        // (text, is_eos), so a text that was already committed during the search is scored again with </s> at the end
        if (fin.wplen[tid] > 0) commit_word(fin, tid, true);
        else s.commit_lm[tid] = lm_accumulate_eos(q, fin.lmscore[tid], lm_score(lm, fin.ctx[tid], fin.nctx[tid], lm.eos));
    }
    __syncthreads();
    if (tid == 0) {
        const int n = s.nbeam;
        int bi = 0; double bs = -INFINITY;
        for (int i = 0; i < n; ++i) {
            bool first = true;
            for (int qq = 0; qq < i; ++qq) if (fin.fhash[qq] == fin.fhash[i]) { first = false; break; }
            if (!first) continue;
            double acc = fin.score[i];
            for (int qq = i + 1; qq < n; ++qq) if (fin.fhash[qq] == fin.fhash[i]) acc = logaddexp_d(acc, fin.score[qq]);
            if (LM) acc = __dadd_rn(acc, s.commit_lm[i]);
            if (acc > bs) { bs = acc; bi = i; }
        }
        int* oid = out_ids + (size_t)b * T;
        int len = 0, k = bi;
        for (int t = Tb - 1; t >= 0; --t) {
            const int app = bps[(size_t)t * BW_MAX + k];
            if (app != SYM_NONE) oid[len++] = app;         // reversed
            k = bpp[(size_t)t * BW_MAX + k];
        }
        for (int i = 0; i < len / 2; ++i) { const int x = oid[i]; oid[i] = oid[len - 1 - i]; oid[len - 1 - i] = x; }
        if (len > 0 && oid[len - 1] == space_id) --len;    // whitespace-normalised text: no trailing space
        for (int i = len; i < T; ++i) oid[i] = -1;
        out_len[b] = len;
        if (out_score) out_score[b] = (float)bs;
    }
}

// batched LM queries (parity tests of the device trie walk): ctx [N, MAX_ORDER-1] oldest -> newest, nctx [N], word [N]
__global__ void lm_score_kernel(const DeviceLM lm, const int* __restrict__ ctx, const int* __restrict__ nctx,
                                const int* __restrict__ word, double* __restrict__ out, int N)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int c[MAX_ORDER - 1];
    for (int j = 0; j < MAX_ORDER - 1; ++j) c[j] = ctx[(size_t)i * (MAX_ORDER - 1) + j];
    out[i] = lm_score(lm, c, nctx[i], word[i]);
}

}  // namespace beam
}  // namespace vasr

// ---------------------------------------------------------------------------------------------------------------------
// host side: LM handle + entry points
// ---------------------------------------------------------------------------------------------------------------------
struct vasr_lm {
    vasr::beam::DeviceLM dev{};
    std::vector<void*> allocs;
    int device = 0;
};

namespace {
template <typename T>
int upload(vasr_lm* h, const T* src, size_t n, const T** dst)
{
    void* p = nullptr;
    VASR_CUDA_OK(cudaMalloc(&p, n * sizeof(T) + 16));
    h->allocs.push_back(p);
    VASR_CUDA_OK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = (const T*)p;
    return VASR_OK;
}
}  // namespace

extern "C" uint64_t vasr_lm_hash_labels(const int32_t* label_ids, int n)
{
    unsigned long long h = vasr::beam::H0;
    for (int i = 0; i < n; ++i) h = vasr::beam::mix(h, label_ids[i]);
    return h;
}

extern "C" void vasr_lm_destroy(vasr_lm* h)
{
    if (!h) return;
    for (void* p : h->allocs) cudaFree(p);
    delete h;
}

extern "C" int vasr_lm_create(const vasr_lm_arrays* a, vasr_lm** out)
{
    using namespace vasr;
    using namespace vasr::beam;
    VASR_REQUIRE(a && out, "vasr_lm_create: null argument");
    VASR_REQUIRE(a->order >= 2 && a->order <= MAX_ORDER, "vasr_lm_create: n-gram order must be in [2, %d] (got %d)", MAX_ORDER, a->order);
    VASR_REQUIRE(a->counts && a->uni_prob && a->uni_backoff && a->uni_next && a->long_word && a->long_prob,
                 "vasr_lm_create: null array");
    VASR_REQUIRE(a->vocab > 0 && (uint64_t)a->vocab == a->counts[0], "vasr_lm_create: vocab does not match counts[0]");
    VASR_REQUIRE(a->bos >= 0 && a->bos < a->vocab && a->eos >= 0 && a->eos < a->vocab, "vasr_lm_create: <s>/</s> id out of range");
    VASR_REQUIRE(a->vocab_keys && a->vocab_vals && a->vocab_slots > 0 && (a->vocab_slots & (a->vocab_slots - 1)) == 0,
                 "vasr_lm_create: vocabulary table must have a power-of-two number of slots");
    for (int k = 0; k < a->order - 2; ++k)
        VASR_REQUIRE(a->mid_word[k] && a->mid_prob[k] && a->mid_backoff[k] && a->mid_next[k], "vasr_lm_create: null array for order %d", k + 2);
    for (int k = 1; k < a->order; ++k)
        VASR_REQUIRE(a->counts[k] < 0xffffffffull, "vasr_lm_create: more than 2^32 n-grams of order %d", k + 1);
    vasr_lm* h = new vasr_lm();
    cudaGetDevice(&h->device);
    DeviceLM& d = h->dev;
    d.order = a->order; d.vocab = a->vocab; d.bos = a->bos; d.eos = a->eos; d.vmask = a->vocab_slots - 1;
    int rc = VASR_OK;
    auto fail = [&](int code) { vasr_lm_destroy(h); return code; };
    if ((rc = upload(h, a->uni_prob, (size_t)a->vocab, &d.uni_prob))) return fail(rc);
    if ((rc = upload(h, a->uni_backoff, (size_t)a->vocab, &d.uni_backoff))) return fail(rc);
    if ((rc = upload(h, a->uni_next, (size_t)a->vocab + 1, &d.uni_next))) return fail(rc);
    for (int k = 0; k < a->order - 2; ++k) {
        const size_t n = (size_t)a->counts[k + 1];
        if ((rc = upload(h, a->mid_word[k], n, &d.mid_word[k]))) return fail(rc);
        if ((rc = upload(h, a->mid_prob[k], n, &d.mid_prob[k]))) return fail(rc);
        if ((rc = upload(h, a->mid_backoff[k], n, &d.mid_backoff[k]))) return fail(rc);
        if ((rc = upload(h, a->mid_next[k], n + 1, &d.mid_next[k]))) return fail(rc);
    }
    const size_t nl = (size_t)a->counts[a->order - 1];
    if ((rc = upload(h, a->long_word, nl, &d.long_word))) return fail(rc);
    if ((rc = upload(h, a->long_prob, nl, &d.long_prob))) return fail(rc);
    const unsigned long long* vk = nullptr;
    if ((rc = upload(h, (const unsigned long long*)a->vocab_keys, (size_t)a->vocab_slots, &vk))) return fail(rc);
    d.vkeys = vk;
    if ((rc = upload(h, a->vocab_vals, (size_t)a->vocab_slots, &d.vvals))) return fail(rc);
    *out = h;
    return VASR_OK;
}

extern "C" int vasr_lm_order(const vasr_lm* h) { return h ? h->dev.order : 0; }

extern "C" int vasr_lm_score_batch(const vasr_lm* h, const int32_t* ctx, const int32_t* nctx, const int32_t* word,
                                   double* out, int N, void* stream)
{
    using namespace vasr;
    VASR_REQUIRE(h && ctx && nctx && word && out, "vasr_lm_score_batch: null argument");
    VASR_REQUIRE(N > 0, "vasr_lm_score_batch: N must be positive (got %d)", N);
    vasr::beam::lm_score_kernel<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(h->dev, ctx, nctx, word, out, N);
    VASR_LAUNCH_OK("lm_score_kernel");
    return VASR_OK;
}

extern "C" size_t vasr_ctc_beam_workspace_bytes(int B, int T)
{
    if (B <= 0 || T <= 0) return 0;
    return (size_t)2 * B * T * vasr::beam::BW_MAX;
}

extern "C" size_t vasr_ctc_beam_lm_workspace_bytes(int B, int T, int beam_width)
{
    if (B <= 0 || T <= 0 || beam_width <= 0) return 0;
    return ((size_t)2 * B * T * vasr::beam::BW_MAX + 255) / 256 * 256;     // same back-pointer records as without LM
}

static int beam_launch(const float* log_probs, const int32_t* frames, int B, int T, int V1, int blank, int space_id, int beam_width,
                       float token_min_logp, float beam_prune_logp, const vasr_lm* lm, double alpha, double beta,
                       double unk_score_offset, void* workspace, size_t workspace_bytes,
                       int32_t* out_ids, int32_t* out_len, float* out_score, void* stream, const char* who)
{
    using namespace vasr;
    using namespace vasr::beam;
    VASR_REQUIRE(log_probs && workspace && out_ids && out_len, "%s: null argument", who);
    VASR_REQUIRE(B > 0 && T > 0, "%s: B and T must be positive (got %d, %d)", who, B, T);
    VASR_REQUIRE(V1 >= 2 && V1 <= 128, "%s: classes (+blank) must be in [2, 128] (got %d)", who, V1);
    VASR_REQUIRE(beam_width >= 1 && beam_width <= BW_MAX, "%s: beam_width must be in [1, %d] (got %d)", who, BW_MAX, beam_width);
    VASR_REQUIRE(blank >= 0 && blank < V1, "%s: blank id out of range", who);
    VASR_REQUIRE(space_id < V1, "%s: space id out of range", who);
    const size_t need = lm ? vasr_ctc_beam_lm_workspace_bytes(B, T, beam_width) : vasr_ctc_beam_workspace_bytes(B, T);
    if (workspace_bytes < need)
        return set_error(VASR_ENOMEM, "%s: workspace %zu < required %zu bytes", who, workspace_bytes, need);
    static bool attr_set = false;
    const size_t smem = sizeof(Smem);
    if (!attr_set) {
        VASR_CUDA_OK(cudaFuncSetAttribute(beam_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VASR_CUDA_OK(cudaFuncSetAttribute(beam_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* bp_parent = (unsigned char*)workspace;
    unsigned char* bp_sym = bp_parent + (size_t)B * T * BW_MAX;
    LmParams q{};
    if (lm) {
        q.alpha = alpha; q.beta = beta; q.unk = unk_score_offset; q.log10_to_ln = 1.0 / log10(M_E);
        beam_kernel<true><<<B, THREADS, smem, st>>>(log_probs, frames, T, V1, blank, space_id, beam_width, token_min_logp,
                                                    beam_prune_logp, bp_parent, bp_sym, out_ids, out_len, out_score, lm->dev, q);
    } else {
        beam_kernel<false><<<B, THREADS, smem, st>>>(log_probs, frames, T, V1, blank, space_id, beam_width, token_min_logp,
                                                     beam_prune_logp, bp_parent, bp_sym, out_ids, out_len, out_score, DeviceLM{}, q);
    }
    VASR_LAUNCH_OK("beam_kernel");
#ifdef VASR_DEV
    if (getenv("VASR_BEAM_PROF")) {
        unsigned long long h[16], z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_beam_prof, sizeof(h));
        cudaMemcpyToSymbol(g_beam_prof, z, sizeof(z));
        fprintf(stderr, "BEAMPROF B=%d T=%d lm=%d | kcycles of CTA 0: lp+clear %llu symbol-rank %llu sync %llu cand-list %llu | lm-commit %llu expand %llu insert %llu "
                        "merge %llu prune %llu select %llu rank+fill %llu end-sync %llu\n",
                B, T, lm ? 1 : 0, h[0] / 1000, h[1] / 1000, h[2] / 1000, h[3] / 1000, h[4] / 1000, h[5] / 1000, h[6] / 1000, h[7] / 1000,
                h[8] / 1000, h[9] / 1000, h[10] / 1000, h[11] / 1000);
    }
#endif
    return VASR_OK;
}

extern "C" int vasr_ctc_beam_search(const float* log_probs, const int32_t* frames, int B, int T, int V1, int blank, int space_id,
                                    int beam_width, float token_min_logp, float beam_prune_logp,
                                    void* workspace, size_t workspace_bytes,
                                    int32_t* out_ids, int32_t* out_len, float* out_score, void* stream)
{
    return beam_launch(log_probs, frames, B, T, V1, blank, space_id, beam_width, token_min_logp, beam_prune_logp, nullptr, 0.0, 0.0,
                       0.0, workspace, workspace_bytes, out_ids, out_len, out_score, stream, "vasr_ctc_beam_search");
}

extern "C" int vasr_ctc_beam_search_lm(const float* log_probs, const int32_t* frames, int B, int T, int V1, int blank, int space_id,
                                       int beam_width, float token_min_logp, float beam_prune_logp,
                                       const vasr_lm* lm, double alpha, double beta, double unk_score_offset,
                                       void* workspace, size_t workspace_bytes,
                                       int32_t* out_ids, int32_t* out_len, float* out_score, void* stream)
{
    VASR_REQUIRE(lm, "vasr_ctc_beam_search_lm: null language model");
    return beam_launch(log_probs, frames, B, T, V1, blank, space_id, beam_width, token_min_logp, beam_prune_logp, lm, alpha, beta,
                       unk_score_offset, workspace, workspace_bytes, out_ids, out_len, out_score, stream,
                       "vasr_ctc_beam_search_lm");
}
