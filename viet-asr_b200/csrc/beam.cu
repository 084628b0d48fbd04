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
// ascending index, expansion in (symbol, beam) order, merge by log-sum-exp in first-seen order (bitonic sort on
// (key, insertion index)), prune at best + beam_prune_logp, keep the beam_width best (ties keep first-seen order).
#include "common.cuh"
#include "kernels.cuh"
#include <math.h>
#include <vector>

namespace vasr {
namespace beam {

constexpr int BW_MAX = 128;             // beam width limit
constexpr int MC = 16;                  // candidate symbols per frame (blank included)
constexpr int NC_MAX = BW_MAX * MC;     // expansion limit per frame
constexpr int THREADS = 256;
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
    const double m = fmax(a, b);
    return m + log(exp(a - m) + exp(b - m));
}
// order-preserving map double -> u64 (ascending)
__device__ __forceinline__ unsigned long long dkey(double x)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}

// ascending bitonic sort of (k1, k2) pairs, n a power of two
__device__ void bitonic_sort(unsigned long long* k1, unsigned int* k2, int n)
{
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += THREADS) {
                const int x = i ^ j;
                if (x > i) {
                    const bool up = (i & k) == 0;
                    const unsigned long long a1 = k1[i], b1 = k1[x];
                    const unsigned int a2 = k2[i], b2 = k2[x];
                    const bool gt = (a1 > b1) || (a1 == b1 && a2 > b2);
                    if (gt == up) { k1[i] = b1; k1[x] = a1; k2[i] = b2; k2[x] = a2; }
                }
            }
            __syncthreads();
        }
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

struct Smem {
    unsigned long long hash[BW_MAX], fhash[BW_MAX];       // fhash: hash of the text without a trailing space
    double score[BW_MAX];
    unsigned char lastsym[BW_MAX], lastkey[BW_MAX];
    unsigned long long nhash[BW_MAX], nfhash[BW_MAX];
    double nscore[BW_MAX];
    unsigned char nlastsym[BW_MAX], nlastkey[BW_MAX];
    // LM state per beam (unused without LM)
    double lmscore[BW_MAX], nlmscore[BW_MAX], commit_lm[BW_MAX];
    unsigned long long wph[BW_MAX], nwph[BW_MAX];
    int wplen[BW_MAX], nwplen[BW_MAX];
    int ctx[BW_MAX][MAX_ORDER - 1], nctx_[BW_MAX][MAX_ORDER - 1], commit_ctx[BW_MAX][MAX_ORDER - 1];
    int nctx[BW_MAX], nnctx[BW_MAX], commit_nctx[BW_MAX];
    int has_space;
    double ucomb[NC_MAX];                // ranking score by unique slot
    float lp[128];
    int cand[MC];
    float candp[MC];
    int ncand, nbeam, nuniq, nkept;
    unsigned long long ckey[NC_MAX];     // merge key, later the ranking key
    unsigned int cidx[NC_MAX];           // insertion index
    double cscore[NC_MAX];               // by insertion index
    unsigned int cmeta[NC_MAX];          // by insertion index: src beam | appended sym << 8 | lastsym << 16 | lastkey << 24
    double uscore[NC_MAX];               // merged score by unique slot
    unsigned int urep[NC_MAX];           // representative insertion index by unique slot
    int scan[THREADS / 32];
};

template <bool LM>
__global__ void __launch_bounds__(THREADS)
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
        s.hash[0] = H0; s.fhash[0] = H0; s.score[0] = 0.0; s.lastsym[0] = SYM_NONE; s.lastkey[0] = KEY_NONE; s.nbeam = 1;
        if (LM) {                                           // start state: <s> (score_boundary=True)
            s.lmscore[0] = 0.0; s.wph[0] = H0; s.wplen[0] = 0;
            s.nctx[0] = nkeep > 0 ? 1 : 0; s.ctx[0][0] = lm.bos;
        }
    }
    __syncthreads();

    // the word a beam would commit if ' ' came next: id lookup, LM score, new context (one thread per beam)
    auto commit_word = [&](int i, bool eos) {
        const int wid_lm = vocab_lookup(lm, s.wph[i]);
        const int w = wid_lm < 0 ? 0 : wid_lm;             // <unk> = 0
        double r = lm_score(lm, s.ctx[i], s.nctx[i], w);
        if (wid_lm < 0) r = __dadd_rn(r, q.unk);            // `word not in kenlm_model`
        int nc = 0;
        if (nkeep > 0) {
            const int have = s.nctx[i];
            const int drop = have + 1 > nkeep ? have + 1 - nkeep : 0;
            for (int j = drop; j < have; ++j) s.commit_ctx[i][nc++] = s.ctx[i][j];
            s.commit_ctx[i][nc++] = w;
        }
        s.commit_nctx[i] = nc;
        if (eos) r = __dadd_rn(r, lm_score(lm, s.commit_ctx[i], nc, lm.eos));
        s.commit_lm[i] = lm_accumulate(q, s.lmscore[i], r);
    };

    for (int t = 0; t < Tb; ++t) {
        // ---- 1. frame log-probs (clipped like log(clip(p, 1e-15, 1))) and candidate symbols -------------------
        if (tid < V1) s.lp[tid] = fminf(fmaxf(lpb[(size_t)t * V1 + tid], clip_lo), 0.f);
        __syncthreads();
        if (wid == 0) {
            // argmax, ties -> lowest index (np.argmax)
            float best = -INFINITY; int bi = 0x7fffffff;
            for (int v = lane; v < V1; v += 32) { const float x = s.lp[v]; if (x > best) { best = x; bi = v; } }
            for (int o = 16; o >= 1; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            // candidates in ascending index; if more than MC qualify keep the argmax plus the MC-1 most probable
            // others, ties to the lower index (documented kernel limit; a peaked CTC posterior never reaches it)
            int cnt = 0;
            for (int v0 = 0; v0 < V1; v0 += 32) {
                const int v = v0 + lane;
                const bool f = v < V1 && (s.lp[v] >= tok_min || v == bi);
                cnt += __popc(__ballot_sync(0xffffffffu, f));
            }
            const bool capped = cnt > MC;
            auto selected = [&](int v) -> bool {
                if (v >= V1) return false;
                if (v == bi) return true;
                const float x = s.lp[v];
                if (x < tok_min) return false;
                if (!capped) return true;
                int rank = 0;                          // qualifying non-argmax symbols ahead of v
                for (int u = 0; u < V1; ++u) {
                    const float y = s.lp[u];
                    if (u != bi && y >= tok_min && (y > x || (y == x && u < v))) ++rank;
                }
                return rank < MC - 1;
            };
            int base = 0; bool sp = false;
            for (int v0 = 0; v0 < V1; v0 += 32) {
                const int v = v0 + lane;
                const bool f = selected(v);
                const unsigned m = __ballot_sync(0xffffffffu, f);
                if (f) { const int p = base + __popc(m & ((1u << lane) - 1u)); s.cand[p] = v; s.candp[p] = s.lp[v]; }
                sp = sp || (__ballot_sync(0xffffffffu, f && v == space_id) != 0u);
                base += __popc(m);
            }
            if (lane == 0) { s.ncand = base; s.has_space = sp ? 1 : 0; }
        }
        __syncthreads();
        const int n = s.nbeam, m = s.ncand, N = n * m;
        int Np = 1; while (Np < N) Np <<= 1;

        // ---- 1b. LM: score the word every beam would commit on ' ' (once per beam and frame) ------------------
        if (LM && s.has_space && tid < n && s.wplen[tid] > 0) {
            commit_word(tid, false);
        }

        // ---- 2. expansion, insertion index = cand * n + beam (symbol-major like the reference loop) ------------
        for (int idx = tid; idx < Np; idx += THREADS) {
            if (idx < N) {
                const int j = idx / n, i = idx - j * n;
                const int c = s.cand[j];
                const unsigned long long h = s.hash[i];
                const int ls = s.lastsym[i], lk = s.lastkey[i];
                unsigned long long nh = h; int nls = ls, nlk, app = SYM_NONE;
                if (c == blank) nlk = KEY_BLANK;
                else if (lk == c) nlk = c;                                     // repeat of the last emitted symbol
                else if (c == space_id) {
                    if (h == H0 || ls == space_id) { nls = space_id; nlk = space_id; }   // leading / repeated space: no new word
                    else { nh = mix(h, c); nls = c; nlk = c; app = c; }
                } else { nh = mix(h, c); nls = c; nlk = c; app = c; }
                s.ckey[idx] = mix(nh, nlk);
                s.cidx[idx] = (unsigned)idx;
                s.cscore[idx] = s.score[i] + (double)s.candp[j];
                s.cmeta[idx] = (unsigned)i | ((unsigned)app << 8) | ((unsigned)nls << 16) | ((unsigned)nlk << 24);
            } else { s.ckey[idx] = ~0ull; s.cidx[idx] = 0xffffffffu; }
        }
        __syncthreads();

        // ---- 3. merge equal states: sort by (key, insertion index), log-sum-exp each run in first-seen order ---
        bitonic_sort(s.ckey, s.cidx, Np);
        // heads -> unique slots (ordered compaction via block scan)
        int carry = 0;
        for (int base = 0; base < Np; base += THREADS) {
            const int p = base + tid;
            const bool head = p < N && (p == 0 || s.ckey[p] != s.ckey[p - 1]);
            const unsigned bal = __ballot_sync(0xffffffffu, head);
            if (lane == 0) s.scan[wid] = __popc(bal);
            __syncthreads();
            int off = carry;
            for (int w = 0; w < wid; ++w) off += s.scan[w];
            if (head) {
                const int u = off + __popc(bal & ((1u << lane) - 1u));
                double acc = s.cscore[s.cidx[p]];
                for (int qq = p + 1; qq < N && s.ckey[qq] == s.ckey[p]; ++qq) acc = logaddexp_d(acc, s.cscore[s.cidx[qq]]);
                s.uscore[u] = acc;
                s.urep[u] = s.cidx[p];
                // ranking score: acoustic, plus lm(text) and the partial-word penalty of the state when an LM is fused
                double comb = acc;
                if (LM) {
                    const unsigned meta = s.cmeta[s.cidx[p]];
                    const int src = meta & 0xff, app = (meta >> 8) & 0xff;
                    double lmv; int wl;
                    if (app == space_id) { lmv = s.commit_lm[src]; wl = 0; }
                    else { lmv = s.lmscore[src]; wl = s.wplen[src] + (app != SYM_NONE ? 1 : 0); }
                    comb = __dadd_rn(acc, __dadd_rn(lmv, partial_penalty(q, wl)));
                }
                s.ucomb[u] = comb;
            }
            int tot = 0;
            for (int w = 0; w < THREADS / 32; ++w) tot += s.scan[w];
            carry += tot;
            __syncthreads();
        }
        const int U = carry;

        // ---- 4. prune at best + beam_prune_logp, rank by (score desc, first-seen order) ------------------------
        double best = -INFINITY;
        for (int u = tid; u < U; u += THREADS) best = fmax(best, s.ucomb[u]);
        for (int o = 16; o >= 1; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
        __shared__ double wbest[THREADS / 32];
        if (lane == 0) wbest[wid] = best;
        __syncthreads();
        best = wbest[0];
        for (int w = 1; w < THREADS / 32; ++w) best = fmax(best, wbest[w]);
        int Up = 1; while (Up < U) Up <<= 1;
        for (int u = tid; u < Up; u += THREADS) {
            if (u < U && s.ucomb[u] >= best + (double)prune) {
                s.ckey[u] = ~dkey(s.ucomb[u]);             // ascending sort == descending score
                s.cidx[u] = s.urep[u];
            } else { s.ckey[u] = ~0ull; s.cidx[u] = 0xffffffffu; }
        }
        __syncthreads();
        bitonic_sort(s.ckey, s.cidx, Up);

        // ---- 5. new beams + back-pointers --------------------------------------------------------------------
        // cidx now holds representative insertion indices in rank order; the merged (acoustic) score of a
        // representative is stored at the representative's insertion slot
        for (int u = tid; u < U; u += THREADS) s.cscore[s.urep[u]] = s.uscore[u];
        __syncthreads();
        int kept = 0;
        for (int k = tid; k < BW_MAX; k += THREADS) {
            const bool ok = k < Up && k < beam_width && s.cidx[k] != 0xffffffffu;
            if (ok) {
                const unsigned rep = s.cidx[k];
                const unsigned meta = s.cmeta[rep];
                const int src = meta & 0xff, app = (meta >> 8) & 0xff;
                const unsigned long long nh = (app == SYM_NONE) ? s.hash[src] : mix(s.hash[src], app);
                s.nhash[k] = nh;
                // text without a trailing space: unchanged by a committed ' ' (the source cannot end in one)
                s.nfhash[k] = (app == SYM_NONE) ? s.fhash[src] : (app == space_id ? s.hash[src] : nh);
                s.nscore[k] = s.cscore[rep];
                s.nlastsym[k] = (unsigned char)((meta >> 16) & 0xff);
                s.nlastkey[k] = (unsigned char)((meta >> 24) & 0xff);
                if (LM) {
                    if (app == space_id) {
                        s.nlmscore[k] = s.commit_lm[src]; s.nwph[k] = H0; s.nwplen[k] = 0;
                        s.nnctx[k] = s.commit_nctx[src];
                        for (int j = 0; j < MAX_ORDER - 1; ++j) s.nctx_[k][j] = s.commit_ctx[src][j];
                    } else {
                        s.nlmscore[k] = s.lmscore[src];
                        s.nwph[k] = (app == SYM_NONE) ? s.wph[src] : mix(s.wph[src], app);
                        s.nwplen[k] = s.wplen[src] + (app != SYM_NONE ? 1 : 0);
                        s.nnctx[k] = s.nctx[src];
                        for (int j = 0; j < MAX_ORDER - 1; ++j) s.nctx_[k][j] = s.ctx[src][j];
                    }
                }
                bpp[(size_t)t * BW_MAX + k] = (unsigned char)src;
                bps[(size_t)t * BW_MAX + k] = (unsigned char)app;
            }
            kept += ok ? 1 : 0;
        }
        // count survivors
        {
            int c = kept;
            for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) s.scan[wid] = c;
        }
        __syncthreads();
        if (tid == 0) { int tot = 0; for (int w = 0; w < THREADS / 32; ++w) tot += s.scan[w]; s.nbeam = tot; }
        for (int k = tid; k < BW_MAX; k += THREADS) {
            s.hash[k] = s.nhash[k]; s.fhash[k] = s.nfhash[k]; s.score[k] = s.nscore[k];
            s.lastsym[k] = s.nlastsym[k]; s.lastkey[k] = s.nlastkey[k];
            if (LM) {
                s.lmscore[k] = s.nlmscore[k]; s.wph[k] = s.nwph[k]; s.wplen[k] = s.nwplen[k]; s.nctx[k] = s.nnctx[k];
                for (int j = 0; j < MAX_ORDER - 1; ++j) s.ctx[k][j] = s.nctx_[k][j];
            }
        }
        __syncthreads();
    }

    // ---- end of utterance: the partial word joins the text; states with equal text merge (first-seen order);
    //      with an LM every final text is scored with the </s> term, the best text wins; back-trace ---------------
    if (LM && tid < s.nbeam) {
        // every final text is scored as the END of the sentence: pyctcdecode >= 0.5 keys its LM cache by
        // (text, is_eos), so a text that was already committed during the search is scored again with </s> at the end
        if (s.wplen[tid] > 0) commit_word(tid, true);
        else s.commit_lm[tid] = lm_accumulate_eos(q, s.lmscore[tid], lm_score(lm, s.ctx[tid], s.nctx[tid], lm.eos));
    }
    __syncthreads();
    if (tid == 0) {
        const int n = s.nbeam;
        int bi = 0; double bs = -INFINITY;
        for (int i = 0; i < n; ++i) {
            bool first = true;
            for (int qq = 0; qq < i; ++qq) if (s.fhash[qq] == s.fhash[i]) { first = false; break; }
            if (!first) continue;
            double acc = s.score[i];
            for (int qq = i + 1; qq < n; ++qq) if (s.fhash[qq] == s.fhash[i]) acc = logaddexp_d(acc, s.score[qq]);
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
