/* CPU restatement in plain C of the integer end of the path - TEST INFRASTRUCTURE ONLY (see oracle/quartznet_oracle.py
 * for the import rules: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use anything under oracle/).
 *
 *  oracle_greedy_argmax   GreedyCTCDecoder.forward (nemo/collections/asr/greedy_ctc_decoder.py:33-36):
 *                         argmax over the class axis of [B, T, V] log-probs; torch.argmax returns the FIRST maximal
 *                         index on ties, NaN compares as the maximum.
 *  oracle_ctc_collapse    __ctc_decoder_predictions_tensor (nemo/collections/asr/helpers.py:7-33): blank = last class;
 *                         emit p iff (p != prev or prev == blank) and p != blank; prev = p; ALL T frames are visited
 *                         (no length truncation, helpers.py:26-30).  Output: ids padded with -1, and the count.
 *
 * Pinned by tests/test_oracle_cpu.py against the texts of the reference-generated golden vectors and against the numpy
 * restatement (oracle/quartznet_oracle.py ctc_collapse).  Build: oracle/c/Makefile (called by __graft_entry__.build()). */
#include <math.h>
#include <stdint.h>

void oracle_greedy_argmax(const float* logp, int B, int T, int V, int64_t* ids)
{
    for (long long bt = 0; bt < (long long)B * T; ++bt) {
        const float* row = logp + bt * V;
        int best = 0;
        float bv = row[0];
        for (int v = 1; v < V; ++v) {
            const float x = row[v];
            /* strict '>' keeps the first maximum; a NaN wins over any number, and the first NaN is kept */
            if ((x > bv) || (isnan(x) && !isnan(bv))) { bv = x; best = v; }
        }
        ids[bt] = best;
    }
}

void oracle_ctc_collapse(const int64_t* ids, int B, int T, int blank, int32_t* out_ids, int32_t* out_len)
{
    for (int b = 0; b < B; ++b) {
        const int64_t* p = ids + (long long)b * T;
        int32_t* o = out_ids + (long long)b * T;
        int n = 0;
        int64_t prev = blank;
        for (int t = 0; t < T; ++t) {
            const int64_t c = p[t];
            if ((c != prev || prev == blank) && c != blank) o[n++] = (int32_t)c;
            prev = c;
        }
        out_len[b] = n;
        for (int t = n; t < T; ++t) o[t] = -1;
    }
}
