// Greedy CTC collapse: drop repeats, then blanks (nemo/collections/asr/helpers.py:20-32).
// The reference uses every frame of the tensor it is given and only ever sees one utterance (infer.py:167-171); in a
// zero-padded batch `frames[b]` restricts utterance b to the frames it would have had alone (null = all T frames).
#include "common.cuh"
#include "kernels.cuh"
#include <math.h>

namespace vasr {

// one CTA (256 threads) per utterance; ordered compaction via ballot + block prefix sum
__global__ void __launch_bounds__(256)
ctc_collapse_kernel(const long long* __restrict__ ids, const int* __restrict__ frames, int T, int blank,
                    int* __restrict__ out_ids, int* __restrict__ out_len)
{
    __shared__ int warp_cnt[8];
    __shared__ int base_s;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long* row = ids + (size_t)b * T;
    int* orow = out_ids + (size_t)b * T;
    const int Tb = frames ? min(max(frames[b], 0), T) : T;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int t0 = 0; t0 < Tb; t0 += 256) {
        const int t = t0 + tid;
        long long p = blank, prev = blank;
        if (t < Tb) {
            p = row[t];
            prev = (t > 0) ? row[t - 1] : (long long)blank;   // `previous` starts as the blank id
        }
        const bool keep = (t < Tb) && (p != blank) && (p != prev || prev == blank);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[wid] = __popc(m);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < wid; ++w) off += warp_cnt[w];
        if (keep) orow[off + __popc(m & ((1u << lane) - 1u))] = (int)p;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < 8; ++w) tot += warp_cnt[w];
            base_s += tot;
        }
        __syncthreads();
    }
    const int n = base_s;
    for (int t = n + tid; t < T; t += 256) orow[t] = -1;
    if (tid == 0) out_len[b] = n;
}

// GreedyCTCDecoder.forward (greedy_ctc_decoder.py:33-36): one warp per row, ties -> lowest index
__global__ void __launch_bounds__(256)
greedy_argmax_kernel(const float* __restrict__ logp, int N, int V, long long* __restrict__ ids)
{
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= N) return;
    const float* r = logp + (size_t)row * V;
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int v = lane; v < V; v += 32) {
        const float x = r[v];
        if (x > best || (x != x && bi == 0x7fffffff)) { best = x; bi = v; }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) ids[row] = (long long)(bi == 0x7fffffff ? 0 : bi);
}

int launch_ctc_collapse(const long long* ids, const int* frames, int B, int T, int blank, int* out_ids, int* out_len, cudaStream_t st)
{
    ctc_collapse_kernel<<<B, 256, 0, st>>>(ids, frames, T, blank, out_ids, out_len);
    VASR_LAUNCH_OK("ctc_collapse_kernel");
    return VASR_OK;
}

}  // namespace vasr

extern "C" int vasr_greedy_argmax(const float* log_probs, int N, int V, int64_t* ids, void* stream)
{
    using namespace vasr;
    VASR_REQUIRE(log_probs && ids, "vasr_greedy_argmax: null argument");
    VASR_REQUIRE(N > 0 && V > 0, "vasr_greedy_argmax: N and V must be positive (got %d, %d)", N, V);
    greedy_argmax_kernel<<<ceil_div(N, 8), 256, 0, (cudaStream_t)stream>>>(log_probs, N, V, (long long*)ids);
    VASR_LAUNCH_OK("greedy_argmax_kernel");
    return VASR_OK;
}

extern "C" int vasr_ctc_collapse(const int64_t* ids, const int32_t* frames, int B, int T, int blank,
                                 int32_t* out_ids, int32_t* out_len, void* stream)
{
    using namespace vasr;
    VASR_REQUIRE(ids && out_ids && out_len, "vasr_ctc_collapse: null argument");
    VASR_REQUIRE(B > 0 && T > 0, "vasr_ctc_collapse: B and T must be positive (got %d, %d)", B, T);
    return launch_ctc_collapse((const long long*)ids, frames, B, T, blank, out_ids, out_len, (cudaStream_t)stream);
}
