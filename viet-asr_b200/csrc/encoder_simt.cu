// CUDA-core (fp32 FMA) encoder kernels over channels-last activations [B, T, C]:
//   * lens_kernel      - sequence-length chain of MaskedConv1d.get_seq_len (parts/jasper.py:108-111)
//   * dw_conv_kernel   - depthwise conv (groups = C) with register sliding window
//   * pw_gemm_kernel   - 1x1 conv as an fp32 SGEMM with the BN shift / residual GEMM / ReLU /
//                        length mask fused in the epilogue
//   * decoder_kernel   - the CTC head: skinny fp32 GEMM (cp.async double-buffered) + bias + log-softmax +
//                        greedy argmax (jasper.py:253-254, greedy_ctc_decoder.py:35); used by every gemm mode
// lens/dw_conv/pw_gemm are the exact-order fp32 path (VASR_GEMM_FP32_SIMT) and the oracle the tcgen05 path is
// debugged against on the device.
#include "common.cuh"
#include "kernels.cuh"
#include <float.h>

namespace vasr {

// ---------------------------------------------------------------------------------------------
// lens: stage 0 = trunc(seq_len); stage s+1 = trunc((len_s + 2p - d(k-1) - 1) / stride + 1)
// The chain only changes at strided layers; every other layer maps len -> len.
// enc_len (float) = value after the last conv, like the reference's float lengths.
// ---------------------------------------------------------------------------------------------
__global__ void lens_kernel(const long long* __restrict__ seq_len, int B, int b0, int nb, int n_stage,
                            const int* __restrict__ st_k, const int* __restrict__ st_s,
                            const int* __restrict__ st_d, const int* __restrict__ st_p,
                            int* __restrict__ lens /*[n_stage+1][B]*/, float* __restrict__ enc_len)
{
    const int b = b0 + blockIdx.x * blockDim.x + threadIdx.x;     // utterances [b0, b0 + nb) of a batch of B
    if (b >= b0 + nb) return;
    long long li = seq_len[b];
    float lf = (float)li;
    lens[b] = (int)li;
    for (int s = 0; s < n_stage; ++s) {
        // true division on an integer tensor -> float32 (parts/jasper.py:108-111), then the next
        // MaskedConv1d truncates with .to(long) (:115)
        lf = (float)(li + 2 * st_p[s] - st_d[s] * (st_k[s] - 1) - 1) / (float)st_s[s] + 1.0f;
        li = (long long)lf;
        lens[(size_t)(s + 1) * B + b] = (int)li;
    }
    // every later conv ('same' padded, stride 1) maps the truncated integer length to itself as a
    // float, so the encoder's returned length is float(trunc(.)) e.g. 250.0
    (void)lf;
    if (enc_len) enc_len[b] = (float)li;
}

// ---------------------------------------------------------------------------------------------
// depthwise conv, channels-last.  out[b,t,c] = sum_k w[k][c] * x[b, t*S + k*D - pad, c]
// The input is already zero for t >= len_in (producer invariant); the output is zeroed for
// t >= len_out because the following pointwise MaskedConv1d masks its input (parts/jasper.py:116).
// Thread = one channel, R outputs (t0 + r*D) with the input window held in registers.
// ---------------------------------------------------------------------------------------------
template <int K, int S, int D, int R>
__global__ void __launch_bounds__(128)
dw_conv_kernel(const float* __restrict__ x, const float* __restrict__ w /*[K][C]*/,
               float* __restrict__ y, int C, int T_in, int T_out, int pad,
               const int* __restrict__ len_in, const int* __restrict__ len_out)
{
    // window element j  <->  input time  u = t0*S - pad + j*G ,  G = D (S==1) or 1 (S==2,D==1)
    constexpr int G = (S == 1) ? D : 1;
    constexpr int RS = (S * D) / G;   // window step per output r
    constexpr int KS = D / G;         // window step per tap k
    constexpr int WIN = (R - 1) * RS + (K - 1) * KS + 1;
    const int c = blockIdx.y * 128 + threadIdx.x;
    const int b = blockIdx.z;
    if (c >= C) return;
    // blockIdx.x enumerates (chunk, phase): outputs t = chunk*R*D + phase + r*D
    const int chunk = blockIdx.x / D, phase = blockIdx.x % D;
    const int t0 = chunk * R * D + phase;
    if (t0 >= T_out) return;
    const int lin = min(len_in[b], T_in);
    const int lout = len_out[b];
    const float* xb = x + (size_t)b * T_in * C + c;
    float win[WIN];
    const int u0 = t0 * S - pad;
#pragma unroll
    for (int j = 0; j < WIN; ++j) {
        const int u = u0 + j * G;
        win[j] = (u >= 0 && u < lin) ? __ldg(xb + (size_t)u * C) : 0.f;
    }
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float wk = __ldg(w + (size_t)k * C + c);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fmaf(wk, win[r * RS + k * KS], acc[r]);
    }
    float* yb = y + (size_t)b * T_out * C + c;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int t = t0 + r * D;
        if (t < T_out) yb[(size_t)t * C] = (t < lout) ? acc[r] : 0.f;
    }
}

template <int K, int S, int D, int R>
static int launch_dw_t(const float* x, const float* w, float* y, int B, int C, int T_in, int T_out, int pad,
                       const int* len_in, const int* len_out, cudaStream_t st)
{
    dim3 grid(ceil_div(T_out, R * D) * D, ceil_div(C, 128), B);
    dw_conv_kernel<K, S, D, R><<<grid, 128, 0, st>>>(x, w, y, C, T_in, T_out, pad, len_in, len_out);
    VASR_LAUNCH_OK("dw_conv_kernel");
    return VASR_OK;
}

int launch_dw_conv(const float* x, const float* w, float* y, int B, int C, int T_in, int T_out,
                   int K, int S, int D, int pad, const int* len_in, const int* len_out, cudaStream_t st)
{
#define DW_CASE(k, s, d, r) \
    if (K == k && S == s && D == d) return launch_dw_t<k, s, d, r>(x, w, y, B, C, T_in, T_out, pad, len_in, len_out, st)
    DW_CASE(33, 2, 1, 8);
    DW_CASE(33, 1, 1, 16);
    DW_CASE(39, 1, 1, 16);
    DW_CASE(51, 1, 1, 16);
    DW_CASE(63, 1, 1, 16);
    DW_CASE(75, 1, 1, 16);
    DW_CASE(87, 1, 2, 16);
    DW_CASE(11, 1, 1, 16);   // small kernels used by unit tests / other Jasper variants
    DW_CASE(11, 2, 1, 8);
    DW_CASE(13, 1, 1, 16);
    DW_CASE(15, 1, 2, 16);
    DW_CASE(17, 1, 1, 16);
    DW_CASE(21, 1, 1, 16);
    DW_CASE(25, 1, 1, 16);
    DW_CASE(29, 1, 2, 16);
#undef DW_CASE
    return set_error(VASR_EINVAL,
                     "depthwise conv (kernel=%d, stride=%d, dilation=%d) is not a built shape; built: "
                     "k in {11,13,17,21,25,33,39,51,63,75} s=1 d=1, (11|33, s=2), (15|29|87, d=2)", K, S, D);
}

// ---------------------------------------------------------------------------------------------
// pointwise conv = SGEMM.  Y[n, co] = sum_ci X[n, ci] W[co, ci] (+ sum_cj R[n, cj] Wr[co, cj]) + shift[co]
// X: [N, Cin] (N = B*T rows, channels-last), W: [Cout, Cin] (BN scale folded).  Both K-major.
// CTA tile 128 x 128, K chunk 16, 256 threads, 8 x 8 register tile per thread.
// ---------------------------------------------------------------------------------------------
constexpr int GM = 128, GN = 128, GK = 16, GLD = GM + 4;

enum { EPI_CONV = 0 };

struct PwArgs {
    const float* X; const float* W; int Cin;
    const float* R; const float* Wr; int Cres;     // optional residual GEMM (R may be null)
    const float* shift;                            // [Cout]  (decoder: bias)
    float* Y; int N; int Cout; int T;              // rows N = B*T
    const int* len;                                // [B] valid frames (mask) or null
    int relu; int mask_tail;                       // mask_tail: zero rows t >= len[b]
};

template <int TN>
__device__ __forceinline__ void gemm_accumulate(const float* __restrict__ A, int lda, int a_rows, int row0,
                                                const float* __restrict__ Bm, int ldb, int b_rows, int col0,
                                                int Kdim, float (&acc)[8][TN], float* As, float* Bs)
{
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    for (int k0 = 0; k0 < Kdim; k0 += GK) {
        // global -> smem (transposed to [k][m]); each thread moves 2 float4 of A and 2 of B
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * 256;           // 0..511
            const int r = idx >> 2, kq = (idx & 3) * 4;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            if (row0 + r < a_rows) va = __ldg(reinterpret_cast<const float4*>(A + (size_t)(row0 + r) * lda + k0 + kq));
            if (r < 16 * TN && col0 + r < b_rows) vb = __ldg(reinterpret_cast<const float4*>(Bm + (size_t)(col0 + r) * ldb + k0 + kq));
            As[(kq + 0) * GLD + r] = va.x; As[(kq + 1) * GLD + r] = va.y;
            As[(kq + 2) * GLD + r] = va.z; As[(kq + 3) * GLD + r] = va.w;
            Bs[(kq + 0) * GLD + r] = vb.x; Bs[(kq + 1) * GLD + r] = vb.y;
            Bs[(kq + 2) * GLD + r] = vb.z; Bs[(kq + 3) * GLD + r] = vb.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(As + k * GLD + ty * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(As + k * GLD + ty * 8 + 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bb[TN];
#pragma unroll
            for (int j = 0; j < TN; j += 2) {
                const float2 bv = *reinterpret_cast<const float2*>(Bs + k * GLD + tx * TN + j);
                bb[j] = bv.x; bb[j + 1] = bv.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
}

// TN = output columns per thread: CTA tile = 128 rows x 16*TN columns
template <int EPI, int TN>
__global__ void __launch_bounds__(256)
pw_gemm_kernel(PwArgs p)
{
    __shared__ __align__(16) float As[GK * GLD];
    __shared__ __align__(16) float Bs[GK * GLD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * GM;      // rows n
    const int col0 = blockIdx.y * (16 * TN);      // output channels
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    gemm_accumulate<TN>(p.X, p.Cin, p.N, row0, p.W, p.Cin, p.Cout, col0, p.Cin, acc, As, Bs);
    if (p.R) gemm_accumulate<TN>(p.R, p.Cres, p.N, row0, p.Wr, p.Cres, p.Cout, col0, p.Cres, acc, As, Bs);

    {
        static_assert(TN == 8, "conv epilogue is written for 8 columns per thread");
        float sh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int co = col0 + tx * 8 + j;
            sh[j] = (co < p.Cout) ? __ldg(p.shift + co) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int n = row0 + ty * 8 + i;
            if (n >= p.N) continue;
            bool live = true;
            if (p.mask_tail) {
                const int b = n / p.T, t = n - b * p.T;
                live = t < p.len[b];
            }
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float o = acc[i][j] + sh[j];
                if (p.relu) o = fmaxf(o, 0.f);
                v[j] = live ? o : 0.f;
            }
            float* yrow = p.Y + (size_t)n * p.Cout + col0 + tx * 8;
            if (col0 + tx * 8 + 7 < p.Cout) {
                *reinterpret_cast<float4*>(yrow) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(yrow + 4) = make_float4(v[4], v[5], v[6], v[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (col0 + tx * 8 + j < p.Cout) yrow[j] = v[j];
            }
        }
    }
}

int launch_pw_gemm(const float* X, const float* W, int Cin, const float* R, const float* Wr, int Cres,
                   const float* shift, float* Y, int B, int T, int Cout, const int* len, int relu,
                   int mask_tail, cudaStream_t st)
{
    VASR_REQUIRE(Cin % GK == 0 && (R == nullptr || Cres % GK == 0),
                 "pointwise conv: input channels must be a multiple of %d (got %d / %d)", GK, Cin, Cres);
    PwArgs p{};
    p.X = X; p.W = W; p.Cin = Cin; p.R = R; p.Wr = Wr; p.Cres = Cres; p.shift = shift;
    p.Y = Y; p.N = B * T; p.Cout = Cout; p.T = T; p.len = len; p.relu = relu; p.mask_tail = mask_tail;
    dim3 grid(ceil_div(p.N, GM), ceil_div(Cout, GN));
    pw_gemm_kernel<EPI_CONV, 8><<<grid, 256, 0, st>>>(p);
    VASR_LAUNCH_OK("pw_gemm_kernel<conv>");
    return VASR_OK;
}

// ---------------------------------------------------------------------------------------------
// CTC decoder head: logits = enc[N, Cin] x W[V1, Cin]^T + bias, log-softmax over the V1 classes, greedy argmax
// (JasperDecoderForCTC.forward, jasper.py:253-254; greedy_ctc_decoder.py:35).  A skinny fp32 GEMM (V1 = 29 ... 128
// columns): CTA = RM*16 rows x TN*16 columns, K in chunks of 32 through a 2-stage cp.async ring (the loads of chunk
// k + 1 are in flight while chunk k is multiplied), operands K-major in shared memory with a 4-float pad so that the
// 16-byte loads along k are conflict-free.  Every accumulator sums its products in ascending k, like pw_gemm_kernel.
// Thread (ty, tx): rows ty*RM + i, classes tx + 16*j  ->  a class row lives in the 16 lanes that share ty.
// ---------------------------------------------------------------------------------------------
constexpr int DK = 32, DLD = DK + 4;

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int bytes = valid ? 16 : 0;                          // 0 source bytes: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

template <int TN, int RM>
__global__ void __launch_bounds__(256, TN <= 2 ? 4 : (TN <= 4 ? 3 : 2))
decoder_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
               int Cin, int V1, int N, float* __restrict__ logp, long long* __restrict__ ids)
{
    constexpr int DM = 16 * RM, DN = 16 * TN;
    extern __shared__ __align__(16) float dsm[];
    float* As = dsm;                               // [2][DM][DLD]
    float* Bs = dsm + 2 * DM * DLD;                // [2][DN][DLD]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * DM;

    auto load_stage = [&](int buf, int k0) {
        for (int c = tid; c < DM * (DK / 4); c += 256) {
            const int r = c >> 3, q = c & 7;
            const bool ok = row0 + r < N && k0 + q * 4 < Cin;          // rows past N and the K tail are zero-filled
            cp_async16(As + ((size_t)buf * DM + r) * DLD + q * 4, X + (ok ? (size_t)(row0 + r) * Cin + k0 + q * 4 : 0), ok);
        }
        for (int c = tid; c < DN * (DK / 4); c += 256) {
            const int r = c >> 3, q = c & 7;
            const bool ok = r < V1 && k0 + q * 4 < Cin;
            cp_async16(Bs + ((size_t)buf * DN + r) * DLD + q * 4, W + (ok ? (size_t)r * Cin + k0 + q * 4 : 0), ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    float acc[RM][TN];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = (Cin + DK - 1) / DK;
    load_stage(0, 0);
    for (int kt = 0; kt < nk; ++kt) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                       // chunk kt has landed; chunk kt - 1's buffer is free
        if (kt + 1 < nk) load_stage((kt + 1) & 1, (kt + 1) * DK);
        const float* Ab = As + (size_t)(kt & 1) * DM * DLD + (size_t)ty * RM * DLD;
        const float* Bb = Bs + (size_t)(kt & 1) * DN * DLD + (size_t)tx * DLD;
#pragma unroll
        for (int k4 = 0; k4 < DK / 4; ++k4) {
            float4 a[RM], b[TN];
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i] = *reinterpret_cast<const float4*>(Ab + i * DLD + k4 * 4);
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = *reinterpret_cast<const float4*>(Bb + (size_t)j * 16 * DLD + k4 * 4);
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
                }
        }
    }

    // epilogue: bias, log-softmax over the class row (16 lanes x TN), greedy argmax (ties -> lowest index, torch.argmax)
    float bs[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int co = tx + 16 * j;
        bs[j] = (co < V1) ? __ldg(bias + co) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int n = row0 + ty * RM + i;
        float v[TN];
        float mx = -FLT_MAX;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            v[j] = acc[i][j] + bs[j];
            if (tx + 16 * j < V1) mx = fmaxf(mx, v[j]);
        }
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.f;
#pragma unroll
        for (int j = 0; j < TN; ++j)
            if (tx + 16 * j < V1) se += expf(v[j] - mx);
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        const float lse = logf(se);
        float best = -FLT_MAX; int bi = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            v[j] = (v[j] - mx) - lse;
            const int co = tx + 16 * j;
            if (co < V1 && (v[j] > best)) { best = v[j]; bi = co; }
        }
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (n < N) {
            if (logp) {
                float* lrow = logp + (size_t)n * V1;
#pragma unroll
                for (int j = 0; j < TN; ++j)
                    if (tx + 16 * j < V1) lrow[tx + 16 * j] = v[j];
            }
            if (ids && tx == 0) ids[n] = (long long)bi;
        }
    }
}

template <int TN, int RM>
static int launch_decoder_t(const float* enc, const float* W, const float* bias, int Cin, int V1, int N, float* logp,
                            long long* ids, cudaStream_t st)
{
    constexpr size_t smem = (size_t)2 * (16 * RM + 16 * TN) * DLD * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        VASR_CUDA_OK(cudaFuncSetAttribute(decoder_kernel<TN, RM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    decoder_kernel<TN, RM><<<ceil_div(N, 16 * RM), 256, smem, st>>>(enc, W, bias, Cin, V1, N, logp, ids);
    VASR_LAUNCH_OK("decoder_kernel");
    return VASR_OK;
}

int launch_decoder(const float* enc, const float* W, const float* bias, int Cin, int V1,
                   int N, float* logp, long long* ids, cudaStream_t st)
{
    VASR_REQUIRE(V1 <= GN, "decoder: at most %d classes (incl. blank) are supported (got %d)", GN, V1);
    VASR_REQUIRE(Cin % GK == 0, "decoder: feat_in must be a multiple of %d (got %d)", GK, Cin);     // (the kernel needs 4)
    // 64-row CTAs (4 rows per thread): <= 64 registers at TN = 2, i.e. 4 CTAs per SM.  128-row CTAs (8 rows per
    // thread, a better FMA : shared-load ratio) need 128 registers -> 2 CTAs per SM and 1.7 waves on 256 x 5 s:
    // measured 0.167 ms against 0.217 ms for the round-1 kernel
    if (V1 <= 32) return launch_decoder_t<2, 4>(enc, W, bias, Cin, V1, N, logp, ids, st);
    if (V1 <= 64) return launch_decoder_t<4, 4>(enc, W, bias, Cin, V1, N, logp, ids, st);
    return launch_decoder_t<8, 4>(enc, W, bias, Cin, V1, N, logp, ids, st);
}

int launch_lens(const long long* seq_len, int B, int b0, int nb, int n_stage, const int* st_k, const int* st_s,
                const int* st_d, const int* st_p, int* lens, float* enc_len, cudaStream_t st)
{
    lens_kernel<<<ceil_div(nb, 128), 128, 0, st>>>(seq_len, B, b0, nb, n_stage, st_k, st_s, st_d, st_p, lens, enc_len);
    VASR_LAUNCH_OK("lens_kernel");
    return VASR_OK;
}

}  // namespace vasr
