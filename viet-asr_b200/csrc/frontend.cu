// Log-mel front end: pre-emphasis + reflect-padded STFT (n_fft = 512) + power + slaney mel
// + log, then per-utterance per-feature normalisation and masking.
// Restates FilterbankFeatures.forward (nemo/collections/asr/parts/features.py:245-301)
// and normalize_batch (:17-30) as two kernels over channels-last [B, T_f, nfilt] output.
#include "common.cuh"
#include <math.h>
#include <vector>

namespace vasr {

constexpr int NFFT = 512;
constexpr int NBINS = NFFT / 2 + 1;
constexpr int FE_FRAMES = 16;            // frames per CTA (8 packed-complex FFTs)
constexpr int FE_PAIRS = FE_FRAMES / 2;
constexpr int FE_THREADS = 256;          // = NFFT/2 butterflies per stage

}  // namespace vasr

struct vasr_frontend {
    vasr_frontend_cfg cfg;
    int max_nz = 0;
    float* d_window = nullptr;     // [win]
    float2* d_twiddle = nullptr;   // [256] exp(-2 pi i k / 512)
    int* d_mel_start = nullptr;    // [nfilt]
    int* d_mel_cnt = nullptr;      // [nfilt]
    float* d_mel_w = nullptr;      // [nfilt][max_nz]
    size_t stft_smem = 0;
    int pad_per_utterance = 0;     // vasr_frontend_set_padding: reflect every row at its own length instead of at L
};

namespace vasr {

// K1: grid (ceil(T_frames / 16), B), block 256.
// smem: z[FE_PAIRS][512] float2 | tw[256] float2 | seg[(FE_FRAMES-1)*hop + win]
__global__ void __launch_bounds__(FE_THREADS)
stft_mel_kernel(const float* __restrict__ wave, const long long* __restrict__ length, long long L, int T_frames, int T_out,
                const float* __restrict__ window, const float2* __restrict__ twiddle,
                const int* __restrict__ mel_start, const int* __restrict__ mel_cnt,
                const float* __restrict__ mel_w, int max_nz, int nfilt, int win, int hop,
                float preemph, float guard, float* __restrict__ logmel)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int seg_len = (FE_FRAMES - 1) * hop + win;
    float2* z = reinterpret_cast<float2*>(smem_raw);                 // [FE_PAIRS][NFFT]
    float2* tw = z + FE_PAIRS * NFFT;                                // [NFFT/2]
    float* seg = reinterpret_cast<float*>(tw + NFFT / 2);            // [seg_len]

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * FE_FRAMES;
    const float* x = wave + (long long)b * L;
    // end of the signal for the reflection: the padded row (torch.stft on the [B, L] tensor, features.py:181-188), or -
    // per-utterance mode - this utterance's own length, i.e. what it sees when the reference transcribes it alone
    const long long Lr = length ? max(min(length[b], L), 1ll) : L;

    // ---- stage 0: pre-emphasised, reflect-padded signal segment ------------------
    // frame t uses y[hop*t - win/2 + m], m in [0, win)   (torch.stft centre padding n_fft/2,
    // window zero-padded (n_fft - win)/2 on each side -> only the centre `win` samples count)
    const long long i_base = (long long)hop * t0 - win / 2;
    for (int s = tid; s < seg_len; s += FE_THREADS) {
        long long i = i_base + s;
        if (i < 0) i = -i;
        if (i >= Lr) i = 2 * (Lr - 1) - i;
        float v = 0.f;
        if (i >= 0 && i < Lr) {
            v = x[i];
            if (i > 0) v = v - preemph * x[i - 1];   // features.py:255
        }
        seg[s] = v;
    }
    tw[tid] = twiddle[tid];
    __syncthreads();

    // ---- stage 1: windowed frames, two real frames packed into one complex FFT ----
    const int left = (NFFT - win) / 2;
    for (int idx = tid; idx < FE_PAIRS * NFFT; idx += FE_THREADS) {
        const int p = idx / NFFT, n = idx % NFFT;
        float2 v = make_float2(0.f, 0.f);
        const int m = n - left;
        if (m >= 0 && m < win) {
            const float w = window[m];
            v.x = w * seg[(2 * p) * hop + m];
            v.y = w * seg[(2 * p + 1) * hop + m];
        }
        z[idx] = v;
    }
    __syncthreads();

    // ---- stage 2: 512-point complex FFT = three radix-8 Stockham passes (natural order in and out) -----------
    // 64 threads per transform (one radix-8 butterfly each per pass), 4 transforms at a time; a pass reads its
    // 8 points (stride 64) into registers, everyone syncs, then writes them back to the same buffer at the
    // auto-sort positions - 3 shared-memory round trips instead of the 9 of a radix-2 network.
    {
        const int ft = tid >> 6;                 // transform within the group of 4
        const int i = tid & 63;                  // butterfly index
#pragma unroll 1
        for (int grp = 0; grp < FE_PAIRS / 4; ++grp) {
            float2* zp = z + (grp * 4 + ft) * NFFT;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const int p = (pass == 0) ? 1 : (pass == 1 ? 8 : 64);
                const int k = i & (p - 1);
                const int j = ((i - k) << 3) + k;
                float2 u[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) u[r] = zp[i + 64 * r];
                // twiddle u[r] *= exp(-2 pi i * k * r / (8 p)) = tw[(k * r) * (512 / (8 p))]  (table: 256 entries of W512)
                if (pass > 0) {
                    const int step = NFFT / (8 * p);                 // 8 for p = 8, 1 for p = 64
#pragma unroll
                    for (int r = 1; r < 8; ++r) {
                        const int e = (k * r * step) & (NFFT - 1);   // exponent of W512, < 512
                        float2 w = tw[e & (NFFT / 2 - 1)];
                        if (e >= NFFT / 2) { w.x = -w.x; w.y = -w.y; }   // W^(e) = -W^(e - 256)
                        const float2 a = u[r];
                        u[r] = make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
                    }
                }
                // 8-point DFT in registers (three radix-2 levels, decimation in frequency, outputs bit-reversed)
#define CADD(a, b) make_float2((a).x + (b).x, (a).y + (b).y)
#define CSUB(a, b) make_float2((a).x - (b).x, (a).y - (b).y)
#define MULMI(a) make_float2((a).y, -(a).x)                         /* a * (-i) */
                const float h = 0.70710678118654752440f;
                float2 a0 = CADD(u[0], u[4]), a4 = CSUB(u[0], u[4]);
                float2 a1 = CADD(u[1], u[5]), a5 = CSUB(u[1], u[5]);
                float2 a2 = CADD(u[2], u[6]), a6 = CSUB(u[2], u[6]);
                float2 a3 = CADD(u[3], u[7]), a7 = CSUB(u[3], u[7]);
                a5 = make_float2(h * (a5.x + a5.y), h * (a5.y - a5.x));      // * W8^1 = (1 - i)/sqrt2
                a6 = MULMI(a6);                                               // * W8^2 = -i
                a7 = make_float2(h * (a7.y - a7.x), -h * (a7.x + a7.y));     // * W8^3 = (-1 - i)/sqrt2
                float2 b0 = CADD(a0, a2), b2 = CSUB(a0, a2), b1 = CADD(a1, a3), b3 = CSUB(a1, a3);
                float2 b4 = CADD(a4, a6), b6 = CSUB(a4, a6), b5 = CADD(a5, a7), b7 = CSUB(a5, a7);
                b3 = MULMI(b3); b7 = MULMI(b7);
                float2 v[8];
                v[0] = CADD(b0, b1); v[4] = CSUB(b0, b1); v[2] = CADD(b2, b3); v[6] = CSUB(b2, b3);
                v[1] = CADD(b4, b5); v[5] = CSUB(b4, b5); v[3] = CADD(b6, b7); v[7] = CSUB(b6, b7);
#undef CADD
#undef CSUB
#undef MULMI
                __syncthreads();                                     // every butterfly of this pass has read its inputs
#pragma unroll
                for (int r = 0; r < 8; ++r) zp[j + r * p] = v[r];
                __syncthreads();
            }
        }
    }

    // ---- stage 3: untangle the two spectra, power spectrum re^2 + im^2 -------------
    float pa[FE_PAIRS], pb[FE_PAIRS], pa_ny[FE_PAIRS], pb_ny[FE_PAIRS];
    {
        const int k = tid;                                   // bins 0..255
        const unsigned rk = (unsigned)k;                     // the Stockham passes leave the spectrum in natural order
        const unsigned rn = (unsigned)((NFFT - k) & (NFFT - 1));
#pragma unroll
        for (int p = 0; p < FE_PAIRS; ++p) {
            const float2 zk = z[p * NFFT + rk], zn = z[p * NFFT + rn];
            const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
            const float br = 0.5f * (zk.y + zn.y), bi = 0.5f * (zn.x - zk.x);
            pa[p] = ar * ar + ai * ai;
            pb[p] = br * br + bi * bi;
            pa_ny[p] = 0.f; pb_ny[p] = 0.f;
        }
        if (tid == 0) {                                      // Nyquist bin 256 (its own mirror)
            const unsigned r256 = 256u;
#pragma unroll
            for (int p = 0; p < FE_PAIRS; ++p) {
                const float2 zk = z[p * NFFT + r256];
                pa_ny[p] = zk.x * zk.x;                      // A[256] = Re z, B[256] = Im z
                pb_ny[p] = zk.y * zk.y;
            }
        }
    }
    __syncthreads();
    float* P = reinterpret_cast<float*>(z);                   // [FE_FRAMES][NBINS] over the z buffer
#pragma unroll
    for (int p = 0; p < FE_PAIRS; ++p) {
        P[(2 * p) * NBINS + tid] = pa[p];
        P[(2 * p + 1) * NBINS + tid] = pb[p];
        if (tid == 0) {
            P[(2 * p) * NBINS + 256] = pa_ny[p];
            P[(2 * p + 1) * NBINS + 256] = pb_ny[p];
        }
    }
    __syncthreads();

    // ---- stage 4: sparse mel projection + log --------------------------------------
    for (int idx = tid; idx < FE_FRAMES * nfilt; idx += FE_THREADS) {
        const int f = idx / nfilt, j = idx % nfilt;
        const int t = t0 + f;
        if (t >= T_frames) continue;
        const int st = mel_start[j], cnt = mel_cnt[j];
        const float* wj = mel_w + (size_t)j * max_nz;
        const float* Pf = P + f * NBINS + st;
        float acc = 0.f;
        for (int q = 0; q < cnt; ++q) acc = fmaf(__ldg(wj + q), Pf[q], acc);
        logmel[((size_t)b * T_out + t) * nfilt + j] = logf(acc + guard);   // features.py:266-271
    }
}

// K2: one CTA of 1024 threads per utterance (thread = mel bin x time phase; 256 threads left the two dependent
// passes latency-bound).  mean / unbiased std over t < seq, (x-mean)/(std+1e-5), zero t >= seq.
constexpr int NORM_THREADS = 1024;
__global__ void __launch_bounds__(NORM_THREADS)
normalize_kernel(float* __restrict__ feat, const long long* __restrict__ length, long long* __restrict__ seq_out,
                 int T_frames, int T_out, int nfilt, int hop)
{
    __shared__ float red[NORM_THREADS];
    __shared__ float s_mean[128], s_std[128];
    const int tid = threadIdx.x;
    const int b = blockIdx.x;
    const int phases = NORM_THREADS / nfilt;
    const int j = tid % nfilt, ph = tid / nfilt;
    // seq = ceil(float(len) / hop) (features.py:238-239, float32 arithmetic like torch)
    const float lenf = (float)length[b];
    long long seq = (long long)ceilf(lenf / (float)hop);
    if (tid == 0) seq_out[b] = seq;
    int nvalid = (int)(seq < (long long)T_frames ? seq : (long long)T_frames);
    float* xb = feat + (size_t)b * T_out * nfilt;

    float s = 0.f;
    for (int t = ph; t < nvalid; t += phases) s += xb[(size_t)t * nfilt + j];
    red[tid] = s;
    __syncthreads();
    if (ph == 0) {
        float tot = 0.f;
        for (int q = 0; q < phases; ++q) tot += red[q * nfilt + j];
        s_mean[j] = tot / (float)nvalid;
    }
    __syncthreads();
    const float mean = s_mean[j];
    float ss = 0.f;
    for (int t = ph; t < nvalid; t += phases) {
        const float d = xb[(size_t)t * nfilt + j] - mean;
        ss = fmaf(d, d, ss);
    }
    red[tid] = ss;
    __syncthreads();
    if (ph == 0) {
        float tot = 0.f;
        for (int q = 0; q < phases; ++q) tot += red[q * nfilt + j];
        s_std[j] = sqrtf(tot / (float)(nvalid - 1)) + 1e-5f;     // unbiased, + CONSTANT (features.py:14,25)
    }
    __syncthreads();
    const float sd = s_std[j];
    for (int t = ph; t < T_out; t += phases) {
        float v = 0.f;                                            // pad_value / masked tail (:287-300)
        if (t < nvalid) v = (xb[(size_t)t * nfilt + j] - mean) / sd;
        xb[(size_t)t * nfilt + j] = v;
    }
}

}  // namespace vasr

// ---------------------------------------------------------------------------------
extern "C" int vasr_frontend_create(const vasr_frontend_cfg* cfg, const float* window_host,
                                    const float* mel_fb_host, vasr_frontend** out)
{
    using namespace vasr;
    VASR_REQUIRE(cfg && window_host && mel_fb_host && out, "vasr_frontend_create: null argument");
    VASR_REQUIRE(cfg->n_fft == NFFT, "vasr_frontend_create: only n_fft=512 is built (got %d)", cfg->n_fft);
    VASR_REQUIRE(cfg->n_window_size > 0 && cfg->n_window_size <= NFFT && cfg->n_window_size % 2 == 0,
                 "vasr_frontend_create: n_window_size must be even and in (0, 512] (got %d)", cfg->n_window_size);
    VASR_REQUIRE(cfg->n_window_stride > 0, "vasr_frontend_create: n_window_stride must be a positive int (got %d)",
                 cfg->n_window_stride);
    VASR_REQUIRE(cfg->nfilt > 0 && cfg->nfilt <= 128 && 256 % cfg->nfilt == 0,
                 "vasr_frontend_create: features must divide 256 and be <= 128 (got %d)", cfg->nfilt);
    VASR_REQUIRE(cfg->pad_to >= 0, "vasr_frontend_create: pad_to must be >= 0 (got %d)", cfg->pad_to);
    vasr_frontend* fe = new vasr_frontend();
    fe->cfg = *cfg;
    const int nf = cfg->nfilt;
    std::vector<int> start(nf), cnt(nf);
    int max_nz = 1;
    for (int j = 0; j < nf; ++j) {
        int lo = NBINS, hi = -1;
        for (int k = 0; k < NBINS; ++k)
            if (mel_fb_host[(size_t)j * NBINS + k] != 0.f) { if (k < lo) lo = k; hi = k; }
        if (hi < 0) { lo = 0; hi = -1; }
        start[j] = lo; cnt[j] = hi - lo + 1;
        if (cnt[j] > max_nz) max_nz = cnt[j];
    }
    std::vector<float> packed((size_t)nf * max_nz, 0.f);
    for (int j = 0; j < nf; ++j)
        for (int q = 0; q < cnt[j]; ++q) packed[(size_t)j * max_nz + q] = mel_fb_host[(size_t)j * NBINS + start[j] + q];
    std::vector<float2> tw(NFFT / 2);
    for (int k = 0; k < NFFT / 2; ++k) {
        const double a = -2.0 * M_PI * (double)k / (double)NFFT;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    fe->max_nz = max_nz;
    VASR_CUDA_OK(cudaMalloc(&fe->d_window, sizeof(float) * cfg->n_window_size));
    VASR_CUDA_OK(cudaMalloc(&fe->d_twiddle, sizeof(float2) * NFFT / 2));
    VASR_CUDA_OK(cudaMalloc(&fe->d_mel_start, sizeof(int) * nf));
    VASR_CUDA_OK(cudaMalloc(&fe->d_mel_cnt, sizeof(int) * nf));
    VASR_CUDA_OK(cudaMalloc(&fe->d_mel_w, sizeof(float) * packed.size()));
    VASR_CUDA_OK(cudaMemcpy(fe->d_window, window_host, sizeof(float) * cfg->n_window_size, cudaMemcpyHostToDevice));
    VASR_CUDA_OK(cudaMemcpy(fe->d_twiddle, tw.data(), sizeof(float2) * NFFT / 2, cudaMemcpyHostToDevice));
    VASR_CUDA_OK(cudaMemcpy(fe->d_mel_start, start.data(), sizeof(int) * nf, cudaMemcpyHostToDevice));
    VASR_CUDA_OK(cudaMemcpy(fe->d_mel_cnt, cnt.data(), sizeof(int) * nf, cudaMemcpyHostToDevice));
    VASR_CUDA_OK(cudaMemcpy(fe->d_mel_w, packed.data(), sizeof(float) * packed.size(), cudaMemcpyHostToDevice));
    const int seg_len = (FE_FRAMES - 1) * cfg->n_window_stride + cfg->n_window_size;
    fe->stft_smem = sizeof(float2) * FE_PAIRS * NFFT + sizeof(float2) * NFFT / 2 + sizeof(float) * seg_len;
    VASR_REQUIRE(fe->stft_smem <= 200 * 1024, "vasr_frontend_create: window_stride %d too large", cfg->n_window_stride);
    VASR_CUDA_OK(cudaFuncSetAttribute(stft_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fe->stft_smem));
    *out = fe;
    return VASR_OK;
}

extern "C" void vasr_frontend_destroy(vasr_frontend* fe)
{
    if (!fe) return;
    cudaFree(fe->d_window); cudaFree(fe->d_twiddle); cudaFree(fe->d_mel_start);
    cudaFree(fe->d_mel_cnt); cudaFree(fe->d_mel_w);
    delete fe;
}

extern "C" int vasr_frontend_set_padding(vasr_frontend* fe, int per_utterance)
{
    if (!fe) return vasr::set_error(VASR_EINVAL, "vasr_frontend_set_padding: null handle");
    fe->pad_per_utterance = per_utterance ? 1 : 0;
    return VASR_OK;
}

extern "C" int vasr_frontend_num_frames(const vasr_frontend* fe, int64_t L)
{
    if (!fe || L < 0) return vasr::set_error(VASR_EINVAL, "vasr_frontend_num_frames: bad argument");
    int64_t t = 1 + L / fe->cfg.n_window_stride;
    const int p = fe->cfg.pad_to;
    if (p > 0 && t % p != 0) t += p - t % p;
    return (int)t;
}

extern "C" int vasr_frontend_forward(vasr_frontend* fe, const float* wave, const int64_t* length,
                                     int B, int64_t L, float* feat, int64_t* seq_len, void* stream)
{
    using namespace vasr;
    VASR_REQUIRE(fe && wave && length && feat && seq_len, "vasr_frontend_forward: null argument");
    VASR_REQUIRE(B > 0, "vasr_frontend_forward: batch must be positive (got %d)", B);
    VASR_REQUIRE(L > NFFT / 2, "vasr_frontend_forward: reflect padding needs more than %d samples (got %lld)",
                 NFFT / 2, (long long)L);
    cudaStream_t st = (cudaStream_t)stream;
    const int hop = fe->cfg.n_window_stride;
    const int T_frames = (int)(1 + L / hop);
    const int T_out = vasr_frontend_num_frames(fe, L);
    dim3 grid(ceil_div(T_frames, FE_FRAMES), B);
    stft_mel_kernel<<<grid, FE_THREADS, fe->stft_smem, st>>>(
        wave, fe->pad_per_utterance ? (const long long*)length : nullptr, (long long)L, T_frames, T_out, fe->d_window, fe->d_twiddle, fe->d_mel_start, fe->d_mel_cnt,
        fe->d_mel_w, fe->max_nz, fe->cfg.nfilt, fe->cfg.n_window_size, hop, fe->cfg.preemph,
        fe->cfg.log_zero_guard, feat);
    VASR_LAUNCH_OK("stft_mel_kernel");
    normalize_kernel<<<B, NORM_THREADS, 0, st>>>(feat, (const long long*)length, (long long*)seq_len,
                                        T_frames, T_out, fe->cfg.nfilt, hop);
    VASR_LAUNCH_OK("normalize_kernel");
    return VASR_OK;
}
