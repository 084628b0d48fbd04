// Audio ingest in front of the hot path: int16 PCM -> float32 and band-limited sample-rate conversion on the GPU.
//
// Restates what the reference's callers do on the CPU before `VietASR.transcribe`:
//   * soundfile/librosa float32 convention  x / 2^15                      (nemo/collections/asr/parts/segment.py:61-74)
//   * `librosa.load(path, sr=16000)`                                      (infer.py:200, app.py:66,82)
//       -> librosa.resample(res_type="kaiser_best") -> resampy.resample: Kaiser-windowed sinc interpolation with a
//          512-entries-per-zero-crossing filter table and linear interpolation between table entries, output length
//          int(n * ratio), then librosa's fix_length to ceil(n * ratio) (zero padded).
// librosa / resampy are un-vendored third-party packages that are absent from this image: the algorithm is restated
// from the published one (oracle/resample_oracle.py is the scalar restatement used by the tests) - parity with the
// packages themselves is UNPINNED.  resampy 0.2.x (the release of the reference's era) advances the input time of
// successive output samples by a running fp64 sum `time_register += 1/ratio`; because its filter-table stride is
// truncated to an integer (int(scale * num_table)) the interpolator is not continuous where the time crosses an
// integer, so the rounding of that running sum is observable for ratios whose increment is not a dyadic fraction
// (44.1 kHz -> 16 kHz).  The sum is therefore reproduced exactly: `time_table_kernel` (one thread, sequential fp64
// adds, once per ratio / length, cached in the handle) writes time_register[t], the resampling kernel reads it.
#include "common.cuh"
#include "kernels.cuh"

struct vasr_resampler {
    float* d_win = nullptr;     // [n_win] right half of the windowed sinc
    float* d_delta = nullptr;   // [n_win] forward differences (last = 0)
    int n_win = 0;
    int num_table = 0;          // table entries per zero crossing
    double* d_time = nullptr;   // [time_cap] resampy's running sum time_register[t] for time_inc
    long long time_cap = 0;
    double time_inc = 0.0;
    cudaEvent_t time_ready = nullptr;   // recorded after the table was (re)built; later calls on any stream wait for it
};

namespace vasr {

__global__ void pcm16_to_float_kernel(const int16_t* __restrict__ x, const long long* __restrict__ len, long long L,
                                      float* __restrict__ y)
{
    const int b = blockIdx.y;
    const long long n = len[b];
    const int16_t* xr = x + (size_t)b * L;
    float* yr = y + (size_t)b * L;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < L; i += (long long)gridDim.x * blockDim.x)
        yr[i] = (i < n) ? (float)xr[i] * (1.0f / 32768.0f) : 0.0f;
}

template <typename TIn>
__device__ __forceinline__ float load_sample(const TIn* x, long long i);
template <> __device__ __forceinline__ float load_sample<float>(const float* x, long long i) { return __ldg(x + i); }
template <> __device__ __forceinline__ float load_sample<int16_t>(const int16_t* x, long long i)
{
    return (float)__ldg(x + i) * (1.0f / 32768.0f);
}

// time_register[0] = 0, time_register[t] = time_register[t-1] + inc: the exact sequence of fp64 roundings of resampy's
// sample loop (interpn.resample_f).  Sequential by construction; the same table serves every utterance of a batch.
__global__ void time_table_kernel(double* __restrict__ tbl, long long n, double inc)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double t = 0.0;
    for (long long i = 0; i < n; ++i) {
        tbl[i] = t;
        t = __dadd_rn(t, inc);
    }
}

// one thread per output sample: left wing (x[n], x[n-1], ...) + right wing (x[n+1], ...) of the interpolation filter
template <typename TIn>
__global__ void resample_kernel(const TIn* __restrict__ x, const long long* __restrict__ len_in, long long L_in,
                                float* __restrict__ y, long long* __restrict__ len_out, long long L_out,
                                double ratio, const double* __restrict__ time_tbl, const float* __restrict__ win,
                                const float* __restrict__ delta, int n_win, int num_table, float gain)
{
    const int b = blockIdx.y;
    const long long n_orig = len_in[b];
    // resampy writes int(n * ratio) samples, librosa pads / trims to ceil(n * ratio)
    const long long n_res = (long long)((double)n_orig * ratio);
    const long long n_fix = (long long)ceil((double)n_orig * ratio);
    if (blockIdx.x == 0 && threadIdx.x == 0) len_out[b] = n_fix < L_out ? n_fix : L_out;
    const TIn* xr = x + (size_t)b * L_in;
    float* yr = y + (size_t)b * L_out;
    const double scale = ratio < 1.0 ? ratio : 1.0;
    const int index_step = (int)(scale * num_table);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < L_out; t += (long long)gridDim.x * blockDim.x) {
        float acc = 0.f;
        if (t < n_res) {
            const double time_register = time_tbl[t];
            const long long n = (long long)time_register;
            // left wing
            double frac = scale * (time_register - (double)n);
            double index_frac = frac * num_table;
            int offset = (int)index_frac;
            float eta = (float)(index_frac - offset);
            long long i_max = (n_win - offset) / index_step;
            if (n + 1 < i_max) i_max = n + 1;
            for (long long i = 0; i < i_max; ++i) {
                const int idx = offset + (int)i * index_step;
                const float w = fmaf(eta, __ldg(delta + idx), __ldg(win + idx));
                acc = fmaf(w, load_sample<TIn>(xr, n - i), acc);
            }
            // right wing
            frac = scale - frac;
            index_frac = frac * num_table;
            offset = (int)index_frac;
            eta = (float)(index_frac - offset);
            long long k_max = (n_win - offset) / index_step;
            if (n_orig - n - 1 < k_max) k_max = n_orig - n - 1;
            for (long long k = 0; k < k_max; ++k) {
                const int idx = offset + (int)k * index_step;
                const float w = fmaf(eta, __ldg(delta + idx), __ldg(win + idx));
                acc = fmaf(w, load_sample<TIn>(xr, n + k + 1), acc);
            }
            acc *= gain;
        }
        yr[t] = acc;
    }
}

}  // namespace vasr

extern "C" int vasr_resampler_create(const float* interp_win_host, int n_win, int num_table, vasr_resampler** out)
{
    using namespace vasr;
    VASR_REQUIRE(interp_win_host && out, "vasr_resampler_create: null argument");
    VASR_REQUIRE(n_win > 1 && num_table > 0 && (n_win - 1) % num_table == 0,
                 "vasr_resampler_create: the half window must hold num_zeros * num_table + 1 samples (got %d, %d)", n_win, num_table);
    vasr_resampler* rs = new vasr_resampler();
    rs->n_win = n_win; rs->num_table = num_table;
    std::vector<float> d((size_t)n_win, 0.f);
    for (int i = 0; i + 1 < n_win; ++i) d[i] = interp_win_host[i + 1] - interp_win_host[i];
    cudaError_t e = cudaMalloc(&rs->d_win, sizeof(float) * n_win);
    if (e == cudaSuccess) e = cudaMalloc(&rs->d_delta, sizeof(float) * n_win);
    if (e == cudaSuccess) e = cudaMemcpy(rs->d_win, interp_win_host, sizeof(float) * n_win, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(rs->d_delta, d.data(), sizeof(float) * n_win, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(rs->d_win); cudaFree(rs->d_delta); delete rs;
        return set_error(VASR_ECUDA, "vasr_resampler_create: %s", cudaGetErrorString(e));
    }
    *out = rs;
    return VASR_OK;
}

extern "C" void vasr_resampler_destroy(vasr_resampler* rs)
{
    if (!rs) return;
    cudaFree(rs->d_win); cudaFree(rs->d_delta); cudaFree(rs->d_time);
    if (rs->time_ready) cudaEventDestroy(rs->time_ready);
    delete rs;
}

extern "C" int64_t vasr_resample_out_len(int64_t n_in, int sr_in, int sr_out)
{
    if (n_in < 0 || sr_in <= 0 || sr_out <= 0) return vasr::set_error(VASR_EINVAL, "vasr_resample_out_len: bad argument");
    return (int64_t)ceil((double)n_in * ((double)sr_out / (double)sr_in));
}

extern "C" int vasr_pcm16_to_float(const int16_t* pcm, const int64_t* length, int B, int64_t L, float* wave, void* stream)
{
    using namespace vasr;
    VASR_REQUIRE(pcm && length && wave, "vasr_pcm16_to_float: null argument");
    VASR_REQUIRE(B > 0 && L > 0, "vasr_pcm16_to_float: B and L must be positive (got %d, %lld)", B, (long long)L);
    dim3 grid((unsigned)std::min<int64_t>(ceil_div64(L, 256), 1184), (unsigned)B);
    pcm16_to_float_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pcm, (const long long*)length, (long long)L, wave);
    VASR_LAUNCH_OK("pcm16_to_float_kernel");
    return VASR_OK;
}

extern "C" int vasr_resample(vasr_resampler* rs, const void* x, int pcm16, const int64_t* len_in, int B, int64_t L_in,
                             int sr_in, int sr_out, float* y, int64_t* len_out, int64_t L_out, void* stream)
{
    using namespace vasr;
    VASR_REQUIRE(rs && x && len_in && y && len_out, "vasr_resample: null argument");
    VASR_REQUIRE(B > 0 && L_in > 0 && L_out > 0, "vasr_resample: B, L_in and L_out must be positive");
    VASR_REQUIRE(sr_in > 0 && sr_out > 0, "vasr_resample: sample rates must be positive (got %d, %d)", sr_in, sr_out);
    const double ratio = (double)sr_out / (double)sr_in;
    VASR_REQUIRE((int)((ratio < 1.0 ? ratio : 1.0) * rs->num_table) >= 1,
                 "vasr_resample: ratio %d -> %d is below the filter table resolution", sr_in, sr_out);
    VASR_REQUIRE(L_out >= vasr_resample_out_len(L_in, sr_in, sr_out),
                 "vasr_resample: L_out %lld < ceil(L_in * ratio) = %lld", (long long)L_out,
                 (long long)vasr_resample_out_len(L_in, sr_in, sr_out));
    // resampy scales the low-pass by the ratio when down-sampling (interp_win *= sample_ratio)
    const float gain = ratio < 1.0 ? (float)ratio : 1.0f;
    dim3 grid((unsigned)std::min<int64_t>(ceil_div64(L_out, 128), 2368), (unsigned)B);
    cudaStream_t st = (cudaStream_t)stream;
    const double inc = 1.0 / ratio;
    if (rs->time_inc != inc || rs->time_cap < L_out) {
        // (re)build the time table on this stream; the old table may still be read by earlier launches: cudaFree
        // synchronises the device before releasing it
        const long long cap = std::max<long long>((long long)L_out, rs->time_inc == inc ? 2 * rs->time_cap : 0);
        double* t = nullptr;
        cudaError_t e = cudaMalloc(&t, sizeof(double) * (size_t)cap);
        if (e != cudaSuccess) return set_error(VASR_ECUDA, "vasr_resample: time table (%lld entries): %s", cap, cudaGetErrorString(e));
        cudaFree(rs->d_time);
        rs->d_time = t; rs->time_cap = cap; rs->time_inc = inc;
        time_table_kernel<<<1, 32, 0, st>>>(rs->d_time, cap, inc);
        VASR_LAUNCH_OK("time_table_kernel");
        if (!rs->time_ready) VASR_CUDA_OK(cudaEventCreateWithFlags(&rs->time_ready, cudaEventDisableTiming));
        VASR_CUDA_OK(cudaEventRecord(rs->time_ready, st));
    } else if (rs->time_ready) {
        // the table may have been built on another stream: order this stream's reads after it
        VASR_CUDA_OK(cudaStreamWaitEvent(st, rs->time_ready, 0));
    }
    if (pcm16)
        resample_kernel<int16_t><<<grid, 128, 0, st>>>((const int16_t*)x, (const long long*)len_in, (long long)L_in, y,
                                                      (long long*)len_out, (long long)L_out, ratio, rs->d_time, rs->d_win,
                                                      rs->d_delta, rs->n_win, rs->num_table, gain);
    else
        resample_kernel<float><<<grid, 128, 0, st>>>((const float*)x, (const long long*)len_in, (long long)L_in, y,
                                                    (long long*)len_out, (long long)L_out, ratio, rs->d_time, rs->d_win,
                                                    rs->d_delta, rs->n_win, rs->num_table, gain);
    VASR_LAUNCH_OK("resample_kernel");
    return VASR_OK;
}
