// Micro-benchmark of the depthwise inner loop (rolling register window, FFMA2, shared-memory operands) in isolation:
// cycles per 32-channel x 128-step chunk for different R (outputs per thread) and warps per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int R, int NW>      // NW dw warps; each chunk = 16 channel pairs x 128 time steps
__global__ void __launch_bounds__(NW * 32, 1) k(const float* gx, float* out, long long* cyc, int chunks)
{
    extern __shared__ float2 sm[];                       // window [128 + K - 1 rows][16 pairs] + taps [K][16]
    constexpr int ROWS = 128 + K - 1, XP = 16;
    float2* xs0 = sm; float2* wp0 = sm + ROWS * XP;
    for (int i = threadIdx.x; i < (ROWS + K) * XP; i += blockDim.x) sm[i] = make_float2(gx[i % 4096], gx[(i * 7) % 4096]);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cp = lane & 15;
    constexpr int TG = 128 / R;                          // time groups per chunk
    constexpr int THREADS_PER_CHUNK = TG * 16;
    const int tid_in = threadIdx.x % THREADS_PER_CHUNK;
    const int tw = (tid_in >> 4) * R;
    const float2* xs = xs0 + cp; const float2* wp = wp0 + cp;
    float2 tot = make_float2(0.f, 0.f);
    const int my_chunks = chunks / ((NW * 32) / THREADS_PER_CHUNK);
    long long t0 = clock64();
    for (int c = 0; c < my_chunks; ++c) {
        constexpr int P = 4, WN = R + P, LAST_ROW = K - 1 + R - 1;
        float2 win[WN], wq[P], acc[R];
#pragma unroll
        for (int j = 0; j < WN; ++j) win[j] = xs[(tw + j) * XP];
#pragma unroll
        for (int j = 0; j < P; ++j) wq[j] = wp[j * XP];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int kk = 0; kk < K; ++kk) {
            const float2 wk = wq[kk % P];
            if (kk + P < K) wq[kk % P] = wp[(kk + P) * XP];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = __ffma2_rn(wk, win[(kk + r) % WN], acc[r]);
            const int row = kk + WN;
            if (row <= LAST_ROW) win[kk % WN] = xs[(tw + row) * XP];
        }
#pragma unroll
        for (int r = 0; r < R; ++r) { tot.x += acc[r].x; tot.y += acc[r].y; }
        __syncwarp();
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = tot.x + tot.y;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K, int R, int NW>
void run(const float* gx, float* out, long long* cyc)
{
    const int chunks = 512;
    const size_t smem = (128 + K - 1 + K) * 16 * sizeof(float2);
    cudaFuncSetAttribute(k<K, R, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<K, R, NW><<<148, NW * 32, smem>>>(gx, out, cyc, chunks);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double ffma2_per_smsp = 128.0 * 16 * K / 32 / 4;   // warp-FFMA2 per chunk per SMSP
    printf("K=%d R=%2d warps=%2d: %.0f cycles per chunk (FFMA2 floor %.0f at 2 cyc/instr)  err=%s\n", K, R, NW, c / chunks,
           ffma2_per_smsp * 2, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float *gx, *out; long long* cyc;
    cudaMalloc(&gx, 4096 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = (i % 97) * 0.01f; cudaMemcpy(gx, h, sizeof(h), cudaMemcpyHostToDevice);
    run<63, 8, 8>(gx, out, cyc);
    run<63, 8, 16>(gx, out, cyc);
    run<63, 16, 4>(gx, out, cyc);
    run<63, 16, 8>(gx, out, cyc);
    run<63, 16, 16>(gx, out, cyc);
    run<63, 32, 4>(gx, out, cyc);
    run<63, 32, 8>(gx, out, cyc);
    run<33, 8, 8>(gx, out, cyc);
    run<33, 16, 8>(gx, out, cyc);
    return 0;
}
