// Micro-benchmark: FP32 FMA issue rate per SM sub-partition on sm_100a (scalar FFMA vs packed FFMA2),
// as a function of resident warps per SMSP.  Used to decide whether the depthwise stage is pipe- or latency-bound.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: FFMA with 16 independent accumulators ; 1: FFMA2 with 8 independent float2 accumulators
__global__ void k(float* out, long long* cyc, int iters, float a0)
{
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
    float2 w[4] = {make_float2(a0, a0 * 0.5f), make_float2(a0 * 0.25f, a0 * 0.75f), make_float2(a0 * 1.5f, a0), make_float2(a0 * .3f, a0 * .7f)};
    float2 x[4] = {make_float2(0.5f, 0.25f), make_float2(0.125f, 0.375f), make_float2(0.625f, 0.875f), make_float2(.1f, .2f)};
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 1) acc[i] = __ffma2_rn(w[u & 3], x[(u + i) & 3], acc[i]);
                else {
                    acc[i].x = fmaf(w[u & 3].x, x[(u + i) & 3].x, acc[i].x);
                    acc[i].y = fmaf(w[u & 3].y, x[(u + i) & 3].y, acc[i].y);
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
    const int iters = 2000;
    for (int mode = 0; mode < 2; ++mode)
        for (int warps = 4; warps <= 32; warps *= 2) {
            if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters, 1.0001f); else k<1><<<148, warps * 32>>>(out, cyc, iters, 1.0001f);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
            const double fma_per_thread = (double)iters * 64 * (mode == 1 ? 2 : 2);   // 64 packed = 128 scalar FMAs per iteration in both modes
            const double inst_per_warp = (double)iters * (mode == 1 ? 64 : 128);
            printf("mode=%s warps/SM=%2d (per SMSP %d): %.0f cycles, %.2f cycles per instr per SMSP, %.1f FMA lanes/clk/SM\n",
                   mode ? "FFMA2" : "FFMA ", warps, warps / 4, c, c / (inst_per_warp * warps / 4.0), fma_per_thread * warps * 32 / c);
        }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
