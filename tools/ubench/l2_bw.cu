// L2 -> SM bandwidth of the whole chip for an L2-resident working set, three ways:
//   ldg   : every thread streams 16-byte ld.global.cg loads
//   bulk  : one thread per CTA streams 16 KiB cp.async.bulk copies into a shared-memory ring (the TMA path the encoder
//           kernels use for weights and activation windows)
//   mixed : bulk loads + plain 16-byte stores to a second buffer (the encoder's load : store ratio is ~4 : 1)
// The encoder's segment kernels move ~1.3 MB per 128 x 512 tile through this path; this tool measures the ceiling.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o l2_bw l2_bw.cu
// Run:   ./l2_bw [MiB working set, default 48]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

__global__ void __launch_bounds__(1024, 1) ldg_kernel(const float4* __restrict__ buf, size_t n_vec, int iters, float* sink)
{
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int it = 0; it < iters; ++it) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x + (size_t)it * 977 * 1024 % n_vec;
        for (size_t k = 0; k < n_vec / stride; k += 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                size_t idx = (i + (k + u) * stride) % n_vec;
                asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(buf + idx));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].w;
        }
    }
    if (acc == 123.456f) *sink = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int CHUNK = 16384, STAGES = 8;
// one producer thread streams CHUNK-byte bulk copies into a ring; consumer warps release the stages (and optionally
// store 16-byte vectors to `out`)
__global__ void __launch_bounds__(160, 1) bulk_kernel(const unsigned char* __restrict__ buf, size_t bytes, int chunks_per_cta, float4* out, int store_every)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full[STAGES], empty[STAGES];
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(full + i)), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(empty + i)), "r"(4));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t n_chunks = bytes / CHUNK;
    if (warp == 4) {
        if (lane == 0) {
            for (int c = 0; c < chunks_per_cta; ++c) {
                const int s = c % STAGES; const uint32_t ph = (c / STAGES) & 1;
                uint32_t done = 0;
                while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(empty + s)), "r"(ph ^ 1) : "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(full + s)), "r"(CHUNK) : "memory");
                const size_t ci = ((size_t)blockIdx.x * 7919 + (size_t)c * gridDim.x) % n_chunks;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(smem + (size_t)s * CHUNK)), "l"(buf + ci * CHUNK), "r"(CHUNK), "r"(smem_u32(full + s)) : "memory");
            }
        }
    } else {
        for (int c = 0; c < chunks_per_cta; ++c) {
            const int s = c % STAGES; const uint32_t ph = (c / STAGES) & 1;
            uint32_t done = 0;
            while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(full + s)), "r"(ph) : "memory");
            if (store_every && (c % store_every) == 0) {
                // CHUNK bytes of stores per store_every chunks loaded, spread over the 128 consumer threads
                float4* o = out + ((size_t)blockIdx.x * chunks_per_cta + c) % (n_chunks) * (CHUNK / 16);
                for (int i = threadIdx.x; i < CHUNK / 16; i += 128) o[i] = make_float4(1.f, 2.f, 3.f, 4.f);
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + s)) : "memory");
        }
    }
}

int main(int argc, char** argv)
{
    const size_t mib = argc > 1 ? (size_t)atoi(argv[1]) : 48;
    const size_t bytes = mib << 20;
    int dev = 0, sms = 0, clk = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    unsigned char *buf, *out; float* sink;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&out, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 1, bytes)); CK(cudaMemset(out, 0, bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms;
    printf("working set %zu MiB, %d SMs, max SM clock %d MHz\n", mib, sms, clk / 1000);
    {   // ---- ldg
        const int iters = 20;
        ldg_kernel<<<sms, 1024>>>((const float4*)buf, bytes / 16, 2, sink);       // warm L2
        CK(cudaEventRecord(e0));
        ldg_kernel<<<sms, 1024>>>((const float4*)buf, bytes / 16, iters, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        const double moved = (double)(bytes / 16 / ((size_t)sms * 1024) / 4 * 4) * sms * 1024 * 16 * iters;
        printf("ldg.cg 16B   : %8.1f GB/s  (%.2f ms)\n", moved / ms / 1e6, ms);
    }
    CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * CHUNK));
    for (int store_every : {0, 4, 2}) {
        const int chunks = 4000;
        bulk_kernel<<<sms, 160, STAGES * CHUNK>>>(buf, bytes, 200, (float4*)out, store_every);
        CK(cudaEventRecord(e0));
        bulk_kernel<<<sms, 160, STAGES * CHUNK>>>(buf, bytes, chunks, (float4*)out, store_every);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        const double ld = (double)chunks * sms * CHUNK, st = store_every ? ld / store_every : 0;
        printf("bulk 16KiB x%d stages, stores 1:%d : load %8.1f GB/s + store %7.1f GB/s = %8.1f GB/s  (%.2f ms)\n", STAGES, store_every,
               ld / ms / 1e6, st / ms / 1e6, (ld + st) / ms / 1e6, ms);
    }
    CK(cudaGetLastError());
    return 0;
}
