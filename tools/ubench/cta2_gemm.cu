// Bring-up check of the 2-CTA tensor-core primitives the round-2 encoder kernel plans to use ("Round-2
// kernel plan"): a cluster of two CTAs computes D[256 x 256] = A[256 x K] * B[256 x K]^T (fp16 in, fp32 out) with
//   * tcgen05.alloc / dealloc .cta_group::2 (one warp in each CTA),
//   * TMA loads issued by BOTH CTAs (each loads its 128 rows of A and its 128-row half of B) with
//     cp.async.bulk.tensor.2d.cta_group::2 whose completion lands on the LEADER's mbarrier (peer bit masked off),
//   * one thread of the leader issuing tcgen05.mma.cta_group::2.kind::f16 (M = 256 over both SMs, N = 256, K = 16),
//   * tcgen05.commit.cta_group::2 ... multicast::cluster releasing a barrier in both CTAs,
//   * each CTA's epilogue warps reading their 128 TMEM lanes and writing their 128 rows of D.
// Operand layout is the one of csrc/encoder_tc.cu (K-major, 64-byte rows, SWIZZLE_64B, chunks of 32 k).
// Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o cta2_gemm cta2_gemm.cu -lcuda
// Run:   ./cta2_gemm        (prints max |D - ref| and PASS/FAIL; exits non-zero on failure)
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int MT = 256, NT = 256, KT = 64, KC = 32;
constexpr int PART = 128 * KC * 2;                 // [128 rows x 64 B] = 8 KiB
constexpr int NCH = KT / KC;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;        // clears the CTA-rank bit of a shared::cluster address (cute: Sm100MmaPeerBitMask)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol mistake must end in a diagnosable failure, not a hung GPU
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t parity)
{
    for (long long i = 0; i < 20000000ll; ++i)
        if (mbar_try(b, parity)) return true;
    return false;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ void umma_2sm(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_2sm_multicast(uint64_t* bar, uint16_t mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// D = f32, A = B = f16, K-major both, N = 256, M = 256 (pair)
constexpr uint32_t IDESC = (1u << 4) | ((256u >> 3) << 17) | ((256u >> 4) << 24);

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
gemm2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ D, int* __restrict__ status)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_s = smem;                              // [NCH][128 x 64 B]
    unsigned char* b_s = smem + NCH * PART;                 // [NCH][128 x 64 B]  (this CTA's half of the 256 N rows)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * NCH * PART);
    uint64_t* acc_full = full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int pair = blockIdx.x >> 1;

    if (threadIdx.x == 0) {
        mbar_init(full, 1);
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // both CTAs load; all bytes are accounted on the leader's barrier
        if (rank == 0) mbar_expect_tx(full, 2u * 2u * NCH * PART);
        for (int c = 0; c < NCH; ++c) {
            tma_load_2d_2sm(a_s + c * PART, &tmA, c * KC, pair * MT + (int)rank * 128, full);
            tma_load_2d_2sm(b_s + c * PART, &tmB, c * KC, (int)rank * 128, full);
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        if (!mbar_wait(full, 0)) { atomicExch(status, 1); }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < NCH; ++c)
            for (int ks = 0; ks < KC / 16; ++ks)
                umma_2sm(tmem, make_desc_sw64(smem_u32(a_s + c * PART) + ks * 32), make_desc_sw64(smem_u32(b_s + c * PART) + ks * 32),
                         IDESC, (c > 0 || ks > 0) ? 1u : 0u);
        commit_2sm_multicast(acc_full, 3);
    } else if (warp >= 2) {
        const int q = warp & 3;                                  // TMEM lane quarter of this warp
        const bool ok = mbar_wait(acc_full, 0);
        if (!ok && lane == 0) atomicExch(status, 2 + (int)rank);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ok) {
            const int row = pair * MT + (int)rank * 128 + q * 32 + lane;
            for (int col0 = 0; col0 < NT; col0 += 32) {
                uint32_t r[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col0) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int j = 0; j < 32; ++j) D[(size_t)row * NT + col0 + j] = __uint_as_float(r[j]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();                                               // nobody frees TMEM / exits while the peer still uses it
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main()
{
    const int pairs = 4;                                         // 4 clusters: also checks that every pair lands on a TPC
    const int M = pairs * MT;
    std::vector<__half> hA((size_t)M * KT), hB((size_t)NT * KT);
    std::vector<float> fA(hA.size()), fB(hB.size());
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((int)(s >> 20) % 17 - 8) / 8.0f; };
    for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2half(rnd()); fA[i] = __half2float(hA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2half(rnd()); fB[i] = __half2float(hB[i]); }
    __half *dA, *dB; float* dD; int* dS;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, (size_t)M * NT * 4)); CK(cudaMalloc(&dS, 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, (size_t)M * NT * 4)); CK(cudaMemset(dS, 0, 4));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fn;
    CUtensorMap tmA, tmB;
    cuuint32_t es[2] = {1, 1}, box[2] = {KC, 128};
    cuuint64_t str[1] = {KT * 2};
    cuuint64_t dimA[2] = {KT, (cuuint64_t)M}, dimB[2] = {KT, NT};
    if (enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, dimA, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dimB, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("tensor map encode failed\n"); return 2; }
    const size_t smem = 2 * NCH * PART + 64;
    CK(cudaFuncSetAttribute(gemm2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm2<<<2 * pairs, 192, smem>>>(tmA, tmB, dD, dS);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> hD((size_t)M * NT); int st = 0;
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    double mx = 0; long long bad = 0;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < NT; ++j) {
            double ref = 0;
            for (int k = 0; k < KT; ++k) ref += (double)fA[(size_t)i * KT + k] * fB[(size_t)j * KT + k];
            const double d = fabs((double)hD[(size_t)i * NT + j] - ref);
            if (!(d <= 1e-3)) ++bad;
            if (d > mx || d != d) mx = d;
        }
    printf("cta_group::2 GEMM %dx%dx%d on %d clusters: status %d, max |D - ref| = %.3g, mismatches %lld -> %s\n", M, NT, KT, pairs, st, mx,
           bad, (st == 0 && bad == 0) ? "PASS" : "FAIL");
    return (st == 0 && bad == 0) ? 0 : 1;
}
