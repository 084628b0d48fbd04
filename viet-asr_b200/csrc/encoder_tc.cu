// Fused QuartzNet sub-block for sm_100a:   depthwise conv (CUDA cores, register sliding window)
//   -> fp16 hi/lo split written straight into the swizzled smem A operand
//   -> 1x1 conv(s) on the 5th-gen tensor cores (tcgen05.mma kind::f16, fp32 accumulators in TMEM)
//   -> BN shift + residual (second GEMM chain into the same TMEM tile) + ReLU + length mask epilogue.
// Restates JasperBlock.forward (nemo/collections/asr/parts/jasper.py:408-448) sub-block by sub-block.
//
// Tile: one utterance b, TN = 128 output time steps, up to 512 output channels
//       D[t, co] (M = 128 t, N = 256 co per MMA) = Act[t, ci] (A, K-major, written by the depthwise warps)
//       x W[co, ci] (B, K-major, TMA), chunked over ci in KC = 32 (64-byte fp16 rows -> SWIZZLE_64B).
//       Time is on the TMEM lanes, so an epilogue thread owns one output row.
// Precision: fp16 operands with a 2-term split  x = hi + lo  on both sides and three products
//       hi*hi + lo*hi + hi*lo  (fp32-grade, mode 1) or hi*hi only (mode 2).  Weights are pre-scaled per layer
//       by a power of two so that hi/lo stay in fp16's normal range; the scale is undone in the epilogue.
// Three kernels share the pipeline:
//   segment_pair_kernel  a run of stride-1 sub-blocks of one width, persistent 2-CTA clusters, cta_group::2 MMAs,
//                        8 depthwise warps in two groups (the throughput path: 75 of the 78 sub-blocks of 15x5)
//   segment_kernel       the same run on single CTAs with 32-row tiles (latency mode for small batches)
//   subblock_kernel      one sub-block per launch (stride-2 first block, the dilated K = 87 layer, the final 1x1)
#include "common.cuh"
#include "kernels.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <vector>
#include <type_traits>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>

namespace vasr {
namespace tc {

constexpr int TN = 128;                 // output time steps per CTA
constexpr int KC = 32;                  // input channels per chunk
constexpr int NDW = 4;                  // depthwise warps of the single-CTA kernels: thread = channel pair x 16 outputs
constexpr int NEPI = 8;                 // epilogue warps: any 8 consecutive warps cover each TMEM lane quarter (warp & 3) twice
constexpr int NTHREADS = (NDW + 3 + NEPI) * 32;            // 480: 4 depthwise, window / weight / MMA warps, 8 epilogue
constexpr int MAX_STAGES = 8;           // upper bound of the activation-window / operand ring depths
constexpr int SCHED = 4;                // depth of the tile ring
constexpr int SCHED_CONSUMERS = 1 /*weights*/ + 1 /*MMA*/ + NDW + NEPI;
constexpr int PART_BYTES = 128 * KC * 2;   // activation operand: [128 t rows x 64 B] fp16 = 8 KiB per part
constexpr int W_PART = 256 * KC * 2;       // weight operand:     [256 co rows x 64 B] fp16 = 16 KiB per part
constexpr int MAX_CO_CTA = 512;
constexpr int EPI_STAGE_BYTES = TN * 32 * 4;   // one 128-row x 32-channel fp32 output slice (SWIZZLE_128B)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static int g_num_sms = 0;

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    // try_wait suspends the warp in hardware for a bounded time; parking the waiting roles with nanosleep on top of
    // it was measured to make no difference (profiles/r1_v10_spin_sweep.log)
    // (watchdog: a barrier that never completes - a protocol error - traps instead of hanging the device)
    unsigned spins = 0;
    while (!mbar_try(bar, parity)) { if (++spins > (1u << 28)) __trap(); }
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// plain (non-tensor) bulk copy global -> shared, completion counted on an mbarrier
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA store shared -> global (bulk async-group completion)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(tm), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16 operands, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile [rows x 64 B], SWIZZLE_64B: 8-row groups of 512 B (SBO), canonical
// ((8,n),2):((4,SBO),1) in 16-byte units (cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::K>).
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);            // start address        bits [0,14)
    d |= (uint64_t)1 << 16;                                  // leading byte offset  bits [16,30) (unused for swizzled K-major)
    d |= (uint64_t)(512u >> 4) << 32;                        // stride byte offset   bits [32,46): 8 rows x 64 B
    d |= (uint64_t)1 << 46;                                  // descriptor version 1 (sm_100)
    d |= (uint64_t)4 << 61;                                  // layout type: SWIZZLE_64B
    return d;
}
// byte offset of element (row, 16-byte chunk c16 in 0..3) inside a SWIZZLE_64B K-major tile
__device__ __forceinline__ uint32_t sw64_offset(int row, int c16)
{
    return (uint32_t)((row >> 3) * 512 + (row & 7) * 64 + ((c16 ^ ((row >> 1) & 3)) << 4));
}

// instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=256 (cute UMMA::InstrDescriptor)
constexpr uint32_t IDESC_F16_M128_N256 = (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) |
                                         ((256u >> 3) << 17) | ((128u >> 4) << 24);

struct Params {
    const float* dw_w;       // [Cin/32][K][32] fp32 depthwise taps (ones for a plain 1x1 conv)
    const float* shift;      // [Cout]
    float wscale_inv;        // 2^-s of the layer's power-of-two weight pre-scale
    float* out;              // [B, T_out, Cout] with an explicit batch stride (elements)
    long long out_bstride;
    const int* len_out;      // [B]
    int* tile_counter;       // dynamic tile scheduler (zeroed before the launch)
    int* status;             // range guard: bit 0 is set when a depthwise output does not fit fp16 (zeroed per forward)
    int Cin, Cres, Cout, T_out, pad;
    int n_main, n_res;       // chunks of 32 input channels
    int nN;                  // 256-column (output channel) N blocks per tile (1 or 2)
    int n_xbox, xbox_rows, x_stage_bytes, x_w_off;   // x_w_off: offset of the chunk's depthwise taps in a stage
    int xstages, bstages;
    int relu, mask_tail, aslots;
    int b0;                  // first utterance of this launch (sub-batch on its own stream)
    int n_tt, n_utt, n_cg;   // tiles: time tiles per utterance x utterances x output-channel groups
    unsigned long long* prof; // developer builds: [32] role cycle counters (VASR_TC_PROF=1), else null
    int tma_epi;              // 1: epilogue stages 128 x 32 output slices in smem and stores them with TMA
    int dbg;                  // developer builds: ablation bits (VASR_TC_DBG): 1 = skip MMAs, 2 = skip depthwise FMAs, 4 = skip epilogue stores
};

// Depthwise FIR of one chunk for one thread: channel pair `xs`/`wp` (already offset by the pair), R consecutive outputs
// starting at window row tw.  Rolling register window, fully unrolled over the taps: window slot of (output r, tap k)
// holds row tw + r + k*D.  Each tap consumes R FFMA2 and refills D rows + 1 tap weight that are needed P taps later,
// so shared-memory loads are spread evenly between the FMAs and only R + D*P rows are live.
// Depthwise tap storage of one 32-channel chunk.  VASR_DW_TAP128 = 1: taps are stored in pairs,
// [ceil(K/2)][16 channel pairs][tap 2j: ch0 ch1 | tap 2j+1: ch0 ch1], so that one 16-byte shared-memory load feeds two
// taps of a thread's channel pair (1.5 loads per tap instead of 2: fewer loads in flight per scoreboard).
// VASR_DW_TAP128 = 0: [K][32 channels], one 8-byte load per tap.
#ifndef VASR_DW_TAP128
#define VASR_DW_TAP128 0
#endif
__host__ __device__ constexpr int tap_floats(int K) { return VASR_DW_TAP128 ? ((K + 1) / 2) * 2 * KC : K * KC; }
// tap k of the channel pair whose taps start at `wp` (already offset by the pair)
__device__ __forceinline__ float2 tap_load(const float2* __restrict__ wp, int k)
{
#if VASR_DW_TAP128
    return wp[(size_t)(k >> 1) * KC + (k & 1)];               // pair stride: 16 pairs x 4 floats = KC float2
#else
    return wp[(size_t)k * (KC / 2)];
#endif
}
__device__ __forceinline__ const float2* tap_base(const void* taps, int cp)
{
#if VASR_DW_TAP128
    return reinterpret_cast<const float2*>(taps) + 2 * cp;
#else
    return reinterpret_cast<const float2*>(taps) + cp;
#endif
}

template <int K, int D, int R>
__device__ __forceinline__ void dw_chunk_s1(const float2* __restrict__ xs, const float2* __restrict__ wp, int tw, float2 (&acc)[R])
{
    constexpr int XP = KC / 2;
    // prefetch distance in taps.  Every tap issues two shared-memory loads and a consumer waits on the scoreboard of
    // its load; with 2P + 1 loads in flight and 6 scoreboards per warp, P > 2 makes loads share scoreboards, so a
    // consumer also waits for younger loads and the effective distance shrinks (profiles/: stall_short_sb)
#ifndef VASR_DW_PREFETCH
#define VASR_DW_PREFETCH 4
#endif
    constexpr int P = VASR_DW_PREFETCH;
    constexpr int WN = R + D * P;
    constexpr int LAST_ROW = (K - 1) * D + R - 1;
    float2 win[WN];
#pragma unroll
    for (int j = 0; j < WN; ++j) win[j] = (j <= LAST_ROW) ? xs[(size_t)(tw + j) * XP] : make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
#if VASR_DW_TAP128
    constexpr int K2 = (K + 1) / 2;
#ifndef VASR_DW_TAPPAIRS
#define VASR_DW_TAPPAIRS 2
#endif
    constexpr int PW = VASR_DW_TAPPAIRS;                       // tap pairs in flight
    const float4* wp4 = reinterpret_cast<const float4*>(wp);
    float4 wq[PW];
#pragma unroll
    for (int j = 0; j < PW; ++j) wq[j] = (j < K2) ? wp4[(size_t)j * (KC / 2)] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float4 w4 = wq[(k >> 1) % PW];
        const float2 wk = (k & 1) ? make_float2(w4.z, w4.w) : make_float2(w4.x, w4.y);
        if (((k & 1) || k + 1 == K) && (k >> 1) + PW < K2) wq[(k >> 1) % PW] = wp4[(size_t)((k >> 1) + PW) * (KC / 2)];
#else
    float2 wq[P];
#pragma unroll
    for (int j = 0; j < P; ++j) wq[j] = (j < K) ? tap_load(wp, j) : make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float2 wk = wq[k % P];
        if (k + P < K) wq[k % P] = tap_load(wp, k + P);
#endif
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = __ffma2_rn(wk, win[(k * D + r) % WN], acc[r]);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const int row = k * D + d + WN;                  // replaces the row that just went dead
            if (row <= LAST_ROW) win[(k * D + d) % WN] = xs[(size_t)(tw + row) * XP];
        }
    }
}

// Role cycle counters and ablation switches exist in developer builds only (make dev -> libvasr_b200_dev.so,
// -DVASR_DEV); the product library carries neither the counters nor any branch on them.
#ifdef VASR_DEV
#define PROF_DECL() unsigned long long pacc[4] = {0ull, 0ull, 0ull, 0ull}
#define PROF_BEGIN() long long _pt = clock64()
#define PROF_ADD(i) do { long long _n = clock64(); pacc[i] += (unsigned long long)(_n - _pt); _pt = _n; } while (0)
#define DBG_ON(bit) ((p.dbg & (bit)) != 0)
#else
#define PROF_DECL() do {} while (0)
#define PROF_BEGIN() do {} while (0)
#define PROF_ADD(i) do {} while (0)
#define DBG_ON(bit) false
#endif

// Persistent CTA (one per SM): tiles are claimed from an atomic counter by the scheduler thread and published
// to the other roles through a small shared-memory ring; all operand rings and the TMEM accumulator buffers
// keep running across tiles, so the epilogue of tile i overlaps the depthwise/MMA work of tile i+1.
template <int K, int S, int D, int NPART>
__global__ void __launch_bounds__(NTHREADS, 1)
subblock_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_r,
                const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                const __grid_constant__ CUtensorMap tm_r_hi, const __grid_constant__ CUtensorMap tm_r_lo,
                const __grid_constant__ CUtensorMap tm_out, const Params p)
{
    constexpr int WARP_X = NDW, WARP_A = NDW + 1, WARP_MMA = NDW + 2, WARP_EPI = NDW + 3;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve-up: every operand tile must be 1024-byte aligned.  The dynamic window starts right after the driver's
    // 1 KiB reservation, i.e. aligned; the budget has no slack for a round-up, so fail loudly if that ever changes.
    if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
    unsigned char* smem = smem_raw;
    constexpr int A_SLOT = W_PART * NPART, B_STAGE = PART_BYTES * NPART;   // a_ring = weight slots, b_ring = activation stages
    unsigned char* a_ring = smem;
    unsigned char* b_ring = a_ring + (size_t)p.aslots * A_SLOT;
    unsigned char* x_ring = b_ring + (size_t)p.bstages * B_STAGE;
    unsigned char* epi_stage = x_ring + (size_t)p.xstages * p.x_stage_bytes;      // [2][128 rows x 128 B] when tma_epi
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + (p.tma_epi ? 2 * EPI_STAGE_BYTES : 0));
    const int XSTAGES = p.xstages, BSTAGES = p.bstages;
    uint64_t* full_x = bars;                 // [xstages]  TMA -> dw warps
    uint64_t* empty_x = full_x + MAX_STAGES; // [xstages]  dw warps -> TMA
    uint64_t* full_b = empty_x + MAX_STAGES; // [bstages]  dw warps -> MMA
    uint64_t* empty_b = full_b + MAX_STAGES; // [bstages]  MMA (commit) -> dw warps
    uint64_t* full_a = empty_b + MAX_STAGES; // [aslots]   TMA -> MMA
    uint64_t* empty_a = full_a + 16;         // [aslots]   MMA (commit) -> TMA
    uint64_t* acc_full = empty_a + 16;       // [2]        MMA (commit) -> epilogue
    uint64_t* acc_empty = acc_full + 2;      // [2]        epilogue -> MMA
    uint64_t* sched_full = acc_empty + 2;    // [SCHED]    scheduler -> roles
    uint64_t* sched_empty = sched_full + SCHED;  // [SCHED] roles -> scheduler
    int* tile_ring = reinterpret_cast<int*>(sched_empty + SCHED);   // [SCHED]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tile_ring + SCHED);
    float* ep_shift = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 1024);   // [MAX_CO_CTA] BN shift of this tile's channels

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    PROF_DECL();
    const int nchunks = p.n_main + p.n_res;
    const int n_tiles = p.n_tt * p.n_utt * p.n_cg;
    const int acc_cols = p.nN * 256;                       // TMEM columns of one accumulator buffer
    const int nbuf = (acc_cols <= 256) ? 2 : 1;            // double-buffered when it fits the 512 columns

    if (threadIdx.x == 0) {
        for (int i = 0; i < XSTAGES; ++i) { mbar_init(full_x + i, 1); mbar_init(empty_x + i, NDW); }
        for (int i = 0; i < BSTAGES; ++i) { mbar_init(full_b + i, NDW); mbar_init(empty_b + i, 1); }
        for (int i = 0; i < p.aslots; ++i) { mbar_init(full_a + i, 1); mbar_init(empty_a + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, NEPI); }
        for (int i = 0; i < SCHED; ++i) { mbar_init(sched_full + i, 1); mbar_init(sched_empty + i, SCHED_CONSUMERS); }
        fence_barrier_init();
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (warp == WARP_X && lane == 0) { tma_prefetch_desc(&tm_x); if (p.n_res) tma_prefetch_desc(&tm_r); }
    if (warp == WARP_EPI && lane == 0 && p.tma_epi) tma_prefetch_desc(&tm_out);
    if (warp == WARP_A && lane == 0) {
        tma_prefetch_desc(&tm_w_hi);
        if (NPART == 2) tma_prefetch_desc(&tm_w_lo);
        if (p.n_res) { tma_prefetch_desc(&tm_r_hi); if (NPART == 2) tma_prefetch_desc(&tm_r_lo); }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // tile id -> (output-channel group, utterance, time tile)
    auto decode = [&](int tile, int& co0, int& b, int& t0) {
        const int tt = tile % p.n_tt;
        const int r = tile / p.n_tt;
        b = p.b0 + r % p.n_utt;
        co0 = (r / p.n_utt) * (p.nN * 256);
        t0 = tt * TN;
    };
    // consumer side of the tile ring: returns the tile id (or -1 = no more work)
    auto next_tile = [&](int ti) -> int {
        const int slot = ti % SCHED;
        mbar_wait(sched_full + slot, (ti / SCHED) & 1);
        const int tile = tile_ring[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(sched_empty + slot);
        return tile;
    };

    if (warp == WARP_X) {
        // ======== scheduler + TMA producer of the activation window (fp32, [rows x 32 ch], no swizzle) ========
        if (lane == 0) {
            int gc = 0;
            for (int ti = 0;; ++ti) {
                const int slot = ti % SCHED;
                mbar_wait(sched_empty + slot, ((ti / SCHED) & 1) ^ 1);
                int tile = atomicAdd(p.tile_counter, 1);
                if (tile >= n_tiles) tile = -1;
                tile_ring[slot] = tile;
                mbar_arrive(sched_full + slot);
                if (tile < 0) break;
                int co0, b, t0;
                decode(tile, co0, b, t0);
                for (int c = 0; c < nchunks; ++c, ++gc) {
                    const int s = gc % XSTAGES;
                    PROF_BEGIN();
                    mbar_wait(empty_x + s, ((gc / XSTAGES) & 1) ^ 1);
                    PROF_ADD(0);
                    unsigned char* dst = x_ring + (size_t)s * p.x_stage_bytes;
                    if (c < p.n_main) {
                        mbar_arrive_expect_tx(full_x + s, (uint32_t)(p.n_xbox * p.xbox_rows * KC * 4 + tap_floats(K) * 4));
                        for (int j = 0; j < p.n_xbox; ++j)
                            tma_load_3d(dst + (size_t)j * p.xbox_rows * KC * 4, &tm_x, c * KC,
                                        t0 * S - p.pad + j * p.xbox_rows, b, full_x + s);
                        // the chunk's depthwise taps [K][32] ride in the same stage (no exposed global-load latency)
                        bulk_load(dst + p.x_w_off, p.dw_w + (size_t)c * tap_floats(K), (uint32_t)(tap_floats(K) * 4), full_x + s);
                    } else {
                        mbar_arrive_expect_tx(full_x + s, (uint32_t)(TN * KC * 4));
                        tma_load_3d(dst, &tm_r, (c - p.n_main) * KC, t0, b, full_x + s);
                    }
                }
            }
        }
    } else if (warp == WARP_A) {
        // ======== TMA producer: weight slots [256 co x 32 ci] fp16 hi (+ lo), SWIZZLE_64B ========
        int slot = 0; uint32_t ph = 0;
        for (int ti = 0;; ++ti) {
            const int tile = next_tile(ti);
            if (tile < 0) break;
            int co0, b, t0;
            decode(tile, co0, b, t0);
            if (lane == 0) {
                for (int c = 0; c < nchunks; ++c) {
                    const bool res = c >= p.n_main;
                    const int ci0 = (res ? c - p.n_main : c) * KC;
                    for (int m = 0; m < p.nN; ++m) {
                        PROF_BEGIN();
                        mbar_wait(empty_a + slot, ph ^ 1);
                        PROF_ADD(0);
                        mbar_arrive_expect_tx(full_a + slot, (uint32_t)A_SLOT);
                        unsigned char* dst = a_ring + (size_t)slot * A_SLOT;
                        tma_load_2d(dst, res ? &tm_r_hi : &tm_w_hi, ci0, co0 + m * 256, full_a + slot);
                        if (NPART == 2) tma_load_2d(dst + W_PART, res ? &tm_r_lo : &tm_w_lo, ci0, co0 + m * 256, full_a + slot);
                        if (++slot == p.aslots) { slot = 0; ph ^= 1; }
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == WARP_MMA) {
        // ======== tcgen05.mma issuer (one thread) ========
        int slot = 0; uint32_t ph = 0;
        int gc = 0;
        for (int ti = 0;; ++ti) {
            const int tile = next_tile(ti);
            if (tile < 0) break;
            if (lane == 0) {
                const int ab = ti % nbuf;
                PROF_BEGIN();
                mbar_wait(acc_empty + ab, ((ti / nbuf) & 1) ^ 1);     // epilogue has drained this accumulator buffer
                PROF_ADD(2);
                tcgen05_fence_after();
                for (int c = 0; c < nchunks; ++c, ++gc) {
                    const int sb = gc % BSTAGES;
                    mbar_wait(full_b + sb, (gc / BSTAGES) & 1);
                    PROF_ADD(0);
                    tcgen05_fence_after();
                    const uint32_t b_addr = smem_u32(b_ring + (size_t)sb * B_STAGE);
                    for (int m = 0; m < p.nN; ++m) {
                        mbar_wait(full_a + slot, ph);
                        PROF_ADD(1);
                        tcgen05_fence_after();
                        const uint32_t w_addr = smem_u32(a_ring + (size_t)slot * A_SLOT);
                        const uint32_t d = tmem_base + (uint32_t)(ab * acc_cols + m * 256);
#pragma unroll
                        for (int ks = 0; ks < KC / 16; ++ks) {
                            // A operand = activations (M = 128 time rows), B operand = weights (N = 256 channels)
                            const uint64_t x_hi = make_desc_sw64(b_addr + ks * 32);
                            const uint64_t w_hi = make_desc_sw64(w_addr + ks * 32);
                            if (DBG_ON(1)) continue;
                            umma_f16(d, x_hi, w_hi, IDESC_F16_M128_N256, (c > 0 || ks > 0) ? 1u : 0u);
                            if (NPART == 2) {
                                const uint64_t x_lo = make_desc_sw64(b_addr + PART_BYTES + ks * 32);
                                const uint64_t w_lo = make_desc_sw64(w_addr + W_PART + ks * 32);
                                umma_f16(d, x_lo, w_hi, IDESC_F16_M128_N256, 1u);
                                umma_f16(d, x_hi, w_lo, IDESC_F16_M128_N256, 1u);
                            }
                        }
                        tcgen05_commit(empty_a + slot);        // weight slot reusable once these MMAs retire
                        PROF_ADD(3);
                        if (++slot == p.aslots) { slot = 0; ph ^= 1; }
                    }
                    tcgen05_commit(empty_b + sb);              // activation stage reusable
                }
                tcgen05_commit(acc_full + ab);                 // accumulators of this tile complete
            }
            __syncwarp();
        }
    } else if (warp >= WARP_EPI) {
        // ======== epilogue (8 warps): TMEM -> registers -> +shift, ReLU, mask -> global (channels-last) ========
        // TMEM lane = time row, column = output channel: a thread owns one output row of the tile and writes
        // 32 consecutive channels (128 bytes) per tcgen05.ld; the two warps of a lane quarter split the columns.
        const int q = warp & 3;                                // TMEM lane quarter this warp may access
        const int half = (warp - WARP_EPI) >> 2;               // which half of every N block's column groups
        const int row = q * 32 + lane;                         // tile row (time) owned by this thread
        const bool issuer = (q == 0 && lane == 0);             // one thread per half-group drives its TMA stores
        unsigned char* stage = epi_stage + half * EPI_STAGE_BYTES;
        const int nslice = p.nN * 4;                           // 32-column slices handled by this half-group
        const float wsc = p.wscale_inv;
        int cur_co0 = -1;
        for (int ti = 0;; ++ti) {
            const int tile = next_tile(ti);
            if (tile < 0) break;
            int co0, b, t0;
            decode(tile, co0, b, t0);
            if (co0 != cur_co0) {                              // (re)load the per-channel BN shift of this channel group
                named_bar_sync(3, NEPI * 32);                  // nobody still reads the previous group's values
                for (int i = (warp - WARP_EPI) * 32 + lane; i < p.nN * 256; i += NEPI * 32) ep_shift[i] = __ldg(p.shift + co0 + i);
                named_bar_sync(3, NEPI * 32);
                cur_co0 = co0;
            }
            const int ab = ti % nbuf;
            const int t = t0 + row;
            const bool row_ok = t < p.T_out;
            const bool live = !(p.mask_tail && t >= p.len_out[b]);
            float* orow = p.out + (size_t)b * p.out_bstride + (size_t)(row_ok ? t : 0) * p.Cout + co0;
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * acc_cols);
            auto slice_col = [&](int sidx) { return (sidx >> 2) * 256 + (half * 4 + (sidx & 3)) * 32; };
            // +shift (BN), ReLU, length mask on one 32-channel slice held in registers, then out
            auto finish = [&](uint32_t (&rg)[32], int col0) {
                // hoisted above the staging stores (see segment_kernel)
                constexpr bool HOIST = true;
                const float4* sh4 = reinterpret_cast<const float4*>(ep_shift + col0);
                float4 shv[HOIST ? 8 : 1];
                if (HOIST) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) shv[HOIST ? i : 0] = sh4[i];
                }
                if (p.tma_epi) {
                    if (issuer) bulk_wait_read0();             // previous slice has left the staging buffer
                    named_bar_sync(1 + half, 128);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 sh = HOIST ? shv[HOIST ? i : 0] : sh4[i];
                    float4 v;
                    v.x = fmaf(__uint_as_float(rg[4 * i + 0]), wsc, sh.x);
                    v.y = fmaf(__uint_as_float(rg[4 * i + 1]), wsc, sh.y);
                    v.z = fmaf(__uint_as_float(rg[4 * i + 2]), wsc, sh.z);
                    v.w = fmaf(__uint_as_float(rg[4 * i + 3]), wsc, sh.w);
                    if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    if (!live) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.tma_epi)                             // 128-byte rows, 16-byte chunks XOR-swizzled by row (SWIZZLE_128B)
                        *reinterpret_cast<float4*>(stage + row * 128 + ((i ^ (row & 7)) << 4)) = v;
                    else if (row_ok && !DBG_ON(4))
                        *reinterpret_cast<float4*>(orow + col0 + 4 * i) = v;
                }
                if (p.tma_epi) {
                    fence_proxy_async();
                    named_bar_sync(1 + half, 128);
                    if (issuer && !DBG_ON(4)) { tma_store_3d(&tm_out, stage, co0 + col0, t0, b); bulk_commit(); }
                }
            };
            PROF_BEGIN();
            mbar_wait(acc_full + ab, (ti / nbuf) & 1);
            PROF_ADD(0);
            tcgen05_fence_after();
            // one 32-column slice in registers at a time: with 608 threads the register budget (<= 104) does not
            // allow a second slice in flight without spilling to local memory, which costs far more than the
            // exposed tcgen05.ld latency
            uint32_t ra[32];
#pragma unroll 1
            for (int sidx = 0; sidx < nslice; ++sidx) {
                tmem_ld_32x32b_x32(tbase + (uint32_t)slice_col(sidx), ra);
                tmem_ld_wait();
                if (sidx + 1 == nslice) {                      // every TMEM read of this tile has completed
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + ab);
                }
                finish(ra, slice_col(sidx));
            }
            PROF_ADD(1);
        }
        if (p.tma_epi && issuer) bulk_wait_all0();
    } else if (warp < NDW) {
        // ======== depthwise producers (warps 0..NDW-1) ========
        // thread = one channel PAIR (packed fp32x2 FMAs, FFMA2) x R = 16 outputs:
        //   cp = lane & 15 -> channels 2cp, 2cp+1 of the chunk;  tg = 2*warp + (lane >> 4) -> t = R*tg + r
        // Every tap costs one window row + one tap weight from shared memory per R FFMA2.
        constexpr int R = TN / (2 * NDW);
        constexpr int NB = (K % 3 == 0 && K > 17) ? 3 : 1;     // taps are processed in NB register-window blocks
        constexpr int KB = K / NB;
        constexpr int XP = KC / 2;                               // float2 per window row
        const int cp = lane & 15;
        const int tw = (warp * 2 + (lane >> 4)) * R;
        float amax = 0.f;
        int gc = 0;
        for (int ti = 0;; ++ti) {
            const int tile = next_tile(ti);
            if (tile < 0) break;
            int co0, b, t0;
            decode(tile, co0, b, t0);
            const int len_mid = p.len_out[b];    // the 1x1 conv masks its input rows t >= len (parts/jasper.py:116)
            for (int c = 0; c < nchunks; ++c, ++gc) {
                const int sx = gc % XSTAGES, sb = gc % BSTAGES;
                PROF_BEGIN();
                mbar_wait(full_x + sx, (gc / XSTAGES) & 1);
                PROF_ADD(0);
                const float2* xs = reinterpret_cast<const float2*>(x_ring + (size_t)sx * p.x_stage_bytes) + cp;
                float2 acc[R];
                if (DBG_ON(2)) {
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = xs[(size_t)(tw + r) * XP];
                } else if (c < p.n_main) {
                    // depthwise taps of this chunk [K][32 ch], staged in shared memory next to the window
                    const float2* wp = tap_base(x_ring + (size_t)sx * p.x_stage_bytes + p.x_w_off, cp);
                    if (S == 1) {
                        dw_chunk_s1<K, D, R>(xs, wp, tw, acc);
                    } else {
                        // stride 2 (first block only, 2 chunks): tap-blocked window, row of (r, k) = (tw + r) * S + k
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
                        // (8 outputs at a time: the strided window of 16 would not fit the register budget)
#pragma unroll
                        for (int h = 0; h < R / 8; ++h) {
#pragma unroll 1
                            for (int kb = 0; kb < K; kb += KB) {
                                constexpr int WIN = 7 * S + KB;
                                float2 win[WIN];
#pragma unroll
                                for (int j = 0; j < WIN; ++j) win[j] = xs[(size_t)((tw + 8 * h) * S + kb + j) * XP];
#pragma unroll
                                for (int kk = 0; kk < KB; ++kk) {
                                    const float2 wk = tap_load(wp, kb + kk);
#pragma unroll
                                    for (int r = 0; r < 8; ++r) acc[8 * h + r] = __ffma2_rn(wk, win[r * S + kk], acc[8 * h + r]);
                                }
                            }
                        }
                    }
                } else {
                    // residual branch: the block input itself (1x1 conv only)
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = xs[(size_t)(tw + r) * XP];
                }
                // the depthwise output is not zero beyond len; the following MaskedConv1d zeroes it
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (t0 + tw + r >= len_mid) acc[r] = make_float2(0.f, 0.f);

                // range guard: |x| >= 65520 rounds to inf in fp16 (checked once per thread at the end of the kernel)
#pragma unroll
                for (int r = 0; r < R; ++r) amax = fmaxf(amax, fmaxf(fabsf(acc[r].x), fabsf(acc[r].y)));
                PROF_ADD(1);
                mbar_wait(empty_b + sb, ((gc / BSTAGES) & 1) ^ 1);
                PROF_ADD(2);
                unsigned char* bh = b_ring + (size_t)sb * B_STAGE;
                // sw64_offset(tw + r, cp >> 2) with tw a multiple of 8: everything but the XOR-ed 16-byte chunk is a
                // compile-time function of r -> four base pointers per chunk, immediate offsets per row
                unsigned char* bq[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) bq[q] = bh + (tw >> 3) * 512 + (cp & 3) * 4 + ((((uint32_t)cp >> 2) ^ (uint32_t)q) << 4);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t off = (uint32_t)((r >> 3) * 512 + (r & 7) * 64);
                    unsigned char* bh = bq[(r >> 1) & 3];
                    const __half2 h = __floats2half2_rn(acc[r].x, acc[r].y);
                    *reinterpret_cast<__half2*>(bh + off) = h;
                    if (NPART == 2) {
                        const float2 hf = __half22float2(h);
                        *reinterpret_cast<__half2*>(bh + PART_BYTES + off) = __floats2half2_rn(acc[r].x - hf.x, acc[r].y - hf.y);
                    }
                }
                fence_proxy_async();             // generic-proxy smem writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) { mbar_arrive(full_b + sb); mbar_arrive(empty_x + sx); }
                PROF_ADD(3);
            }
        }
        if (amax >= 65520.f) atomicOr(p.status, 1);          // a depthwise output left the fp16 range (see vasr_encoder_check)
    }
#ifdef VASR_DEV
    if (p.prof && lane == 0) {
        // slots: dw(warp 0) 0..3 = wait full_x / compute / wait empty_b / store; X producer 4 = wait empty_x;
        // A producer 5 = wait empty_a; MMA 6..9 = wait full_b / wait full_a / wait acc_empty / issue+commit; epilogue 10..11 = wait acc_full / drain
        int base = -1;
        if (warp == 0) base = 0; else if (warp == WARP_X) base = 4; else if (warp == WARP_A) base = 5;
        else if (warp == WARP_MMA) base = 6; else if (warp == WARP_EPI) base = 10;
        if (base >= 0)
            for (int i = 0; i < 4; ++i)
                if (pacc[i]) atomicAdd(p.prof + base + i, pacc[i]);
        if (warp == 1) atomicAdd(p.prof + 15, 1ull);
    }
#endif
    tcgen05_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

// =====================================================================================================================
// Multi-layer persistent kernel ("segment"): a run of consecutive stride-1 / dilation-1 separable sub-blocks with the
// same output width and the same number of frames executes as ONE launch.  The work list is layer-major
// (layer, utterance, time tile); a tile of layer l only needs the tiles of layer l-1 of the SAME utterance (the
// depthwise halo never leaves the utterance), so instead of a kernel boundary per layer there is one
// release/acquire counter per (layer, utterance): the epilogue bumps it once its TMA stores have completed, the tile
// scheduler waits for it before it issues the TMA loads of a dependent tile.  All rings, the TMEM buffers and the
// roles keep running across layers: no per-layer launch, pipeline fill/drain or partial last wave.
// =====================================================================================================================
struct alignas(128) LayerDesc {
    CUtensorMap tm_x, tm_r, tm_w_hi, tm_w_lo, tm_r_hi, tm_r_lo, tm_out;
    const float* dw_w;       // [Cin/32][K][32]
    const float* shift;      // [Cout]
    const int* len_out;      // [B]
    float* out;              // [B, T, Cout] output with an explicit batch stride (elements): direct stores of the pair kernel
    long long out_bstride;
    float wscale_inv;
    int K, n_main, n_res, relu, mask_tail, pad;
    int n_xbox, xbox_rows, x_w_off;
};

struct SegParams {
    const LayerDesc* layers;
    int n_layers;
    int* tile_counter;       // zeroed before the launch
    int* status;             // range guard: bit 0 is set when a depthwise output does not fit fp16 (zeroed per forward)
    int* done;               // [n_layers][done_stride] completion counters (zeroed): 2 * n_tt per finished (layer, utterance)
    int done_stride;
    int T_out, nN;
    int x_stage_bytes, xstages, bstages, aslots;
    int epi_bufs;            // pair kernel: 4 KiB store-staging buffers per epilogue warp (2..4)
    int b0, n_tt, n_utt;
    unsigned long long* prof;
    int dbg;
};

__device__ __forceinline__ int ld_acquire_gpu(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// cross-layer dependency wait with a watchdog (~seconds): a protocol error must end in a launch failure that the
// host reports, never in a hung GPU
__device__ __forceinline__ void dep_wait(const int* flag, int need)
{
    unsigned spins = 0;
    while (ld_acquire_gpu(flag) < need) {
        __nanosleep(40);
        if (++spins > (1u << 26)) __trap();
    }
}

#define SEG_K_SWITCH(Kv, STMT)                                   \
    switch (Kv) {                                                \
        case 11: { constexpr int KK = 11; STMT; } break;         \
        case 33: { constexpr int KK = 33; STMT; } break;         \
        case 39: { constexpr int KK = 39; STMT; } break;         \
        case 51: { constexpr int KK = 51; STMT; } break;         \
        case 63: { constexpr int KK = 63; STMT; } break;         \
        default: { constexpr int KK = 75; STMT; } break;         \
    }
static bool seg_kernel_size(int K) { return K == 11 || K == 33 || K == 39 || K == 51 || K == 63 || K == 75; }

// Single-CTA version (480 threads: one depthwise warp per scheduler, all four on the same chunk), dynamic tile
// scheduler.  TR = valid time rows per tile: 32 = "latency mode" for small batches (four times as many tiles per layer,
// each with a quarter of the depthwise work per thread, so a lone utterance spreads over 4x the SMs and a layer takes
// ~1/3 of the time; the MMA still computes M = 128 rows - rows >= TR of the operand are stale and their accumulator
// rows are never stored); 128 = the shape of the pair kernel, used where a cluster launch is not possible.
template <int NPART, int TR>
__global__ void __launch_bounds__(NTHREADS, 1)
segment_kernel(const SegParams p)
{
    static_assert(TR == TN || TR == 32, "tile rows: 128 or 32");
    constexpr int GW = NDW;                                 // depthwise warps that share a chunk
    constexpr int WARP_X = NDW, WARP_A = NDW + 1, WARP_MMA = NDW + 2, WARP_EPI = NDW + 3;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
    unsigned char* smem = smem_raw;
    constexpr int A_SLOT = W_PART * NPART, B_STAGE = PART_BYTES * NPART;
    unsigned char* a_ring = smem;
    unsigned char* b_ring = a_ring + (size_t)p.aslots * A_SLOT;
    unsigned char* x_ring = b_ring + (size_t)p.bstages * B_STAGE;
    unsigned char* epi_stage = x_ring + (size_t)p.xstages * p.x_stage_bytes;      // [2][128 rows x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + 2 * EPI_STAGE_BYTES);
    const int XSTAGES = p.xstages, BSTAGES = p.bstages;
    uint64_t* full_x = bars;
    uint64_t* empty_x = full_x + MAX_STAGES;
    uint64_t* full_b = empty_x + MAX_STAGES;
    uint64_t* empty_b = full_b + MAX_STAGES;
    uint64_t* full_a = empty_b + MAX_STAGES;
    uint64_t* empty_a = full_a + 16;
    uint64_t* acc_full = empty_a + 16;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* sched_full = acc_empty + 2;
    uint64_t* sched_empty = sched_full + SCHED;
    int* tile_ring = reinterpret_cast<int*>(sched_empty + SCHED);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tile_ring + SCHED);
    float* ep_shift = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 1024);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    PROF_DECL();
    const int tpl = p.n_tt * p.n_utt;                      // tiles per layer
    const int n_items = tpl * p.n_layers;
    const int acc_cols = p.nN * 256;
    const int nbuf = (acc_cols <= 256) ? 2 : 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < XSTAGES; ++i) { mbar_init(full_x + i, 1); mbar_init(empty_x + i, GW); }
        for (int i = 0; i < BSTAGES; ++i) { mbar_init(full_b + i, GW); mbar_init(empty_b + i, 1); }
        for (int i = 0; i < p.aslots; ++i) { mbar_init(full_a + i, 1); mbar_init(empty_a + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, NEPI); }
        for (int i = 0; i < SCHED; ++i) { mbar_init(sched_full + i, 1); mbar_init(sched_empty + i, SCHED_CONSUMERS); }
        fence_barrier_init();
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // item -> (layer, utterance, time tile)
    auto decode = [&](int item, int& l, int& b, int& t0) {
        l = item / tpl;
        const int r = item - l * tpl;
        b = p.b0 + r / p.n_tt;
        t0 = (r % p.n_tt) * TR;
    };
    auto next_tile = [&](int ti) -> int {
        const int slot = ti % SCHED;
        mbar_wait(sched_full + slot, (ti / SCHED) & 1);
        const int tile = tile_ring[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(sched_empty + slot);
        return tile;
    };

    if (warp >= NDW) {
    if (warp == WARP_X) {
        // ======== scheduler (+ cross-layer dependency wait) + TMA producer of the activation window ========
        if (lane == 0) {
            int s = 0; uint32_t xph = 0;
            for (int ti = 0;; ++ti) {
                const int slot = ti % SCHED;
                mbar_wait(sched_empty + slot, ((ti / SCHED) & 1) ^ 1);
                int tile = atomicAdd(p.tile_counter, 1);
                if (tile >= n_items) tile = -1;
                tile_ring[slot] = tile;
                mbar_arrive(sched_full + slot);
                if (tile < 0) break;
                int l, b, t0;
                decode(tile, l, b, t0);
                const LayerDesc* L = p.layers + l;
                if (l > 0) {
                    // every tile of layer l-1 of this utterance has been stored (both epilogue halves of each time tile)
                    const int* flag = p.done + (size_t)(l - 1) * p.done_stride + b;
                    const int need = 2 * p.n_tt;
                    PROF_BEGIN();
                    dep_wait(flag, need);
                    PROF_ADD(1);
                    fence_proxy_async_all();        // the TMA (async proxy) reads below are ordered after the acquire
                }
                const int K = L->K, n_main = L->n_main, nch = L->n_main + L->n_res;
                const int n_xbox = L->n_xbox, xbox_rows = L->xbox_rows, x_w_off = L->x_w_off, pad = L->pad;
                const float* dw_w = L->dw_w;
                for (int c = 0; c < nch; ++c) {
                    PROF_BEGIN();
                    mbar_wait(empty_x + s, xph ^ 1);
                    PROF_ADD(0);
                    unsigned char* dst = x_ring + (size_t)s * p.x_stage_bytes;
                    if (DBG_ON(16)) {
                        mbar_arrive(full_x + s);
                    } else if (c < n_main) {
                        mbar_arrive_expect_tx(full_x + s, (uint32_t)(n_xbox * xbox_rows * KC * 4 + tap_floats(K) * 4));
                        for (int j = 0; j < n_xbox; ++j)
                            tma_load_3d(dst + (size_t)j * xbox_rows * KC * 4, &L->tm_x, c * KC, t0 - pad + j * xbox_rows, b, full_x + s);
                        bulk_load(dst + x_w_off, dw_w + (size_t)c * tap_floats(K), (uint32_t)(tap_floats(K) * 4), full_x + s);
                    } else {
                        mbar_arrive_expect_tx(full_x + s, (uint32_t)(TR * KC * 4));
                        tma_load_3d(dst, &L->tm_r, (c - n_main) * KC, t0, b, full_x + s);
                    }
                    if (++s == XSTAGES) { s = 0; xph ^= 1; }
                }
            }
        }
    } else if (warp == WARP_A) {
        // ======== TMA producer: weight slots ========
        int slot = 0; uint32_t ph = 0;
        for (int ti = 0;; ++ti) {
            const int tile = next_tile(ti);
            if (tile < 0) break;
            if (lane == 0) {
                int l, b, t0;
                decode(tile, l, b, t0);
                const LayerDesc* L = p.layers + l;
                const int n_main = L->n_main, nch = L->n_main + L->n_res;
                for (int c = 0; c < nch; ++c) {
                    const bool res = c >= n_main;
                    const int ci0 = (res ? c - n_main : c) * KC;
                    for (int m = 0; m < p.nN; ++m) {
                        PROF_BEGIN();
                        mbar_wait(empty_a + slot, ph ^ 1);
                        PROF_ADD(0);
                        unsigned char* dst = a_ring + (size_t)slot * A_SLOT;
                        if (DBG_ON(8)) {
                            mbar_arrive(full_a + slot);
                        } else {
                            mbar_arrive_expect_tx(full_a + slot, (uint32_t)A_SLOT);
                            tma_load_2d(dst, res ? &L->tm_r_hi : &L->tm_w_hi, ci0, m * 256, full_a + slot);
                            if (NPART == 2) tma_load_2d(dst + W_PART, res ? &L->tm_r_lo : &L->tm_w_lo, ci0, m * 256, full_a + slot);
                        }
                        if (++slot == p.aslots) { slot = 0; ph ^= 1; }
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == WARP_MMA) {
        // ======== tcgen05.mma issuer (one thread) ========
        int slot = 0; uint32_t ph = 0;
        int sb = 0; uint32_t bph = 0;
        int ab = 0; uint32_t accph = 0;
        for (int ti = 0;; ++ti) {
            const int tile = next_tile(ti);
            if (tile < 0) break;
            if (lane == 0) {
                int l, b, t0;
                decode(tile, l, b, t0);
                const int nch = p.layers[l].n_main + p.layers[l].n_res;
                PROF_BEGIN();
                mbar_wait(acc_empty + ab, accph ^ 1);
                PROF_ADD(2);
                tcgen05_fence_after();
                for (int c = 0; c < nch; ++c) {
                    mbar_wait(full_b + sb, bph);
                    PROF_ADD(0);
                    tcgen05_fence_after();
                    const uint32_t b_addr = smem_u32(b_ring + (size_t)sb * B_STAGE);
                    for (int m = 0; m < p.nN; ++m) {
                        mbar_wait(full_a + slot, ph);
                        PROF_ADD(1);
                        tcgen05_fence_after();
                        const uint32_t w_addr = smem_u32(a_ring + (size_t)slot * A_SLOT);
                        const uint32_t d = tmem_base + (uint32_t)(ab * acc_cols + m * 256);
#pragma unroll
                        for (int ks = 0; ks < KC / 16; ++ks) {
                            const uint64_t x_hi = make_desc_sw64(b_addr + ks * 32);
                            const uint64_t w_hi = make_desc_sw64(w_addr + ks * 32);
                            if (DBG_ON(1)) continue;
                            umma_f16(d, x_hi, w_hi, IDESC_F16_M128_N256, (c > 0 || ks > 0) ? 1u : 0u);
                            if (NPART == 2) {
                                const uint64_t x_lo = make_desc_sw64(b_addr + PART_BYTES + ks * 32);
                                const uint64_t w_lo = make_desc_sw64(w_addr + W_PART + ks * 32);
                                umma_f16(d, x_lo, w_hi, IDESC_F16_M128_N256, 1u);
                                umma_f16(d, x_hi, w_lo, IDESC_F16_M128_N256, 1u);
                            }
                        }
                        tcgen05_commit(empty_a + slot);
                        PROF_ADD(3);
                        if (++slot == p.aslots) { slot = 0; ph ^= 1; }
                    }
                    tcgen05_commit(empty_b + sb);
                    if (++sb == BSTAGES) { sb = 0; bph ^= 1; }
                }
                tcgen05_commit(acc_full + ab);
                if (++ab == nbuf) { ab = 0; accph ^= 1; }
            }
            __syncwarp();
        }
    } else if (warp >= WARP_EPI) {
        // ======== epilogue (8 warps): TMEM -> +shift, ReLU, mask -> smem staging -> TMA store -> completion counter ========
        const int q = warp & 3;
        const int half = (warp - WARP_EPI) >> 2;
        const int row = q * 32 + lane;
        const bool issuer = (q == 0 && lane == 0);
        unsigned char* stage = epi_stage + half * EPI_STAGE_BYTES;
        const int nslice = p.nN * 4;
        int cur_l = -1;
        float wsc = 1.f; int relu = 0, mask_tail = 1; const int* len_out = nullptr;
        int ab = 0; uint32_t accph = 0;
        for (int ti = 0;; ++ti) {
            const int tile = next_tile(ti);
            if (tile < 0) break;
            int l, b, t0;
            decode(tile, l, b, t0);
            const LayerDesc* L = p.layers + l;
            if (l != cur_l) {                                  // per-channel BN shift + scalars of the new layer
                named_bar_sync(3, NEPI * 32);
                const float* shift = L->shift;
                for (int i = (warp - WARP_EPI) * 32 + lane; i < p.nN * 256; i += NEPI * 32) ep_shift[i] = __ldg(shift + i);
                wsc = L->wscale_inv; relu = L->relu; mask_tail = L->mask_tail; len_out = L->len_out;
                named_bar_sync(3, NEPI * 32);
                cur_l = l;
            }
            const int t = t0 + row;
            const bool live = !(mask_tail && t >= len_out[b]);
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * acc_cols);
            auto slice_col = [&](int sidx) { return (sidx >> 2) * 256 + (half * 4 + (sidx & 3)) * 32; };
            PROF_BEGIN();
            mbar_wait(acc_full + ab, accph);
            PROF_ADD(0);
            tcgen05_fence_after();
            const int ab_cur = ab;
            if (++ab == nbuf) { ab = 0; accph ^= 1; }
            uint32_t ra[32];
#pragma unroll 1
            for (int sidx = 0; sidx < nslice; ++sidx) {
                const int col0 = slice_col(sidx);
                tmem_ld_32x32b_x32(tbase + (uint32_t)col0, ra);
                // the slice's BN shift (shared-memory broadcast) is fetched while the TMEM load is in flight; loading it
                // inside the loop below would chain every LDS behind the previous staging STS (possible alias) and
                // expose its latency eight times per slice
                constexpr int NH = 8;
                const float4* sh4 = reinterpret_cast<const float4*>(ep_shift + col0);
                float4 shv[NH];
#pragma unroll
                for (int i = 0; i < NH; ++i) shv[i] = sh4[i];
                tmem_ld_wait();
                if (sidx + 1 == nslice) {                      // every TMEM read of this tile has completed
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + ab_cur);
                }
                if (issuer) bulk_wait_read0();                 // previous slice has left the staging buffer
                named_bar_sync(1 + half, 128);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 sh = shv[i % NH];
                    float4 v;
                    v.x = fmaf(__uint_as_float(ra[4 * i + 0]), wsc, sh.x);
                    v.y = fmaf(__uint_as_float(ra[4 * i + 1]), wsc, sh.y);
                    v.z = fmaf(__uint_as_float(ra[4 * i + 2]), wsc, sh.z);
                    v.w = fmaf(__uint_as_float(ra[4 * i + 3]), wsc, sh.w);
                    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    if (!live) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!DBG_ON(32)) *reinterpret_cast<float4*>(stage + row * 128 + ((i ^ (row & 7)) << 4)) = v;
                }
                fence_proxy_async();
                named_bar_sync(1 + half, 128);
                if (issuer && !DBG_ON(4)) { tma_store_3d(&L->tm_out, stage, col0, t0, b); bulk_commit(); }
            }
            if (issuer) {
                // this half's part of the tile is in global memory: publish it to the tiles of the next layer
                bulk_wait_all0();
                fence_proxy_async_all();
                __threadfence();
                atomicAdd(p.done + (size_t)l * p.done_stride + b, 1);
            }
            PROF_ADD(1);
        }
    }
    } else {
        // ======== depthwise producers ========
        constexpr int R = TR / (2 * GW);
        constexpr int XP = KC / 2;
        const int wg = warp & (GW - 1);
        const int cp = lane & 15;
        const int tw = (wg * 2 + (lane >> 4)) * R;
        int sx = 0, sb = 0; uint32_t xph = 0, bph = 0;
        float amax = 0.f;
        int gc = 0;                                         // running chunk index over all tiles (both groups count all)
        for (int ti = 0;; ++ti) {
            const int tile = next_tile(ti);
            if (tile < 0) break;
            int l, b, t0;
            decode(tile, l, b, t0);
            const LayerDesc* L = p.layers + l;
            const int K = L->K, n_main = L->n_main, nch = L->n_main + L->n_res, x_w_off = L->x_w_off;
            const int len_mid = L->len_out[b];
            const bool tail_tile = t0 + TR > len_mid;      // only tiles that straddle the utterance's end need the row mask
            for (int c = 0; c < nch; ++c, ++gc) {
                PROF_BEGIN();
                mbar_wait(full_x + sx, xph);
                PROF_ADD(0);
                const float2* xs = reinterpret_cast<const float2*>(x_ring + (size_t)sx * p.x_stage_bytes) + cp;
                float2 acc[R];
                if (c < n_main) {
                    const float2* wp = tap_base(x_ring + (size_t)sx * p.x_stage_bytes + x_w_off, cp);
                    SEG_K_SWITCH(K, (dw_chunk_s1<KK, 1, R>(xs, wp, tw, acc)));
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = xs[(size_t)(tw + r) * XP];
                }
                if (tail_tile) {
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (t0 + tw + r >= len_mid) acc[r] = make_float2(0.f, 0.f);
                }
                // range guard: |x| >= 65520 rounds to inf in fp16 (checked once per thread at the end of the kernel)
#pragma unroll
                for (int r = 0; r < R; ++r) amax = fmaxf(amax, fmaxf(fabsf(acc[r].x), fabsf(acc[r].y)));
                PROF_ADD(1);
                mbar_wait(empty_b + sb, bph ^ 1);
                PROF_ADD(2);
                unsigned char* bh0 = b_ring + (size_t)sb * B_STAGE;
                unsigned char* bq[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) bq[q] = bh0 + (tw >> 3) * 512 + (cp & 3) * 4 + ((((uint32_t)cp >> 2) ^ (uint32_t)q) << 4);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (DBG_ON(64)) break;
                    // (tw is a multiple of 8 only when R >= 8; the 4-output latency tiles take the general formula)
                    const uint32_t off = (R >= 8) ? (uint32_t)((r >> 3) * 512 + (r & 7) * 64)
                                                  : sw64_offset(tw + r, cp >> 2) + (cp & 3) * 4;
                    unsigned char* bh = (R >= 8) ? bq[(r >> 1) & 3] : bh0;
                    const __half2 h = __floats2half2_rn(acc[r].x, acc[r].y);
                    *reinterpret_cast<__half2*>(bh + off) = h;
                    if (NPART == 2) {
                        const float2 hf = __half22float2(h);
                        *reinterpret_cast<__half2*>(bh + PART_BYTES + off) = __floats2half2_rn(acc[r].x - hf.x, acc[r].y - hf.y);
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { mbar_arrive(full_b + sb); mbar_arrive(empty_x + sx); }
                if (++sx == XSTAGES) { sx = 0; xph ^= 1; }
                if (++sb == BSTAGES) { sb = 0; bph ^= 1; }
                PROF_ADD(3);
            }
        }
        if (amax >= 65520.f) atomicOr(p.status, 1);          // a depthwise output left the fp16 range (see vasr_encoder_check)
    }
#ifdef VASR_DEV
    if (p.prof && lane == 0) {
        // slots as in subblock_kernel; slot 5 (xprod) additionally: cycles spent waiting for a cross-layer dependency
        int base = -1;
        if (warp == 0) base = 0; else if (warp == WARP_X) base = 4; else if (warp == WARP_A) base = 12;
        else if (warp == WARP_MMA) base = 6; else if (warp == WARP_EPI) base = 10;
        if (base >= 0)
            for (int i = 0; i < 4; ++i)
                if (pacc[i]) atomicAdd(p.prof + base + i, pacc[i]);
        if (warp == 1) atomicAdd(p.prof + 15, 1ull);
    }
#endif
    tcgen05_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

// =====================================================================================================================
// CTA-pair segment kernel (the throughput path): a cluster of two CTAs works on two tiles of the same layer at once
// (tiles 2u and 2u+1 of the layer's list).  Each CTA does everything for ITS tile - window TMA, depthwise, operand
// stores, epilogue - but the 1x1 convolutions are issued by the leader as tcgen05.mma.cta_group::2 (M = 256 over both
// SMs): the weight block of an instruction (256 output channels x 32 input channels) is split between the CTAs, so a
// weight slot is 16 KiB instead of 32, every SM streams half of the weights from L2 and the tensor core of each SM
// reads half of the B operand.  The shared memory saved - and the 32 KiB of store staging that the register-direct
// epilogue does not need - pay for the third window stage that the two-group depthwise needs and for deeper weight
// prefetch.
//
// Depthwise: 8 warps in two groups (warps 0-3 / 4-7) that take alternate chunks, i.e. two depthwise warps per SM
// sub-partition whose FMA streams fill each other's stalls and fp16-split/store phases.  512 threads, so every
// thread may use 128 registers (no setmaxnreg re-partitioning needed): warps 0-7 depthwise, 8 window producer,
// 9 weight producer, 10 MMA issuer, 11 idle, 12-15 epilogue (one per TMEM lane quarter).
//
// Work list: static.  Item i = (layer, pair of tiles); cluster c executes items c, c + n_clusters, ... in increasing
// order.  An item only depends on items with a smaller index (previous layer, same utterance), and every cluster works
// through its list in order, so the smallest unfinished item can always run: no deadlock as long as the kernel's
// clusters become resident eventually (grid <= SM count).  A layer with an odd number of tiles pairs its last tile
// with a duplicate of itself: the peer computes it too but neither stores nor publishes it.
//
// Cross-CTA protocol (barriers live at the same offsets in both CTAs; "L:" = the leader's copy is the one in use):
//   L:full_b[stage]   8 arrivals: the 4 depthwise warps of the chunk's group in EACH cta (remote arrive for rank 1)
//   empty_b, empty_a, acc_full              released in both CTAs by tcgen05.commit ... multicast::cluster (mask 0b11)
//   L:full_a[slot]    both CTAs' weight TMA loads (cp.async.bulk.tensor ... cta_group::2) complete on it
//   L:acc_empty       16 arrivals: the epilogue warps of both CTAs
// Layer-to-layer hand-over: every epilogue warp bumps the (layer, utterance) counter after its stores (release at gpu
// scope); a tile of the next layer starts once all NEPI x n_tt warps of the utterance's previous layer have done so.
// Remote arrives use the default (release.cta) semantics like CUTLASS's ClusterBarrier::arrive: what they publish is
// either consumed by the arriving CTA's own tensor core (operand stage: generic-proxy stores + fence.proxy.async in
// the writing CTA) or is no data at all (accumulator drained).  A cluster-scope release costs a MEMBAR.ALL.GPU per
// arrive (measured: +1600 cycles per chunk in the depthwise store phase).
// =====================================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // shared::cluster address -> same offset in the even (leader) CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_2sm(uint64_t* bar)          // arrives on `bar` in BOTH CTAs of the pair
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// D = f32, A = B = f16, both K-major, N = 256, M = 256 (128 rows in each CTA of the pair)
constexpr uint32_t IDESC_F16_M256_N256 = (1u << 4) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
constexpr int W_HALF = 128 * KC * 2;            // this CTA's half of a weight block: [128 co x 64 B] fp16 = 8 KiB per part
constexpr int PAIR_THREADS = 512;              // 8 depthwise warps, window / weight / MMA warps (+ 1 idle), 4 epilogue warps
constexpr int NEPI_PAIR = 4;                    // one epilogue warp per TMEM lane quarter
constexpr int EPI_WARP_BYTES = 32 * 32 * 4;     // one 32-row x 32-channel fp32 block (SWIZZLE_128B); two per epilogue warp

template <int NPART>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
segment_pair_kernel(const SegParams p)
{
    constexpr int NDW_PAIR = 8, GW = 4;                    // depthwise warps (two groups of GW)
    constexpr int WARP_X = 8, WARP_A = 9, WARP_MMA = 10, WARP_EPI = 12;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
    unsigned char* smem = smem_raw;
    constexpr int A_SLOT = W_HALF * NPART, B_STAGE = PART_BYTES * NPART;
    unsigned char* a_ring = smem;
    unsigned char* b_ring = a_ring + (size_t)p.aslots * A_SLOT;
    unsigned char* x_ring = b_ring + (size_t)p.bstages * B_STAGE;
    unsigned char* epi_stage = x_ring + (size_t)p.xstages * p.x_stage_bytes;      // [4 warps][2][32 rows x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + (size_t)NEPI_PAIR * p.epi_bufs * EPI_WARP_BYTES);
    const int XSTAGES = p.xstages, BSTAGES = p.bstages;
    uint64_t* full_x = bars;
    uint64_t* empty_x = full_x + MAX_STAGES;
    uint64_t* full_b = empty_x + MAX_STAGES;
    uint64_t* empty_b = full_b + MAX_STAGES;
    uint64_t* full_a = empty_b + MAX_STAGES;
    uint64_t* empty_a = full_a + 16;
    uint64_t* acc_full = empty_a + 16;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* ep_shift = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 1024);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    PROF_DECL();
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int tpl = p.n_tt * p.n_utt;                      // tiles per layer
    const int ppl = (tpl + 1) >> 1;                        // tile pairs per layer (the last one may hold a duplicate)
    const int n_items = ppl * p.n_layers;
    const int item0 = (int)(blockIdx.x >> 1), item_step = (int)(gridDim.x >> 1);
    const int acc_cols = p.nN * 256;
    const int nbuf = (acc_cols <= 256) ? 2 : 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < XSTAGES; ++i) { mbar_init(full_x + i, 1); mbar_init(empty_x + i, GW); }
        for (int i = 0; i < BSTAGES; ++i) { mbar_init(full_b + i, 2 * GW); mbar_init(empty_b + i, 1); }
        for (int i = 0; i < p.aslots; ++i) { mbar_init(full_a + i, 1); mbar_init(empty_a + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 2 * NEPI_PAIR); }
        fence_barrier_init();
    }
    if (warp == WARP_MMA) {                                // one warp of each CTA, same warp id in both
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                                    // barriers of both CTAs are initialised before any remote arrive
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // item -> (layer, utterance, time tile) of THIS cta's tile; dup = the layer's odd last tile, computed by both CTAs
    auto decode = [&](int item, int& l, int& b, int& t0, bool& dup) {
        l = item / ppl;
        int r = 2 * (item - l * ppl) + (int)rank;
        dup = r >= tpl;
        if (dup) r = tpl - 1;
        b = p.b0 + r / p.n_tt;
        t0 = (r % p.n_tt) * TN;
    };

    if (warp >= NDW_PAIR) {
    if (warp == WARP_X) {
        // ======== TMA producer of this CTA's activation windows (+ cross-layer dependency wait) ========
        if (lane == 0) {
            int s = 0; uint32_t xph = 0;
            for (int item = item0; item < n_items; item += item_step) {
                int l, b, t0; bool dup;
                decode(item, l, b, t0, dup);
                const LayerDesc* L = p.layers + l;
                if (l > 0) {
                    // every tile of layer l-1 of this utterance has been stored (all four epilogue warps of each time tile)
                    const int* flag = p.done + (size_t)(l - 1) * p.done_stride + b;
                    const int need = NEPI_PAIR * p.n_tt;
                    PROF_BEGIN();
                    dep_wait(flag, need);
                    PROF_ADD(1);
                    fence_proxy_async_all();        // the TMA (async proxy) reads below are ordered after the acquire
                }
                const int K = L->K, n_main = L->n_main, nch = L->n_main + L->n_res;
                const int n_xbox = L->n_xbox, xbox_rows = L->xbox_rows, x_w_off = L->x_w_off, pad = L->pad;
                const float* dw_w = L->dw_w;
                for (int c = 0; c < nch; ++c) {
                    PROF_BEGIN();
                    mbar_wait(empty_x + s, xph ^ 1);
                    PROF_ADD(0);
                    unsigned char* dst = x_ring + (size_t)s * p.x_stage_bytes;
                    if (DBG_ON(16)) {                                  // ablation: no window traffic (stale windows)
                        mbar_arrive(full_x + s);
                    } else if (c < n_main) {
                        mbar_arrive_expect_tx(full_x + s, (uint32_t)(n_xbox * xbox_rows * KC * 4 + tap_floats(K) * 4));
                        for (int j = 0; j < n_xbox; ++j)
                            tma_load_3d(dst + (size_t)j * xbox_rows * KC * 4, &L->tm_x, c * KC, t0 - pad + j * xbox_rows, b, full_x + s);
                        bulk_load(dst + x_w_off, dw_w + (size_t)c * tap_floats(K), (uint32_t)(tap_floats(K) * 4), full_x + s);
                    } else {
                        mbar_arrive_expect_tx(full_x + s, (uint32_t)(TN * KC * 4));
                        tma_load_3d(dst, &L->tm_r, (c - n_main) * KC, t0, b, full_x + s);
                    }
                    if (++s == XSTAGES) { s = 0; xph ^= 1; }
                }
            }
        }
    } else if (warp == WARP_A) {
        // ======== weight producer: this CTA's 128 of the 256 output channels of every block ========
        if (lane == 0) {
            int slot = 0; uint32_t ph = 0;
            for (int item = item0; item < n_items; item += item_step) {
                const LayerDesc* L = p.layers + item / ppl;
                const int n_main = L->n_main, nch = L->n_main + L->n_res;
                for (int c = 0; c < nch; ++c) {
                    const bool res = c >= n_main;
                    const int ci0 = (res ? c - n_main : c) * KC;
                    for (int m = 0; m < p.nN; ++m) {
                        PROF_BEGIN();
                        mbar_wait(empty_a + slot, ph ^ 1);                    // released in both CTAs by the multicast commit
                        PROF_ADD(0);
                        // all bytes of the slot (both halves) are accounted on the leader's barrier
                        if (DBG_ON(8)) {                               // ablation: no weight traffic (stale operands)
                            if (leader) mbar_arrive(full_a + slot);
                            if (++slot == p.aslots) { slot = 0; ph ^= 1; }
                            continue;
                        }
                        if (leader) mbar_arrive_expect_tx(full_a + slot, (uint32_t)(2 * A_SLOT));
                        unsigned char* dst = a_ring + (size_t)slot * A_SLOT;
                        const int co = m * 256 + (int)rank * 128;
                        tma_load_2d_2sm(dst, res ? &L->tm_r_hi : &L->tm_w_hi, ci0, co, full_a + slot);
                        if (NPART == 2) tma_load_2d_2sm(dst + W_HALF, res ? &L->tm_r_lo : &L->tm_w_lo, ci0, co, full_a + slot);
                        if (++slot == p.aslots) { slot = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // ======== tcgen05.mma.cta_group::2 issuer: one thread of the leader ========
        if (leader && lane == 0) {
            int slot = 0; uint32_t ph = 0;
            int sb = 0; uint32_t bph = 0;
            int ab = 0; uint32_t accph = 0;
            for (int item = item0; item < n_items; item += item_step) {
                const LayerDesc* L = p.layers + item / ppl;
                const int nch = L->n_main + L->n_res;
                PROF_BEGIN();
                mbar_wait(acc_empty + ab, accph ^ 1);            // both epilogues have drained this buffer
                PROF_ADD(2);
                tcgen05_fence_after();
                for (int c = 0; c < nch; ++c) {
                    mbar_wait(full_b + sb, bph);                  // both CTAs' depthwise warps have published the chunk
                    PROF_ADD(0);
                    tcgen05_fence_after();
                    const uint32_t b_addr = smem_u32(b_ring + (size_t)sb * B_STAGE);
                    for (int m = 0; m < p.nN; ++m) {
                        mbar_wait(full_a + slot, ph);
                        PROF_ADD(1);
                        tcgen05_fence_after();
                        const uint32_t w_addr = smem_u32(a_ring + (size_t)slot * A_SLOT);
                        const uint32_t d = tmem_base + (uint32_t)(ab * acc_cols + m * 256);
#pragma unroll
                        for (int ks = 0; ks < KC / 16; ++ks) {
                            const uint64_t x_hi = make_desc_sw64(b_addr + ks * 32);
                            const uint64_t w_hi = make_desc_sw64(w_addr + ks * 32);
                            if (DBG_ON(1)) continue;
                            umma_f16_2sm(d, x_hi, w_hi, IDESC_F16_M256_N256, (c > 0 || ks > 0) ? 1u : 0u);
                            if (NPART == 2) {
                                const uint64_t x_lo = make_desc_sw64(b_addr + PART_BYTES + ks * 32);
                                const uint64_t w_lo = make_desc_sw64(w_addr + W_HALF + ks * 32);
                                umma_f16_2sm(d, x_lo, w_hi, IDESC_F16_M256_N256, 1u);
                                umma_f16_2sm(d, x_hi, w_lo, IDESC_F16_M256_N256, 1u);
                            }
                        }
                        tcgen05_commit_2sm(empty_a + slot);
                        PROF_ADD(3);
                        if (++slot == p.aslots) { slot = 0; ph ^= 1; }
                    }
                    tcgen05_commit_2sm(empty_b + sb);
                    if (++sb == BSTAGES) { sb = 0; bph ^= 1; }
                }
                tcgen05_commit_2sm(acc_full + ab);
                if (++ab == nbuf) { ab = 0; accph ^= 1; }
            }
        }
    } else if (warp >= WARP_EPI) {
        // ======== epilogue of this CTA's 128 rows: TMEM -> +shift, ReLU, mask -> per-warp smem staging -> TMA store ========
        // One warp per TMEM lane quarter; TMEM lane = time row, so a thread owns one output row.  Per 32-column block:
        // tcgen05.ld.32x32b.x32 -> registers -> 128-byte SWIZZLE_128B rows in a 4 KiB buffer PRIVATE to the warp -> one
        // TMA store (32 rows x 32 channels, full 128-byte lines; rows beyond T are clipped by the tensor map).  Two
        // buffers per warp and the next block's TMEM load issued before the current block is finished, so the only
        // synchronisation is inside the warp: no block-wide barrier sits between TMEM and the store.  (Measured:
        // block-wide staging with named barriers drained a 128 x 512 tile in 18 k cycles, register-direct 8-byte stores
        // in 17 k - the LSU pays per 128-byte line touched; the TMA store itself costs < 2 k.)  The accumulator is
        // single-buffered for 512 channels, so the drain time is what the MMA issuer waits for at every tile.
        const int q = warp & 3;
        const int row = q * 32 + lane;                         // tile row (time) owned by this thread
        unsigned char* stage = epi_stage + (size_t)q * p.epi_bufs * EPI_WARP_BYTES;
        const uint32_t nbufs = (uint32_t)p.epi_bufs;
        const uint32_t acc_empty_leader = mapa_u32(acc_empty, 0);
        const int nblk = p.nN * 8;                             // 32-column blocks per tile
        const uint32_t row_off = (uint32_t)lane * 128u, row_sw = (uint32_t)lane & 7u;
        int cur_l = -1;
        float2 wsc2 = make_float2(1.f, 1.f); const int* len_out = nullptr;
        int ab = 0; uint32_t accph = 0;
        uint32_t sbuf = 0;                                     // staging buffer of the next block (round robin)
        for (int item = item0; item < n_items; item += item_step) {
            int l, b, t0; bool dup;
            decode(item, l, b, t0, dup);
            const LayerDesc* L = p.layers + l;
            if (l != cur_l) {                                  // per-channel BN shift + scalars of the new layer
                named_bar_sync(3, NEPI_PAIR * 32);
                const float* shift = L->shift;
                for (int i = (warp - WARP_EPI) * 32 + lane; i < p.nN * 256; i += NEPI_PAIR * 32) ep_shift[i] = __ldg(shift + i);
                wsc2 = make_float2(L->wscale_inv, L->wscale_inv); len_out = L->len_out;
                named_bar_sync(3, NEPI_PAIR * 32);
                cur_l = l;
            }
            // rows t >= len are stored as zeros (every segment layer masks its tail)
            const bool live = (t0 + row) < len_out[b];
            const bool masked = __any_sync(0xffffffffu, !live);          // warp-uniform
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * acc_cols);
            PROF_BEGIN();
            mbar_wait(acc_full + ab, accph);
            PROF_ADD(0);
            tcgen05_fence_after();
            const int ab_cur = ab;
            if (++ab == nbuf) { ab = 0; accph ^= 1; }
            // finish one 32-column block held in registers: +shift, ReLU, (mask), stage, TMA store.  The epilogue is a single
            // warp per scheduler next to two depthwise warps - its drain time is its instruction count - so tiles that lie
            // entirely inside the utterance (all but the last one) run without the per-value row mask.
            auto drain = [&](auto masked_tag) {
                constexpr bool MASKED = decltype(masked_tag)::value;
                auto finish = [&](const uint32_t (&rg)[32], int col0) {
                    // the block's BN shift (shared-memory broadcast) is fetched BEFORE the warp synchronises on the staging
                    // buffer: behind the __syncwarp the loads could not be hoisted and every block exposed their latency
                    // (ncu: 23 % of the drain's stall samples sat on the first FFMA2 after these loads)
                    const float4* sh4 = reinterpret_cast<const float4*>(ep_shift + col0);
                    float4 shv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) shv[i] = sh4[i];
                    unsigned char* buf = stage + sbuf * EPI_WARP_BYTES;
                    // the store issued `nbufs` blocks ago has left this buffer
                    if (lane == 0) { if (nbufs == 2) bulk_wait_read<1>(); else if (nbufs == 3) bulk_wait_read<2>(); else bulk_wait_read<3>(); }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 sh = shv[i];
                        float2 v0 = __ffma2_rn(make_float2(__uint_as_float(rg[4 * i + 0]), __uint_as_float(rg[4 * i + 1])), wsc2, make_float2(sh.x, sh.y));
                        float2 v1 = __ffma2_rn(make_float2(__uint_as_float(rg[4 * i + 2]), __uint_as_float(rg[4 * i + 3])), wsc2, make_float2(sh.z, sh.w));
                        float4 v = make_float4(fmaxf(v0.x, 0.f), fmaxf(v0.y, 0.f), fmaxf(v1.x, 0.f), fmaxf(v1.y, 0.f));
                        if (MASKED && !live) v = make_float4(0.f, 0.f, 0.f, 0.f);
                        *reinterpret_cast<float4*>(buf + row_off + (((uint32_t)i ^ row_sw) << 4)) = v;   // SWIZZLE_128B
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && !dup && !DBG_ON(4)) { tma_store_3d(&L->tm_out, buf, col0, t0 + q * 32, b); bulk_commit(); }
                    if (++sbuf == nbufs) sbuf = 0;
                };
                uint32_t ra[32], rb[32];
                tmem_ld_32x32b_x32(tbase, ra);
                tmem_ld_wait();
#pragma unroll 1
                for (int blk = 0; blk < nblk; blk += 2) {          // nblk is even
                    tmem_ld_32x32b_x32(tbase + (uint32_t)((blk + 1) * 32), rb);
                    finish(ra, blk * 32);
                    tmem_ld_wait();
                    if (blk + 2 < nblk) tmem_ld_32x32b_x32(tbase + (uint32_t)((blk + 2) * 32), ra);
                    else {                                         // every TMEM read of this tile has completed
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote(acc_empty_leader + 8u * (uint32_t)ab_cur);
                    }
                    finish(rb, (blk + 1) * 32);
                    tmem_ld_wait();
                }
            };
            if (masked) drain(std::true_type{}); else drain(std::false_type{});
            if (lane == 0 && !dup) {
                // this warp's part of the tile is in global memory: publish it to the tiles of the next layer
                bulk_wait_all0();
                fence_proxy_async_all();
                __threadfence();
                atomicAdd(p.done + (size_t)l * p.done_stride + b, 1);
            }
            PROF_ADD(1);
        }
        if (lane == 0) bulk_wait_all0();
    }
    } else {
        // ======== depthwise producers, two groups on alternate chunks ========
        constexpr int R = TN / (2 * GW);
        constexpr int XP = KC / 2;
        const int grp = warp >> 2;
        const int wg = warp & (GW - 1);
        const int cp = lane & 15;
        const int tw = (wg * 2 + (lane >> 4)) * R;
        const uint32_t full_b_leader = mapa_u32(full_b, 0);
        int sx = 0, sb = 0; uint32_t xph = 0, bph = 0;
        float amax = 0.f;
        int gc = 0;                                         // running chunk index over all tiles (both groups count all)
        for (int item = item0; item < n_items; item += item_step) {
            int l, b, t0; bool dup;
            decode(item, l, b, t0, dup);
            const LayerDesc* L = p.layers + l;
            const int K = L->K, n_main = L->n_main, nch = L->n_main + L->n_res, x_w_off = L->x_w_off;
            const int len_mid = L->len_out[b];
            const bool tail_tile = t0 + TN > len_mid;      // only tiles that straddle the utterance's end need the row mask
            for (int c = 0; c < nch; ++c, ++gc) {
                if ((gc & 1) != grp) {                     // the other group's chunk: just keep the ring positions in step
                    if (++sx == XSTAGES) { sx = 0; xph ^= 1; }
                    if (++sb == BSTAGES) { sb = 0; bph ^= 1; }
                    continue;
                }
                PROF_BEGIN();
                mbar_wait(full_x + sx, xph);
                PROF_ADD(0);
                const float2* xs = reinterpret_cast<const float2*>(x_ring + (size_t)sx * p.x_stage_bytes) + cp;
                float2 acc[R];
                if (c < n_main && !DBG_ON(2)) {
                    const float2* wp = tap_base(x_ring + (size_t)sx * p.x_stage_bytes + x_w_off, cp);
                    SEG_K_SWITCH(K, (dw_chunk_s1<KK, 1, R>(xs, wp, tw, acc)));
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = xs[(size_t)(tw + r) * XP];
                }
                if (tail_tile) {
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (t0 + tw + r >= len_mid) acc[r] = make_float2(0.f, 0.f);
                }
                // range guard: |x| >= 65520 rounds to inf in fp16 (checked once per thread at the end of the kernel)
#pragma unroll
                for (int r = 0; r < R; ++r) amax = fmaxf(amax, fmaxf(fabsf(acc[r].x), fabsf(acc[r].y)));
                PROF_ADD(1);
                mbar_wait(empty_b + sb, bph ^ 1);                  // multicast commit of the leader's MMA thread
                PROF_ADD(2);
                unsigned char* bh0 = b_ring + (size_t)sb * B_STAGE;
                unsigned char* bq[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) bq[q] = bh0 + (tw >> 3) * 512 + (cp & 3) * 4 + ((((uint32_t)cp >> 2) ^ (uint32_t)q) << 4);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t off = (uint32_t)((r >> 3) * 512 + (r & 7) * 64);
                    unsigned char* bh = bq[(r >> 1) & 3];
                    const __half2 h = __floats2half2_rn(acc[r].x, acc[r].y);
                    *reinterpret_cast<__half2*>(bh + off) = h;
                    if (NPART == 2) {
                        const float2 hf = __half22float2(h);
                        *reinterpret_cast<__half2*>(bh + PART_BYTES + off) = __floats2half2_rn(acc[r].x - hf.x, acc[r].y - hf.y);
                    }
                }
                // generic-proxy stores -> async proxy (this SM's tensor core reads this stage), then the arrive on the
                // leader's barrier
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (leader) mbar_arrive(full_b + sb); else mbar_arrive_remote(full_b_leader + 8u * (uint32_t)sb);
                    mbar_arrive(empty_x + sx);
                }
                if (++sx == XSTAGES) { sx = 0; xph ^= 1; }
                if (++sb == BSTAGES) { sb = 0; bph ^= 1; }
                PROF_ADD(3);
            }
        }
        if (amax >= 65520.f) atomicOr(p.status, 1);          // a depthwise output left the fp16 range (see vasr_encoder_check)
    }
#ifdef VASR_DEV
    if (p.prof && lane == 0) {
        // leader CTA: dw group 0 (warp 0) 0..3, window producer 4..5, weight producer 12, MMA 6..9, epilogue 10..11;
        // peer CTA: dw group 1 (warp 4) -> 16..19, epilogue -> 20..21, window producer 22..23
        int base = -1;
        if (leader) {
            if (warp == 0) base = 0; else if (warp == WARP_X) base = 4; else if (warp == WARP_A) base = 12;
            else if (warp == WARP_MMA) base = 6; else if (warp == WARP_EPI) base = 10;
        } else {
            if (warp == 4) base = 16; else if (warp == WARP_EPI) base = 20; else if (warp == WARP_X) base = 22;
        }
        if (base >= 0)
            for (int i = 0; i < 4; ++i)
                if (pacc[i]) atomicAdd(p.prof + base + i, pacc[i]);
        if (warp == 1 && leader) atomicAdd(p.prof + 15, 1ull);
    }
#endif
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();            // no CTA frees TMEM or exits while its peer may still arrive on its barriers / read its smem
    if (warp == WARP_MMA) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    }
}

// ------------------------------------------------------------------------------------------ host side
static int encode_tm(CUtensorMap* tm, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
                     const cuuint64_t* strides_bytes, const cuuint32_t* box, CUtensorMapSwizzle sw)
{
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(tm, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(VASR_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return VASR_OK;
}

// activations [B, T, C] fp32 -> 3-D map (C, T, B), box (32, rows, 1), no swizzle, zero OOB fill
static int encode_act(CUtensorMap* tm, const float* base, int B, int T, int C, long long bstride, int box_rows)
{
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)C * 4, (cuuint64_t)bstride * 4};
    cuuint32_t box[3] = {(cuuint32_t)KC, (cuuint32_t)box_rows, 1};
    return encode_tm(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}
// output [B, T, C] fp32 -> 3-D map (C, T, B), box (32, rows, 1), SWIZZLE_128B (rows beyond T are clipped by the TMA)
static int encode_out(CUtensorMap* tm, const float* base, int B, int T, int C, long long bstride, int box_rows = TN)
{
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)C * 4, (cuuint64_t)bstride * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
    return encode_tm(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
}
// weights [Cout, Cin] fp16 -> 2-D map (Cin, Cout), box (32, 256), SWIZZLE_64B
static int encode_w(CUtensorMap* tm, const __half* base, int Cout, int Cin, int box_rows = 256)
{
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)Cout};
    cuuint64_t str[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)box_rows};
    return encode_tm(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
}

struct KernelEntry { int K, S, D; const void* fn[2]; };   // [0] f16x3 (hi + lo parts), [1] f16x1
#define TC_ENTRY(k, s, d) {k, s, d, {(const void*)subblock_kernel<k, s, d, 2>, (const void*)subblock_kernel<k, s, d, 1>}}
static const KernelEntry g_kernels[] = {
    TC_ENTRY(1, 1, 1), TC_ENTRY(33, 2, 1), TC_ENTRY(33, 1, 1), TC_ENTRY(39, 1, 1), TC_ENTRY(51, 1, 1),
    TC_ENTRY(63, 1, 1), TC_ENTRY(75, 1, 1), TC_ENTRY(87, 1, 2), TC_ENTRY(11, 1, 1), TC_ENTRY(11, 2, 1), TC_ENTRY(15, 1, 2),
};
static const KernelEntry* find_kernel(int K, int S, int D)
{
    for (const KernelEntry& e : g_kernels)
        if (e.K == K && e.S == S && e.D == D) return &e;
    return nullptr;
}
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int SMEM_FIXED = 1024 /*barriers*/ + MAX_CO_CTA * 4 /*BN shift*/;

// Developer switches (libvasr_b200_dev.so only; the product library has no environment switches on this path):
//   VASR_TC_PROF=1 role cycle counters, VASR_TC_DBG=bits ablations, VASR_TC_PAIR=0 single-CTA 128-row kernel instead
//   of the pair kernel, VASR_TC_LAT=0 no 32-row latency tiles, VASR_TC_RINGS=x,b,a ring depths of the pair kernel.
#ifdef VASR_DEV
static int dev_env(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
#else
static constexpr int dev_env(const char*, int dflt) { return dflt; }
#endif

static void x_geometry(int K, int S, int D, int* n_xbox, int* xbox_rows, int* w_off, int* stage_bytes, int tn = TN)
{
    const int rows = (tn - 1) * S + (K - 1) * D + 1;
    int nb = (rows + 255) / 256, br = (rows + nb - 1) / nb;
    br = (br + 7) / 8 * 8;
    int bytes = nb * br * KC * 4;
    if (bytes < tn * KC * 4) bytes = tn * KC * 4;            // residual / identity chunks load one tile of rows
    *n_xbox = nb; *xbox_rows = br; *w_off = bytes;
    bytes += tap_floats(K) * 4;                               // + the chunk's depthwise taps
    *stage_bytes = (bytes + 1023) / 1024 * 1024;
}

// ring depths of the single-CTA kernels under the 227 KiB budget: 2 activation-window stages (3 when there is room), 2
// activation-operand stages (3 when there is room), and as many 256-channel weight slots as fit (at most four chunks).
static void pick_rings(int npart, int x_stage_bytes, int nN, int epi_bytes, int* xstages, int* bstages, int* aslots)
{
    const int w_slot = W_PART * npart, b_stage = PART_BYTES * npart;
    const int overhead = SMEM_FIXED + epi_bytes;
    int xs = 2, bs = 2;
    int slots = (SMEM_LIMIT - overhead - bs * b_stage - xs * x_stage_bytes) / w_slot;
    if (slots > 4 * nN) slots = 4 * nN;
    if (slots > 16) slots = 16;
    int left = SMEM_LIMIT - overhead - bs * b_stage - xs * x_stage_bytes - slots * w_slot;
    if (left >= x_stage_bytes) { xs = 3; left -= x_stage_bytes; }
    if (left >= b_stage) { bs = 3; left -= b_stage; }
    *xstages = xs; *bstages = bs; *aslots = slots;
}
// ring depths of the pair kernel: each depthwise group holds a window stage while it computes, so a third stage is what
// lets the loads run ahead (with two, every chunk exposes its TMA latency); three operand stages let the depthwise
// warps work through the accumulator drain at the end of a tile; the rest goes to 16 KiB (half-block) weight slots.
static bool pick_rings_pair(int npart, int x_stage_bytes, int nN, int* xstages, int* bstages, int* aslots, int* epi_bufs)
{
    const int w_slot = W_HALF * npart, b_stage = PART_BYTES * npart, e_buf = NEPI_PAIR * EPI_WARP_BYTES;
    int xs = 3, bs = 3, eb = 2;
    int slots = (SMEM_LIMIT - SMEM_FIXED - eb * e_buf - bs * b_stage - xs * x_stage_bytes) / w_slot;
    if (slots < nN) { bs = 2; slots = (SMEM_LIMIT - SMEM_FIXED - eb * e_buf - bs * b_stage - xs * x_stage_bytes) / w_slot; }
    if (slots < nN) return false;
    if (slots > nN + 1) slots = nN + 1;                     // one chunk of weights + one slot of prefetch; then staging
    int left = SMEM_LIMIT - SMEM_FIXED - eb * e_buf - bs * b_stage - xs * x_stage_bytes - slots * w_slot;
    while (eb < 4 && left >= e_buf) { ++eb; left -= e_buf; }
    while (slots < 4 * nN && slots < 16 && left >= w_slot) { ++slots; left -= w_slot; }
    *xstages = xs; *bstages = bs; *aslots = slots; *epi_bufs = eb;
    return true;
}

}  // namespace tc

static std::atomic<unsigned long long> g_layer_uid{1};

int tc_init()
{
    using namespace tc;
    if (g_encode) return VASR_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    VASR_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
        return set_error(VASR_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    int dev = 0;
    VASR_CUDA_OK(cudaGetDevice(&dev));
    VASR_CUDA_OK(cudaDeviceGetAttribute(&tc::g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    for (const KernelEntry& e : g_kernels)
        for (int i = 0; i < 2; ++i)
            VASR_CUDA_OK(cudaFuncSetAttribute(e.fn[i], cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    const void* seg_fns[] = {(const void*)segment_kernel<2, 128>, (const void*)segment_kernel<1, 128>,
                             (const void*)segment_kernel<2, 32>, (const void*)segment_kernel<1, 32>,
                             (const void*)segment_pair_kernel<2>, (const void*)segment_pair_kernel<1>};
    for (const void* f : seg_fns) VASR_CUDA_OK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    g_encode = (EncodeTiledFn)fn;
    return VASR_OK;
}

bool subblock_tc_supported(const SubBlock& sb)
{
    using namespace tc;
    const int K = sb.separable ? sb.kernel : 1;
    if (!find_kernel(K, sb.stride, sb.dilation)) return false;
    if (sb.cin % KC != 0 || (sb.has_res && sb.res_cin % KC != 0)) return false;
    if (sb.cout % 256 != 0) return false;                      // N blocks of 256 output channels
    if (sb.cout > MAX_CO_CTA && sb.cout % MAX_CO_CTA != 0) return false;
    if (sb.has_res && sb.stride != 1) return false;
    return true;
}

// weights: one power-of-two pre-scale per layer, fp16 hi/lo split, TMA maps
int tc_prepare_layer(SubBlock& sb, const float* w_main, const float* w_res, const float* dw_kc,
                     std::vector<void*>& allocs)
{
    using namespace tc;
    const int Co = sb.cout, Ci = sb.cin, Cr = sb.has_res ? sb.res_cin : 0;
    sb.uid = g_layer_uid.fetch_add(1);
    std::vector<__half> mh((size_t)Co * Ci), ml((size_t)Co * Ci), rh((size_t)Co * Cr), rl((size_t)Co * Cr);
    // one power-of-two pre-scale per layer: the largest |w| lands in [2^13, 2^14), so hi and lo = w - hi stay in
    // fp16's normal range for every weight within 2^17 of the maximum (smaller ones lose lo bits that are below
    // fp32's own resolution of the dot product anyway)
    float mx = 0.f;
    for (size_t i = 0; i < (size_t)Co * Ci; ++i) mx = fmaxf(mx, fabsf(w_main[i]));
    for (size_t i = 0; i < (size_t)Co * Cr; ++i) mx = fmaxf(mx, fabsf(w_res[i]));
    int ex = 0;
    if (mx > 0.f && isfinite(mx)) frexpf(mx, &ex);          // mx = f * 2^ex, f in [0.5, 1)
    const int sh_bits = (mx > 0.f) ? 14 - ex : 0;
    const float sc = ldexpf(1.f, sh_bits);
    sb.wscale_inv_scalar = ldexpf(1.f, -sh_bits);
    auto split = [&](float v, __half& h, __half& l) {
        const float x = v * sc;
        h = __float2half_rn(x);
        l = __float2half_rn(x - __half2float(h));
    };
    for (size_t i = 0; i < (size_t)Co * Ci; ++i) split(w_main[i], mh[i], ml[i]);
    for (size_t i = 0; i < (size_t)Co * Cr; ++i) split(w_res[i], rh[i], rl[i]);
    auto up = [&](const void* src, size_t bytes, void** dst) -> int {
        void* d = nullptr;
        VASR_CUDA_OK(cudaMalloc(&d, bytes ? bytes : 16));
        allocs.push_back(d);
        if (bytes) VASR_CUDA_OK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
        *dst = d;
        return VASR_OK;
    };
    int rc;
    if ((rc = up(mh.data(), sizeof(__half) * mh.size(), &sb.pw_h))) return rc;
    if ((rc = up(ml.data(), sizeof(__half) * ml.size(), &sb.pw_l))) return rc;
    static_assert(sizeof(CUtensorMap) == sizeof(sb.tm_w_hi), "tensor map storage");
    if ((rc = encode_w((CUtensorMap*)sb.tm_w_hi, (const __half*)sb.pw_h, Co, Ci))) return rc;
    if ((rc = encode_w((CUtensorMap*)sb.tm_w_lo, (const __half*)sb.pw_l, Co, Ci))) return rc;
    if (Cr) {
        if ((rc = up(rh.data(), sizeof(__half) * rh.size(), &sb.res_h))) return rc;
        if ((rc = up(rl.data(), sizeof(__half) * rl.size(), &sb.res_l))) return rc;
        if ((rc = encode_w((CUtensorMap*)sb.tm_r_hi, (const __half*)sb.res_h, Co, Cr))) return rc;
        if ((rc = encode_w((CUtensorMap*)sb.tm_r_lo, (const __half*)sb.res_l, Co, Cr))) return rc;
    } else {
        memcpy(sb.tm_r_hi, sb.tm_w_hi, sizeof(sb.tm_w_hi));
        memcpy(sb.tm_r_lo, sb.tm_w_lo, sizeof(sb.tm_w_lo));
    }
    {   // depthwise taps re-packed per 32-channel chunk (layout: tap_floats / tap_load); identity for a plain 1x1 conv
        const int K = sb.separable ? sb.kernel : 1;
        const int per_chunk = tap_floats(K);
        std::vector<float> pk((size_t)(Ci / KC) * per_chunk, 0.0f);
        for (int c = 0; c < Ci; ++c)
            for (int k = 0; k < K; ++k) {
                const float v = sb.separable ? dw_kc[(size_t)k * Ci + c] : 1.0f;
                const int cc = c % KC;
#if VASR_DW_TAP128
                const size_t idx = (size_t)(c / KC) * per_chunk + (size_t)(k >> 1) * 2 * KC + (cc >> 1) * 4 + (k & 1) * 2 + (cc & 1);
#else
                const size_t idx = (size_t)(c / KC) * per_chunk + (size_t)k * KC + cc;
#endif
                pk[idx] = v;
            }
        if ((rc = up(pk.data(), sizeof(float) * pk.size(), (void**)&sb.dw_tc))) return rc;
    }
    return VASR_OK;
}

#ifdef VASR_DEV
// role cycle counters of the last launch (developer builds): [32] device words, printed after a stream sync
static unsigned long long* dev_prof_buffer(cudaStream_t st)
{
    static int on = -1;
    static unsigned long long* d = nullptr;
    if (on < 0) {
        on = tc::dev_env("VASR_TC_PROF", 0) > 0 ? 1 : 0;
        if (on && cudaMalloc(&d, 32 * sizeof(unsigned long long)) != cudaSuccess) on = 0;
    }
    if (!on) return nullptr;
    cudaMemsetAsync(d, 0, 32 * sizeof(unsigned long long), st);
    return d;
}
static void dev_prof_print(const char* head, const unsigned long long* d, cudaStream_t st, bool pair)
{
    unsigned long long h[32];
    if (cudaStreamSynchronize(st) != cudaSuccess || cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return;
    const double n = (double)(h[15] ? h[15] : 1);
    fprintf(stderr, "%s | dw: wait_x %.0f comp %.0f wait_b %.0f store %.0f | xprod wait %.0f dep %.0f | aprod wait %.0f | mma: wait_b %.0f wait_a %.0f wait_acc %.0f issue %.0f | epi: wait %.0f drain %.0f (kcycles per CTA)\n",
            head, h[0] / n / 1e3, h[1] / n / 1e3, h[2] / n / 1e3, h[3] / n / 1e3, h[4] / n / 1e3, h[5] / n / 1e3, h[12] / n / 1e3,
            h[6] / n / 1e3, h[7] / n / 1e3, h[8] / n / 1e3, h[9] / n / 1e3, h[10] / n / 1e3, h[11] / n / 1e3);
    if (pair)
        fprintf(stderr, "   peer CTA | dw group 1: wait_x %.0f comp %.0f wait_b %.0f store %.0f | epi: wait %.0f drain %.0f | xprod wait %.0f dep %.0f\n",
                h[16] / n / 1e3, h[17] / n / 1e3, h[18] / n / 1e3, h[19] / n / 1e3, h[20] / n / 1e3, h[21] / n / 1e3, h[22] / n / 1e3, h[23] / n / 1e3);
}
#endif

int launch_subblock_tc(SubBlock& sb, const float* x, long long x_bstride, const float* res_in, long long r_bstride,
                       float* y, long long y_bstride, int B, int T_in,
                       int T_out, const int* len_in, const int* len_out, int split3, int b0, int nb,
                       int* tile_counter, int* status, int grid_limit, cudaStream_t st)
{
    using namespace tc;
    (void)len_in;
    const int K = sb.separable ? sb.kernel : 1;
    const KernelEntry* ke = find_kernel(K, sb.stride, sb.dilation);
    if (!ke || !subblock_tc_supported(sb))
        return set_error(VASR_EINVAL, "tcgen05 path: sub-block (cin=%d cout=%d k=%d s=%d d=%d) is not a built shape",
                         sb.cin, sb.cout, sb.kernel, sb.stride, sb.dilation);
    const int npart = split3 ? 2 : 1;
    Params p{};
    p.dw_w = sb.dw_tc; p.shift = sb.shift; p.wscale_inv = sb.wscale_inv_scalar; p.out = y; p.out_bstride = y_bstride; p.len_out = len_out;
    p.Cin = sb.cin; p.Cres = sb.has_res ? sb.res_cin : 0; p.Cout = sb.cout; p.T_out = T_out;
    p.pad = sb.separable ? sb.pad : 0;
    p.n_main = sb.cin / KC; p.n_res = sb.has_res ? sb.res_cin / KC : 0;
    const int co_cta = sb.cout > MAX_CO_CTA ? MAX_CO_CTA : sb.cout;
    p.nN = co_cta / 256;
    x_geometry(K, sb.stride, sb.dilation, &p.n_xbox, &p.xbox_rows, &p.x_w_off, &p.x_stage_bytes);
    p.relu = sb.relu ? 1 : 0; p.mask_tail = sb.final_layer ? 0 : 1;
    // TMA-store epilogue when its 32 KiB of staging still leaves one chunk of weight slots; direct stores otherwise
    p.tma_epi = 1;
    pick_rings(npart, p.x_stage_bytes, p.nN, 2 * EPI_STAGE_BYTES, &p.xstages, &p.bstages, &p.aslots);
    if (p.aslots < p.nN) {
        p.tma_epi = 0;
        pick_rings(npart, p.x_stage_bytes, p.nN, 0, &p.xstages, &p.bstages, &p.aslots);
    }
    if (p.aslots < p.nN) return set_error(VASR_EINVAL, "tcgen05 path: shared memory budget exceeded (k=%d)", K);
    const size_t smem = (size_t)p.aslots * W_PART * npart + (size_t)p.bstages * PART_BYTES * npart +
                        (size_t)p.xstages * p.x_stage_bytes + SMEM_FIXED + (p.tma_epi ? 2 * EPI_STAGE_BYTES : 0);
    p.b0 = b0;
    // activation tensor maps cover the whole batch and are cached per layer (pointers/shapes rarely change)
    int rc;
    if (sb.tmc_x != x || sb.tmc_B != B || sb.tmc_T != T_in || sb.tmc_xs != x_bstride) {
        if ((rc = encode_act((CUtensorMap*)sb.tm_x, x, B, T_in, sb.cin, x_bstride, p.xbox_rows))) return rc;
        sb.tmc_x = x; sb.tmc_B = B; sb.tmc_T = T_in; sb.tmc_xs = x_bstride;
        sb.tmc_r = nullptr;
    }
    if (sb.has_res) {
        if (sb.tmc_r != res_in || sb.tmc_rs != r_bstride) {
            if ((rc = encode_act((CUtensorMap*)sb.tm_r, res_in, B, T_in, sb.res_cin, r_bstride, TN))) return rc;
            sb.tmc_r = res_in; sb.tmc_rs = r_bstride;
        }
    }
    p.tile_counter = tile_counter; p.status = status;
    p.n_tt = ceil_div(T_out, TN); p.n_utt = nb; p.n_cg = sb.cout / co_cta;
    const int n_tiles = p.n_tt * p.n_utt * p.n_cg;
    if (sb.tmc_y != y || sb.tmc_yB != B || sb.tmc_yT != T_out || sb.tmc_ys != y_bstride) {
        if ((rc = encode_out((CUtensorMap*)sb.tm_y, y, B, T_out, sb.cout, y_bstride))) return rc;
        sb.tmc_y = y; sb.tmc_yB = B; sb.tmc_yT = T_out; sb.tmc_ys = y_bstride;
    }
    int max_ctas = g_num_sms;                                        // persistent: at most one CTA per SM
    if (grid_limit > 0 && grid_limit < max_ctas) max_ctas = grid_limit;   // concurrent sub-batch kernels share the SMs
    dim3 grid(n_tiles < max_ctas ? n_tiles : max_ctas, 1, 1);
    void* args[] = {(void*)sb.tm_x, (void*)(sb.has_res ? sb.tm_r : sb.tm_x), (void*)sb.tm_w_hi, (void*)sb.tm_w_lo,
                    (void*)sb.tm_r_hi, (void*)sb.tm_r_lo, (void*)sb.tm_y, (void*)&p};
#ifdef VASR_DEV
    p.prof = dev_prof_buffer(st);
    p.dbg = dev_env("VASR_TC_DBG", 0);
#endif
    VASR_CUDA_OK(cudaLaunchKernel(ke->fn[split3 ? 0 : 1], grid, dim3(NTHREADS), args, smem, st));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
#ifdef VASR_DEV
    if (p.prof) {
        char head[160];
        snprintf(head, sizeof(head), "TCPROF k=%d cin=%d cout=%d res=%d tiles=%d ctas=%d xs=%d bs=%d as=%d", K, sb.cin, sb.cout,
                 (int)sb.has_res, n_tiles, (int)grid.x, p.xstages, p.bstages, p.aslots);
        dev_prof_print(head, p.prof, st, false);
    }
#endif
    return VASR_OK;
}

// ------------------------------------------------------------------------------------------ segment launch

// Device descriptor tables (tensor maps + per-layer scalars) are cached.  The key holds EVERYTHING a table is built
// from: each layer's identity, its input / residual / output pointers with their batch strides (the strides depend on
// T_f, the segment's T does not: T_f = 2k and 2k - 1 give the same T) and its length table, plus batch, frames and the
// kernel variant.
struct SegLayerKey {
    unsigned long long uid; const void* x; const void* res; const void* y; const void* len_out; long long xs, rs, ys;
    bool operator==(const SegLayerKey& o) const
    {
        return uid == o.uid && x == o.x && res == o.res && y == o.y && len_out == o.len_out && xs == o.xs && rs == o.rs && ys == o.ys;
    }
};
struct SegCacheEntry {
    std::vector<SegLayerKey> key; int B, T, variant;
    tc::LayerDesc* d_desc;
};
static std::vector<SegCacheEntry> g_seg_cache;

bool segment_tc_layer_ok(const SubBlock& sb)
{
    using namespace tc;
    if (!sb.separable || sb.stride != 1 || sb.dilation != 1 || sb.final_layer) return false;
    if (!seg_kernel_size(sb.kernel)) return false;
    if (!subblock_tc_supported(sb)) return false;
    if (sb.cout > MAX_CO_CTA) return false;
    return true;
}

// how a run of layers is executed for a given sub-batch shape
struct SegPlan {
    int tr;                  // valid rows per tile: 128, or 32 (latency mode)
    bool pair;               // CTA-pair kernel (cta_group::2)
    int x_stage_bytes, xstages, bstages, aslots, epi_bufs;
    size_t smem;
};
static int seg_x_stage_bytes(const SegLayer* L, int n, int tr)
{
    int mx = 0;
    for (int i = 0; i < n; ++i) {
        int nb, br, wo, sbytes;
        tc::x_geometry(L[i].sb->kernel, 1, 1, &nb, &br, &wo, &sbytes, tr);
        if (sbytes > mx) mx = sbytes;
    }
    return mx;
}
static bool plan_segment(const SegLayer* L, int n, int npart, int T, int nb, SegPlan* pl)
{
    using namespace tc;
    const int nN = L[0].sb->cout / 256;
    // latency mode: when the 128-row tiles of a layer would occupy less than half of the SMs, use 32-row tiles
    const bool lat = dev_env("VASR_TC_LAT", 1) != 0 && T > 0 && ceil_div(T, TN) * nb * 2 <= g_num_sms;
    pl->tr = lat ? 32 : TN;
    pl->pair = false;
    pl->x_stage_bytes = seg_x_stage_bytes(L, n, pl->tr);
    if (!lat && dev_env("VASR_TC_PAIR", 1) != 0 && g_num_sms >= 2 &&
        pick_rings_pair(npart, pl->x_stage_bytes, nN, &pl->xstages, &pl->bstages, &pl->aslots, &pl->epi_bufs)) {
        pl->pair = true;
        auto pair_smem = [&](int xs, int bs, int as, int eb) {
            return (size_t)as * W_HALF * npart + (size_t)bs * PART_BYTES * npart + (size_t)xs * pl->x_stage_bytes + SMEM_FIXED +
                   (size_t)eb * NEPI_PAIR * EPI_WARP_BYTES;
        };
#ifdef VASR_DEV
        if (const char* e = getenv("VASR_TC_RINGS")) {         // x,b,a,e: window / operand stages, weight slots, staging buffers
            int xs = 0, bs = 0, as = 0, eb = 0;
            if (sscanf(e, "%d,%d,%d,%d", &xs, &bs, &as, &eb) == 4 && xs >= 2 && xs <= MAX_STAGES && bs >= 2 && bs <= MAX_STAGES && as >= nN &&
                as <= 16 && eb >= 2 && eb <= 4 && pair_smem(xs, bs, as, eb) <= (size_t)SMEM_LIMIT) {
                pl->xstages = xs; pl->bstages = bs; pl->aslots = as; pl->epi_bufs = eb;
            } else fprintf(stderr, "VASR_TC_RINGS=%s does not fit this segment (x stage %d bytes): ignored\n", e, pl->x_stage_bytes);
        }
#endif
        pl->smem = pair_smem(pl->xstages, pl->bstages, pl->aslots, pl->epi_bufs);
        return true;
    }
    pick_rings(npart, pl->x_stage_bytes, nN, 2 * EPI_STAGE_BYTES, &pl->xstages, &pl->bstages, &pl->aslots);
    pl->smem = (size_t)pl->aslots * W_PART * npart + (size_t)pl->bstages * PART_BYTES * npart +
               (size_t)pl->xstages * pl->x_stage_bytes + SMEM_FIXED + 2 * EPI_STAGE_BYTES;
    return pl->aslots >= nN;
}

bool segment_tc_ok(const SegLayer* L, int n, int split3)
{
    using namespace tc;
    if (n < 2) return false;
    for (int i = 0; i < n; ++i) {
        if (!segment_tc_layer_ok(*L[i].sb)) return false;
        if (L[i].sb->cout != L[0].sb->cout) return false;
    }
    // every variant must be able to run the segment: the single-CTA 128-row rings are the tightest
    int xs, bs, as;
    pick_rings(split3 ? 2 : 1, seg_x_stage_bytes(L, n, TN), L[0].sb->cout / 256, 2 * EPI_STAGE_BYTES, &xs, &bs, &as);
    return as >= L[0].sb->cout / 256;
}

int launch_segment_tc(const SegLayer* L, int n, int B, int T, int split3, int b0, int nb,
                      int* tile_counter, int* status, int* done, int done_stride, int grid_limit, cudaStream_t st)
{
    using namespace tc;
    if (!segment_tc_ok(L, n, split3)) return set_error(VASR_EINVAL, "tcgen05 path: layers do not form a segment");
    const int npart = split3 ? 2 : 1;
    SegPlan pl{};
    if (!plan_segment(L, n, npart, T, nb, &pl)) return set_error(VASR_EINVAL, "tcgen05 path: shared memory budget exceeded for the segment");
    const int tr = pl.tr;
    const bool pair = pl.pair;
    SegParams p{};
    p.x_stage_bytes = pl.x_stage_bytes; p.xstages = pl.xstages; p.bstages = pl.bstages; p.aslots = pl.aslots; p.epi_bufs = pl.epi_bufs;
    p.nN = L[0].sb->cout / 256;
    const int variant = pair ? 256 : tr;                      // key of the descriptor cache (weight / output boxes differ)
    std::vector<SegLayerKey> key((size_t)n);
    for (int i = 0; i < n; ++i)
        key[i] = SegLayerKey{L[i].sb->uid, L[i].x, L[i].sb->has_res ? L[i].res : nullptr, L[i].y, L[i].len_out, L[i].xs,
                             L[i].sb->has_res ? L[i].rs : 0, L[i].ys};
    LayerDesc* d_desc = nullptr;
    for (const SegCacheEntry& e : g_seg_cache)
        if (e.B == B && e.T == T && e.variant == variant && e.key == key) { d_desc = e.d_desc; break; }
    if (!d_desc) {
        std::vector<LayerDesc> h((size_t)n);
        int rc;
        for (int i = 0; i < n; ++i) {
            const SubBlock& sb = *L[i].sb;
            LayerDesc& d = h[i];
            memset(&d, 0, sizeof(d));
            int stage_bytes;
            x_geometry(sb.kernel, 1, 1, &d.n_xbox, &d.xbox_rows, &d.x_w_off, &stage_bytes, tr);
            if ((rc = encode_act(&d.tm_x, L[i].x, B, T, sb.cin, L[i].xs, d.xbox_rows))) return rc;
            if (sb.has_res) { if ((rc = encode_act(&d.tm_r, L[i].res, B, T, sb.res_cin, L[i].rs, tr))) return rc; }
            else d.tm_r = d.tm_x;
            if (pair) {                                       // each CTA of a pair loads 128 of the 256 rows of a weight block
                if ((rc = encode_w(&d.tm_w_hi, (const __half*)sb.pw_h, sb.cout, sb.cin, 128))) return rc;
                if ((rc = encode_w(&d.tm_w_lo, (const __half*)sb.pw_l, sb.cout, sb.cin, 128))) return rc;
                if (sb.has_res) {
                    if ((rc = encode_w(&d.tm_r_hi, (const __half*)sb.res_h, sb.cout, sb.res_cin, 128))) return rc;
                    if ((rc = encode_w(&d.tm_r_lo, (const __half*)sb.res_l, sb.cout, sb.res_cin, 128))) return rc;
                } else { d.tm_r_hi = d.tm_w_hi; d.tm_r_lo = d.tm_w_lo; }
                if ((rc = encode_out(&d.tm_out, L[i].y, B, T, sb.cout, L[i].ys, 32))) return rc;   // per-warp 32-row stores
            } else {
                memcpy(&d.tm_w_hi, sb.tm_w_hi, sizeof(CUtensorMap)); memcpy(&d.tm_w_lo, sb.tm_w_lo, sizeof(CUtensorMap));
                memcpy(&d.tm_r_hi, sb.tm_r_hi, sizeof(CUtensorMap)); memcpy(&d.tm_r_lo, sb.tm_r_lo, sizeof(CUtensorMap));
                if ((rc = encode_out(&d.tm_out, L[i].y, B, T, sb.cout, L[i].ys, tr))) return rc;
            }
            d.dw_w = sb.dw_tc; d.shift = sb.shift; d.len_out = L[i].len_out; d.wscale_inv = sb.wscale_inv_scalar;
            d.out = L[i].y; d.out_bstride = L[i].ys;
            d.K = sb.kernel; d.n_main = sb.cin / KC; d.n_res = sb.has_res ? sb.res_cin / KC : 0;
            d.relu = sb.relu ? 1 : 0; d.mask_tail = 1; d.pad = sb.pad;
            if (pair && !sb.relu) return set_error(VASR_EINVAL, "tcgen05 path: the pair kernel's epilogue applies ReLU unconditionally");
        }
        if (g_seg_cache.size() >= 64) {                      // bounded: drop everything once nothing can still be reading it
            VASR_CUDA_OK(cudaDeviceSynchronize());
            for (SegCacheEntry& e : g_seg_cache) cudaFree(e.d_desc);
            g_seg_cache.clear();
        }
        VASR_CUDA_OK(cudaMalloc((void**)&d_desc, sizeof(LayerDesc) * (size_t)n));
        VASR_CUDA_OK(cudaMemcpy(d_desc, h.data(), sizeof(LayerDesc) * (size_t)n, cudaMemcpyHostToDevice));
        g_seg_cache.push_back(SegCacheEntry{key, B, T, variant, d_desc});
    }
    p.layers = d_desc; p.n_layers = n;
    p.tile_counter = tile_counter; p.status = status; p.done = done; p.done_stride = done_stride;
    p.T_out = T; p.b0 = b0; p.n_tt = ceil_div(T, tr); p.n_utt = nb;
    const long long tpl = (long long)p.n_tt * p.n_utt;
    const long long n_items = (pair ? (tpl + 1) / 2 : tpl) * n;         // work items: tiles, or pairs of tiles
    int max_ctas = g_num_sms;
    if (grid_limit > 0 && grid_limit < max_ctas) max_ctas = grid_limit;
    unsigned grid_x;
    if (pair) {
        long long clusters = max_ctas / 2;
        if (clusters < 1) clusters = 1;
        if (n_items < clusters) clusters = n_items;
        grid_x = (unsigned)(2 * clusters);
    } else {
        grid_x = (unsigned)(n_items < max_ctas ? n_items : max_ctas);
    }
    dim3 grid(grid_x, 1, 1);
#ifdef VASR_DEV
    p.prof = dev_prof_buffer(st);
    p.dbg = dev_env("VASR_TC_DBG", 0);
#endif
    void* args[] = {(void*)&p};
    const void* fn = pair     ? (split3 ? (const void*)segment_pair_kernel<2> : (const void*)segment_pair_kernel<1>)
                     : tr != TN ? (split3 ? (const void*)segment_kernel<2, 32> : (const void*)segment_kernel<1, 32>)
                              : (split3 ? (const void*)segment_kernel<2, 128> : (const void*)segment_kernel<1, 128>);
    VASR_CUDA_OK(cudaLaunchKernel(fn, grid, dim3(pair ? PAIR_THREADS : NTHREADS), args, pl.smem, st));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
#ifdef VASR_DEV
    if (p.prof) {
        char head[200];
        snprintf(head, sizeof(head), "TCSEG%s layers=%d k=%d..%d cout=%d items=%lld ctas=%d tr=%d xs=%d bs=%d as=%d eb=%d", pair ? "(pair)" : "", n,
                 L[0].sb->kernel, L[n - 1].sb->kernel, L[0].sb->cout, n_items, (int)grid.x, tr, p.xstages, p.bstages, p.aslots, p.epi_bufs);
        dev_prof_print(head, p.prof, st, pair);
    }
#endif
    return VASR_OK;
}

}  // namespace vasr
