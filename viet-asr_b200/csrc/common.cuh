// Shared helpers for the vasr_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string>
#include <atomic>

#include "../../include/vasr_b200.h"

namespace vasr {

// thread-local error text returned by vasr_last_error()
std::string& last_error_ref();
int set_error(int code, const char* fmt, ...);
extern std::atomic<long long> g_launch_count;

#define VASR_CUDA_OK(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess)                                                          \
            return vasr::set_error(VASR_ECUDA, "%s failed: %s (%s:%d)", #expr,          \
                                   cudaGetErrorString(_e), __FILE__, __LINE__);         \
    } while (0)

#define VASR_LAUNCH_OK(name)                                                            \
    do {                                                                                \
        vasr::g_launch_count.fetch_add(1, std::memory_order_relaxed);                   \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess)                                                          \
            return vasr::set_error(VASR_ECUDA, "launch of %s failed: %s (%s:%d)", name, \
                                   cudaGetErrorString(_e), __FILE__, __LINE__);         \
    } while (0)

#define VASR_REQUIRE(cond, ...)                                                         \
    do {                                                                                \
        if (!(cond)) return vasr::set_error(VASR_EINVAL, __VA_ARGS__);                  \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Developer switches exist in libvasr_b200_dev.so only (-DVASR_DEV, `make dev`); the product library never reads the
// environment on the compute path.
#ifdef VASR_DEV
static inline int dev_env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
#else
static inline int dev_env_int(const char*, int dflt) { return dflt; }
#endif
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------
// per-layer description shared by the SIMT and the tcgen05 encoder paths
// ---------------------------------------------------------------------------
struct SubBlock {
    int cin, cout;
    int kernel, stride, dilation, pad;
    bool separable;   // dw + pw ; otherwise plain 1x1 conv (kernel == 1)
    bool relu;        // ReLU directly after BN (every sub-block; for the last sub-block of a
                      // residual block the ReLU comes after the residual add - same epilogue)
    bool has_res;     // last sub-block of a residual block: + W_r * block_input
    int res_cin;
    bool final_layer; // last layer of the encoder: tail frames are NOT zeroed
    int len_stage_in; // index into the device length table (input resolution)
    int len_stage_out;
    // device weights (fp32), BN folded:
    float* dw_w = nullptr;     // [kernel][cin]   (transposed for channels-last)
    float* pw_w = nullptr;     // [cout][cin]     scale-folded, K-major
    float* res_w = nullptr;    // [cout][res_cin] scale-folded
    float* shift = nullptr;    // [cout]  (BN shift, + residual BN shift)
    // tcgen05 operands: fp16 hi / lo split of the (power-of-two pre-scaled) weights, [cout][cin],
    // the inverse pre-scale per output channel, and the TMA descriptors (CUtensorMap, 128 B each)
    void* pw_h = nullptr; void* pw_l = nullptr;
    void* res_h = nullptr; void* res_l = nullptr;
    float wscale_inv_scalar = 1.f;   // 2^-s of the layer's power-of-two weight pre-scale (tcgen05 path)
    float* dw_tc = nullptr;    // [cin/32][kernel][32] depthwise taps for the fused kernel
    alignas(64) unsigned char tm_w_hi[128];
    alignas(64) unsigned char tm_w_lo[128];
    alignas(64) unsigned char tm_r_hi[128];
    alignas(64) unsigned char tm_r_lo[128];
    // cached activation tensor maps (whole batch) and the key they were encoded for
    alignas(64) unsigned char tm_x[128];
    alignas(64) unsigned char tm_r[128];
    const void* tmc_x = nullptr; const void* tmc_r = nullptr; int tmc_B = 0, tmc_T = 0;
    alignas(64) unsigned char tm_y[128];      // output tensor map of the TMA-store epilogue
    const void* tmc_y = nullptr; int tmc_yB = 0, tmc_yT = 0;
    long long tmc_xs = 0, tmc_rs = 0, tmc_ys = 0;   // batch strides the cached maps were encoded with
    unsigned long long uid = 0;                      // unique per prepared layer (keys the segment descriptor cache)
};

}  // namespace vasr
