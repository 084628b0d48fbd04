// Internal launch wrappers shared between translation units.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include <vector>

namespace vasr {

int launch_lens(const long long* seq_len, int B, int b0, int nb, int n_stage, const int* st_k, const int* st_s,
                const int* st_d, const int* st_p, int* lens, float* enc_len, cudaStream_t st);

int launch_dw_conv(const float* x, const float* w, float* y, int B, int C, int T_in, int T_out,
                   int K, int S, int D, int pad, const int* len_in, const int* len_out, cudaStream_t st);

int launch_pw_gemm(const float* X, const float* W, int Cin, const float* R, const float* Wr, int Cres,
                   const float* shift, float* Y, int B, int T, int Cout, const int* len, int relu,
                   int mask_tail, cudaStream_t st);

int launch_decoder(const float* enc, const float* W, const float* bias, int Cin, int V1,
                   int N, float* logp, long long* ids, cudaStream_t st);

int launch_ctc_collapse(const long long* ids, const int* frames, int B, int T, int blank, int* out_ids, int* out_len,
                        cudaStream_t st);

// tcgen05 fused sub-block (encoder_tc.cu); returns VASR_EINVAL when the shape is not built
int launch_subblock_tc(SubBlock& sb, const float* x, long long x_bstride, const float* res_in, long long r_bstride,
                       float* y, long long y_bstride, int B, int T_in,
                       int T_out, const int* len_in, const int* len_out, int split3, int b0, int nb,
                       int* tile_counter, int* status, int grid_limit, cudaStream_t st);
bool subblock_tc_supported(const SubBlock& sb);
// multi-layer persistent launch over a run of stride-1 separable sub-blocks with the same cout and frame count
struct SegLayer {
    SubBlock* sb;
    const float* x; long long xs;      // input [B, T, cin] + batch stride (elements)
    const float* res; long long rs;    // residual-branch input (block input) or null
    float* y; long long ys;            // output [B, T, cout]
    const int* len_out;                // [B]
};
bool segment_tc_layer_ok(const SubBlock& sb);
bool segment_tc_ok(const SegLayer* L, int n, int split3);
int launch_segment_tc(const SegLayer* L, int n, int B, int T, int split3, int b0, int nb,
                      int* tile_counter, int* status, int* done, int done_stride, int grid_limit, cudaStream_t st);
int tc_init();
// w_main [cout][cin], w_res [cout][res_cin] (or null): BN-scale-folded fp32 weights on the host
int tc_prepare_layer(SubBlock& sb, const float* w_main, const float* w_res, const float* dw_kc /*[K][cin] or null*/,
                     std::vector<void*>& allocs);

}  // namespace vasr
