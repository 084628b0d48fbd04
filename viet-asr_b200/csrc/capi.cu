// C ABI: model lifecycle (state-dict loading, BatchNorm folding, operand packing), the encoder
// layer schedule over channels-last activations, decoder, and the host-buffer whole-path call.
// Interfaces replaced: see include/vasr_b200.h.
#include "common.cuh"
#include "kernels.cuh"
#include <map>
#include <vector>
#include <string>
#include <string.h>
#include <math.h>
#include <stdlib.h>

namespace vasr {

std::string& last_error_ref() { static thread_local std::string e; return e; }
std::atomic<long long> g_launch_count{0};

int set_error(int code, const char* fmt, ...)
{
    char buf[1024];
    va_list ap; va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}

struct HostTensor { std::vector<int64_t> dims; std::vector<float> data; };

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr size_t TILE_COUNTER_BYTES = 8192;   // one int per (layer, sub-batch) launch of the persistent kernels; the last int is the range-guard status word
constexpr size_t STATUS_WORD = TILE_COUNTER_BYTES / sizeof(int) - 1;
// tile counters + the (layer, utterance) completion counters of the multi-layer segment kernels; zeroed per forward
static inline size_t sync_region_bytes(size_t n_layers, int B) { return TILE_COUNTER_BYTES + align_up(n_layers * (size_t)B * sizeof(int), 256); }

}  // namespace vasr

struct vasr_model {
    std::vector<vasr_block_cfg> blocks;
    int feat_in = 0, num_classes = 0;
    std::map<std::string, vasr::HostTensor> host;
    bool finalized = false;
    int gemm_mode = VASR_GEMM_FP32_SIMT;
    std::vector<vasr::SubBlock> layers;
    std::vector<void*> allocs;
    int n_stage = 0;
    int *d_st_k = nullptr, *d_st_s = nullptr, *d_st_d = nullptr, *d_st_p = nullptr;
    float* d_dec_w = nullptr; float* d_dec_b = nullptr;
    int out_channels = 0;
    int cmax = 0;       // widest activation that lives in the workspace
    // sub-batch streams of the tensor-core encoder path (tail of one sub-batch's layer overlaps the next layer
    // of another: removes the wave-quantisation loss of 1 CTA/SM kernels)
    std::vector<cudaStream_t> sub_streams;
    std::vector<cudaEvent_t> sub_events;
    cudaEvent_t fork_event = nullptr;
    // vasr_transcribe_host: H2D of sub-batch i+1 overlaps the front end + encoder of sub-batch i
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t host_start = nullptr;
    cudaEvent_t copied[8] = {}, ready[8] = {};
    int max_sub = 2;   // measured on B200: 2 sub-batches beat 1 (tail overlap) and 4 (weight re-streaming)
    // scratch of vasr_transcribe_host (grown on demand)
    void* scratch = nullptr; size_t scratch_bytes = 0;
    size_t last_ws_off = 0, last_ws_bytes = 0; int last_B = 0;   // encoder workspace of the last vasr_transcribe_host_to_device
};

namespace vasr {

static int dev_upload(vasr_model* m, const std::vector<float>& h, float** out)
{
    float* d = nullptr;
    VASR_CUDA_OK(cudaMalloc(&d, sizeof(float) * h.size()));
    m->allocs.push_back(d);
    VASR_CUDA_OK(cudaMemcpy(d, h.data(), sizeof(float) * h.size(), cudaMemcpyHostToDevice));
    *out = d;
    return VASR_OK;
}

static int dev_upload_i(vasr_model* m, const std::vector<int>& h, int** out)
{
    int* d = nullptr;
    VASR_CUDA_OK(cudaMalloc(&d, sizeof(int) * (h.size() ? h.size() : 1)));
    m->allocs.push_back(d);
    if (!h.empty()) VASR_CUDA_OK(cudaMemcpy(d, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice));
    *out = d;
    return VASR_OK;
}

static int get_tensor(vasr_model* m, const std::string& name, std::initializer_list<int64_t> want,
                      const HostTensor** out)
{
    auto it = m->host.find(name);
    if (it == m->host.end())
        return set_error(VASR_EINVAL, "Missing key(s) in state_dict: \"%s\"", name.c_str());
    const HostTensor& t = it->second;
    std::vector<int64_t> w(want);
    bool ok = t.dims.size() == w.size();
    for (size_t i = 0; ok && i < w.size(); ++i) ok = t.dims[i] == w[i];
    if (!ok) {
        std::string got, exp;
        for (auto d : t.dims) got += std::to_string(d) + ",";
        for (auto d : w) exp += std::to_string(d) + ",";
        return set_error(VASR_EINVAL, "size mismatch for %s: checkpoint has [%s] but the model expects [%s]",
                         name.c_str(), got.c_str(), exp.c_str());
    }
    *out = &t;
    return VASR_OK;
}

// BatchNorm1d(eps=1e-3) in eval mode -> per-channel (scale, shift)   (parts/jasper.py:392)
static int bn_fold(vasr_model* m, const std::string& prefix, int c, std::vector<float>& scale, std::vector<float>& shift)
{
    const HostTensor *g, *b, *mu, *var;
    int rc;
    if ((rc = get_tensor(m, prefix + ".weight", {c}, &g))) return rc;
    if ((rc = get_tensor(m, prefix + ".bias", {c}, &b))) return rc;
    if ((rc = get_tensor(m, prefix + ".running_mean", {c}, &mu))) return rc;
    if ((rc = get_tensor(m, prefix + ".running_var", {c}, &var))) return rc;
    scale.resize(c); shift.resize(c);
    for (int i = 0; i < c; ++i) {
        const double s = (double)g->data[i] / sqrt((double)var->data[i] + 1e-3);
        scale[i] = (float)s;
        shift[i] = (float)((double)b->data[i] - (double)mu->data[i] * s);
    }
    return VASR_OK;
}

// fold the BN scale into a [cout][cin] 1x1 weight (host) and upload the fp32 copy for the CUDA-core path
static int fold_and_upload(vasr_model* m, const HostTensor* w, const std::vector<float>& scale, int cout, int cin,
                           std::vector<float>& folded, float** d_w)
{
    folded.resize((size_t)cout * cin);
    for (int o = 0; o < cout; ++o)
        for (int i = 0; i < cin; ++i) folded[(size_t)o * cin + i] = w->data[(size_t)o * cin + i] * scale[o];
    return dev_upload(m, folded, d_w);
}

}  // namespace vasr

// ---------------------------------------------------------------------------------------------
extern "C" int vasr_abi_version(void) { return VASR_ABI_VERSION; }
extern "C" const char* vasr_last_error(void) { return vasr::last_error_ref().c_str(); }
extern "C" int64_t vasr_launch_count(void) { return (int64_t)vasr::g_launch_count.load(); }

extern "C" int vasr_model_create(const vasr_block_cfg* blocks, int n_blocks, int feat_in,
                                 int num_classes_with_blank, vasr_model** out)
{
    using namespace vasr;
    VASR_REQUIRE(out && n_blocks >= 0 && (blocks || n_blocks == 0), "vasr_model_create: null block list");
    VASR_REQUIRE(feat_in > 0 && feat_in % 16 == 0, "vasr_model_create: feat_in must be a positive multiple of 16 (got %d)", feat_in);
    // 0 = encoder-only handle (no decoder head); n_blocks == 0 = decoder-only handle
    VASR_REQUIRE(num_classes_with_blank >= 0 && num_classes_with_blank <= 128 && num_classes_with_blank != 1,
                 "vasr_model_create: num_classes (+blank) must be 0 or in [2, 128] (got %d)", num_classes_with_blank);
    VASR_REQUIRE(n_blocks > 0 || num_classes_with_blank > 0, "vasr_model_create: neither encoder blocks nor a decoder head");
    for (int b = 0; b < n_blocks; ++b) {
        const vasr_block_cfg& c = blocks[b];
        VASR_REQUIRE(c.filters > 0 && c.filters % 16 == 0, "block %d: filters must be a positive multiple of 16 (got %d)", b, c.filters);
        VASR_REQUIRE(c.repeat >= 1, "block %d: repeat must be >= 1 (got %d)", b, c.repeat);
        VASR_REQUIRE(c.kernel >= 1 && c.kernel % 2 == 1, "block %d: kernel must be odd (got %d)", b, c.kernel);
        VASR_REQUIRE(c.stride >= 1 && c.dilation >= 1, "block %d: stride/dilation must be >= 1", b);
        // parts/jasper.py:61-62
        VASR_REQUIRE(!(c.stride > 1 && c.dilation > 1), "Only stride OR dilation may be greater than 1 (block %d)", b);
        VASR_REQUIRE(c.separable || c.kernel == 1,
                     "block %d: non-separable blocks are only built for kernel=1 (got %d)", b, c.kernel);
        VASR_REQUIRE(c.stride == 1 || (c.repeat == 1 && !c.residual && c.separable),
                     "block %d: a strided block must be separable with repeat=1 and no residual", b);
    }
    vasr_model* m = new vasr_model();
    if (n_blocks > 0) m->blocks.assign(blocks, blocks + n_blocks);
    m->feat_in = feat_in;
    m->num_classes = num_classes_with_blank;
    *out = m;
    return VASR_OK;
}

extern "C" void vasr_model_destroy(vasr_model* m)
{
    if (!m) return;
    for (void* p : m->allocs) cudaFree(p);
    if (m->scratch) cudaFree(m->scratch);
    for (cudaStream_t s : m->sub_streams) cudaStreamDestroy(s);
    for (cudaEvent_t e : m->sub_events) cudaEventDestroy(e);
    if (m->fork_event) cudaEventDestroy(m->fork_event);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    if (m->host_start) cudaEventDestroy(m->host_start);
    for (int i = 0; i < 8; ++i) { if (m->copied[i]) cudaEventDestroy(m->copied[i]); if (m->ready[i]) cudaEventDestroy(m->ready[i]); }
    delete m;
}

extern "C" int vasr_model_load_tensor(vasr_model* m, const char* name, const float* data,
                                      const int64_t* dims, int ndim, int is_device)
{
    using namespace vasr;
    VASR_REQUIRE(m && name && dims && ndim >= 0 && ndim <= 4, "vasr_model_load_tensor: bad argument");
    const std::string key(name);
    if (key.size() >= 19 && key.compare(key.size() - 19, 19, "num_batches_tracked") == 0) return VASR_OK;
    size_t n = 1;
    HostTensor t;
    for (int i = 0; i < ndim; ++i) { VASR_REQUIRE(dims[i] >= 0, "negative dim"); t.dims.push_back(dims[i]); n *= (size_t)dims[i]; }
    VASR_REQUIRE(data || n == 0, "vasr_model_load_tensor: null data for %s", name);
    t.data.resize(n);
    if (n) {
        if (is_device) VASR_CUDA_OK(cudaMemcpy(t.data.data(), data, n * sizeof(float), cudaMemcpyDeviceToHost));
        else memcpy(t.data.data(), data, n * sizeof(float));
    }
    m->host[key] = std::move(t);
    m->finalized = false;
    return VASR_OK;
}

extern "C" int vasr_model_finalize(vasr_model* m, int gemm_mode)
{
    using namespace vasr;
    VASR_REQUIRE(m, "vasr_model_finalize: null model");
    VASR_REQUIRE(gemm_mode >= VASR_GEMM_FP32_SIMT && gemm_mode <= VASR_GEMM_F16X1,
                 "vasr_model_finalize: unknown gemm_mode %d", gemm_mode);
    for (void* p : m->allocs) cudaFree(p);
    m->allocs.clear(); m->layers.clear();
    int rc;
    std::vector<int> st_k, st_s, st_d, st_p;
    int cin = m->feat_in, stage = 0, cmax = 0;
    const int nb = (int)m->blocks.size();
    for (int b = 0; b < nb; ++b) {
        const vasr_block_cfg& c = m->blocks[b];
        const int block_cin = cin;
        const int per = c.separable ? 5 : 4;
        int ci = cin;
        for (int r = 0; r < c.repeat; ++r) {
            SubBlock sb{};
            sb.cin = ci; sb.cout = c.filters; sb.kernel = c.kernel; sb.stride = c.stride; sb.dilation = c.dilation;
            sb.pad = (c.dilation > 1) ? (c.dilation * c.kernel) / 2 - 1 : c.kernel / 2;   // parts/jasper.py:60-65
            sb.separable = c.separable != 0;
            sb.relu = true;
            sb.has_res = (c.residual != 0) && (r == c.repeat - 1);
            sb.res_cin = block_cin;
            sb.final_layer = (b == nb - 1) && (r == c.repeat - 1);
            sb.len_stage_in = stage;
            const std::string pre = "encoder." + std::to_string(b) + ".mconv.";
            const HostTensor* w;
            std::vector<float> scale, shift, dwt;
            if (c.separable) {
                const HostTensor* dw;
                if ((rc = get_tensor(m, pre + std::to_string(per * r) + ".conv.weight", {ci, 1, c.kernel}, &dw))) return rc;
                if ((rc = get_tensor(m, pre + std::to_string(per * r + 1) + ".conv.weight", {c.filters, ci, 1}, &w))) return rc;
                if ((rc = bn_fold(m, pre + std::to_string(per * r + 2), c.filters, scale, shift))) return rc;
                dwt.resize((size_t)c.kernel * ci);
                for (int ch = 0; ch < ci; ++ch)
                    for (int k = 0; k < c.kernel; ++k) dwt[(size_t)k * ci + ch] = dw->data[(size_t)ch * c.kernel + k];
                if ((rc = dev_upload(m, dwt, &sb.dw_w))) return rc;
                if (c.stride > 1) {
                    st_k.push_back(c.kernel); st_s.push_back(c.stride); st_d.push_back(c.dilation); st_p.push_back(sb.pad);
                    ++stage;
                }
            } else {
                if ((rc = get_tensor(m, pre + std::to_string(per * r) + ".conv.weight", {c.filters, ci, 1}, &w))) return rc;
                if ((rc = bn_fold(m, pre + std::to_string(per * r + 1), c.filters, scale, shift))) return rc;
            }
            sb.len_stage_out = stage;
            std::vector<float> f_main, f_res;
            if ((rc = fold_and_upload(m, w, scale, c.filters, ci, f_main, &sb.pw_w))) return rc;
            if (sb.has_res) {
                const HostTensor* wr;
                std::vector<float> rscale, rshift;
                const std::string rp = "encoder." + std::to_string(b) + ".res.0.";
                if ((rc = get_tensor(m, rp + "0.conv.weight", {c.filters, block_cin, 1}, &wr))) return rc;
                if ((rc = bn_fold(m, rp + "1", c.filters, rscale, rshift))) return rc;
                if ((rc = fold_and_upload(m, wr, rscale, c.filters, block_cin, f_res, &sb.res_w))) return rc;
                for (int i = 0; i < c.filters; ++i) shift[i] += rshift[i];
            }
            if (gemm_mode != VASR_GEMM_FP32_SIMT) {
                if ((rc = tc_init())) return rc;
                VASR_REQUIRE(subblock_tc_supported(sb),
                             "tcgen05 path: sub-block (cin=%d cout=%d k=%d s=%d d=%d) is not a built shape",
                             sb.cin, sb.cout, sb.kernel, sb.stride, sb.dilation);
                if ((rc = tc_prepare_layer(sb, f_main.data(), sb.has_res ? f_res.data() : nullptr,
                                           c.separable ? dwt.data() : nullptr, m->allocs))) return rc;
            }
            if ((rc = dev_upload(m, shift, &sb.shift))) return rc;
            if (!sb.final_layer) cmax = std::max(cmax, sb.cout);
            cmax = std::max(cmax, sb.cin);
            m->layers.push_back(sb);
            ci = c.filters;
        }
        cin = c.filters;
    }
    m->out_channels = cin;
    m->cmax = cmax;
    m->n_stage = stage;
    if ((rc = dev_upload_i(m, st_k, &m->d_st_k))) return rc;
    if ((rc = dev_upload_i(m, st_s, &m->d_st_s))) return rc;
    if ((rc = dev_upload_i(m, st_d, &m->d_st_d))) return rc;
    if ((rc = dev_upload_i(m, st_p, &m->d_st_p))) return rc;
    m->d_dec_w = m->d_dec_b = nullptr;
    if (m->num_classes > 0) {
        const HostTensor *dw, *db;
        if ((rc = get_tensor(m, "decoder_layers.0.weight", {m->num_classes, cin, 1}, &dw))) return rc;
        if ((rc = get_tensor(m, "decoder_layers.0.bias", {m->num_classes}, &db))) return rc;
        if ((rc = dev_upload(m, dw->data, &m->d_dec_w))) return rc;
        if ((rc = dev_upload(m, db->data, &m->d_dec_b))) return rc;
    }
    m->gemm_mode = gemm_mode;
    m->finalized = true;
    VASR_CUDA_OK(cudaDeviceSynchronize());
    return VASR_OK;
}

extern "C" int vasr_model_gemm_mode(const vasr_model* m) { return m ? m->gemm_mode : -1; }

extern "C" int vasr_model_out_frames(const vasr_model* m, int T_f)
{
    if (!m || T_f <= 0) return vasr::set_error(VASR_EINVAL, "vasr_model_out_frames: bad argument");
    int t = T_f;
    for (const vasr_block_cfg& c : m->blocks) {
        const int pad = (c.dilation > 1) ? (c.dilation * c.kernel) / 2 - 1 : c.kernel / 2;
        for (int r = 0; r < c.repeat; ++r) t = (t + 2 * pad - c.dilation * (c.kernel - 1) - 1) / c.stride + 1;
    }
    return t;
}

extern "C" int vasr_model_out_channels(const vasr_model* m)
{
    if (!m) return -1;
    return m->blocks.empty() ? m->feat_in : m->blocks.back().filters;
}
extern "C" int vasr_model_num_classes(const vasr_model* m) { return m ? m->num_classes : -1; }

extern "C" size_t vasr_encoder_workspace_bytes(const vasr_model* m, int B, int T_f)
{
    if (!m || !m->finalized || B <= 0 || T_f <= 0) return 0;
    // 3 rotating activation buffers + 1 depthwise buffer, each [B, T_f, cmax] worst case
    // (T never grows along the stack), + the length table
    const size_t act = vasr::align_up((size_t)B * T_f * m->cmax * sizeof(float), 256);
    const size_t lens = vasr::align_up((size_t)(m->n_stage + 1) * B * sizeof(int), 256);
    return lens + vasr::sync_region_bytes(m->layers.size(), B) + 4 * act;
}

namespace vasr {
// number of sub-batch streams the tensor path uses for a batch (1 = single stream)
static int pick_nsub(const vasr_model* m, int B, int T_f)
{
    if (m->gemm_mode == VASR_GEMM_FP32_SIMT) return 1;
    int want = dev_env_int("VASR_SUBSTREAMS", m->max_sub);
    if (want < 1) want = 1;
    if (want > 8) want = 8;
    const int tiles_per_utt = ceil_div(vasr_model_out_frames(m, T_f), 128);
    // a sub-batch must still fill the machine on its own: the persistent kernels of two sub-batches cannot share an SM
    // (shared memory), so they only overlap at their tails.  Measured (profiles/r2_experiments.md): 128 x 5 s as 2 x 64
    // takes 5.41 ms, as 1 x 128 3.96 ms; from 256 tiles per sub-batch on, two sub-batches win by ~2 %
    while (want > 1 && (B / want) * tiles_per_utt < 256) --want;
    return want;
}
}  // namespace vasr

// sub_ready: optional [nsub] events; when given, sub-batch s waits for sub_ready[s] only (its features are ready)
// instead of everything enqueued on `stream` so far, the caller has already zeroed the tile counters and
// ordered the workspace against earlier work (vasr_transcribe_host).
static int encoder_forward_impl(vasr_model* m, const float* feat, const int64_t* seq_len, int B, int T_f,
                                float* enc, float* enc_len, void* workspace, size_t workspace_bytes,
                                void* stream, const cudaEvent_t* sub_ready)
{
    using namespace vasr;
    VASR_REQUIRE(m && feat && seq_len && enc && workspace, "vasr_encoder_forward: null argument");
    if (!m->finalized) return set_error(VASR_ESTATE, "vasr_encoder_forward: weights not finalized (restore_from first)");
    if (m->blocks.empty()) return set_error(VASR_ESTATE, "vasr_encoder_forward: this handle was created without encoder blocks");
    VASR_REQUIRE(B > 0 && T_f > 0, "vasr_encoder_forward: B and T must be positive (got %d, %d)", B, T_f);
    const size_t need = vasr_encoder_workspace_bytes(m, B, T_f);
    if (workspace_bytes < need)
        return set_error(VASR_ENOMEM, "vasr_encoder_forward: workspace %zu < required %zu bytes", workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    int* lens = (int*)ws;
    const size_t sync_b = sync_region_bytes(m->layers.size(), B);
    const size_t lens_b = align_up((size_t)(m->n_stage + 1) * B * sizeof(int), 256) + sync_b;
    int* counters = (int*)(ws + lens_b - sync_b);
    int* done = (int*)(ws + lens_b - sync_b + TILE_COUNTER_BYTES);   // [layer][B] completion counters (segment kernels)
    const size_t act = align_up((size_t)B * T_f * m->cmax * sizeof(float), 256);
    float* P[3] = {(float*)(ws + lens_b), (float*)(ws + lens_b + act), (float*)(ws + lens_b + 2 * act)};
    float* DW = (float*)(ws + lens_b + 3 * act);
    int rc;
    VASR_REQUIRE(m->layers.size() * 8 < STATUS_WORD, "too many layers for the tile-counter table");
    int* status = counters + STATUS_WORD;
    if (!sub_ready) {
        if (m->gemm_mode != VASR_GEMM_FP32_SIMT) VASR_CUDA_OK(cudaMemsetAsync(counters, 0, sync_b, st));
        if ((rc = launch_lens((const long long*)seq_len, B, 0, B, m->n_stage, m->d_st_k, m->d_st_s, m->d_st_d, m->d_st_p,
                              lens, enc_len, st))) return rc;
    }

    // ---- sub-batch plan: contiguous utterance ranges, each on its own stream (tensor path only)
    const bool tc = m->gemm_mode != VASR_GEMM_FP32_SIMT;
    const int nsub = pick_nsub(m, B, T_f);
    if (sub_ready && nsub < 2) return set_error(VASR_ESTATE, "internal: pipelined host path needs >= 2 sub-batches");
    if (nsub > 1) {
        while ((int)m->sub_streams.size() < nsub) {
            cudaStream_t s2; cudaEvent_t e2;
            VASR_CUDA_OK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
            VASR_CUDA_OK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            m->sub_streams.push_back(s2); m->sub_events.push_back(e2);
        }
        if (!m->fork_event) VASR_CUDA_OK(cudaEventCreateWithFlags(&m->fork_event, cudaEventDisableTiming));
        if (!sub_ready) {
            VASR_CUDA_OK(cudaEventRecord(m->fork_event, st));
            for (int s2 = 0; s2 < nsub; ++s2) VASR_CUDA_OK(cudaStreamWaitEvent(m->sub_streams[s2], m->fork_event, 0));
        } else {
            for (int s2 = 0; s2 < nsub; ++s2) {
                const int b0 = (int)((long long)B * s2 / nsub), b1 = (int)((long long)B * (s2 + 1) / nsub);
                VASR_CUDA_OK(cudaStreamWaitEvent(m->sub_streams[s2], sub_ready[s2], 0));
                if ((rc = launch_lens((const long long*)seq_len, B, b0, b1 - b0, m->n_stage, m->d_st_k, m->d_st_s,
                                      m->d_st_d, m->d_st_p, lens, enc_len, m->sub_streams[s2]))) return rc;
            }
        }
    }

    const int grid_limit = 0;       // every launch may use all SMs (persistent kernels of the sub-batches take turns)
    // ---- plan: buffers and shapes of every layer (3 rotating buffers; the output never aliases the layer's input
    // or the block input).  Tensor path: every utterance owns a FIXED region of each rotating buffer (batch stride
    // T_f * cmax), whatever the layer's T and C.  Sub-batches run on their own streams and may be several layers
    // apart (layers differ in C), so a per-layer [B, T, C] packing would let one sub-batch's output overlap the rows
    // another sub-batch is still reading.  The same invariant makes the multi-layer segment kernels safe: a tile only
    // ever touches its own utterance's regions, and it starts after every tile of the previous layer of that utterance.
    struct Plan { size_t li; const float* cur; const float* res; float* out; int T, T_out; long long xs, rs, ys; const int* len_in; const int* len_out; };
    std::vector<Plan> plan;
    {
        const float* cur = feat;
        const float* block_in = feat;
        int T = T_f;
        size_t li = 0;
        const long long ustride = (long long)T_f * m->cmax;
        for (size_t b = 0; b < m->blocks.size(); ++b) {
            const vasr_block_cfg& c = m->blocks[b];
            block_in = cur;
            for (int r = 0; r < c.repeat; ++r, ++li) {
                SubBlock& sb = m->layers[li];
                const int T_out = (T + 2 * sb.pad - sb.dilation * (sb.kernel - 1) - 1) / sb.stride + 1;
                float* out = nullptr;
                if (sb.final_layer) out = enc;
                else
                    for (int q = 0; q < 3; ++q)
                        if (P[q] != cur && P[q] != block_in) { out = P[q]; break; }
                auto bstride_of = [&](const float* ptr, int t, int ch) -> long long {
                    return (ptr == feat || ptr == enc) ? (long long)t * ch : ustride;
                };
                Plan pl;
                pl.li = li; pl.cur = cur; pl.res = sb.has_res ? block_in : nullptr; pl.out = out; pl.T = T; pl.T_out = T_out;
                pl.xs = bstride_of(cur, T, sb.cin); pl.rs = bstride_of(block_in, T, sb.res_cin); pl.ys = bstride_of(out, T_out, sb.cout);
                pl.len_in = lens + (size_t)sb.len_stage_in * B; pl.len_out = lens + (size_t)sb.len_stage_out * B;
                plan.push_back(pl);
                cur = out;
                T = T_out;
            }
        }
    }
    const int mega = dev_env_int("VASR_TC_MEGA", 1);       // developer builds: 0 = one launch per sub-block
    const int split3 = m->gemm_mode == VASR_GEMM_F16X3;
    auto run_layers = [&]() -> int {
    for (size_t i = 0; i < plan.size();) {
        // longest run of layers that one segment kernel can execute
        size_t j = i;
        std::vector<SegLayer> seg;
        if (tc && mega) {
            while (j < plan.size()) {
                const Plan& pl = plan[j];
                SubBlock& sb = m->layers[pl.li];
                if (!segment_tc_layer_ok(sb) || pl.T != pl.T_out) break;
                if (!seg.empty() && (sb.cout != seg[0].sb->cout || pl.T != plan[i].T)) break;
                seg.push_back(SegLayer{&sb, pl.cur, pl.xs, pl.res, pl.rs, pl.out, pl.ys, pl.len_out});
                if (!segment_tc_ok(seg.data(), (int)seg.size(), split3) && seg.size() >= 2) { seg.pop_back(); break; }
                ++j;
            }
            if (seg.size() < 2 || !segment_tc_ok(seg.data(), (int)seg.size(), split3)) { seg.clear(); j = i; }
        }
        if (!seg.empty()) {
            for (int s2 = 0; s2 < nsub; ++s2) {
                const int b0 = (int)((long long)B * s2 / nsub), b1 = (int)((long long)B * (s2 + 1) / nsub);
                if (b1 == b0) continue;
                if ((rc = launch_segment_tc(seg.data(), (int)seg.size(), B, plan[i].T, split3, b0, b1 - b0,
                                            counters + plan[i].li * 8 + s2, status, done + plan[i].li * (size_t)B, B, grid_limit,
                                            nsub > 1 ? m->sub_streams[s2] : st))) return rc;
            }
            i = j;
            continue;
        }
        const Plan& pl = plan[i];
        SubBlock& sb = m->layers[pl.li];
        if (tc) {
            for (int s2 = 0; s2 < nsub; ++s2) {
                const int b0 = (int)((long long)B * s2 / nsub), b1 = (int)((long long)B * (s2 + 1) / nsub);
                if (b1 == b0) continue;
                if ((rc = launch_subblock_tc(sb, pl.cur, pl.xs, pl.res, pl.rs, pl.out, pl.ys, B, pl.T, pl.T_out, pl.len_in, pl.len_out,
                                             split3, b0, b1 - b0, counters + pl.li * 8 + s2, status, grid_limit,
                                             nsub > 1 ? m->sub_streams[s2] : st))) return rc;
            }
        } else {
            const float* gin = pl.cur;
            if (sb.separable) {
                if ((rc = launch_dw_conv(pl.cur, sb.dw_w, DW, B, sb.cin, pl.T, pl.T_out, sb.kernel, sb.stride,
                                         sb.dilation, sb.pad, pl.len_in, pl.len_out, st))) return rc;
                gin = DW;
            }
            if ((rc = launch_pw_gemm(gin, sb.pw_w, sb.cin, pl.res, sb.res_w, sb.res_cin, sb.shift, pl.out, B, pl.T_out,
                                     sb.cout, pl.len_out, sb.relu ? 1 : 0, sb.final_layer ? 0 : 1, st))) return rc;
        }
        ++i;
    }
    return VASR_OK;
    };
    rc = run_layers();
    // join the sub-batch streams back into the caller's stream - also when a launch failed half-way, so that the
    // caller's stream never runs ahead of work that was already enqueued on them
    if (nsub > 1)
        for (int s2 = 0; s2 < nsub; ++s2) {
            const cudaError_t e1 = cudaEventRecord(m->sub_events[s2], m->sub_streams[s2]);
            const cudaError_t e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(st, m->sub_events[s2], 0) : e1;
            if (e2 != cudaSuccess && rc == VASR_OK) rc = set_error(VASR_ECUDA, "joining sub-batch stream %d failed: %s", s2, cudaGetErrorString(e2));
        }
    return rc;
}

extern "C" int vasr_encoder_forward(vasr_model* m, const float* feat, const int64_t* seq_len, int B, int T_f,
                                    float* enc, float* enc_len, void* workspace, size_t workspace_bytes,
                                    void* stream)
{
    return encoder_forward_impl(m, feat, seq_len, B, T_f, enc, enc_len, workspace, workspace_bytes, stream, nullptr);
}

extern "C" int vasr_encoder_check(vasr_model* m, void* workspace, size_t workspace_bytes, int B, void* stream)
{
    using namespace vasr;
    VASR_REQUIRE(m && workspace && B > 0, "vasr_encoder_check: bad argument");
    if (!m->finalized) return set_error(VASR_ESTATE, "vasr_encoder_check: weights not finalized");
    if (m->gemm_mode == VASR_GEMM_FP32_SIMT) { VASR_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream)); return VASR_OK; }
    const size_t lens_b = align_up((size_t)(m->n_stage + 1) * B * sizeof(int), 256);
    VASR_REQUIRE(workspace_bytes >= lens_b + TILE_COUNTER_BYTES, "vasr_encoder_check: workspace too small");
    int flag = 0;
    VASR_CUDA_OK(cudaMemcpyAsync(&flag, (const char*)workspace + lens_b + STATUS_WORD * sizeof(int), sizeof(int),
                                 cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    VASR_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    if (flag)
        return set_error(VASR_ERANGE, "encoder: a depthwise-convolution output exceeded the fp16 range (|x| >= 65520) of the "
                                      "f16x3/f16x1 operand format; results are invalid - use gemm_mode fp32 for this checkpoint");
    return VASR_OK;
}

extern "C" int vasr_decoder_forward(vasr_model* m, const float* enc, int B, int T_e,
                                    float* log_probs, int64_t* ids, void* stream)
{
    using namespace vasr;
    VASR_REQUIRE(m && enc, "vasr_decoder_forward: null argument");
    if (!m->finalized) return set_error(VASR_ESTATE, "vasr_decoder_forward: weights not finalized (restore_from first)");
    if (!m->d_dec_w) return set_error(VASR_ESTATE, "vasr_decoder_forward: this handle was created without a decoder head");
    VASR_REQUIRE(B > 0 && T_e > 0, "vasr_decoder_forward: B and T must be positive (got %d, %d)", B, T_e);
    return launch_decoder(enc, m->d_dec_w, m->d_dec_b, m->out_channels, m->num_classes, B * T_e, log_probs,
                          (long long*)ids, (cudaStream_t)stream);
}

// whole path from HOST waveforms; results either copied back to host buffers (then the call synchronises and checks the
// range guard) or left in caller-owned DEVICE buffers (everything stays enqueued on `stream`)
static int transcribe_impl(vasr_frontend* fe, vasr_model* m, const float* wave_host,
                           const int64_t* length_host, int B, int64_t L,
                           int32_t* out_ids, int32_t* out_len, bool out_on_host, void* stream)
{
    using namespace vasr;
    int32_t* out_ids_host = out_ids; int32_t* out_len_host = out_len;
    VASR_REQUIRE(fe && m && wave_host && length_host && out_ids && out_len, "vasr_transcribe_host: null argument");
    if (!m->finalized) return set_error(VASR_ESTATE, "vasr_transcribe_host: weights not finalized");
    if (m->blocks.empty() || !m->d_dec_w) return set_error(VASR_ESTATE, "vasr_transcribe_host: needs a handle with encoder and decoder");
    VASR_REQUIRE(B > 0, "vasr_transcribe_host: batch must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    const int T_f = vasr_frontend_num_frames(fe, L);
    if (T_f < 0) return T_f;
    const int T_e = vasr_model_out_frames(m, T_f);
    const int C = m->out_channels;
    const size_t ws_b = vasr_encoder_workspace_bytes(m, B, T_f);
    // scratch layout
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    const size_t o_wave = take((size_t)B * L * sizeof(float));
    const size_t o_len = take((size_t)B * sizeof(int64_t));
    const size_t o_feat = take((size_t)B * T_f * m->feat_in * sizeof(float));
    const size_t o_seq = take((size_t)B * sizeof(int64_t));
    const size_t o_enc = take((size_t)B * T_e * C * sizeof(float));
    const size_t o_elen = take((size_t)B * sizeof(float));
    const size_t o_ids = take((size_t)B * T_e * sizeof(int64_t));
    const size_t o_oid = take((size_t)B * T_e * sizeof(int32_t));
    const size_t o_olen = take((size_t)B * sizeof(int32_t));
    const size_t o_frames = take((size_t)B * sizeof(int32_t));
    const size_t o_ws = take(ws_b);
    if (off > m->scratch_bytes) {
        VASR_CUDA_OK(cudaStreamSynchronize(st));
        if (m->scratch) cudaFree(m->scratch);
        m->scratch = nullptr; m->scratch_bytes = 0;
        VASR_CUDA_OK(cudaMalloc(&m->scratch, off));
        m->scratch_bytes = off;
    }
    char* s = (char*)m->scratch;
    int rc;
    const int nsub = pick_nsub(m, B, T_f);
    // the encoder runs the batch as ONE launch chain (nsub = 1) unless every sub-batch fills the machine; the copy of
    // the waveforms is still cut into chunks so that the front end of chunk i runs while chunk i + 1 crosses PCIe
    const int nchunk = nsub >= 2 ? nsub : ((size_t)B * L * sizeof(float) >= ((size_t)4 << 20) ? (B < 4 ? B : 4) : 1);
    if (nchunk >= 2 && !m->copy_stream) {
        VASR_CUDA_OK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
        VASR_CUDA_OK(cudaEventCreateWithFlags(&m->host_start, cudaEventDisableTiming));
        for (int i = 0; i < 8; ++i) {
            VASR_CUDA_OK(cudaEventCreateWithFlags(&m->copied[i], cudaEventDisableTiming));
            VASR_CUDA_OK(cudaEventCreateWithFlags(&m->ready[i], cudaEventDisableTiming));
        }
    }
    if (nchunk < 2) {
        VASR_CUDA_OK(cudaMemcpyAsync(s + o_wave, wave_host, (size_t)B * L * sizeof(float), cudaMemcpyHostToDevice, st));
        VASR_CUDA_OK(cudaMemcpyAsync(s + o_len, length_host, (size_t)B * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        if ((rc = vasr_frontend_forward(fe, (const float*)(s + o_wave), (const int64_t*)(s + o_len), B, L,
                                        (float*)(s + o_feat), (int64_t*)(s + o_seq), st))) return rc;
        if ((rc = vasr_encoder_forward(m, (const float*)(s + o_feat), (const int64_t*)(s + o_seq), B, T_f,
                                       (float*)(s + o_enc), (float*)(s + o_elen), s + o_ws, ws_b, st))) return rc;
    } else if (nsub < 2) {
        VASR_CUDA_OK(cudaEventRecord(m->host_start, st));                 // scratch is free once earlier work on st is done
        VASR_CUDA_OK(cudaStreamWaitEvent(m->copy_stream, m->host_start, 0));
        for (int h = 0; h < nchunk; ++h) {
            const int b0 = (int)((long long)B * h / nchunk), b1 = (int)((long long)B * (h + 1) / nchunk);
            VASR_CUDA_OK(cudaMemcpyAsync(s + o_wave + (size_t)b0 * L * sizeof(float), wave_host + (size_t)b0 * L,
                                         (size_t)(b1 - b0) * L * sizeof(float), cudaMemcpyHostToDevice, m->copy_stream));
            VASR_CUDA_OK(cudaMemcpyAsync(s + o_len + (size_t)b0 * sizeof(int64_t), length_host + b0,
                                         (size_t)(b1 - b0) * sizeof(int64_t), cudaMemcpyHostToDevice, m->copy_stream));
            VASR_CUDA_OK(cudaEventRecord(m->copied[h], m->copy_stream));
        }
        for (int h = 0; h < nchunk; ++h) {
            const int b0 = (int)((long long)B * h / nchunk), b1 = (int)((long long)B * (h + 1) / nchunk);
            VASR_CUDA_OK(cudaStreamWaitEvent(st, m->copied[h], 0));
            if ((rc = vasr_frontend_forward(fe, (const float*)(s + o_wave) + (size_t)b0 * L, (const int64_t*)(s + o_len) + b0,
                                            b1 - b0, L, (float*)(s + o_feat) + (size_t)b0 * T_f * m->feat_in,
                                            (int64_t*)(s + o_seq) + b0, st))) return rc;
        }
        if ((rc = vasr_encoder_forward(m, (const float*)(s + o_feat), (const int64_t*)(s + o_seq), B, T_f,
                                       (float*)(s + o_enc), (float*)(s + o_elen), s + o_ws, ws_b, st))) return rc;
    } else {
        // software pipeline over the encoder's sub-batches: the waveforms of sub-batch i+1 cross PCIe on a copy
        // stream while sub-batch i runs its front end (on `st`) and its encoder layers (on its sub-stream)
        // tile counters live at the start of the encoder workspace (after the length table)
        const size_t lens_b = align_up((size_t)(m->n_stage + 1) * B * sizeof(int), 256);
        VASR_CUDA_OK(cudaMemsetAsync(s + o_ws + lens_b, 0, sync_region_bytes(m->layers.size(), B), st));
        VASR_CUDA_OK(cudaEventRecord(m->host_start, st));                 // scratch is free once earlier work on st is done
        VASR_CUDA_OK(cudaStreamWaitEvent(m->copy_stream, m->host_start, 0));
        for (int h = 0; h < nsub; ++h) {
            const int b0 = (int)((long long)B * h / nsub), b1 = (int)((long long)B * (h + 1) / nsub);
            VASR_CUDA_OK(cudaMemcpyAsync(s + o_wave + (size_t)b0 * L * sizeof(float), wave_host + (size_t)b0 * L,
                                         (size_t)(b1 - b0) * L * sizeof(float), cudaMemcpyHostToDevice, m->copy_stream));
            VASR_CUDA_OK(cudaMemcpyAsync(s + o_len + (size_t)b0 * sizeof(int64_t), length_host + b0,
                                         (size_t)(b1 - b0) * sizeof(int64_t), cudaMemcpyHostToDevice, m->copy_stream));
            VASR_CUDA_OK(cudaEventRecord(m->copied[h], m->copy_stream));
        }
        for (int h = 0; h < nsub; ++h) {
            const int b0 = (int)((long long)B * h / nsub), b1 = (int)((long long)B * (h + 1) / nsub);
            VASR_CUDA_OK(cudaStreamWaitEvent(st, m->copied[h], 0));
            if ((rc = vasr_frontend_forward(fe, (const float*)(s + o_wave) + (size_t)b0 * L, (const int64_t*)(s + o_len) + b0,
                                            b1 - b0, L, (float*)(s + o_feat) + (size_t)b0 * T_f * m->feat_in,
                                            (int64_t*)(s + o_seq) + b0, st))) return rc;
            VASR_CUDA_OK(cudaEventRecord(m->ready[h], st));
        }
        if ((rc = encoder_forward_impl(m, (const float*)(s + o_feat), (const int64_t*)(s + o_seq), B, T_f,
                                       (float*)(s + o_enc), (float*)(s + o_elen), s + o_ws, ws_b, st, m->ready))) return rc;
    }
    if ((rc = vasr_decoder_forward(m, (const float*)(s + o_enc), B, T_e, nullptr, (int64_t*)(s + o_ids), st))) return rc;
    // every utterance is collapsed over the frames it would have had alone (the reference transcribes one utterance per
    // call, infer.py:167-171): T_e of its own length, not of the batch's padded length
    std::vector<int32_t> frames((size_t)B);
    for (int b = 0; b < B; ++b) {
        int64_t lb = length_host[b];
        if (lb > L) lb = L;
        const int tf = lb > 0 ? vasr_frontend_num_frames(fe, lb) : 0;
        frames[b] = tf > 0 ? vasr_model_out_frames(m, tf) : 0;
    }
    VASR_CUDA_OK(cudaMemcpyAsync(s + o_frames, frames.data(), (size_t)B * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    if ((rc = vasr_ctc_collapse((const int64_t*)(s + o_ids), (const int32_t*)(s + o_frames), B, T_e, m->num_classes - 1,
                                (int32_t*)(s + o_oid), (int32_t*)(s + o_olen), st))) return rc;
    if (!out_on_host) {
        // results stay on the device (e.g. for an NCCL gather); nothing is synchronised - the caller checks the range
        // guard with vasr_transcribe_check once it has synchronised anyway
        VASR_CUDA_OK(cudaMemcpyAsync(out_ids, s + o_oid, (size_t)B * T_e * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        VASR_CUDA_OK(cudaMemcpyAsync(out_len, s + o_olen, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        m->last_ws_off = o_ws; m->last_ws_bytes = ws_b; m->last_B = B;
        return VASR_OK;
    }
    VASR_CUDA_OK(cudaMemcpyAsync(out_ids_host, s + o_oid, (size_t)B * T_e * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    VASR_CUDA_OK(cudaMemcpyAsync(out_len_host, s + o_olen, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    VASR_CUDA_OK(cudaStreamSynchronize(st));
    if ((rc = vasr_encoder_check(m, s + o_ws, ws_b, B, st))) return rc;      // VASR_ERANGE on fp16 overflow (tensor-core modes)
#ifdef VASR_DEV
    if (const char* dump = getenv("VASR_HOST_DUMP")) {       // debugging aid: intermediate tensors of the host route
        auto dump_buf = [&](const char* name, size_t off_b, size_t bytes) {
            std::vector<char> h(bytes);
            cudaMemcpy(h.data(), s + off_b, bytes, cudaMemcpyDeviceToHost);
            std::string path = std::string(dump) + "/" + name;
            FILE* f = fopen(path.c_str(), "wb");
            if (f) { fwrite(h.data(), 1, bytes, f); fclose(f); }
        };
        dump_buf("feat.bin", o_feat, (size_t)B * T_f * m->feat_in * sizeof(float));
        dump_buf("seq.bin", o_seq, (size_t)B * sizeof(int64_t));
        dump_buf("enc.bin", o_enc, (size_t)B * T_e * C * sizeof(float));
        dump_buf("ids.bin", o_ids, (size_t)B * T_e * sizeof(int64_t));
        dump_buf("lens.bin", o_ws, (size_t)(m->n_stage + 1) * B * sizeof(int));
    }
#endif
    return VASR_OK;
}

extern "C" int vasr_transcribe_host(vasr_frontend* fe, vasr_model* m, const float* wave_host,
                                    const int64_t* length_host, int B, int64_t L,
                                    int32_t* out_ids_host, int32_t* out_len_host, void* stream)
{
    return transcribe_impl(fe, m, wave_host, length_host, B, L, out_ids_host, out_len_host, true, stream);
}

extern "C" int vasr_transcribe_host_to_device(vasr_frontend* fe, vasr_model* m, const float* wave_host,
                                              const int64_t* length_host, int B, int64_t L,
                                              int32_t* out_ids_dev, int32_t* out_len_dev, void* stream)
{
    return transcribe_impl(fe, m, wave_host, length_host, B, L, out_ids_dev, out_len_dev, false, stream);
}

extern "C" int vasr_transcribe_check(vasr_model* m, void* stream)
{
    using namespace vasr;
    VASR_REQUIRE(m, "vasr_transcribe_check: null model");
    if (!m->scratch || m->last_B <= 0) { VASR_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream)); return VASR_OK; }
    return vasr_encoder_check(m, (char*)m->scratch + m->last_ws_off, m->last_ws_bytes, m->last_B, stream);
}
