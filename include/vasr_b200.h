/*
 * vasr_b200.h - C ABI of the B200-native VietASR CTC inference hot path.
 *
 * Plain C: opaque handles, raw device/host pointers, sizes.  No torch types.
 * Every entry point returns 0 on success and a negative vasr_status on
 * failure; vasr_last_error() gives the thread-local message the host shim
 * raises as ValueError (VASR_EINVAL) or RuntimeError (everything else).
 * Unless a function says "host", every data pointer is a DEVICE pointer the
 * caller owns and keeps alive; work is enqueued on the cudaStream_t passed as
 * `stream` (a void*, 0 = legacy default stream) and no entry point
 * synchronises the device except vasr_model_finalize and the *_host calls.
 * Handles (vasr_frontend, vasr_model, vasr_lm, vasr_resampler) are NOT thread-safe and are single-stream: a handle
 * owns scratch buffers, helper streams and cached descriptors, so concurrent calls on one handle from two host
 * threads, or interleaved calls on two streams without the caller ordering them, are undefined.  Use one handle
 * per module instance per host thread per device (the reference shares one instance across requests without a
 * lock, app.py:22-28; a server doing that must serialise its calls).
 *
 * Activations inside the library are channels-last fp32:
 *     features [B, T_f, F]   activations [B, T, C]   log-probs [B, T_e, V+1]
 * i.e. the reference's [B, C, T] tensors transposed (the host shim hands them
 * back to NeMo-style callers as transposed *views*, no copy).
 *
 * Each entry point replaces one reference interface (paths relative to the
 * dangvansam/viet-asr checkout):
 *
 *   vasr_frontend_*      FilterbankFeatures.__init__/forward,
 *                        nemo/collections/asr/parts/features.py:113-236, 245-301;
 *                        normalize_batch :17-30; wrapper
 *                        AudioToMelSpectrogramPreprocessor.forward,
 *                        nemo/collections/asr/audio_preprocessing.py:78-87, 314-383
 *   vasr_model_create    JasperEncoder.__init__ + JasperDecoderForCTC.__init__,
 *                        nemo/collections/asr/jasper.py:136-196, 242-251
 *   vasr_model_load_tensor / vasr_model_finalize
 *                        TrainableNM.restore_from -> load_state_dict,
 *                        nemo/backends/pytorch/nm.py:97-103
 *   vasr_encoder_forward JasperEncoder.forward, jasper.py:198-204
 *                        (JasperBlock.forward parts/jasper.py:408-448,
 *                         MaskedConv1d.forward parts/jasper.py:113-132)
 *   vasr_decoder_forward JasperDecoderForCTC.forward jasper.py:253-254 +
 *                        GreedyCTCDecoder.forward greedy_ctc_decoder.py:33-36
 *   vasr_ctc_collapse    __ctc_decoder_predictions_tensor,
 *                        nemo/collections/asr/helpers.py:7-33
 *   vasr_transcribe_host VietASR.transcribe, infer.py:167-171 (greedy wiring,
 *                        infer.py:113), batched; host buffers in, ids out
 */
#ifndef VASR_B200_H
#define VASR_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VASR_ABI_VERSION 2

typedef enum vasr_status {
    VASR_OK = 0,
    VASR_EINVAL = -1,     /* bad argument / unsupported configuration  (ValueError)   */
    VASR_ECUDA = -2,      /* CUDA runtime / driver error               (RuntimeError) */
    VASR_ESTATE = -3,     /* call order (e.g. forward before finalize) (RuntimeError) */
    VASR_ENOMEM = -4,     /* workspace too small / allocation failed   (RuntimeError) */
    VASR_ERANGE = -5      /* an activation left the range of the f16x3 / f16x1 operand format (RuntimeError) */
} vasr_status;

/* precision of the 1x1 (pointwise / residual / final) GEMMs of the encoder */
typedef enum vasr_gemm_mode {
    VASR_GEMM_FP32_SIMT = 0,  /* fp32 FMA on CUDA cores (exact-order reference path)          */
    VASR_GEMM_F16X3     = 1,  /* tcgen05 kind::f16, fp16 hi/lo split on both operands, 3 products:
                                 fp32-grade (parity mode)                                        */
    VASR_GEMM_F16X1     = 2   /* tcgen05 kind::f16, single product (fast mode, ~5e-4 rel logits) */
} vasr_gemm_mode;

/* One Jasper block, fields as in the YAML `jasper:` list (jasper.py:27-66). */
typedef struct vasr_block_cfg {
    int32_t filters;    /* output channels                                   */
    int32_t repeat;     /* sub-blocks                                        */
    int32_t kernel;     /* conv kernel size (odd)                            */
    int32_t stride;     /* >1 only with dilation == 1                        */
    int32_t dilation;
    int32_t residual;   /* 0/1 : 1x1 conv + BN of the block input is added   */
    int32_t separable;  /* 1: depthwise + pointwise ; 0: plain conv, kernel must be 1 */
} vasr_block_cfg;

/* Front-end configuration (ctor kwargs of AudioToMelSpectrogramPreprocessor that
 * matter on the inference path; dither = 0). */
typedef struct vasr_frontend_cfg {
    int32_t n_window_size;    /* 320  */
    int32_t n_window_stride;  /* 160  */
    int32_t n_fft;            /* 512 (only 512 is built)                     */
    int32_t nfilt;            /* 64                                          */
    float   preemph;          /* 0.97                                        */
    float   log_zero_guard;   /* 2^-24, 'add' guard                          */
    int32_t pad_to;           /* 0 (infer.py:90) or a positive multiple      */
} vasr_frontend_cfg;

typedef struct vasr_frontend vasr_frontend;
typedef struct vasr_model vasr_model;

/* ---- misc ------------------------------------------------------------- */
int         vasr_abi_version(void);
const char* vasr_last_error(void);
/* number of kernels this library has launched in this process (all handles) */
int64_t     vasr_launch_count(void);

/* ---- front end -------------------------------------------------------- */
/* window: host float[n_window_size] (torch.hann_window(n, periodic=False));
 * mel_fb: host float[nfilt * (n_fft/2+1)] row-major (librosa slaney basis). */
int  vasr_frontend_create(const vasr_frontend_cfg* cfg, const float* window_host,
                          const float* mel_fb_host, vasr_frontend** out);
void vasr_frontend_destroy(vasr_frontend* fe);
/* Where the STFT's centre padding reflects (torch.stft(center=True), features.py:181-188):
 *   per_utterance = 0 (default): at the end of the batch-padded row of L samples - exactly what the reference computes for
 *                      a [B, L] tensor; the last feature frame of a SHORTER utterance then sees the zero padding;
 *   per_utterance = 1: at every utterance's own length[b] - what that utterance sees when the reference transcribes it
 *                      alone (infer.py:167-171 only ever runs one utterance per call), so a transcript does not depend on
 *                      what else is in the batch.  Frames at or beyond ceil(length/hop) are zeroed in both modes. */
int  vasr_frontend_set_padding(vasr_frontend* fe, int per_utterance);
/* T_f = 1 + L / hop, then padded up to a multiple of pad_to if pad_to > 0 */
int  vasr_frontend_num_frames(const vasr_frontend* fe, int64_t L);
/* wave [B, L] f32, length [B] i64  ->  feat [B, T_f, nfilt] f32 (normalised, masked),
 * seq_len [B] i64 = ceil(length / hop).  Reflect padding is over the batch-padded row
 * of L samples, like torch.stft(center=True) on [B, L]; requires L > n_fft/2.        */
int  vasr_frontend_forward(vasr_frontend* fe, const float* wave, const int64_t* length,
                           int B, int64_t L, float* feat, int64_t* seq_len, void* stream);

/* ---- acoustic model --------------------------------------------------- */
int  vasr_model_create(const vasr_block_cfg* blocks, int n_blocks, int feat_in,
                       int num_classes_with_blank, vasr_model** out);
void vasr_model_destroy(vasr_model* m);
/* name = state-dict key ("encoder.3.mconv.1.conv.weight", "decoder_layers.0.bias", ...);
 * data = fp32, host or device (is_device), dims = shape.  "num_batches_tracked" is ignored. */
int  vasr_model_load_tensor(vasr_model* m, const char* name, const float* data,
                            const int64_t* dims, int ndim, int is_device);
/* folds BatchNorm (eps = 1e-3) into the 1x1 weights, packs operands; errors name the first missing key */
int  vasr_model_finalize(vasr_model* m, int gemm_mode);
int  vasr_model_gemm_mode(const vasr_model* m);
/* T_e for T_f feature frames */
int  vasr_model_out_frames(const vasr_model* m, int T_f);
int  vasr_model_out_channels(const vasr_model* m);
int  vasr_model_num_classes(const vasr_model* m);
size_t vasr_encoder_workspace_bytes(const vasr_model* m, int B, int T_f);

/* feat [B, T_f, feat_in], seq_len [B] i64 -> enc [B, T_e, C_out] f32 (tail frames NOT
 * zeroed, like the reference), enc_len [B] f32.  workspace: >= vasr_encoder_workspace_bytes. */
int  vasr_encoder_forward(vasr_model* m, const float* feat, const int64_t* seq_len, int B, int T_f,
                          float* enc, float* enc_len, void* workspace, size_t workspace_bytes,
                          void* stream);
/* Range guard of the tensor-core modes.  f16x3 / f16x1 feed the 1x1 convolutions with fp16 operands; a depthwise
 * output beyond +-65504 cannot be represented and would silently turn into inf/NaN and then - through ReLU - into
 * zeros.  Every encoder kernel records that event in a status word inside `workspace`.  vasr_encoder_check
 * synchronises `stream`, reads the word of the last vasr_encoder_forward that used `workspace` and returns
 * VASR_ERANGE if it is set (VASR_OK otherwise; always VASR_OK in fp32 mode).  vasr_transcribe_host checks it itself. */
int  vasr_encoder_check(vasr_model* m, void* workspace, size_t workspace_bytes, int B, void* stream);
/* enc [B, T_e, C_out] -> log_probs [B, T_e, V+1] f32 (may be NULL), ids [B, T_e] i64 greedy argmax */
int  vasr_decoder_forward(vasr_model* m, const float* enc, int B, int T_e,
                          float* log_probs, int64_t* ids, void* stream);

/* GreedyCTCDecoder.forward on foreign log-probs: log_probs [N, V] f32 -> ids [N] i64 (ties -> lowest index) */
int  vasr_greedy_argmax(const float* log_probs, int N, int V, int64_t* ids, void* stream);

/* ---- greedy CTC collapse ---------------------------------------------- */
/* ids [B, T] i64 -> out_ids [B, T] i32 (collapsed, -1 padded), out_len [B] i32.
 * frames [B] i32 (device) or NULL: utterance b is collapsed over its first min(frames[b], T) frames.  NULL = all T
 * frames, which is what the reference does with the single utterance it is given (helpers.py:26-30); in a
 * zero-padded batch pass vasr_model_out_frames(frontend frames of utterance b) so that every utterance is decoded
 * over exactly the frames it would have had alone. */
int  vasr_ctc_collapse(const int64_t* ids, const int32_t* frames, int B, int T, int blank,
                       int32_t* out_ids, int32_t* out_len, void* stream);

/* ---- CTC prefix beam search, no language model ------------------------- */
/* BeamSearchDecoderWithLM.forward with lm_path=None (beam_search_decoder.py:95-102 -> pyctcdecode decode()).
 * log_probs [B, T, V] f32, frames [B] i32 or NULL as in vasr_ctc_collapse (NULL: all T frames are used, like the
 * reference with its batch of one, beam_search_decoder.py:96-101), blank = V-1 for NeMo vocabularies,
 * space_id = index of ' ' in the vocabulary or -1.  out_ids [B, T] i32 (best text as symbol ids, -1 padded,
 * single spaces, no leading space), out_len [B] i32, out_score [B] f32 (log score, may be NULL).
 * beam_width <= 128; pyctcdecode defaults: token_min_logp = -5, beam_prune_logp = -10. */
size_t vasr_ctc_beam_workspace_bytes(int B, int T);
int  vasr_ctc_beam_search(const float* log_probs, const int32_t* frames, int B, int T, int V, int blank, int space_id,
                          int beam_width, float token_min_logp, float beam_prune_logp,
                          void* workspace, size_t workspace_bytes,
                          int32_t* out_ids, int32_t* out_len, float* out_score, void* stream);

/* ---- n-gram language model fused into the beam search ------------------- */
/* Replaces the kenlm.Model that pyctcdecode.build_ctcdecoder(vocab, kenlm_model_path=lm_path, alpha, beta) opens
 * (beam_search_decoder.py:82-87; infer.py:184-191: 3-gram-lm.binary, beam 100, alpha 0.5, beta 1.5).
 * The host decodes the KenLM binary (viet-asr_b200/kenlm_binary.py) into the flat HOST arrays below; vasr_lm_create
 * copies them to the current device.  Reverse trie: node of n-gram (w1..wn) is reached by wn -> w(n-1) -> ...;
 * prob/backoff are log10.  *_next[i] .. *_next[i+1] is the child range of node i in the next order's arrays.
 * vocab_keys/vals: open-addressing table (linear probing, 0 = empty, power-of-two slots) from
 * vasr_lm_hash_labels(label ids of a word) to its word id; words missing from the table score as <unk> (id 0). */
typedef struct vasr_lm vasr_lm;
typedef struct vasr_lm_arrays {
    int32_t order;                      /* 2..5 */
    int32_t vocab;                      /* == counts[0] */
    int32_t bos, eos;                   /* ids of <s>, </s> */
    const uint64_t* counts;             /* [order] n-grams per order */
    const float* uni_prob; const float* uni_backoff; const uint32_t* uni_next;      /* [vocab], [vocab], [vocab+1] */
    const int32_t* mid_word[3]; const float* mid_prob[3]; const float* mid_backoff[3];   /* orders 2..order-1: [counts[k+1]] */
    const uint32_t* mid_next[3];                                                     /* [counts[k+1] + 1] */
    const int32_t* long_word; const float* long_prob;                                /* order `order`: [counts[order-1]] */
    const uint64_t* vocab_keys; const int32_t* vocab_vals; int32_t vocab_slots;
} vasr_lm_arrays;
int  vasr_lm_create(const vasr_lm_arrays* host_arrays, vasr_lm** out);
void vasr_lm_destroy(vasr_lm* lm);
int  vasr_lm_order(const vasr_lm* lm);
uint64_t vasr_lm_hash_labels(const int32_t* label_ids, int n);   /* host function: key of a word in vocab_keys */
/* log10 P(word | ctx) with back-off (kenlm Model.BaseScore) for N queries on the device: ctx [N, 4] i32 oldest ->
 * newest (first nctx[i] <= order-1 entries used), word [N] i32, out [N] f64 - all device pointers. */
int  vasr_lm_score_batch(const vasr_lm* lm, const int32_t* ctx, const int32_t* nctx, const int32_t* word,
                         double* out, int N, void* stream);
/* Beam search with LM shallow fusion (pyctcdecode decode() with a kenlm model and no unigram list):
 * candidates ranked by acoustic + sum over words (alpha * ln P_lm(word | history) + beta) + unk_score_offset for a
 * non-empty partial word (x len/6 beyond 6 characters); out-of-vocabulary words are charged unk_score_offset (log10)
 * before alpha; a text first completed at the end of the utterance also gets P(</s>).  pyctcdecode default
 * unk_score_offset = -10.  Labels must be single characters.  out_score = acoustic + LM score of the best text. */
size_t vasr_ctc_beam_lm_workspace_bytes(int B, int T, int beam_width);
int  vasr_ctc_beam_search_lm(const float* log_probs, const int32_t* frames, int B, int T, int V, int blank, int space_id,
                             int beam_width, float token_min_logp, float beam_prune_logp,
                             const vasr_lm* lm, double alpha, double beta, double unk_score_offset,
                             void* workspace, size_t workspace_bytes,
                             int32_t* out_ids, int32_t* out_len, float* out_score, void* stream);

/* ---- audio ingest: PCM decode + sample-rate conversion ------------------- */
/* What the reference's callers do on the CPU before VietASR.transcribe: int16 PCM -> float32 (x / 2^15, the
 * soundfile/librosa convention, nemo/collections/asr/parts/segment.py:61-74) and librosa.load(path, sr=16000)
 * (infer.py:200, app.py:66,82) = resampy "kaiser_best" windowed-sinc interpolation, output length
 * ceil(n * sr_out / sr_in).  librosa/resampy are un-vendored: restated from the published algorithm, parity with
 * the packages UNPINNED (oracle/resample_oracle.py).
 * interp_win_host: right half of the low-pass (num_zeros * num_table + 1 samples, num_table per zero crossing). */
typedef struct vasr_resampler vasr_resampler;
int  vasr_resampler_create(const float* interp_win_host, int n_win, int num_table, vasr_resampler** out);
void vasr_resampler_destroy(vasr_resampler* rs);
int64_t vasr_resample_out_len(int64_t n_in, int sr_in, int sr_out);
/* pcm [B, L] i16, length [B] i64 -> wave [B, L] f32 (zero beyond length)                                         */
int  vasr_pcm16_to_float(const int16_t* pcm, const int64_t* length, int B, int64_t L, float* wave, void* stream);
/* x [B, L_in] (f32, or i16 PCM when pcm16 != 0), len_in [B] i64 -> y [B, L_out] f32 (zero beyond len_out),
 * len_out [B] i64; L_out >= vasr_resample_out_len(L_in, sr_in, sr_out).  All device pointers.  The handle caches
 * resampy's running-sum time table per ratio (rebuilt on the stream when the ratio changes or L_out grows): one
 * handle per caller thread.                                                                                       */
int  vasr_resample(vasr_resampler* rs, const void* x, int pcm16, const int64_t* len_in, int B, int64_t L_in,
                   int sr_in, int sr_out, float* y, int64_t* len_out, int64_t L_out, void* stream);

/* ---- whole path, HOST buffers (the reference-facing plugin call) ------- */
/* wave_host [B, L] f32 and length_host [B] i64 in (pinned or pageable) host memory;
 * out_ids_host [B, T_e] i32 (-1 padded) and out_len_host [B] i32 in host memory.
 * Copies H2D, runs front end -> encoder -> decoder -> collapse, copies D2H, synchronises.
 * Device scratch is cached inside the model handle and grows on demand.              */
int  vasr_transcribe_host(vasr_frontend* fe, vasr_model* m, const float* wave_host,
                          const int64_t* length_host, int B, int64_t L,
                          int32_t* out_ids_host, int32_t* out_len_host, void* stream);
/* Same, but the collapsed ids stay in caller-owned DEVICE buffers out_ids_dev [B, T_e] / out_len_dev [B] and nothing is
 * synchronised: the pipelined host->device copy of the waveforms and the compute are enqueued on `stream` (and the
 * handle's helper streams, joined back into it).  For data-parallel runs: every rank feeds its own host shard and the
 * ids are gathered over NCCL (replaces the .cpu()/all_gather of the reference's PtActions._infer,
 * nemo/backends/pytorch/actions.py:594-612, 784-811).  vasr_transcribe_check synchronises `stream` and reports the
 * range guard (VASR_ERANGE) of the last such call. */
int  vasr_transcribe_host_to_device(vasr_frontend* fe, vasr_model* m, const float* wave_host,
                                    const int64_t* length_host, int B, int64_t L,
                                    int32_t* out_ids_dev, int32_t* out_len_dev, void* stream);
int  vasr_transcribe_check(vasr_model* m, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VASR_B200_H */
