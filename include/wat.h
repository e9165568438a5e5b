/* libwat — C ABI of the B200-native Whisper-AT tagging path.
 *
 * The reference (chat-prompt/whisper-at) is pure Python and has no FFI; its boundary for this path is the
 * Python API.  Each entry point below replaces the arithmetic behind one reference call so a
 * maintainer can bind it with ctypes:
 *
 *   wat_logmel      <- whisper_at.audio.log_mel_spectrogram        package/whisper-at/whisper_at/audio.py:110-157
 *   wat_encoder     <- AudioEncoder.forward (x, all_x)             package/whisper-at/whisper_at/model.py:156-177
 *   wat_tltr        <- ATModel.forward                             package/whisper-at/whisper_at/model.py:351-379
 *   wat_tag         <- the per-window body of transcribe()         package/whisper-at/whisper_at/transcribe.py:127,241-263
 *   wat_tag_host    <- same, host buffers in / host logits out (what a non-torch caller binds)
 *   wat_tag_pcm16 / wat_tag_host_pcm16 <- same for the int16 PCM that load_audio decodes (audio.py:26-63)
 *   wat_create / wat_set_weight / wat_finalize  <- Whisper.__init__ + load_state_dict   model.py:224-246, __init__.py:184-191
 *
 * Conventions: plain pointers and sizes only; device pointers are owned by the caller (e.g. torch tensors);
 * `stream` is a cudaStream_t passed as void*; every function returns 0 on success or a negative wat_status,
 * and wat_last_error() returns a thread-local message.  No C++ exception crosses the boundary.  A handle is
 * bound to the CUDA device that was current at wat_create and is not re-entrant (one handle per host thread).  Calls on
 * one handle share its workspace: they are ordered on the device in call order even when they are issued on different
 * streams (an event dependency is inserted), and every entry point restores the caller's current device before returning.
 * There is no CPU fallback: without a CUDA device every compute call returns WAT_ERR_CUDA.
 */
#ifndef WAT_H_
#define WAT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WAT_ABI_VERSION 2

#if defined(__GNUC__)
#define WAT_API __attribute__((visibility("default")))
#else
#define WAT_API
#endif

typedef struct wat_handle wat_handle;

enum wat_status {
  WAT_OK = 0,
  WAT_ERR_INVALID = -1,   /* bad argument / shape (the reference raises AssertionError / RuntimeError) */
  WAT_ERR_CUDA = -2,      /* CUDA runtime or driver error, or no device */
  WAT_ERR_STATE = -3,     /* wrong call order: weights missing, not finalized */
  WAT_ERR_NOMEM = -4
};

enum wat_precision {
  WAT_FP32 = 0,           /* every contraction in fp32 SIMT FMA: the correctness mode (logits within 1e-3) */
  WAT_BF16 = 1            /* bf16 operands, fp32 accumulate on tcgen05 tensor cores; fp32 residual stream */
};

typedef struct wat_config {
  int32_t n_mels;         /* 80 or 128 */
  int32_t n_audio_ctx;    /* 1500 */
  int32_t n_audio_state;  /* d */
  int32_t n_audio_head;   /* d / 64 */
  int32_t n_audio_layer;  /* L */
  int32_t at_low_compute; /* 0: tl_tr_1_8, 1: tl_down_tr_512_1_8  (model.py:243-246) */
  int32_t at_dim;         /* 512 when at_low_compute, ignored otherwise */
  int32_t n_class;        /* 527 */
  int32_t precision;      /* enum wat_precision */
  int32_t max_batch;      /* clips processed per internal chunk (workspace is sized for this); 0 = default */
} wat_config;

/* Head variants of the TL-TR training recipe (src/whisper_at_train/models.py:49-200).  LW_TR / LW_DOWN_TR are the
 * package's tl_tr / tl_down_tr heads (what wat_create builds); the others are the paper's baselines. */
enum wat_head_mode {
  WAT_HEAD_LW_TR = 0,        /* 'lw_tr_{t}_{l}'            time transformer per layer, layer transformer, mlp      models.py:178-187 */
  WAT_HEAD_LW_DOWN_TR = 1,   /* 'lw_down_tr_{dim}_{t}_{l}' + LN/linear down-projection first                      models.py:190-200 */
  WAT_HEAD_MEAN_MLP = 2,     /* 'mean_mlp'   mean over layers and time, mlp                                        models.py:113-117 */
  WAT_HEAD_LAST_MLP = 3,     /* 'last_mlp'   last layer, mean over time, mlp                                       models.py:120-124 */
  WAT_HEAD_WA_MLP = 4,       /* 'wa_mlp'     mean over time, learned layer weights / their sum, mlp                models.py:127-132 */
  WAT_HEAD_MEAN_TR = 5,      /* 'mean_tr_{h}' mean over layers, time transformer, mean over time, mlp              models.py:135-140 */
  WAT_HEAD_LAST_TR = 6,      /* 'last_tr_{h}'                                                                      models.py:143-148 */
  WAT_HEAD_WA_TR = 7,        /* 'wa_tr_{h}'                                                                        models.py:151-157 */
  WAT_HEAD_WA_DOWN_TR = 8    /* 'wa_down_tr_{dim}_{h}'                                                             models.py:160-167 */
};

typedef struct wat_head_config {
  int32_t rep_dim;        /* d of the pooled encoder states (multiple of 128) */
  int32_t n_layer;        /* L */
  int32_t inter_dim;      /* transformer width of the *_down_* modes (multiple of 128); ignored otherwise */
  int32_t n_class;        /* label_dim */
  int32_t mode;           /* enum wat_head_mode */
  int32_t n_time_head;    /* heads of time_tr  (the {t} / {h} of the mode string) */
  int32_t n_layer_head;   /* heads of layer_tr (the {l} of the mode string); ignored by the baselines */
  int32_t precision;      /* enum wat_precision */
  int32_t max_batch;      /* clips per internal chunk; 0 = default */
} wat_head_config;

WAT_API int wat_abi_version(void);
WAT_API const char* wat_last_error(void);

/* Model lifetime.  Weights are passed by their reference state_dict key ("encoder.blocks.0.attn.query.weight",
 * "at_model.time_tr.mlp.0.bias", ...), as contiguous fp32 HOST arrays; decoder.* keys are accepted and ignored. */
WAT_API int wat_create(const wat_config* cfg, wat_handle** out);
WAT_API int wat_set_weight(wat_handle* h, const char* key, const float* host_data, int64_t numel);
WAT_API int wat_finalize(wat_handle* h);   /* checks every tensor of the tagging path is present, packs for the device */
WAT_API int wat_destroy(wat_handle* h);

/* log-mel of B clips.  pcm: device fp32, clip c at pcm + c*clip_stride; n_valid (HOST int32[B] or NULL = all
 * n_samples) real samples per clip, `n_pad` zero samples appended (the reference's `padding`); the first
 * n_frames frames are written to mel_out [B, n_mels, n_frames] fp32 (the reference's layout).  The `max - 8`
 * clamp spans the whole padded signal of a clip (clamp_scope 0: one call of the reference per clip) or the
 * whole batch (clamp_scope 1: what the reference computes when handed a 2-D tensor, audio.py:155).
 * Needs no weights: valid on a handle that has not been finalized. */
WAT_API int wat_logmel(wat_handle* h, const float* pcm, int64_t clip_stride, const int32_t* n_valid, int32_t n_samples,
               int32_t n_pad, int32_t B, int32_t n_frames, int32_t clamp_scope, float* mel_out, void* stream);

/* encoder: mel [B, n_mels, 3000] fp32 device -> pooled_out [B, L, 75, d] fp32 (every clip, not only clip 0);
 * x_out (may be NULL) [B, 1500, d] fp32 = ln_post(x). */
WAT_API int wat_encoder(wat_handle* h, const float* mel, int32_t B, float* pooled_out, float* x_out, void* stream);

/* TL-TR head on pooled[:, :, t_start:t_start+t_len, :] of a [B, L, t_total, d] fp32 device tensor with decision
 * window dw = int(at_time_res * 2.5); logits_out [B, ceil(t_len/dw), n_class] fp32 device.
 * Limit: 1 <= dw <= 128 (at_time_res <= 51.2 s); larger windows return WAT_ERR_INVALID (the reference would zero-pad one
 * window of dw rows, model.py:360-364, but a 30 s segment only has 75 pooled rows). */
WAT_API int wat_tltr(wat_handle* h, const float* pooled, int32_t B, int32_t t_total, int32_t t_start, int32_t t_len,
             int32_t dw, float* logits_out, void* stream);

/* A head-only handle for any wat_head_mode: weights by the TLTR module's own state_dict keys ("time_tr.attn.query.weight",
 * "mlp_layer.1.bias", "layer_weight", ...; an "at_model." or "module." prefix is accepted).  Usable with wat_set_weight /
 * wat_finalize / wat_tltr / wat_head_forward / wat_destroy only. */
WAT_API int wat_head_create(const wat_head_config* cfg, wat_handle** out);
/* TLTR.forward (models.py:108-200): audio_rep [B, n_layer, Tp, rep_dim] fp32 device, Tp <= 128 -> logits_out [B, n_class] */
WAT_API int wat_head_forward(wat_handle* h, const float* audio_rep, int32_t B, int32_t Tp, float* logits_out, void* stream);

/* fused mel -> encoder -> head for B clips of <= 480000 samples; logits_out [B, ceil(75/dw), n_class] device */
WAT_API int wat_tag(wat_handle* h, const float* pcm, int64_t clip_stride, const int32_t* n_valid, int32_t n_samples,
            int32_t B, int32_t dw, float* logits_out, void* stream);

/* same with HOST buffers (pinned or pageable): copies in, computes, copies logits out, synchronises */
WAT_API int wat_tag_host(wat_handle* h, const float* pcm_host, int64_t clip_stride, const int32_t* n_valid, int32_t n_samples,
                 int32_t B, int32_t dw, float* logits_host);

/* the same two entry points for 16-bit PCM (what load_audio decodes, audio.py:63: float = int16 / 32768): the samples
 * are scaled inside the mel kernel, so a clip moves 0.96 MB instead of 1.92 MB over PCIe / HBM. Results are bit-identical
 * to the fp32 entry points fed with int16/32768. */
WAT_API int wat_tag_pcm16(wat_handle* h, const int16_t* pcm, int64_t clip_stride, const int32_t* n_valid, int32_t n_samples,
                          int32_t B, int32_t dw, float* logits_out, void* stream);
WAT_API int wat_tag_host_pcm16(wat_handle* h, const int16_t* pcm_host, int64_t clip_stride, const int32_t* n_valid,
                               int32_t n_samples, int32_t B, int32_t dw, float* logits_host);

/* Pipelined host entry points: wat_tag_host == submit + wait.  submit returns as soon as the work is queued (H2D of the PCM on a
 * copy stream into one of two device stages, compute on the handle's stream, D2H of the logits behind it); wait blocks until
 * that call's logits are in logits_host.  With  submit(t+1); wait(t);  in a loop the PCM of the next call crosses PCIe while
 * the current one computes.  pcm_host and logits_host must stay valid (and should be pinned) until the ticket has been waited
 * for; a third submit first waits for the call two tickets back.  Tickets are per handle, start at 1, and waiting twice is
 * harmless.  (This is the double buffering the reference's DataLoader + .to(device) would do for a GPU run; transcribe.py:128.) */
WAT_API int wat_tag_host_submit(wat_handle* h, const float* pcm_host, int64_t clip_stride, const int32_t* n_valid,
                                int32_t n_samples, int32_t B, int32_t dw, float* logits_host, int64_t* ticket);
WAT_API int wat_tag_host_submit_pcm16(wat_handle* h, const int16_t* pcm_host, int64_t clip_stride, const int32_t* n_valid,
                                      int32_t n_samples, int32_t B, int32_t dw, float* logits_host, int64_t* ticket);
WAT_API int wat_tag_host_wait(wat_handle* h, int64_t ticket);

/* introspection */
WAT_API int64_t wat_workspace_bytes(const wat_handle* h);
WAT_API int64_t wat_kernel_launches(const wat_handle* h);   /* kernels launched by this handle since creation */
WAT_API int wat_num_sms(const wat_handle* h);

/* per-kernel-class device timing: when enabled every kernel launch of the handle is bracketed by CUDA events on
 * the launching stream; wat_profile_read synchronises, sums the elapsed ms and launch counts per class
 * (arrays of wat_profile_classes() entries) and clears the records. */
WAT_API int wat_profile(wat_handle* h, int32_t enable);
WAT_API int wat_profile_classes(void);
WAT_API const char* wat_profile_class_name(int32_t i);
WAT_API int wat_profile_read(wat_handle* h, double* ms, int64_t* launches);

/* unit-test hooks for single kernels (device pointers; fp32 in/out, converted internally when tc != 0) */
WAT_API int wat_dbg_gemm(const float* A, const float* W, const float* bias, const float* R, float* C, int32_t M, int32_t N,
                 int32_t K, int32_t act, int32_t tc, void* stream);
/* the bf16-output epilogues as the encoder launches them (A [M,K], W [N,K], C bf16 device; bias fp32):
 * seq_T == 0: C [M,N] = act(A W^T + bias) (act 1 = exact GELU, the fc1 kernel); seq_T > 0: fused-QKV split, N = 3D:
 * C [M,2D] = q|k and vt [M/seq_T, n_head, 64, seq_Tpad] = V transposed per head (keys >= seq_T untouched) */
WAT_API int wat_dbg_gemm_bf16(const void* A, const void* W, const float* bias, void* C, void* vt, int32_t M, int32_t N, int32_t K,
                      int32_t act, int32_t seq_T, int32_t seq_Tpad, int32_t n_head, void* stream);
/* a LayerNorm folded into the GEMM that consumes it, chained as the bf16 encoder does (see DESIGN.md "LayerNorm folding"):
 * x = R + A1 W1^T + bias1 (R / x fp32, or fp16 when x_f16 - the bf16 encoder's residual stream; the epilogue also leaves
 * xb = bf16(x), per-slice row statistics; pooled = 20-row means of xb),
 * then out = act(LN(x; gamma, beta) W2^T + bias2) as rstd (xb W2'^T - mean colsum) + bias2'.  All pointers device. */
WAT_API int wat_dbg_ln_slices(int32_t M, int32_t D, int32_t K1, int32_t pair);
WAT_API int wat_dbg_ln_gemm(const void* A1, const void* W1, const float* bias1, const void* R, void* x, void* xb, float* stats,
                    float* pooled, const float* W2, const float* gamma, const float* beta, const float* bias2, void* out, int32_t M,
                    int32_t D, int32_t K1, int32_t N2, int32_t act, int32_t pair, int32_t x_f16, void* stream);
/* x [B*T, D] fp32, wqkv [3D, D], bqkv [3D] -> out [B*T, D]: fused-QKV GEMM + encoder self-attention (hd 64) */
WAT_API int wat_dbg_attention(const float* x, const float* wqkv, const float* bqkv, float* out, int32_t B, int32_t T,
                      int32_t n_head, int32_t tc, void* stream);
/* tc: 0 fp32 SIMT kernels, 1 tcgen05 kernels (max-free first pass + running-max fallback), 3 same with a clock trace on stderr,
 * 4 tcgen05 with the running-max pass only.  wat_dbg_attention_repeats(): query tiles of the last call that fell back. */
WAT_API int wat_dbg_attention_repeats(void);
WAT_API int wat_dbg_tma_overlap_probe(void);   /* 1 if the driver accepts a tensor map whose row stride < row length */

#ifdef __cplusplus
}
#endif
#endif /* WAT_H_ */
