// Internal launcher interface between the C-ABI layer (wat_api.cu) and the kernel translation units.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace wat {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE setting: a process that drives several GPUs (two handles
// on two devices) must opt in on each of them.  `done_mask` is the launcher's own bit set of configured devices.
template <typename K>
inline cudaError_t opt_in_smem(K kernel, int bytes, unsigned long long& done_mask) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (done_mask & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done_mask |= bit;
  return e;
}

// ---------------------------------------------------------------------------------- mel.cu
struct MelTables {
  int n_mels = 0;
  double2* twiddle = nullptr;   // [400]
  float* window = nullptr;      // [400]
  int* fb_start = nullptr;      // [n_mels]
  int* fb_off = nullptr;        // [n_mels + 1]
  float* fb_w = nullptr;        // [nnz]
};
cudaError_t launch_mel_power(const MelTables& tb, const void* pcm, bool pcm_i16, long long clip_stride, const int* n_valid_arr,
                             int n_valid_all, int n_pad, int B, int n_frames, int n_store, int frames_alloc,
                             float* logspec, float* clip_max, cudaStream_t st);
cudaError_t launch_share_max(float* clip_max, int B, cudaStream_t st);
cudaError_t launch_mel_norm(const float* logspec, const float* clip_max, int B, int n_store, int frames_alloc,
                            int n_mels, int mode, void* out, cudaStream_t st);
cudaError_t launch_mel_to_timemajor(const float* mel, int B, int T, int n_mels, int mode, void* out, cudaStream_t st);

// ---------------------------------------------------------------------------------- simt.cu (fp32 mode + glue)
// C[m, :] = R[m or m % r_mod, :] + act(A[m, :] . W^T + bias)
struct GemmF32 {
  const float* A; long long lda;
  const float* W;                 // [N, K] row-major
  const float* bias;              // [N] or null
  float* C; long long ldc;
  const float* R; long long ldr; int r_mod;   // null | residual (r_mod 0) | table of r_mod rows (positional embedding)
  int M, N, K, act;               // act: 0 none, 1 exact GELU
};
cudaError_t launch_gemm_f32(const GemmF32& g, cudaStream_t st);

// y = LN(x) * gamma + beta over the last dim; out fp32 or bf16
cudaError_t launch_layernorm(const float* x, const float* gamma, const float* beta, int M, int D, void* out,
                             bool out_bf16, cudaStream_t st);

// LN of x [B*T, D] fused with the encoder's 20x average pool of the same rows into pooled[b, layer, :, :]
cudaError_t launch_layernorm_pool20(const float* x, const float* gamma, const float* beta, int B, int T, int D, void* out,
                                    bool out_bf16, float* pooled, int layer, int L, cudaStream_t st);

// encoder self-attention, fp32, head_dim 64: q/k/v rows at stride ld (same row indexing), n_seq sequences of T
cudaError_t launch_attn_f32_hd64(const float* q, const float* k, const float* v, long long ld, float* out,
                                 long long ldo, int n_seq, int T, int n_head, cudaStream_t st);

// short-sequence attention (TL-TR head): n_seq sequences of T <= 128 tokens, n_head heads of hd dims.
// qkv is [n_seq*T, 3*D] (q | k | v); in/out fp32 or bf16.
cudaError_t launch_attn_small(const void* qkv, bool in_bf16, void* out, bool out_bf16, int n_seq, int T, int n_head,
                              int hd, cudaStream_t st);

// mean over groups of `win` consecutive rows: x [n_groups*win, D] (row stride D) -> out row g at out + g*out_stride
// optional xb [n_groups, D] bf16 + stats [n_groups, 1, 2]: the row's bf16 copy and (sum, sum of squares) for a folded LayerNorm
// x / out rows are fp32 or fp16 (in_f16 / out_f16; fp16 = the bf16 mode's residual stream)
cudaError_t launch_group_mean(const void* x, bool in_f16, int n_groups, int win, int D, void* out, bool out_f16, long long out_stride,
                              cudaStream_t st, __nv_bfloat16* xb = nullptr, float* stats = nullptr);
cudaError_t launch_layernorm_f16in(const void* x, const float* gamma, const float* beta, int M, int D, float* out, cudaStream_t st);
// encoder pooling: x [B, 1500, D] -> pooled[b, layer, 0..74, :] with pooled laid out [B, L, 75, D]
cudaError_t launch_pool20(const float* x, int B, int T, int D, int layer, int L, float* pooled, cudaStream_t st);
// the same from the bf16 copy of the residual stream the fc2 epilogue leaves (bf16 mode): half the bytes; rows summed in fp32
// in a fixed order, so a clip's pooled state does not depend on its position in the batch
cudaError_t launch_pool20_bf16(const __nv_bfloat16* xb, int B, int T, int D, int layer, int L, float* pooled, cudaStream_t st);

// k=3, pad=1 im2col over time-major activations: src [B, Tin, C] -> out [B*Tout, 3C], row (b, j) =
// (src[b, j*stride-1], src[b, j*stride], src[b, j*stride+1]) with zero rows outside [0, Tin)
cudaError_t launch_im2col_k3(const void* src, bool bf16, int B, int Tin, int C, int stride, int Tout, void* out,
                             cudaStream_t st);
// fp32 -> bf16 (weights packing)
cudaError_t launch_f32_to_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t st);

// TL-TR window regroup (model.py:360-367): pooled [B, L, Tp, D] -> rows (b, s, l, tau), zero rows past Tp
// baseline heads: reduce the layer axis first (kind 0 = mean, 1 = last layer, 2 = weights w[L] / sum(w)); out rows = (b*S + s)*dw + tau
cudaError_t launch_head_layer_reduce(const float* pooled, int B, int L, int Tp_total, int t_start, int Tp, int dw, int S, int D,
                                     int kind, const float* w, void* out, bool out_f16, cudaStream_t st, __nv_bfloat16* xb = nullptr,
                                     float* stats = nullptr);
cudaError_t launch_head_gather(const float* pooled, int B, int L, int Tp_total, int t_start, int Tp, int dw, int S,
                               int D, void* out, bool out_f16, cudaStream_t st, __nv_bfloat16* xb = nullptr, float* stats = nullptr);
// W' = bf16(W diag(gamma)) [N, K], colsum [N], bias_out [N] = bias + W beta  (LayerNorm folded into the GEMM that consumes it)
cudaError_t launch_fold_ln_weights(const float* W, const float* gamma, const float* beta, const float* bias, int N, int K,
                                   __nv_bfloat16* Wout, float* colsum, float* bias_out, cudaStream_t st);

// ---------------------------------------------------------------------------------- gemm_tc.cu (bf16 tcgen05)
enum TcEpi : int {
  TC_EPI_BF16 = 0,        // out bf16 = act(acc + bias)
  TC_EPI_F32_RES = 1,     // out fp32 = R + act(acc + bias)      (R may alias out; r_mod > 0: table of r_mod rows)
  TC_EPI_QKV = 2,         // cols [0, 2D) -> bf16 [M, 2D] (q | k), cols [2D, 3D) -> V^T [B, H, 64, Tpad]
  TC_EPI_F32 = 3,         // out fp32 = act(acc + bias)
};
struct GemmTc {
  const __nv_bfloat16* A; long long lda;                    // [M, K] bf16, row stride lda elements
  const __nv_bfloat16* W;                                   // [N, K] row-major bf16
  const float* bias;
  void* C; long long ldc;
  const float* R; long long ldr; int r_mod;
  int M, N, K, act, epi;
  // TC_EPI_QKV only:
  __nv_bfloat16* vt; int seq_T; int seq_Tpad; int n_head;  // flat row r -> (b = r / seq_T, t = r % seq_T)
  float q_scale;          // TC_EPI_QKV: factor applied to the q columns [0, D) before rounding (0 = none)
  int force_pair;         // 0: default kernel choice, 1: CTA-pair kernel, -1: single-CTA kernel (tests)
  long long* trace;       // optional device buffer for a clock trace (tests)
  // ---- LayerNorm folding (bf16 mode).  PRODUCER side, fp32 epilogues only (all optional): while the output rows are in
  // registers the epilogue also writes
  //   xb    [M, ldxb] bf16 copy of the rows (the next GEMM's A operand, un-normalised),
  //   stats [M, stats_np, 2] fp32 (sum, sum of squares) per column slice; stats_np = gemm_tc_stats_slices(M, N, K, epi, force_pair).
  __nv_bfloat16* xb; long long ldxb;
  float* stats; int stats_np;
  // CONSUMER side: A holds un-normalised rows; out = rstd (A W'^T - mean colsum) + bias with the rows' mean / rstd from
  // ln_stats [M, ln_np, 2] and W', colsum, bias prepared by launch_fold_ln_weights
  const float* ln_stats; int ln_np; const float* ln_colsum;
  // fp32 epilogues: C is stored / R is read as fp16 with the same leading dimensions in elements (bf16 mode keeps the residual
  // stream in fp16, the dtype the reference's own GPU path uses for it; the arithmetic stays fp32)
  int c_f16, r_f16;
};
cudaError_t launch_gemm_tc(const GemmTc& g, int num_sms, cudaStream_t st);
int gemm_tc_stats_slices(int M, int N, int K, int epi, int force_pair);

// ---------------------------------------------------------------------------------- attn_tc.cu
// V^T buffer: [B, H, VT_ROWS, Tpad] bf16 = V transposed per head (written by the QKV epilogue; K-major B operand of the
// PV MMA).  Keys T..Tpad-1 are never written and must stay zero (their P is exactly 0, but 0 * garbage could be NaN):
// the buffer is zeroed when it is allocated.
constexpr int VT_ROWS = 64;
// qk: [B*T, 2D] bf16 (q | k), vt as above, out: [B*T, D] bf16
// q_prescaled: q already carries AT_QSCALE = 64^-0.5 log2(e) (GemmTc::q_scale), which enables the max-free first pass;
// n_repeat: optional device counter of tiles that fell back to the running-max pass
constexpr float AT_QSCALE = 0.125f * 1.4426950408889634f;
cudaError_t launch_attn_tc(const __nv_bfloat16* qk, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int T,
                           int Tpad, int n_head, bool q_prescaled, cudaStream_t st, long long* trace = nullptr,
                           unsigned int* n_repeat = nullptr);

// driver entry point for cuTensorMapEncodeTiled, resolved once through the runtime
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

}  // namespace wat
