// C ABI of libwat (include/wat.h): handle, weight registry/packing, workspace and the orchestration of the
// tagging path  log-mel -> conv stem -> L residual attention blocks (+20x pooling of every layer) -> TL-TR head.
// Reference call chain: transcribe.py:127,241-263 -> model.py:156-177 (AudioEncoder.forward) -> model.py:351-379.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/wat.h"
#include "kernels.h"

using namespace wat;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(expr)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (expr);                                                                         \
    if (e__ != cudaSuccess) return fail(WAT_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// a kernel launch through one of the launchers: counted, error-checked, optionally bracketed by CUDA events
#define KL(h, expr)                                                                                   \
  do {                                                                                                \
    const int pi__ = (h)->profiling ? prof_begin((h), #expr) : ((h)->prof_override = -1);                                    \
    cudaError_t e__ = (expr);                                                                         \
    (h)->launches++;                                                                                  \
    if (pi__ >= 0) prof_end((h), pi__);                                                               \
    if (e__ != cudaSuccess) return fail(WAT_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

enum ProfClass { PC_MEL = 0, PC_LAYOUT, PC_GEMM, PC_ATTN, PC_LN, PC_POOL, PC_HEAD_ATTN, PC_MEAN,
                 PC_GEMM_QKV, PC_GEMM_OUT, PC_GEMM_FC1, PC_GEMM_FC2, PC_GEMM_HEAD, PC_COUNT };
// "gemm" = conv stem GEMMs; "gemm_head" = every GEMM of the TL-TR head incl. the classifier; the four encoder-block GEMMs have
// their own classes
const char* const kProfNames[PC_COUNT] = {"mel", "layout", "gemm", "attention", "layernorm", "pool", "head_attention", "mean",
                                          "gemm_qkv", "gemm_out", "gemm_fc1", "gemm_fc2", "gemm_head"};

struct ProfRec { int cls; cudaEvent_t a, b; };

// Every entry point runs on the handle's device and puts the caller's current device back on exit (a ctypes caller that
// drives several GPUs from one thread must not have its device switched under it).
struct DeviceGuard {
  int prev = -1, dev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int d) : dev(d) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};
#define ON_DEVICE(h)                                                                                  \
  DeviceGuard guard__((h)->device);                                                                   \
  if (guard__.err != cudaSuccess) return fail(WAT_ERR_CUDA, "cudaSetDevice(%d): %s", (h)->device, cudaGetErrorString(guard__.err))

struct Slot {            // where a state_dict tensor lands on the device
  float* dst = nullptr;
  int64_t numel = 0;
  int conv_c = 0;        // > 0: conv weight [out, c, 3] -> stored [out, 3, c]
  bool set = false;
  bool optional = false;
};

struct BlockW {
  int D = 0, H = 0;
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *bqkv, *bo, *b1, *b2;
  float *wqkv, *wo, *w1, *w2;                       // fp32 [3D,D] [D,D] [4D,D] [D,4D]
  __nv_bfloat16 *wqkv_h = nullptr, *wo_h = nullptr, *w1_h = nullptr, *w2_h = nullptr;
  // bf16 mode: the two LayerNorms are folded into the GEMMs that consume them (gemm_tc.cu): wqkv_h = bf16(Wqkv diag(ln1_g)),
  // w1_h = bf16(W1 diag(ln2_g)), with their column sums and the biases that absorb W beta
  float *cs_qkv = nullptr, *bqkv_f = nullptr, *cs_1 = nullptr, *b1_f = nullptr;
};

struct Buf {
  void* p = nullptr;
  size_t bytes = 0;
};

}  // namespace

struct wat_handle {
  wat_config cfg;
  int device = 0, num_sms = 148;
  bool finalized = false;
  bool bf16 = false;
  int d = 0, H = 0, L = 0, di = 0;
  int64_t launches = 0;
  std::map<std::string, Slot> slots;
  std::vector<void*> owned;                          // every cudaMalloc of weights/tables
  MelTables mel;
  float *conv1_w, *conv1_b, *conv2_w, *conv2_b, *pos, *lnp_g, *lnp_b;
  __nv_bfloat16 *conv1_w_h = nullptr, *conv2_w_h = nullptr;
  std::vector<BlockW> enc;
  BlockW time_tr, layer_tr;
  bool head_only = false;                            // wat_head_create: no mel tables, no encoder
  int head_mode = WAT_HEAD_LW_TR;
  float* layer_w = nullptr;                          // 'wa_*' modes: learned layer weights [L]
  float *down_g = nullptr, *down_b = nullptr, *down_w = nullptr, *down_bias = nullptr, *cls_g, *cls_b, *cls_w, *cls_bias;
  __nv_bfloat16* down_w_h = nullptr;
  // workspace (grow-only)
  int ws_B = 0;
  int64_t rows_cap = 0;
  int head_chunk = 1;
  Buf x, x2, xn, qkv, vt, att, hbuf, logspec, clipmax, melT, pooled, lmean, lnout, nvalid, logits, stats;
  Buf pcm_stage[2];                                              // host entry points: two PCM stages, see tag_host_submit
  cudaStream_t own_stream = nullptr;
  // the workspace is shared by every call on the handle: when a call arrives on another stream than the previous one (e.g.
  // wat_tag on the caller's stream, then wat_tag_host on the handle's own stream) it is ordered after the previous call's work
  cudaStream_t last_stream = nullptr;
  bool last_stream_valid = false;
  cudaEvent_t order_ev = nullptr;
  cudaStream_t copy_stream = nullptr;                            // wat_tag_host*: H2D of PCM pieces, overlapped with the mel kernel
  cudaEvent_t piece_ev[2][8] = {};                               // ... per stage: piece p of the PCM has arrived
  cudaEvent_t done_ev[2] = {};                                   // ... per stage: the call's logits are in the caller's host buffer
  int64_t next_ticket = 1;                                       // ticket t uses stage t & 1; at most two calls in flight
  int64_t slot_ticket[2] = {0, 0};                               // the call that owns the stage, 0 = finished / none
  int64_t ws_bytes = 0;
  // per-kernel-class timing (wat_profile)
  bool profiling = false;
  cudaStream_t cur_stream = nullptr;
  std::vector<ProfRec> prof;
  size_t prof_used = 0;
  int prof_override = -1;                            // class of the next launch (encoder GEMM kinds)
  bool in_head = false;                              // run_head is executing: its GEMMs are profiled as "gemm_head"
};

namespace {

int prof_class(const char* s) {
  if (strstr(s, "gemm")) return PC_GEMM;
  if (strstr(s, "attn_small")) return PC_HEAD_ATTN;
  if (strstr(s, "attn")) return PC_ATTN;
  if (strstr(s, "layernorm")) return PC_LN;       // includes the fused LN + pool of layers 0..L-2
  if (strstr(s, "pool20")) return PC_POOL;
  if (strstr(s, "group_mean")) return PC_MEAN;
  if (strstr(s, "mel_power") || strstr(s, "mel_norm") || strstr(s, "share_max")) return PC_MEL;
  return PC_LAYOUT;
}

int prof_begin(wat_handle* h, const char* what) {
  if (h->prof_used == h->prof.size()) {
    ProfRec r;
    r.cls = 0;
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
    h->prof.push_back(r);
  }
  const int i = (int)h->prof_used++;
  h->prof[i].cls = h->prof_override >= 0 ? h->prof_override : prof_class(what);
  if (h->in_head && h->prof[i].cls == PC_GEMM) h->prof[i].cls = PC_GEMM_HEAD;
  h->prof_override = -1;
  cudaEventRecord(h->prof[i].a, h->cur_stream);
  return i;
}

void prof_end(wat_handle* h, int i) { cudaEventRecord(h->prof[i].b, h->cur_stream); }

template <typename T>
int dalloc(wat_handle* h, T** p, int64_t n) {
  void* q = nullptr;
  CU(cudaMalloc(&q, sizeof(T) * (size_t)(n > 0 ? n : 1)));
  h->owned.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}

int grow(wat_handle* h, Buf& b, size_t bytes, bool zero = false) {
  if (b.bytes >= bytes) return 0;
  if (b.p) { CU(cudaFree(b.p)); h->ws_bytes -= (int64_t)b.bytes; b.p = nullptr; b.bytes = 0; }
  CU(cudaMalloc(&b.p, bytes));
  b.bytes = bytes;
  h->ws_bytes += (int64_t)bytes;
  if (zero) { CU(cudaMemset(b.p, 0, bytes)); CU(cudaDeviceSynchronize()); }   // one-off; later work may run on any (non-blocking) stream
  return 0;
}

int add_slot(wat_handle* h, const std::string& key, float** dst, int64_t numel, int conv_c = 0, bool optional = false) {
  int rc = dalloc(h, dst, numel);
  if (rc) return rc;
  Slot s;
  s.dst = *dst; s.numel = numel; s.conv_c = conv_c; s.optional = optional;
  h->slots[key] = s;
  return 0;
}

void alias_slot(wat_handle* h, const std::string& key, float* dst, int64_t numel) {
  Slot s;
  s.dst = dst; s.numel = numel;
  h->slots[key] = s;
}

int make_block(wat_handle* h, BlockW& b, const std::string& p, int D, int H) {
  b.D = D; b.H = H;
  int rc = 0;
  if ((rc = dalloc(h, &b.wqkv, (int64_t)3 * D * D))) return rc;
  if ((rc = dalloc(h, &b.bqkv, (int64_t)3 * D))) return rc;
  CU(cudaMemset(b.bqkv, 0, sizeof(float) * 3 * D));                    // key has no bias (model.py:66)
  alias_slot(h, p + ".attn.query.weight", b.wqkv, (int64_t)D * D);
  alias_slot(h, p + ".attn.key.weight", b.wqkv + (int64_t)D * D, (int64_t)D * D);
  alias_slot(h, p + ".attn.value.weight", b.wqkv + (int64_t)2 * D * D, (int64_t)D * D);
  alias_slot(h, p + ".attn.query.bias", b.bqkv, D);
  alias_slot(h, p + ".attn.value.bias", b.bqkv + 2 * D, D);
  if ((rc = add_slot(h, p + ".attn.out.weight", &b.wo, (int64_t)D * D))) return rc;
  if ((rc = add_slot(h, p + ".attn.out.bias", &b.bo, D))) return rc;
  if ((rc = add_slot(h, p + ".attn_ln.weight", &b.ln1_g, D))) return rc;
  if ((rc = add_slot(h, p + ".attn_ln.bias", &b.ln1_b, D))) return rc;
  if ((rc = add_slot(h, p + ".mlp.0.weight", &b.w1, (int64_t)4 * D * D))) return rc;
  if ((rc = add_slot(h, p + ".mlp.0.bias", &b.b1, (int64_t)4 * D))) return rc;
  if ((rc = add_slot(h, p + ".mlp.2.weight", &b.w2, (int64_t)4 * D * D))) return rc;
  if ((rc = add_slot(h, p + ".mlp.2.bias", &b.b2, D))) return rc;
  if ((rc = add_slot(h, p + ".mlp_ln.weight", &b.ln2_g, D))) return rc;
  if ((rc = add_slot(h, p + ".mlp_ln.bias", &b.ln2_b, D))) return rc;
  return 0;
}

bool head_has_down(int m) { return m == WAT_HEAD_LW_DOWN_TR || m == WAT_HEAD_WA_DOWN_TR; }
bool head_has_time_tr(int m) { return m != WAT_HEAD_MEAN_MLP && m != WAT_HEAD_LAST_MLP && m != WAT_HEAD_WA_MLP; }
bool head_has_layer_tr(int m) { return m == WAT_HEAD_LW_TR || m == WAT_HEAD_LW_DOWN_TR; }
bool head_has_layer_w(int m) { return m == WAT_HEAD_WA_MLP || m == WAT_HEAD_WA_TR || m == WAT_HEAD_WA_DOWN_TR; }

// parameters of the head (ATModel.__init__ model.py:323-349; TLTR.__init__ src/whisper_at_train/models.py:49-106)
int add_head_slots(wat_handle* h, int n_time_head, int n_layer_head) {
  const int d = h->d, di = h->di, m = h->head_mode;
  int rc;
  if (head_has_down(m)) {
    if ((rc = add_slot(h, "at_model.down_layer.0.weight", &h->down_g, d))) return rc;
    if ((rc = add_slot(h, "at_model.down_layer.0.bias", &h->down_b, d))) return rc;
    if ((rc = add_slot(h, "at_model.down_layer.1.weight", &h->down_w, (int64_t)di * d))) return rc;
    if ((rc = add_slot(h, "at_model.down_layer.1.bias", &h->down_bias, di))) return rc;
  }
  if (head_has_layer_w(m) && (rc = add_slot(h, "at_model.layer_weight", &h->layer_w, h->L))) return rc;
  if (head_has_time_tr(m) && (rc = make_block(h, h->time_tr, "at_model.time_tr", di, n_time_head))) return rc;
  if (head_has_layer_tr(m) && (rc = make_block(h, h->layer_tr, "at_model.layer_tr", di, n_layer_head))) return rc;
  if ((rc = add_slot(h, "at_model.mlp_layer.0.weight", &h->cls_g, di))) return rc;
  if ((rc = add_slot(h, "at_model.mlp_layer.0.bias", &h->cls_b, di))) return rc;
  if ((rc = add_slot(h, "at_model.mlp_layer.1.weight", &h->cls_w, (int64_t)h->cfg.n_class * di))) return rc;
  if ((rc = add_slot(h, "at_model.mlp_layer.1.bias", &h->cls_bias, h->cfg.n_class))) return rc;
  return 0;
}

// slaney mel filterbank, restating librosa.filters.mel(sr=16000, n_fft=400, n_mels) (audio.py:96-101)
double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

int build_mel_tables(wat_handle* h) {
  const int n_mels = h->cfg.n_mels;
  std::vector<double> mel_f(n_mels + 2);
  const double lo = hz_to_mel(0.0), hi = hz_to_mel(8000.0);
  for (int i = 0; i < n_mels + 2; ++i) {
    // np.linspace(lo, hi, n): start + i*step, last point exact
    const double step = (hi - lo) / (n_mels + 1);
    mel_f[i] = mel_to_hz(i == n_mels + 1 ? hi : lo + i * step);
  }
  std::vector<int> start(n_mels), off(n_mels + 1);
  std::vector<float> w;
  for (int m = 0; m < n_mels; ++m) {
    off[m] = (int)w.size();
    int first = -1;
    const double fd0 = mel_f[m + 1] - mel_f[m], fd1 = mel_f[m + 2] - mel_f[m + 1];
    const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
    std::vector<float> row(201);
    int last = -1;
    for (int k = 0; k <= 200; ++k) {
      const double fk = k * (16000.0 / 400.0);
      const double lower = -(mel_f[m] - fk) / fd0, upper = (mel_f[m + 2] - fk) / fd1;
      float v = (float)std::fmax(0.0, std::fmin(lower, upper));   // librosa stores the triangle in float32 ...
      v = (float)(v * enorm);                                      // ... then scales by enorm (float64 -> float32)
      row[k] = v;
      if (v != 0.f) { if (first < 0) first = k; last = k; }
    }
    if (first < 0) { first = 0; last = -1; }
    if (last >= 200) return fail(WAT_ERR_INVALID, "mel filter %d touches bin 200", m);
    start[m] = first;
    for (int k = first; k <= last; ++k) w.push_back(row[k]);
  }
  off[n_mels] = (int)w.size();
  std::vector<double2> tw(400);
  std::vector<float> win(400);
  const double PI = 3.14159265358979323846;
  for (int j = 0; j < 400; ++j) {
    tw[j].x = std::cos(2.0 * PI * j / 400.0);
    tw[j].y = std::sin(2.0 * PI * j / 400.0);
    win[j] = (float)(0.5 - 0.5 * std::cos(2.0 * PI * j / 400.0));
  }
  h->mel.n_mels = n_mels;
  int rc;
  if ((rc = dalloc(h, &h->mel.twiddle, 400))) return rc;
  if ((rc = dalloc(h, &h->mel.window, 400))) return rc;
  if ((rc = dalloc(h, &h->mel.fb_start, n_mels))) return rc;
  if ((rc = dalloc(h, &h->mel.fb_off, n_mels + 1))) return rc;
  if ((rc = dalloc(h, &h->mel.fb_w, (int64_t)w.size()))) return rc;
  CU(cudaMemcpy(h->mel.twiddle, tw.data(), sizeof(double2) * 400, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->mel.window, win.data(), sizeof(float) * 400, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->mel.fb_start, start.data(), sizeof(int) * n_mels, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->mel.fb_off, off.data(), sizeof(int) * (n_mels + 1), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->mel.fb_w, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice));
  return 0;
}

int to_bf16(wat_handle* h, const float* src, __nv_bfloat16** dst, int64_t n) {
  int rc = dalloc(h, dst, n);
  if (rc) return rc;
  KL(h, launch_f32_to_bf16(src, *dst, n, 0));
  return 0;
}

int pack_block_bf16(wat_handle* h, BlockW& b) {
  const int64_t D = b.D;
  int rc;
  if ((rc = dalloc(h, &b.wqkv_h, 3 * D * D)) || (rc = dalloc(h, &b.cs_qkv, 3 * D)) || (rc = dalloc(h, &b.bqkv_f, 3 * D))) return rc;
  KL(h, launch_fold_ln_weights(b.wqkv, b.ln1_g, b.ln1_b, b.bqkv, (int)(3 * D), (int)D, b.wqkv_h, b.cs_qkv, b.bqkv_f, 0));
  if ((rc = to_bf16(h, b.wo, &b.wo_h, D * D))) return rc;
  if ((rc = dalloc(h, &b.w1_h, 4 * D * D)) || (rc = dalloc(h, &b.cs_1, 4 * D)) || (rc = dalloc(h, &b.b1_f, 4 * D))) return rc;
  KL(h, launch_fold_ln_weights(b.w1, b.ln2_g, b.ln2_b, b.b1, (int)(4 * D), (int)D, b.w1_h, b.cs_1, b.b1_f, 0));
  if ((rc = to_bf16(h, b.w2, &b.w2_h, 4 * D * D))) return rc;
  return 0;
}

// row-proportional part of the workspace: the residual stream and every per-row intermediate of a transformer block
int ensure_rows(wat_handle* h, int64_t rows_cap, int Bc) {
  if (rows_cap <= h->rows_cap) return 0;
  const int64_t d = h->d;
  const size_t es = h->bf16 ? 2 : 4;
  int rc;
  if ((rc = grow(h, h->x, sizeof(float) * rows_cap * d))) return rc;
  if ((rc = grow(h, h->x2, sizeof(float) * rows_cap * d))) return rc;
  if ((rc = grow(h, h->xn, es * rows_cap * d))) return rc;
  if ((rc = grow(h, h->qkv, es * rows_cap * 3 * d))) return rc;
  if ((rc = grow(h, h->att, es * rows_cap * d))) return rc;
  // hbuf also holds the conv1 im2col rows [Bc*3000, 3*n_mels], which exceed rows*4d when d < 1.5 n_mels
  const int64_t hbuf_elems = std::max(rows_cap * 4 * d, h->head_only ? (int64_t)0 : (int64_t)Bc * 3000 * 3 * h->cfg.n_mels);
  if ((rc = grow(h, h->hbuf, es * hbuf_elems))) return rc;
  if (h->bf16) {                                                   // per-row (sum, sum of squares) slices for the folded LayerNorms
    const int64_t np_max = std::max<int64_t>(std::max(d, (int64_t)h->di) / 64, 1);
    if ((rc = grow(h, h->stats, sizeof(float) * 2 * rows_cap * np_max))) return rc;
  }
  h->rows_cap = rows_cap;
  return 0;
}

int ensure_ws(wat_handle* h, int B) {
  const int Bc = B < h->cfg.max_batch ? B : h->cfg.max_batch;
  if (Bc <= h->ws_B) return 0;
  const int64_t d = h->d, L = h->L;
  const int64_t rows_enc = h->head_only ? 0 : (int64_t)Bc * 1500;
  const int64_t head_rows_clip = L * 150;                         // S*dw < 75 + dw <= 150 for dw <= 75; Tp <= 128 in wat_head_forward
  // the encoder's rows, and room for the head rows of at least one clip (a head-only handle: of the whole batch).  run_head
  // grows this once it knows the decision window, so that the head of a whole batch runs as one chunk (see there).
  const int64_t rows_cap = std::max(rows_enc, (h->head_only ? (int64_t)Bc : (int64_t)1) * head_rows_clip);
  const size_t es = h->bf16 ? 2 : 4;
  int rc;
  if ((rc = ensure_rows(h, rows_cap, Bc))) return rc;
  if (!h->head_only && (rc = grow(h, h->hbuf, es * (size_t)Bc * 3000 * 3 * h->cfg.n_mels))) return rc;   // conv1 im2col rows (no-op when rows*4d covers them)
  if (h->bf16 && !h->head_only) {
    if ((rc = grow(h, h->vt, (size_t)2 * Bc * h->H * VT_ROWS * 1536, true))) return rc;   // zeroed: the 36 padding keys stay 0
  }
  if (!h->head_only) {
    if ((rc = grow(h, h->logspec, sizeof(float) * (size_t)Bc * 3000 * h->cfg.n_mels))) return rc;
    if ((rc = grow(h, h->clipmax, sizeof(float) * Bc))) return rc;
    if ((rc = grow(h, h->nvalid, sizeof(int) * Bc))) return rc;
    if ((rc = grow(h, h->melT, es * (size_t)Bc * 3000 * h->cfg.n_mels))) return rc;
    if ((rc = grow(h, h->pooled, sizeof(float) * (size_t)Bc * L * 75 * d))) return rc;
  }
  if ((rc = grow(h, h->lmean, sizeof(float) * (size_t)Bc * 75 * d))) return rc;
  if ((rc = grow(h, h->lnout, sizeof(float) * (size_t)Bc * 75 * d))) return rc;
  h->ws_B = Bc;
  h->head_chunk = Bc;
  return 0;
}

// one GEMM in the handle's precision.  out_mode: 0 = activation-typed output (bf16 in bf16 mode), 1 = fp32 (+R)
int gemm(wat_handle* h, const void* A, int64_t lda, const float* Wf, const __nv_bfloat16* Wh, const float* bias, void* C,
         int64_t ldc, const float* R, int64_t ldr, int r_mod, int M, int N, int K, int act, bool out_f32, cudaStream_t st) {
  if (!h->bf16) {
    GemmF32 g;
    g.A = (const float*)A; g.lda = lda; g.W = Wf; g.bias = bias; g.C = (float*)C; g.ldc = ldc;
    g.R = R; g.ldr = ldr; g.r_mod = r_mod; g.M = M; g.N = N; g.K = K; g.act = act;
    KL(h, launch_gemm_f32(g, st));
  } else {
    GemmTc g;
    memset(&g, 0, sizeof(g));
    g.A = (const __nv_bfloat16*)A; g.lda = lda; g.W = Wh; g.bias = bias; g.C = C; g.ldc = ldc;
    g.R = R; g.ldr = ldr; g.r_mod = r_mod; g.M = M; g.N = N; g.K = K; g.act = act;
    g.epi = out_f32 ? (R ? TC_EPI_F32_RES : TC_EPI_F32) : TC_EPI_BF16;
    KL(h, launch_gemm_tc(g, h->num_sms, st));
  }
  return 0;
}

// ResidualAttentionBlock.forward (model.py:128-139) on the fp32 residual stream x [n_seq*T, D]: fp32 mode (SIMT kernels)
int run_block_f32(wat_handle* h, const BlockW& w, float* x, int n_seq, int T, bool encoder, cudaStream_t st,
                  float* pooled = nullptr, int pool_layer = -1) {
  const int D = w.D, rows = n_seq * T;
  int rc;
  // the first LayerNorm also emits the 20x pooled state of the PREVIOUS layer's output (= this block's input)
  if (pooled && pool_layer >= 0) KL(h, launch_layernorm_pool20(x, w.ln1_g, w.ln1_b, n_seq, T, D, h->xn.p, false, pooled, pool_layer, h->L, st));
  else KL(h, launch_layernorm(x, w.ln1_g, w.ln1_b, rows, D, h->xn.p, false, st));
  if (encoder) h->prof_override = PC_GEMM_QKV;
  if ((rc = gemm(h, h->xn.p, D, w.wqkv, w.wqkv_h, w.bqkv, h->qkv.p, 3 * D, nullptr, 0, 0, rows, 3 * D, D, 0, false, st))) return rc;
  if (encoder) {
    const float* q = (const float*)h->qkv.p;
    KL(h, launch_attn_f32_hd64(q, q + D, q + 2 * D, 3 * D, (float*)h->att.p, D, n_seq, T, w.H, st));
  } else {
    KL(h, launch_attn_small(h->qkv.p, false, h->att.p, false, n_seq, T, w.H, D / w.H, st));
  }
  if (encoder) h->prof_override = PC_GEMM_OUT;
  if ((rc = gemm(h, h->att.p, D, w.wo, w.wo_h, w.bo, x, D, x, D, 0, rows, D, D, 0, true, st))) return rc;
  KL(h, launch_layernorm(x, w.ln2_g, w.ln2_b, rows, D, h->xn.p, false, st));
  if (encoder) h->prof_override = PC_GEMM_FC1;
  if ((rc = gemm(h, h->xn.p, D, w.w1, w.w1_h, w.b1, h->hbuf.p, 4 * D, nullptr, 0, 0, rows, 4 * D, D, 1, false, st))) return rc;
  if (encoder) h->prof_override = PC_GEMM_FC2;
  if ((rc = gemm(h, h->hbuf.p, 4 * D, w.w2, w.w2_h, w.b2, x, D, x, D, 0, rows, D, 4 * D, 0, true, st))) return rc;
  return 0;
}

// fp32-output GEMM of the bf16 path whose output rows feed a LayerNorm: besides C (= R + act(A W^T + bias)) the epilogue
// leaves the rows' bf16 copy in h->xn and their (sum, sum of squares) slices in h->stats; returns the slice count in *np
// The residual stream C is kept in fp16 in bf16 mode (the dtype of the reference's own GPU path; the arithmetic of the epilogue
// is fp32): r_f16 says whether R is that stream (true) or an fp32 table such as the positional embedding (false).
int gemm_producer(wat_handle* h, const void* A, int64_t lda, const __nv_bfloat16* W, const float* bias, void* C, const void* R,
                  bool r_f16, int64_t ldr, int r_mod, int M, int N, int K, int act, cudaStream_t st, int* np) {
  GemmTc g;
  memset(&g, 0, sizeof(g));
  g.A = (const __nv_bfloat16*)A; g.lda = lda; g.W = W; g.bias = bias; g.C = C; g.ldc = N; g.R = (const float*)R; g.ldr = ldr; g.r_mod = r_mod;
  g.M = M; g.N = N; g.K = K; g.act = act; g.epi = R ? TC_EPI_F32_RES : TC_EPI_F32;
  g.c_f16 = 1; g.r_f16 = r_f16 ? 1 : 0;
  g.xb = (__nv_bfloat16*)h->xn.p; g.ldxb = N;
  g.stats = (float*)h->stats.p; g.stats_np = gemm_tc_stats_slices(M, N, K, g.epi, 0);
  *np = g.stats_np;
  KL(h, launch_gemm_tc(g, h->num_sms, st));
  return 0;
}

// The same block in bf16 mode.  No LayerNorm kernel runs: x arrives with its bf16 copy in h->xn and `np` statistics slices in
// h->stats (left by whatever produced x; x itself is the fp16 residual stream), both LayerNorms are folded into the QKV / fc1
// GEMMs, and the out-proj / fc2
// epilogues leave the same by-products for the next consumer.  An encoder layer's 20x pooled state (model.py:171-174) is taken
// from the bf16 copy fc2 leaves (pool20_bf16_kernel).
int run_block_bf16(wat_handle* h, const BlockW& w, void* x, int n_seq, int T, bool encoder, cudaStream_t st, int* np,
                   float* pooled = nullptr, int pool_layer = -1) {
  const int D = w.D, rows = n_seq * T;
  int rc;
  GemmTc g;
  memset(&g, 0, sizeof(g));
  g.A = (const __nv_bfloat16*)h->xn.p; g.lda = D; g.W = w.wqkv_h; g.bias = w.bqkv_f; g.C = h->qkv.p;
  g.M = rows; g.N = 3 * D; g.K = D;
  g.ln_stats = (const float*)h->stats.p; g.ln_np = *np; g.ln_colsum = w.cs_qkv;
  if (encoder) {
    g.ldc = 2 * D; g.epi = TC_EPI_QKV;
    g.vt = (__nv_bfloat16*)h->vt.p; g.seq_T = T; g.seq_Tpad = 1536; g.n_head = w.H;
    g.q_scale = AT_QSCALE;                                        // q leaves the epilogue as q * 64^-0.5 * log2(e): attention computes P = 2^S
    h->prof_override = PC_GEMM_QKV;
    KL(h, launch_gemm_tc(g, h->num_sms, st));
    KL(h, launch_attn_tc((const __nv_bfloat16*)h->qkv.p, (const __nv_bfloat16*)h->vt.p, (__nv_bfloat16*)h->att.p, n_seq, T,
                         1536, w.H, true, st));
  } else {
    g.ldc = 3 * D; g.epi = TC_EPI_BF16;
    KL(h, launch_gemm_tc(g, h->num_sms, st));
    KL(h, launch_attn_small(h->qkv.p, true, h->att.p, true, n_seq, T, w.H, D / w.H, st));
  }
  if (encoder) h->prof_override = PC_GEMM_OUT;
  if ((rc = gemm_producer(h, h->att.p, D, w.wo_h, w.bo, x, x, true, D, 0, rows, D, D, 0, st, np))) return rc;
  memset(&g, 0, sizeof(g));
  g.A = (const __nv_bfloat16*)h->xn.p; g.lda = D; g.W = w.w1_h; g.bias = w.b1_f; g.C = h->hbuf.p; g.ldc = 4 * D;
  g.M = rows; g.N = 4 * D; g.K = D; g.act = 1; g.epi = TC_EPI_BF16;
  g.ln_stats = (const float*)h->stats.p; g.ln_np = *np; g.ln_colsum = w.cs_1;
  if (encoder) h->prof_override = PC_GEMM_FC1;
  KL(h, launch_gemm_tc(g, h->num_sms, st));
  if (encoder) h->prof_override = PC_GEMM_FC2;
  if ((rc = gemm_producer(h, h->hbuf.p, 4 * D, w.w2_h, w.b2, x, x, true, D, 0, rows, D, 4 * D, 0, st, np))) return rc;
  if (pooled && pool_layer >= 0) KL(h, launch_pool20_bf16((const __nv_bfloat16*)h->xn.p, n_seq, T, D, pool_layer, h->L, pooled, st));
  return 0;
}

// conv stem + blocks from time-major mel (melT) -> pooled [B, L, 75, d]; optional ln_post(x)
int run_encoder(wat_handle* h, int B, float* pooled, float* x_out, cudaStream_t st) {
  const int d = h->d, nm = h->cfg.n_mels;
  float* x = (float*)h->x.p;
  int rc;
  // conv1 (k3, p1) + GELU : im2col [B*3000, 3*n_mels] -> h1 [B*3000, d]
  KL(h, launch_im2col_k3(h->melT.p, h->bf16, B, 3000, nm, 1, 3000, h->hbuf.p, st));
  if ((rc = gemm(h, h->hbuf.p, 3 * nm, h->conv1_w, h->conv1_w_h, h->conv1_b, h->qkv.p, d, nullptr, 0, 0, B * 3000, d, 3 * nm, 1,
                 false, st))) return rc;
  // conv2 (k3, s2, p1) + GELU + positional embedding : im2col [B*1500, 3d] -> x [B*1500, d] fp32
  KL(h, launch_im2col_k3(h->qkv.p, h->bf16, B, 3000, d, 2, 1500, h->hbuf.p, st));
  if (h->bf16) {
    int np = 0;
    if ((rc = gemm_producer(h, h->hbuf.p, 3 * d, h->conv2_w_h, h->conv2_b, x, h->pos, false, d, 1500, B * 1500, d, 3 * d, 1, st, &np))) return rc;
    for (int l = 0; l < h->L; ++l)
      if ((rc = run_block_bf16(h, h->enc[l], x, B, 1500, true, st, &np, pooled, l))) return rc;
  } else {
    if ((rc = gemm(h, h->hbuf.p, 3 * d, h->conv2_w, h->conv2_w_h, h->conv2_b, x, d, h->pos, d, 1500, B * 1500, d, 3 * d, 1, true, st)))
      return rc;
    for (int l = 0; l < h->L; ++l)
      if ((rc = run_block_f32(h, h->enc[l], x, B, 1500, true, st, pooled, l - 1))) return rc;
    KL(h, launch_pool20(x, B, 1500, d, h->L - 1, h->L, pooled, st));   // last layer: nothing downstream reads x again
  }
  if (x_out) {
    if (h->bf16) KL(h, launch_layernorm_f16in(x, h->lnp_g, h->lnp_b, B * 1500, d, x_out, st));
    else KL(h, launch_layernorm(x, h->lnp_g, h->lnp_b, B * 1500, d, x_out, false, st));
  }
  return 0;
}

int run_head_impl(wat_handle* h, const float* pooled, int B, int t_total, int t_start, int t_len, int dw, float* logits,
                  cudaStream_t st);
int run_head(wat_handle* h, const float* pooled, int B, int t_total, int t_start, int t_len, int dw, float* logits,
             cudaStream_t st) {
  h->in_head = true;
  const int rc = run_head_impl(h, pooled, B, t_total, t_start, t_len, dw, logits, st);
  h->in_head = false;
  return rc;
}
int run_head_impl(wat_handle* h, const float* pooled, int B, int t_total, int t_start, int t_len, int dw, float* logits,
                  cudaStream_t st) {
  const int d = h->d, di = h->di, L = h->L, nc = h->cfg.n_class;
  if (dw < 1 || dw > 128) return fail(WAT_ERR_INVALID, "decision window %d out of range [1,128]", dw);
  if (t_len < 1 || t_len > (h->head_only ? 128 : 75) || t_start < 0 || t_start + t_len > t_total) return fail(WAT_ERR_INVALID, "bad pooled slice");
  const int S = (t_len + dw - 1) / dw;
  const int mode = h->head_mode;
  const bool layerwise = head_has_layer_tr(mode);
  const int64_t rows_clip = (int64_t)S * (layerwise ? L : 1) * dw;
  int rc;
  // One chunk for the whole call when it fits: the head works on S*dw*L rows per clip (1.6x the encoder's 1500 for large-v2 at
  // 10 s), so the row buffers are grown - once per handle and resolution - up to twice the encoder's size; beyond that the
  // clips are split into equal chunks.
  if (!h->head_only) {
    const int64_t want = (int64_t)std::min(B, h->head_chunk) * rows_clip;
    const int64_t limit = std::max<int64_t>(h->rows_cap, (int64_t)2 * h->ws_B * 1500);
    if (want > h->rows_cap && (rc = ensure_rows(h, std::min(want, limit), h->ws_B))) return rc;
  }
  int hc = (int)std::min<int64_t>(h->rows_cap / rows_clip, h->head_chunk);
  if (hc < 1) return fail(WAT_ERR_INVALID, "pooled length %d too long for the workspace", t_len);
  if (hc < B) { const int n_chunks = (B + hc - 1) / hc; hc = (B + n_chunks - 1) / n_chunks; }
  float* x = (float*)h->x.p;
  float* x2 = (float*)h->x2.p;
  for (int b0 = 0; b0 < B; b0 += hc) {
    const int nb = std::min(hc, B - b0);
    const int rows = (int)(nb * rows_clip);
    const float* pin = pooled + (int64_t)b0 * L * t_total * d;
    float* src0 = head_has_down(mode) ? x2 : x;                   // regrouped input; the down-projection lands in x
    // bf16 mode: the rows a transformer block works on (x, and x2 for the layer transformer) are the fp16 residual stream; the
    // regrouped input of the down-projection's LayerNorm kernel and the final means (lmean) stay fp32
    const bool f16 = h->bf16;
    const bool src_f16 = f16 && !head_has_down(mode) && head_has_time_tr(mode);
    // bf16 mode: whoever produces the rows a transformer block starts from also leaves their bf16 copy (h->xn) and their
    // LayerNorm statistics (h->stats, `np` slices) - see run_block_bf16.  The down-projection's own LayerNorm stays a kernel.
    const bool fold = h->bf16 && head_has_time_tr(mode);
    const bool fold_src = fold && !head_has_down(mode);
    __nv_bfloat16* xb = (__nv_bfloat16*)h->xn.p;
    float* stats = (float*)h->stats.p;
    int np = 1;
    if (layerwise) {
      KL(h, launch_head_gather(pin, nb, L, t_total, t_start, t_len, dw, S, d, src0, src_f16, st, fold_src ? xb : nullptr, fold_src ? stats : nullptr));
    } else {                                                      // baselines: the layer axis is reduced first (models.py:113-167)
      const int kind = (mode == WAT_HEAD_LAST_MLP || mode == WAT_HEAD_LAST_TR) ? 1 : head_has_layer_w(mode) ? 2 : 0;
      KL(h, launch_head_layer_reduce(pin, nb, L, t_total, t_start, t_len, dw, S, d, kind, h->layer_w, src0, src_f16, st,
                                     fold_src ? xb : nullptr, fold_src ? stats : nullptr));
    }
    if (head_has_down(mode)) {
      KL(h, launch_layernorm(x2, h->down_g, h->down_b, rows, d, h->xn.p, h->bf16, st));
      if (h->bf16) {
        if ((rc = gemm_producer(h, h->xn.p, d, h->down_w_h, h->down_bias, x, nullptr, false, 0, 0, rows, di, d, 0, st, &np))) return rc;
      } else if ((rc = gemm(h, h->xn.p, d, h->down_w, h->down_w_h, h->down_bias, x, di, nullptr, 0, 0, rows, di, d, 0, true, st))) return rc;
    }
    if (layerwise) {
      if (h->bf16) {
        if ((rc = run_block_bf16(h, h->time_tr, x, nb * S * L, dw, false, st, &np))) return rc;
        KL(h, launch_group_mean(x, true, nb * S * L, dw, di, x2, true, di, st, xb, stats));
        np = 1;
        if ((rc = run_block_bf16(h, h->layer_tr, x2, nb * S, L, false, st, &np))) return rc;
      } else {
        if ((rc = run_block_f32(h, h->time_tr, x, nb * S * L, dw, false, st))) return rc;
        KL(h, launch_group_mean(x, false, nb * S * L, dw, di, x2, false, di, st));
        if ((rc = run_block_f32(h, h->layer_tr, x2, nb * S, L, false, st))) return rc;
      }
      KL(h, launch_group_mean(x2, f16, nb * S, L, di, h->lmean.p, false, di, st));
    } else {
      if (head_has_time_tr(mode)) {
        if (h->bf16) { if ((rc = run_block_bf16(h, h->time_tr, x, nb * S, dw, false, st, &np))) return rc; }
        else if ((rc = run_block_f32(h, h->time_tr, x, nb * S, dw, false, st))) return rc;
      }
      // x is fp16 only when a transformer block (or the bf16 down-projection) wrote it; the *_mlp baselines read the fp32 rows
      const bool x_f16 = f16 && (head_has_time_tr(mode) || head_has_down(mode));
      KL(h, launch_group_mean(x, x_f16, nb * S, dw, di, h->lmean.p, false, di, st));
    }
    KL(h, launch_layernorm((const float*)h->lmean.p, h->cls_g, h->cls_b, nb * S, di, h->lnout.p, false, st));
    GemmF32 g;
    g.A = (const float*)h->lnout.p; g.lda = di; g.W = h->cls_w; g.bias = h->cls_bias;
    g.C = logits + (int64_t)b0 * S * nc; g.ldc = nc; g.R = nullptr; g.ldr = 0; g.r_mod = 0;
    g.M = nb * S; g.N = nc; g.K = di; g.act = 0;
    KL(h, launch_gemm_f32(g, st));
  }
  return 0;
}

// Calls on one handle are serialised on the device even when they come on different streams (they share the workspace).
// The stream of the previous call must still exist (torch's streams live as long as the process).
int order_after_previous(wat_handle* h, cudaStream_t st) {
  if (h->last_stream_valid && h->last_stream != st) {
    if (!h->order_ev) CU(cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming));
    CU(cudaEventRecord(h->order_ev, h->last_stream));
    CU(cudaStreamWaitEvent(st, h->order_ev, 0));
  }
  h->last_stream = st;
  h->last_stream_valid = true;
  h->cur_stream = st;
  return 0;
}

int check_ready(wat_handle* h, bool needs_encoder = true) {
  if (!h) return fail(WAT_ERR_INVALID, "null handle");
  if (!h->finalized) return fail(WAT_ERR_STATE, "wat_finalize has not been called");
  if (needs_encoder && h->head_only) return fail(WAT_ERR_STATE, "head-only handle (wat_head_create): no mel / encoder on it");
  return 0;
}

int mel_frames(wat_handle* h, const void* pcm, bool i16, int64_t clip_stride, const int32_t* n_valid_host, int n_samples, int n_pad,
               int B, int n_frames, cudaStream_t st, int clip0 = 0) {      // clip0: slot of the first clip in logspec / clipmax
  // frames that can see signal: t <= (n + 199) / 160 ; they all feed the per-clip max (audio.py:155)
  const int n_total_frames = (n_samples + n_pad) / 160;
  if (n_frames > n_total_frames) return fail(WAT_ERR_INVALID, "n_frames %d exceeds the %d frames of the padded signal", n_frames, n_total_frames);
  int n_scan = std::min(n_total_frames, (n_samples + 199) / 160 + 1);
  if (n_scan < n_frames) n_scan = n_frames;
  const int* nv_dev = nullptr;
  if (n_valid_host) {
    for (int i = 0; i < B; ++i)
      if (n_valid_host[i] < 0 || n_valid_host[i] > n_samples) return fail(WAT_ERR_INVALID, "n_valid[%d] out of range", i);
    CU(cudaMemcpyAsync((int*)h->nvalid.p + clip0, n_valid_host, sizeof(int) * B, cudaMemcpyHostToDevice, st));
    nv_dev = (const int*)h->nvalid.p + clip0;
  }
  KL(h, launch_mel_power(h->mel, pcm, i16, clip_stride, nv_dev, n_samples, n_pad, B, n_scan, n_frames, n_frames,
                         (float*)h->logspec.p + (size_t)clip0 * n_frames * h->cfg.n_mels, (float*)h->clipmax.p + clip0, st));
  h->launches++;                                                 // the clip_max fill kernel
  return 0;
}

}  // namespace

// =========================================================================================== C ABI
extern "C" {

int wat_abi_version(void) { return WAT_ABI_VERSION; }
const char* wat_last_error(void) { return g_err; }

int wat_create(const wat_config* cfg, wat_handle** out) {
  if (!cfg || !out) return fail(WAT_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->n_mels != 80 && cfg->n_mels != 128) return fail(WAT_ERR_INVALID, "Unsupported n_mels: %d", cfg->n_mels);
  if (cfg->n_audio_ctx != 1500) return fail(WAT_ERR_INVALID, "n_audio_ctx must be 1500");
  if (cfg->n_audio_state <= 0 || cfg->n_audio_state % 128 || cfg->n_audio_head * 64 != cfg->n_audio_state)
    return fail(WAT_ERR_INVALID, "n_audio_state must be a multiple of 128 with head_dim 64 (got d=%d, heads=%d)",
                cfg->n_audio_state, cfg->n_audio_head);
  if (cfg->n_audio_state > 1280)
    return fail(WAT_ERR_INVALID, "n_audio_state %d not supported: the LayerNorm / pooling kernels hold a row in registers (d <= 1280, Whisper large)",
                cfg->n_audio_state);
  if (cfg->n_audio_layer < 1 || cfg->n_audio_layer > 128) return fail(WAT_ERR_INVALID, "bad n_audio_layer");
  if (cfg->precision != WAT_FP32 && cfg->precision != WAT_BF16) return fail(WAT_ERR_INVALID, "bad precision");
  if (cfg->n_class < 1) return fail(WAT_ERR_INVALID, "bad n_class");
  if (cfg->at_low_compute && (cfg->at_dim <= 0 || cfg->at_dim % 128)) return fail(WAT_ERR_INVALID, "at_dim must be a multiple of 128");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(WAT_ERR_CUDA, "no CUDA device: libwat has no CPU fallback");
  wat_handle* h = new wat_handle();
  h->cfg = *cfg;
  if (h->cfg.max_batch <= 0) h->cfg.max_batch = 128;
  h->bf16 = cfg->precision == WAT_BF16;
  h->d = cfg->n_audio_state; h->H = cfg->n_audio_head; h->L = cfg->n_audio_layer;
  h->di = cfg->at_low_compute ? cfg->at_dim : h->d;
  if (h->di % 8) { delete h; return fail(WAT_ERR_INVALID, "head width must be a multiple of 8"); }
  cudaDeviceProp prop;
  if (cudaGetDevice(&h->device) != cudaSuccess || cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) {
    delete h;
    return fail(WAT_ERR_CUDA, "cannot query device");
  }
  h->num_sms = prop.multiProcessorCount;
  if (h->bf16 && prop.major != 10) { delete h; return fail(WAT_ERR_CUDA, "bf16 mode needs an sm_100a device (found sm_%d%d)", prop.major, prop.minor); }
  int rc = 0;
  const int d = h->d, nm = cfg->n_mels;
  do {
    if ((rc = build_mel_tables(h))) break;
    if ((rc = add_slot(h, "encoder.conv1.weight", &h->conv1_w, (int64_t)d * nm * 3, nm))) break;
    if ((rc = add_slot(h, "encoder.conv1.bias", &h->conv1_b, d))) break;
    if ((rc = add_slot(h, "encoder.conv2.weight", &h->conv2_w, (int64_t)d * d * 3, d))) break;
    if ((rc = add_slot(h, "encoder.conv2.bias", &h->conv2_b, d))) break;
    if ((rc = add_slot(h, "encoder.positional_embedding", &h->pos, (int64_t)1500 * d, 0, true))) break;
    if ((rc = add_slot(h, "encoder.ln_post.weight", &h->lnp_g, d))) break;
    if ((rc = add_slot(h, "encoder.ln_post.bias", &h->lnp_b, d))) break;
    h->enc.resize(h->L);
    for (int l = 0; l < h->L && !rc; ++l) rc = make_block(h, h->enc[l], "encoder.blocks." + std::to_string(l), d, h->H);
    if (rc) break;
    h->head_mode = cfg->at_low_compute ? WAT_HEAD_LW_DOWN_TR : WAT_HEAD_LW_TR;
    if ((rc = add_head_slots(h, 1, 8))) break;
    // default positional table (model.py:52-58); callers may overwrite it through wat_set_weight
    std::vector<float> pos((size_t)1500 * d);
    const double inc = std::log(10000.0) / (d / 2 - 1);
    for (int t = 0; t < 1500; ++t)
      for (int i = 0; i < d / 2; ++i) {
        const float inv = std::exp((float)(-inc * i));
        const float st = (float)t * inv;
        pos[(size_t)t * d + i] = std::sin(st);
        pos[(size_t)t * d + d / 2 + i] = std::cos(st);
      }
    if (cudaMemcpy(h->pos, pos.data(), sizeof(float) * pos.size(), cudaMemcpyHostToDevice) != cudaSuccess) { rc = fail(WAT_ERR_CUDA, "pos upload"); break; }
    if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(WAT_ERR_CUDA, "stream create"); break; }
    if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(WAT_ERR_CUDA, "stream create"); break; }
    for (int i = 0; i < 16 && !rc; ++i)
      if (cudaEventCreateWithFlags(&h->piece_ev[i / 8][i % 8], cudaEventDisableTiming) != cudaSuccess) rc = fail(WAT_ERR_CUDA, "event create");
    for (int i = 0; i < 2 && !rc; ++i)
      if (cudaEventCreateWithFlags(&h->done_ev[i], cudaEventDisableTiming) != cudaSuccess) rc = fail(WAT_ERR_CUDA, "event create");
    if (rc) break;
  } while (0);
  if (rc) { wat_destroy(h); return rc; }
  *out = h;
  return WAT_OK;
}

int wat_set_weight(wat_handle* h, const char* key, const float* host_data, int64_t numel) {
  if (!h || !key || !host_data) return fail(WAT_ERR_INVALID, "null argument");
  if (h->finalized) return fail(WAT_ERR_STATE, "handle already finalized");
  ON_DEVICE(h);
  if (!strncmp(key, "decoder.", 8)) return WAT_OK;                // ASR decoder: not on this path
  std::string norm(key);
  if (h->head_only) {                                             // TLTR's own keys: [module.]time_tr.* -> at_model.time_tr.*
    if (!norm.compare(0, 7, "module.")) norm.erase(0, 7);
    if (norm.compare(0, 9, "at_model.")) norm = "at_model." + norm;
  }
  auto it = h->slots.find(norm);
  if (it == h->slots.end()) return fail(WAT_ERR_INVALID, "Unexpected key in state_dict: %s", key);
  Slot& s = it->second;
  if (numel != s.numel) return fail(WAT_ERR_INVALID, "size mismatch for %s: expected %lld elements, got %lld", key, (long long)s.numel, (long long)numel);
  if (s.conv_c > 0) {
    const int c = s.conv_c;
    const int64_t out_ch = numel / (3 * c);
    std::vector<float> tmp((size_t)numel);
    for (int64_t o = 0; o < out_ch; ++o)
      for (int ci = 0; ci < c; ++ci)
        for (int k = 0; k < 3; ++k) tmp[(size_t)(o * 3 + k) * c + ci] = host_data[(size_t)(o * c + ci) * 3 + k];
    CU(cudaMemcpy(s.dst, tmp.data(), sizeof(float) * numel, cudaMemcpyHostToDevice));
  } else {
    CU(cudaMemcpy(s.dst, host_data, sizeof(float) * numel, cudaMemcpyHostToDevice));
  }
  s.set = true;
  return WAT_OK;
}

int wat_finalize(wat_handle* h) {
  if (!h) return fail(WAT_ERR_INVALID, "null handle");
  if (h->finalized) return WAT_OK;
  ON_DEVICE(h);
  for (auto& kv : h->slots)
    if (!kv.second.set && !kv.second.optional) return fail(WAT_ERR_STATE, "Missing key in state_dict: %s", kv.first.c_str());
  if (h->bf16) {
    int rc;
    const int64_t d = h->d;
    if (!h->head_only) {
      if ((rc = to_bf16(h, h->conv1_w, &h->conv1_w_h, d * 3 * h->cfg.n_mels))) return rc;
      if ((rc = to_bf16(h, h->conv2_w, &h->conv2_w_h, d * 3 * d))) return rc;
      for (auto& b : h->enc) if ((rc = pack_block_bf16(h, b))) return rc;
    }
    if (head_has_time_tr(h->head_mode) && (rc = pack_block_bf16(h, h->time_tr))) return rc;
    if (head_has_layer_tr(h->head_mode) && (rc = pack_block_bf16(h, h->layer_tr))) return rc;
    if (head_has_down(h->head_mode) && (rc = to_bf16(h, h->down_w, &h->down_w_h, (int64_t)h->di * d))) return rc;
    CU(cudaDeviceSynchronize());
  }
  h->finalized = true;
  return WAT_OK;
}

int wat_destroy(wat_handle* h) {
  if (!h) return WAT_OK;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  for (void* p : h->owned) cudaFree(p);
  Buf* bufs[] = {&h->x, &h->x2, &h->xn, &h->qkv, &h->vt, &h->att, &h->hbuf, &h->logspec, &h->clipmax, &h->melT,
                 &h->pooled, &h->lmean, &h->lnout, &h->nvalid, &h->logits, &h->pcm_stage[0], &h->pcm_stage[1], &h->stats};
  for (Buf* b : bufs) if (b->p) cudaFree(b->p);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->order_ev) cudaEventDestroy(h->order_ev);
  for (int i = 0; i < 16; ++i) if (h->piece_ev[i / 8][i % 8]) cudaEventDestroy(h->piece_ev[i / 8][i % 8]);
  for (int i = 0; i < 2; ++i) if (h->done_ev[i]) cudaEventDestroy(h->done_ev[i]);
  for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  delete h;
  return WAT_OK;
}

int wat_logmel(wat_handle* h, const float* pcm, int64_t clip_stride, const int32_t* n_valid, int32_t n_samples,
               int32_t n_pad, int32_t B, int32_t n_frames, int32_t clamp_scope, float* mel_out, void* stream) {
  if (!h) return fail(WAT_ERR_INVALID, "null handle");
  if (h->head_only) return fail(WAT_ERR_STATE, "head-only handle (wat_head_create): no mel tables on it");
  ON_DEVICE(h);
  int rc = 0;
  if (!pcm || !mel_out || B < 1 || n_samples < 1 || n_pad < 0 || n_frames < 1) return fail(WAT_ERR_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = order_after_previous(h, st))) return rc;
  Buf &ls = h->logspec, &cm = h->clipmax, &nv = h->nvalid;
  if ((rc = grow(h, ls, sizeof(float) * (size_t)B * n_frames * h->cfg.n_mels))) return rc;
  if ((rc = grow(h, cm, sizeof(float) * B))) return rc;
  if ((rc = grow(h, nv, sizeof(int) * B))) return rc;
  if ((rc = mel_frames(h, pcm, false, clip_stride, n_valid, n_samples, n_pad, B, n_frames, st))) return rc;
  if (clamp_scope == 1 && B > 1) KL(h, launch_share_max((float*)cm.p, B, st));
  KL(h, launch_mel_norm((const float*)ls.p, (const float*)cm.p, B, n_frames, n_frames, h->cfg.n_mels, 0, mel_out, st));
  return WAT_OK;
}

int wat_encoder(wat_handle* h, const float* mel, int32_t B, float* pooled_out, float* x_out, void* stream) {
  int rc = check_ready(h);
  if (rc) return rc;
  ON_DEVICE(h);
  if (!mel || !pooled_out || B < 1) return fail(WAT_ERR_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if ((rc = order_after_previous(h, st))) return rc;
  const int cb = h->cfg.max_batch;
  for (int b0 = 0; b0 < B; b0 += cb) {
    const int nb = std::min(cb, B - b0);
    if ((rc = ensure_ws(h, nb))) return rc;
    KL(h, launch_mel_to_timemajor(mel + (int64_t)b0 * h->cfg.n_mels * 3000, nb, 3000, h->cfg.n_mels, h->bf16 ? 2 : 1, h->melT.p, st));
    if ((rc = run_encoder(h, nb, pooled_out + (int64_t)b0 * h->L * 75 * h->d,
                          x_out ? x_out + (int64_t)b0 * 1500 * h->d : nullptr, st))) return rc;
  }
  return WAT_OK;
}

int wat_tltr(wat_handle* h, const float* pooled, int32_t B, int32_t t_total, int32_t t_start, int32_t t_len, int32_t dw,
             float* logits_out, void* stream) {
  int rc = check_ready(h, false);
  if (rc) return rc;
  ON_DEVICE(h);
  if (!pooled || !logits_out || B < 1) return fail(WAT_ERR_INVALID, "bad argument");
  if ((rc = ensure_ws(h, std::min(B, h->cfg.max_batch)))) return rc;
  if ((rc = order_after_previous(h, (cudaStream_t)stream))) return rc;
  return run_head(h, pooled, B, t_total, t_start, t_len, dw, logits_out, (cudaStream_t)stream);
}

int wat_head_create(const wat_head_config* cfg, wat_handle** out) {
  if (!cfg || !out) return fail(WAT_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->mode < WAT_HEAD_LW_TR || cfg->mode > WAT_HEAD_WA_DOWN_TR) return fail(WAT_ERR_INVALID, "unknown head mode %d", cfg->mode);
  if (cfg->rep_dim <= 0 || cfg->rep_dim % 128) return fail(WAT_ERR_INVALID, "rep_dim must be a multiple of 128");
  if (cfg->n_layer < 1 || cfg->n_layer > 128) return fail(WAT_ERR_INVALID, "bad n_layer");
  if (cfg->n_class < 1) return fail(WAT_ERR_INVALID, "bad n_class");
  if (cfg->precision != WAT_FP32 && cfg->precision != WAT_BF16) return fail(WAT_ERR_INVALID, "bad precision");
  const int di = head_has_down(cfg->mode) ? cfg->inter_dim : cfg->rep_dim;
  if (di <= 0 || di % 128) return fail(WAT_ERR_INVALID, "inter_dim must be a multiple of 128");
  const int nt = head_has_time_tr(cfg->mode) ? cfg->n_time_head : 1, nl = head_has_layer_tr(cfg->mode) ? cfg->n_layer_head : 1;
  if (nt < 1 || nl < 1 || di % nt || di % nl || (di / nt) % 4 || (di / nl) % 4)
    return fail(WAT_ERR_INVALID, "head counts (%d, %d) must divide the transformer width %d into multiples of 4", nt, nl, di);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(WAT_ERR_CUDA, "no CUDA device: libwat has no CPU fallback");
  wat_handle* h = new wat_handle();
  memset(&h->cfg, 0, sizeof(h->cfg));
  h->cfg.n_mels = 80; h->cfg.n_audio_ctx = 1500; h->cfg.n_audio_state = cfg->rep_dim; h->cfg.n_audio_head = cfg->rep_dim / 64;
  h->cfg.n_audio_layer = cfg->n_layer; h->cfg.at_low_compute = head_has_down(cfg->mode); h->cfg.at_dim = di;
  h->cfg.n_class = cfg->n_class; h->cfg.precision = cfg->precision;
  h->cfg.max_batch = cfg->max_batch > 0 ? cfg->max_batch : 128;
  h->bf16 = cfg->precision == WAT_BF16;
  h->d = cfg->rep_dim; h->H = cfg->rep_dim / 64; h->L = cfg->n_layer; h->di = di;
  h->head_only = true;
  h->head_mode = cfg->mode;
  cudaDeviceProp prop;
  if (cudaGetDevice(&h->device) != cudaSuccess || cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) {
    delete h;
    return fail(WAT_ERR_CUDA, "cannot query device");
  }
  h->num_sms = prop.multiProcessorCount;
  if (h->bf16 && prop.major != 10) { delete h; return fail(WAT_ERR_CUDA, "bf16 mode needs an sm_100a device (found sm_%d%d)", prop.major, prop.minor); }
  int rc = add_head_slots(h, nt, nl);
  if (!rc && cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) rc = fail(WAT_ERR_CUDA, "stream create");
  if (rc) { wat_destroy(h); return rc; }
  *out = h;
  return WAT_OK;
}

int wat_head_forward(wat_handle* h, const float* audio_rep, int32_t B, int32_t Tp, float* logits_out, void* stream) {
  if (Tp < 1 || Tp > 128) return fail(WAT_ERR_INVALID, "Tp %d out of range [1,128]", Tp);
  return wat_tltr(h, audio_rep, B, Tp, 0, Tp, Tp, logits_out, stream);      // one window spanning the whole segment
}

// piece_ev / piece_clips: wat_tag_host copies the PCM in pieces of piece_clips clips on its copy stream; the mel kernel of a
// piece waits for that piece's event only, so the H2D of piece p + 1 overlaps the mel of piece p.
static int tag_impl(wat_handle* h, const void* pcm, bool i16, int64_t clip_stride, const int32_t* n_valid, int32_t n_samples,
                    int32_t B, int32_t dw, float* logits_out, cudaStream_t st, const cudaEvent_t* piece_ev = nullptr,
                    int piece_clips = 0) {
  int rc = check_ready(h);
  if (rc) return rc;
  ON_DEVICE(h);
  if (!pcm || !logits_out || B < 1 || n_samples < 1 || n_samples > 480000) return fail(WAT_ERR_INVALID, "bad argument (clips are <= 480000 samples)");
  if (dw < 1 || dw > 128) return fail(WAT_ERR_INVALID, "decision window %d out of range [1,128]", dw);
  if ((rc = order_after_previous(h, st))) return rc;
  const int cb = h->cfg.max_batch;
  const int S = (75 + dw - 1) / dw;
  const size_t esz = i16 ? sizeof(int16_t) : sizeof(float);
  for (int b0 = 0; b0 < B; b0 += cb) {
    const int nb = std::min(cb, B - b0);
    if ((rc = ensure_ws(h, nb))) return rc;
    const void* p0 = reinterpret_cast<const char*>(pcm) + (size_t)b0 * clip_stride * esz;
    if (!piece_ev) {
      if ((rc = mel_frames(h, p0, i16, clip_stride, n_valid ? n_valid + b0 : nullptr, n_samples, 480000, nb, 3000, st))) return rc;
    } else {
      for (int q0 = 0; q0 < nb;) {
        const int piece = (b0 + q0) / piece_clips;
        const int q1 = std::min(nb, (piece + 1) * piece_clips - b0);
        CU(cudaStreamWaitEvent(st, piece_ev[piece], 0));
        if ((rc = mel_frames(h, reinterpret_cast<const char*>(p0) + (size_t)q0 * clip_stride * esz, i16, clip_stride,
                             n_valid ? n_valid + b0 + q0 : nullptr, n_samples, 480000, q1 - q0, 3000, st, q0))) return rc;
        q0 = q1;
      }
    }
    KL(h, launch_mel_norm((const float*)h->logspec.p, (const float*)h->clipmax.p, nb, 3000, 3000, h->cfg.n_mels, h->bf16 ? 2 : 1,
                          h->melT.p, st));
    if ((rc = run_encoder(h, nb, (float*)h->pooled.p, nullptr, st))) return rc;
    if ((rc = run_head(h, (const float*)h->pooled.p, nb, 75, 0, 75, dw, logits_out + (int64_t)b0 * S * h->cfg.n_class, st))) return rc;
  }
  return WAT_OK;
}

// Host entry points.  A call is submitted (H2D of the PCM in pieces on the copy stream, compute on the handle's own stream, D2H
// of the logits behind it) and later waited for.  Two PCM stages alternate, so while call t computes, the PCM of call t + 1 is
// already crossing PCIe: in a submit(t+1) / wait(t) loop the H2D copy disappears behind the previous call's compute.  The compute
// of successive calls is serialised by the stream (they share the workspace); the device logits buffer is shared too, because
// the D2H of call t precedes the head of call t + 1 in stream order.
static int tag_host_submit(wat_handle* h, const void* pcm_host, bool i16, int64_t clip_stride, const int32_t* n_valid,
                           int32_t n_samples, int32_t B, int32_t dw, float* logits_host, int64_t* ticket) {
  int rc = check_ready(h);
  if (rc) return rc;
  ON_DEVICE(h);
  if (!pcm_host || !logits_host || B < 1 || n_samples < 1 || n_samples > 480000 || clip_stride < n_samples)
    return fail(WAT_ERR_INVALID, "bad argument");
  if (dw < 1 || dw > 128) return fail(WAT_ERR_INVALID, "decision window %d out of range [1,128]", dw);
  if (n_valid)
    for (int i = 0; i < B; ++i)
      if (n_valid[i] < 0 || n_valid[i] > n_samples) return fail(WAT_ERR_INVALID, "n_valid[%d] out of range", i);
  const int S = (75 + dw - 1) / dw;
  cudaStream_t st = h->own_stream;
  const int slot = (int)(h->next_ticket & 1);
  // the stage (and its events) belong to call t - 2 until that call has finished; a caller that never waited for it waits here
  if (h->slot_ticket[slot]) { CU(cudaEventSynchronize(h->done_ev[slot])); h->slot_ticket[slot] = 0; }
  const size_t esz = i16 ? sizeof(int16_t) : sizeof(float);
  const size_t pcm_bytes = esz * ((size_t)(B - 1) * clip_stride + n_samples);
  if ((rc = grow(h, h->pcm_stage[slot], pcm_bytes))) return rc;
  if ((rc = grow(h, h->logits, sizeof(float) * (size_t)B * S * h->cfg.n_class))) return rc;
  const int n_piece = B >= 16 ? 8 : 1;
  const int piece_clips = (B + n_piece - 1) / n_piece;
  cudaEvent_t* piece_ev = h->piece_ev[slot];
  for (int p = 0; p < n_piece; ++p) {
    const int c0 = p * piece_clips, c1 = std::min(B, c0 + piece_clips);
    if (c0 >= c1) { CU(cudaEventRecord(piece_ev[p], h->copy_stream)); continue; }
    const size_t off = (size_t)c0 * clip_stride * esz;
    const size_t bytes = esz * ((size_t)(c1 - c0 - 1) * clip_stride + n_samples);
    CU(cudaMemcpyAsync((char*)h->pcm_stage[slot].p + off, (const char*)pcm_host + off, bytes, cudaMemcpyHostToDevice, h->copy_stream));
    CU(cudaEventRecord(piece_ev[p], h->copy_stream));
  }
  if ((rc = tag_impl(h, h->pcm_stage[slot].p, i16, clip_stride, n_valid, n_samples, B, dw, (float*)h->logits.p, st, piece_ev, piece_clips))) {
    // the H2D pieces (and whatever was launched before the failure) may still be in flight: the caller is free to release
    // pcm_host as soon as we return, and a later call reuses the stage
    cudaStreamSynchronize(h->copy_stream);
    cudaStreamSynchronize(st);
    return rc;
  }
  CU(cudaMemcpyAsync(logits_host, h->logits.p, sizeof(float) * (size_t)B * S * h->cfg.n_class, cudaMemcpyDeviceToHost, st));
  CU(cudaEventRecord(h->done_ev[slot], st));
  h->slot_ticket[slot] = h->next_ticket;
  *ticket = h->next_ticket++;
  return WAT_OK;
}

static int tag_host_wait(wat_handle* h, int64_t ticket) {
  if (!h) return fail(WAT_ERR_INVALID, "null handle");
  if (ticket < 1 || ticket >= h->next_ticket) return fail(WAT_ERR_INVALID, "unknown ticket %lld", (long long)ticket);
  const int slot = (int)(ticket & 1);
  if (h->slot_ticket[slot] != ticket) return WAT_OK;              // waited for already, or its stage was re-used (which waits for it)
  ON_DEVICE(h);
  CU(cudaEventSynchronize(h->done_ev[slot]));
  h->slot_ticket[slot] = 0;
  return WAT_OK;
}

static int tag_host_impl(wat_handle* h, const void* pcm_host, bool i16, int64_t clip_stride, const int32_t* n_valid,
                         int32_t n_samples, int32_t B, int32_t dw, float* logits_host) {
  int64_t ticket = 0;
  if (int rc = tag_host_submit(h, pcm_host, i16, clip_stride, n_valid, n_samples, B, dw, logits_host, &ticket)) return rc;
  return tag_host_wait(h, ticket);
}

int wat_tag(wat_handle* h, const float* pcm, int64_t clip_stride, const int32_t* n_valid, int32_t n_samples, int32_t B,
            int32_t dw, float* logits_out, void* stream) {
  return tag_impl(h, pcm, false, clip_stride, n_valid, n_samples, B, dw, logits_out, (cudaStream_t)stream);
}
int wat_tag_pcm16(wat_handle* h, const int16_t* pcm, int64_t clip_stride, const int32_t* n_valid, int32_t n_samples, int32_t B,
                  int32_t dw, float* logits_out, void* stream) {
  return tag_impl(h, pcm, true, clip_stride, n_valid, n_samples, B, dw, logits_out, (cudaStream_t)stream);
}
int wat_tag_host(wat_handle* h, const float* pcm_host, int64_t clip_stride, const int32_t* n_valid, int32_t n_samples,
                 int32_t B, int32_t dw, float* logits_host) {
  return tag_host_impl(h, pcm_host, false, clip_stride, n_valid, n_samples, B, dw, logits_host);
}
int wat_tag_host_pcm16(wat_handle* h, const int16_t* pcm_host, int64_t clip_stride, const int32_t* n_valid, int32_t n_samples,
                       int32_t B, int32_t dw, float* logits_host) {
  return tag_host_impl(h, pcm_host, true, clip_stride, n_valid, n_samples, B, dw, logits_host);
}

int wat_tag_host_submit(wat_handle* h, const float* pcm_host, int64_t clip_stride, const int32_t* n_valid, int32_t n_samples,
                        int32_t B, int32_t dw, float* logits_host, int64_t* ticket) {
  if (!ticket) return fail(WAT_ERR_INVALID, "null ticket");
  return tag_host_submit(h, pcm_host, false, clip_stride, n_valid, n_samples, B, dw, logits_host, ticket);
}
int wat_tag_host_submit_pcm16(wat_handle* h, const int16_t* pcm_host, int64_t clip_stride, const int32_t* n_valid, int32_t n_samples,
                              int32_t B, int32_t dw, float* logits_host, int64_t* ticket) {
  if (!ticket) return fail(WAT_ERR_INVALID, "null ticket");
  return tag_host_submit(h, pcm_host, true, clip_stride, n_valid, n_samples, B, dw, logits_host, ticket);
}
int wat_tag_host_wait(wat_handle* h, int64_t ticket) { return tag_host_wait(h, ticket); }

int64_t wat_workspace_bytes(const wat_handle* h) { return h ? h->ws_bytes : 0; }
int64_t wat_kernel_launches(const wat_handle* h) { return h ? h->launches : 0; }
int wat_num_sms(const wat_handle* h) { return h ? h->num_sms : 0; }

int wat_profile(wat_handle* h, int32_t enable) {
  if (!h) return fail(WAT_ERR_INVALID, "null handle");
  h->profiling = enable != 0;
  h->prof_used = 0;
  return WAT_OK;
}

int wat_profile_classes(void) { return PC_COUNT; }
const char* wat_profile_class_name(int32_t i) { return (i >= 0 && i < PC_COUNT) ? kProfNames[i] : ""; }

int wat_profile_read(wat_handle* h, double* ms, int64_t* launches) {
  if (!h || !ms || !launches) return fail(WAT_ERR_INVALID, "null argument");
  ON_DEVICE(h);
  for (int i = 0; i < PC_COUNT; ++i) { ms[i] = 0.0; launches[i] = 0; }
  for (size_t i = 0; i < h->prof_used; ++i) {
    CU(cudaEventSynchronize(h->prof[i].b));
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, h->prof[i].a, h->prof[i].b));
    ms[h->prof[i].cls] += t;
    launches[h->prof[i].cls]++;
  }
  h->prof_used = 0;
  return WAT_OK;
}

// ------------------------------------------------------------------------------------------- test hooks
int wat_dbg_gemm(const float* A, const float* W, const float* bias, const float* R, float* C, int32_t M, int32_t N,
                 int32_t K, int32_t act, int32_t tc, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!tc) {
    GemmF32 g;
    g.A = A; g.lda = K; g.W = W; g.bias = bias; g.C = C; g.ldc = N; g.R = R; g.ldr = N; g.r_mod = 0;
    g.M = M; g.N = N; g.K = K; g.act = act;
    CU(launch_gemm_f32(g, st));
    return WAT_OK;
  }
  int dev = 0, sms = 148;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  __nv_bfloat16 *Ah = nullptr, *Wh = nullptr;
  CU(cudaMalloc(&Ah, sizeof(__nv_bfloat16) * (size_t)M * K));
  CU(cudaMalloc(&Wh, sizeof(__nv_bfloat16) * (size_t)N * K));
  CU(launch_f32_to_bf16(A, Ah, (int64_t)M * K, st));
  CU(launch_f32_to_bf16(W, Wh, (int64_t)N * K, st));
  GemmTc g;
  memset(&g, 0, sizeof(g));
  g.A = Ah; g.lda = K; g.W = Wh; g.bias = bias; g.C = C; g.ldc = N; g.R = R; g.ldr = N; g.r_mod = 0;
  g.M = M; g.N = N; g.K = K; g.act = act; g.epi = R ? TC_EPI_F32_RES : TC_EPI_F32;
  g.force_pair = tc >= 2 ? 1 : -1;                               // tc: 1 = single-CTA kernel, 2 = CTA-pair kernel, 3 = pair + trace
  __nv_bfloat16* Cb = nullptr;                                   // tc 4: trace of the bf16 (+GELU) epilogue kernel; C is not written
  if (tc == 4) { CU(cudaMalloc(&Cb, sizeof(__nv_bfloat16) * (size_t)M * N)); g.C = Cb; g.epi = TC_EPI_BF16; g.R = nullptr; }
  long long* trace = nullptr;
  if (tc >= 3) { CU(cudaMalloc(&trace, 32 * 8 * 8)); CU(cudaMemsetAsync(trace, 0, 32 * 8 * 8, st)); g.trace = trace; }
  cudaError_t e = launch_gemm_tc(g, sms, st);
  cudaError_t e2 = cudaStreamSynchronize(st);
  if (trace) {
    long long ht[32 * 8];
    cudaMemcpy(ht, trace, sizeof(ht), cudaMemcpyDeviceToHost);
    cudaFree(trace);
    fprintf(stderr, "epilogue warp of cluster 0 (cycles): tile | wait_start acc_ready chunk0 chunk1 chunk2 chunk3\n");
    for (int t = 0; t < 24 && ht[t * 8]; ++t) {
      fprintf(stderr, "tile %2d |", t);
      for (int k = 0; k < 6; ++k) fprintf(stderr, " %7lld", ht[t * 8 + k] ? ht[t * 8 + k] - ht[0] : -1);
      fprintf(stderr, "\n");
    }
    fprintf(stderr, "chunk 1 of tiles 8..15: [tmem ld + wait done -> ] bias_done gelu_done stored  (cycles since chunk start)\n");
    for (int t = 0; t < 8; ++t) {
      const long long* c = ht + 24 * 8 + t * 4;
      if (c[3]) fprintf(stderr, "  bias %6lld  gelu %6lld  store %6lld\n", c[0] - c[3], c[1] - c[3], c[2] - c[3]);
    }
  }
  cudaFree(Ah); cudaFree(Wh);
  if (Cb) cudaFree(Cb);
  if (e != cudaSuccess) return fail(WAT_ERR_CUDA, "launch_gemm_tc: %s", cudaGetErrorString(e));
  if (e2 != cudaSuccess) return fail(WAT_ERR_CUDA, "gemm_tc execution: %s", cudaGetErrorString(e2));
  return WAT_OK;
}

// The bf16-output epilogues of the tcgen05 GEMM exactly as the encoder launches them: A [M, K], W [N, K] bf16 device, bias fp32.
//   seq_T == 0 : C [M, N] bf16 = act(A W^T + bias)           (fc1: act = 1 selects the 16-epilogue-warp GELU kernel)
//   seq_T  > 0 : fused-QKV split epilogue, N = 3D: C [M, 2D] bf16 = (q | k), vt [M / seq_T, n_head, 64, seq_Tpad] = V^T
int wat_dbg_gemm_bf16(const void* A, const void* W, const float* bias, void* C, void* vt, int32_t M, int32_t N, int32_t K,
                      int32_t act, int32_t seq_T, int32_t seq_Tpad, int32_t n_head, void* stream) {
  if (!A || !W || !C || M < 1 || N < 1 || K < 1) return fail(WAT_ERR_INVALID, "bad argument");
  if (seq_T > 0 && (!vt || M % seq_T || n_head * 64 * 3 != N || seq_Tpad < seq_T)) return fail(WAT_ERR_INVALID, "bad QKV shape");
  int dev = 0, sms = 148;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  GemmTc g;
  memset(&g, 0, sizeof(g));
  g.A = (const __nv_bfloat16*)A; g.lda = K; g.W = (const __nv_bfloat16*)W; g.bias = bias; g.C = C;
  g.M = M; g.N = N; g.K = K; g.act = act;
  if (seq_T > 0) {
    g.epi = TC_EPI_QKV; g.ldc = 2 * (N / 3); g.vt = (__nv_bfloat16*)vt; g.seq_T = seq_T; g.seq_Tpad = seq_Tpad; g.n_head = n_head;
  } else {
    g.epi = TC_EPI_BF16; g.ldc = N;
  }
  cudaError_t e = launch_gemm_tc(g, sms, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(WAT_ERR_CUDA, "launch_gemm_tc: %s", cudaGetErrorString(e));
  return WAT_OK;
}

// LayerNorm folded into its consumer, as run_block_bf16 chains it (all pointers device):
//   producer  x = R + A1 W1^T + bias1            A1 [M,K1] bf16, W1 [D,K1] bf16, R / x [M,D] fp32 (fp16 when x_f16: how the
//             bf16 encoder keeps its residual stream); the epilogue also writes
//             xb [M,D] bf16 and stats [M, np, 2] (np = return value of wat_dbg_ln_slices); if pooled != NULL (M % 1500 == 0),
//             pooled [M/1500, 1, 75, D] = 20-row means of xb (pool20_bf16_kernel, as the encoder takes them)
//   consumer  out [M,N2] bf16 = act(LN(x; gamma, beta) W2^T + bias2) computed as rstd (xb W2'^T - mean colsum) + bias2'
// pair: 1 = CTA-pair kernels, -1 = single-CTA kernels.
int wat_dbg_ln_slices(int32_t M, int32_t D, int32_t K1, int32_t pair) { return gemm_tc_stats_slices(M, D, K1, TC_EPI_F32_RES, pair); }
int wat_dbg_ln_gemm(const void* A1, const void* W1, const float* bias1, const void* R, void* x, void* xb, float* stats, float* pooled,
                    const float* W2, const float* gamma, const float* beta, const float* bias2, void* out, int32_t M, int32_t D,
                    int32_t K1, int32_t N2, int32_t act, int32_t pair, int32_t x_f16, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 148;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  __nv_bfloat16* W2h = nullptr;
  float *cs = nullptr, *b2f = nullptr;
  CU(cudaMalloc(&W2h, sizeof(__nv_bfloat16) * (size_t)N2 * D));
  CU(cudaMalloc(&cs, sizeof(float) * N2));
  CU(cudaMalloc(&b2f, sizeof(float) * N2));
  cudaError_t e = launch_fold_ln_weights(W2, gamma, beta, bias2, N2, D, W2h, cs, b2f, st);
  GemmTc g;
  memset(&g, 0, sizeof(g));
  g.A = (const __nv_bfloat16*)A1; g.lda = K1; g.W = (const __nv_bfloat16*)W1; g.bias = bias1; g.C = x; g.ldc = D; g.R = (const float*)R; g.ldr = D;
  g.c_f16 = g.r_f16 = x_f16 ? 1 : 0;                              // R and x in fp16: the bf16 mode's residual stream
  g.M = M; g.N = D; g.K = K1; g.epi = TC_EPI_F32_RES; g.force_pair = pair;
  g.xb = (__nv_bfloat16*)xb; g.ldxb = D; g.stats = stats; g.stats_np = gemm_tc_stats_slices(M, D, K1, TC_EPI_F32_RES, pair);
  long long* trace = nullptr;                                     // WAT_DBG_TRACE=1: clock trace of the producer GEMM's first epilogue warp -> stderr
  if (getenv("WAT_DBG_TRACE") && pair == 1) { CU(cudaMalloc(&trace, 32 * 8 * 8)); CU(cudaMemsetAsync(trace, 0, 32 * 8 * 8, st)); g.trace = trace; }
  if (e == cudaSuccess) e = launch_gemm_tc(g, sms, st);
  if (trace) {
    long long ht[32 * 8];
    cudaStreamSynchronize(st);
    cudaMemcpy(ht, trace, sizeof(ht), cudaMemcpyDeviceToHost);
    cudaFree(trace);
    for (int i = 0; i < 24; ++i)
      fprintf(stderr, "tile %2d: start %8lld  wait_mma %6lld  chunks %6lld %6lld %6lld %6lld\n", i, ht[i * 8] - ht[0], ht[i * 8 + 1] - ht[i * 8],
              ht[i * 8 + 2] - ht[i * 8 + 1], ht[i * 8 + 3] - ht[i * 8 + 2], ht[i * 8 + 4] - ht[i * 8 + 3], ht[i * 8 + 5] - ht[i * 8 + 4]);
  }
  if (e == cudaSuccess && pooled) e = launch_pool20_bf16((const __nv_bfloat16*)xb, M / 1500, 1500, D, 0, 1, pooled, st);
  GemmTc c;
  memset(&c, 0, sizeof(c));
  c.A = (const __nv_bfloat16*)xb; c.lda = D; c.W = W2h; c.bias = b2f; c.C = out; c.ldc = N2; c.M = M; c.N = N2; c.K = D; c.act = act;
  c.epi = TC_EPI_BF16; c.force_pair = pair;
  c.ln_stats = stats; c.ln_np = g.stats_np; c.ln_colsum = cs;
  if (e == cudaSuccess) e = launch_gemm_tc(c, sms, st);
  cudaError_t e2 = cudaStreamSynchronize(st);
  cudaFree(W2h); cudaFree(cs); cudaFree(b2f);
  if (e != cudaSuccess) return fail(WAT_ERR_CUDA, "ln gemm launch: %s", cudaGetErrorString(e));
  if (e2 != cudaSuccess) return fail(WAT_ERR_CUDA, "ln gemm execution: %s", cudaGetErrorString(e2));
  return WAT_OK;
}

// tiles of the last wat_dbg_attention(tc != 0) call that were repeated with the running-max pass
static thread_local int g_dbg_attn_repeats = 0;
int wat_dbg_attention_repeats(void) { return g_dbg_attn_repeats; }

// x [B*T, D] fp32, wqkv [3D, D], bqkv [3D] -> out [B*T, D] fp32 = softmax(q k^T / 8) v per head (QKV GEMM + attention)
int wat_dbg_attention(const float* x, const float* wqkv, const float* bqkv, float* out, int32_t B, int32_t T,
                      int32_t n_head, int32_t tc, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int D = n_head * 64;
  const int64_t rows = (int64_t)B * T;
  if (!tc) {
    float* qkv = nullptr;
    CU(cudaMalloc(&qkv, sizeof(float) * rows * 3 * D));
    GemmF32 g;
    g.A = x; g.lda = D; g.W = wqkv; g.bias = bqkv; g.C = qkv; g.ldc = 3 * D; g.R = nullptr; g.ldr = 0; g.r_mod = 0;
    g.M = (int)rows; g.N = 3 * D; g.K = D; g.act = 0;
    cudaError_t e = launch_gemm_f32(g, st);
    if (e == cudaSuccess) e = launch_attn_f32_hd64(qkv, qkv + D, qkv + 2 * D, 3 * D, out, D, B, T, n_head, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(qkv);
    if (e != cudaSuccess || e2 != cudaSuccess) return fail(WAT_ERR_CUDA, "fp32 attention: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return WAT_OK;
  }
  int dev = 0, sms = 148;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int Tpad = ((T + 127) / 128) * 128;
  __nv_bfloat16 *xh, *wh, *qk, *vt, *oh;
  CU(cudaMalloc(&xh, 2 * rows * D));
  CU(cudaMalloc(&wh, 2 * (size_t)3 * D * D));
  CU(cudaMalloc(&qk, 2 * rows * 2 * D));
  CU(cudaMalloc(&vt, 2 * (size_t)B * n_head * VT_ROWS * Tpad));
  CU(cudaMalloc(&oh, 2 * rows * D));
  CU(cudaMemsetAsync(vt, 0, 2 * (size_t)B * n_head * VT_ROWS * Tpad, st));
  CU(launch_f32_to_bf16(x, xh, rows * D, st));
  CU(launch_f32_to_bf16(wqkv, wh, (int64_t)3 * D * D, st));
  GemmTc g;
  memset(&g, 0, sizeof(g));
  g.A = xh; g.lda = D; g.W = wh; g.bias = bqkv; g.C = qk; g.ldc = 2 * D; g.M = (int)rows; g.N = 3 * D; g.K = D;
  g.epi = TC_EPI_QKV; g.vt = vt; g.seq_T = T; g.seq_Tpad = Tpad; g.n_head = n_head;
  g.q_scale = tc == 4 ? 0.f : AT_QSCALE;                          // tc 4: un-scaled q -> only the running-max pass can run
  cudaError_t e = launch_gemm_tc(g, sms, st);
  long long* trace = nullptr;
  if (tc == 3) { CU(cudaMalloc(&trace, 8192)); CU(cudaMemsetAsync(trace, 0, 8192, st)); }
  unsigned int* n_rep = nullptr;
  CU(cudaMalloc(&n_rep, sizeof(unsigned int)));
  CU(cudaMemsetAsync(n_rep, 0, sizeof(unsigned int), st));
  if (e == cudaSuccess) e = launch_attn_tc(qk, vt, oh, B, T, Tpad, n_head, tc != 4, st, trace, n_rep);
  {
    unsigned int hr = 0;
    cudaStreamSynchronize(st);
    cudaMemcpy(&hr, n_rep, sizeof(hr), cudaMemcpyDeviceToHost);
    cudaFree(n_rep);
    g_dbg_attn_repeats = (int)hr;
  }
  if (trace) {                                                    // per-phase clock trace of one CTA -> stderr
    long long ht[1024];
    cudaStreamSynchronize(st);
    cudaMemcpy(ht, trace, 8192, cudaMemcpyDeviceToHost);
    cudaFree(trace);
    const long long t0 = ht[0];
    fprintf(stderr, "softmax warp (cycles since start): step | wait_s0 s_ready ld_done max_done halfA exp_done arrived\n");
    for (int j = 0; j < (T + 63) / 64; ++j) {
      fprintf(stderr, "sm %2d |", j);
      for (int k = 0; k < 7; ++k) fprintf(stderr, " %7lld", ht[j * 8 + k] ? ht[j * 8 + k] - t0 : -1);
      fprintf(stderr, " %s", ht[j * 8 + 7] ? "R" : " ");
      fprintf(stderr, "   || mma: iter_start qk_issued p_seen v_seen pv_issued k_seen sfree_seen qk_mma_done |");
      for (int k = 0; k < 8; ++k) fprintf(stderr, " %7lld", ht[512 + j * 8 + k] ? ht[512 + j * 8 + k] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  cudaError_t e2 = cudaStreamSynchronize(st);
  if (e == cudaSuccess && e2 == cudaSuccess) {
    // bf16 -> fp32 through a tiny identity: reuse the LN-free path by a cast kernel on the host side
    std::vector<__nv_bfloat16> hb((size_t)rows * D);
    std::vector<float> hf((size_t)rows * D);
    cudaMemcpy(hb.data(), oh, 2 * rows * D, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < hb.size(); ++i) hf[i] = __bfloat162float(hb[i]);
    cudaMemcpy(out, hf.data(), sizeof(float) * rows * D, cudaMemcpyHostToDevice);
  }
  cudaFree(xh); cudaFree(wh); cudaFree(qk); cudaFree(vt); cudaFree(oh);
  if (e != cudaSuccess) return fail(WAT_ERR_CUDA, "tc attention launch: %s", cudaGetErrorString(e));
  if (e2 != cudaSuccess) return fail(WAT_ERR_CUDA, "tc attention execution: %s", cudaGetErrorString(e2));
  return WAT_OK;
}

int wat_dbg_tma_overlap_probe(void) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return -1;
  void* p = nullptr;
  if (cudaMalloc(&p, 1 << 20) != cudaSuccess) return -1;
  CUtensorMap m;
  cuuint64_t dims[2] = {240, 3000};
  cuuint64_t strides[1] = {160};                                 // 80 bf16: rows overlap
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFree(p);
  return r == CUDA_SUCCESS ? 1 : 0;
}

}  // extern "C"
