// Fused log-mel front end (reference: package/whisper-at/whisper_at/audio.py:110-157).
//
//   phase 1  mel_power_kernel : framing (centre / reflect, hop 160) -> hann window -> 400-point real DFT
//            -> |X|^2 -> sparse slaney mel projection -> log10(max(.,1e-10)), plus the per-clip maximum.
//   phase 2  mel_norm_kernel  : max(x, clipmax - 8), (x + 4) / 4, written in the layout the consumer wants.
//
// The 400-point DFT of a frame is a three-stage Cooley-Tukey transform, 400 = 16 x 5 x 5, evaluated in fp64 (what enters is
// the fp32 product sample * window, exactly as torch.stft forms it, so the result is the exact transform of the reference's
// windowed frame; the reference's own fp32 FFT sits ~1e-5 away from it in log-mel units - SURVEY.md §7 "mel tolerance is
// tight").  With n = 25 n1 + n2 and k = k1 + 16 k2:
//   A.  Y[n2][k1] = sum_n1 x[25 n1 + n2] W16^(n1 k1)       16-point DFT of a REAL sequence: only k1 = 0..8 are evaluated
//       T[n2][k1] = Y[n2][k1] W400^(n2 k1)                  (four real 4-point DFTs + 3 complex MACs per k1)
//   B.  X[k1 + 16 k2] = sum_n2 T[n2][k1] W25^(n2 k2)        25-point DFT as 5 x 5 (n2 = 5a + b, k2 = c + 5d):
//       B1. Z[b][c] = W25^(bc) sum_a T[5a + b] W5^(ac)      in place (a thread owns the five slots of residue b)
//       B2. X[c + 5d] = sum_b Z[b][c] W5^(bd)
// Bins k with k mod 16 in 9..15 are never formed: |X[k]|^2 = |X[400 - k]|^2 for real input and (400 - k) mod 16 is in 1..7.
// Only bins 1..199 are needed (columns 0 and 200 of the slaney filterbank are zero).  ~12.5k fp64 FMAs per frame instead of
// the 39.6k (+ 50% twiddle rotations) of the twice-folded direct DFT this kernel used in round 1 (~8.6k with the conjugate-pair
// 5-point transforms below).
#include "common.cuh"
#include "kernels.h"

namespace wat {

constexpr int MEL_F = 10;                          // frames per CTA (even: the mel projection splits them in two halves)
constexpr int MEL_THREADS = 256;
constexpr int MEL_SPAN = MEL_F * 160 + 240;        // samples a CTA touches
constexpr int MEL_BINS = 200;                      // bins 0..199 (200 has zero weight)
constexpr int MEL_K1 = 9;                          // k1 = 0..8 of the 16-point stage

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  // monotone int encoding: works for mixed signs
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, fma(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {          // c + a b
  return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

// out[c] = sum_a in[a] W5^(a c), W5^j = w5[j] (j = 0..4), in conjugate-pair form: W5^((5-a) c) = conj(W5^(a c)), so with
// s_a = in[a] + in[5-a], d_a = in[a] - in[5-a] (a = 1, 2) the terms of a pair are s_a Re(W) + i d_a Im(W):
//   out[1], out[4] = A1 +- i B1,  A1 = in0 + c1 s1 + c2 s2,  B1 = y1 d1 + y2 d2        (c_k, y_k = Re, Im of w5[k])
//   out[2], out[3] = A2 +- i B2,  A2 = in0 + c2 s1 + c1 s2,  B2 = y2 d1 - y1 d2
// 36 fp64 operations instead of the 80 of the plain 5 x 4 complex multiply-adds.
__device__ __forceinline__ void dft5(const double2 (&in)[5], const double2 (&w5)[5], double2 (&out)[5]) {
  const double c1 = w5[1].x, y1 = w5[1].y, c2 = w5[2].x, y2 = w5[2].y;
  const double2 s1 = make_double2(in[1].x + in[4].x, in[1].y + in[4].y), d1 = make_double2(in[1].x - in[4].x, in[1].y - in[4].y);
  const double2 s2 = make_double2(in[2].x + in[3].x, in[2].y + in[3].y), d2 = make_double2(in[2].x - in[3].x, in[2].y - in[3].y);
  out[0] = make_double2(in[0].x + (s1.x + s2.x), in[0].y + (s1.y + s2.y));
  const double2 A1 = make_double2(fma(c2, s2.x, fma(c1, s1.x, in[0].x)), fma(c2, s2.y, fma(c1, s1.y, in[0].y)));
  const double2 A2 = make_double2(fma(c1, s2.x, fma(c2, s1.x, in[0].x)), fma(c1, s2.y, fma(c2, s1.y, in[0].y)));
  const double2 B1 = make_double2(fma(y2, d2.x, y1 * d1.x), fma(y2, d2.y, y1 * d1.y));
  const double2 B2 = make_double2(fma(-y1, d2.x, y2 * d1.x), fma(-y1, d2.y, y2 * d1.y));
  out[1] = make_double2(A1.x - B1.y, A1.y + B1.x);               // A + i B, i B = (-B.y, B.x)
  out[4] = make_double2(A1.x + B1.y, A1.y - B1.x);
  out[2] = make_double2(A2.x - B2.y, A2.y + B2.x);
  out[3] = make_double2(A2.x + B2.y, A2.y - B2.x);
}

// grid: (ceil(n_frames/MEL_F), B).  logspec: [B, frames_alloc, n_mels] (frames >= n_store only feed the max)
__global__ void __launch_bounds__(MEL_THREADS, 3)
mel_power_kernel(const void* __restrict__ pcm, int pcm_i16, long long clip_stride, const int* __restrict__ n_valid_arr,
                 int n_valid_all, int n_pad, int n_frames, int n_store, int frames_alloc, int n_mels,
                 const double2* __restrict__ twiddle,      // [400] (cos, sin)(2 pi j / 400)
                 const float* __restrict__ window,         // [400] periodic hann
                 const int* __restrict__ fb_start,         // [n_mels] first bin of each filter
                 const int* __restrict__ fb_off,           // [n_mels+1] offsets into fb_w
                 const float* __restrict__ fb_w,           // packed non-zero weights
                 float* __restrict__ logspec, float* __restrict__ clip_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* tw = reinterpret_cast<double2*>(smem_raw);                     // [400] W400^j = exp(-2 pi i j / 400)
  double2* Tb = tw + 400;                                                 // [MEL_F][MEL_K1][25]
  float* xs = reinterpret_cast<float*>(Tb + MEL_F * MEL_K1 * 25);         // [MEL_SPAN]
  float* pw = xs + MEL_SPAN;                                              // [MEL_F][MEL_BINS]
  __shared__ float s_max[MEL_THREADS / 32];

  const int clip = blockIdx.y;
  const int t0 = blockIdx.x * MEL_F;
  const int tid = threadIdx.x;
  const int n_valid = n_valid_arr ? n_valid_arr[clip] : n_valid_all;
  const int n_total = n_valid + n_pad;
  // fp32 samples, or int16 PCM scaled by 1/32768 exactly as the reference's load_audio does (audio.py:63)
  const float* xf = reinterpret_cast<const float*>(pcm) + (long long)clip * clip_stride;
  const short* xi = reinterpret_cast<const short*>(pcm) + (long long)clip * clip_stride;

  const int s0 = t0 * 160 - 200;
  {
    // every load of the thread is issued before the first store (the loop used to wait for each sample in turn: 24% of the
    // kernel's stall samples sat on that store)
    constexpr int NLD = (MEL_SPAN + MEL_THREADS - 1) / MEL_THREADS;
    float v[NLD];
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
      const int i = tid + k * MEL_THREADS;
      int s = s0 + i;
      if (s < 0) s = -s;                                                  // reflect about sample 0
      if (s >= n_total) s = 2 * (n_total - 1) - s;                        // reflect about the padded end
      float smp = 0.f;                                                    // appended `padding` samples are zero
      if (i < MEL_SPAN && s >= 0 && s < n_valid) smp = pcm_i16 ? (float)__ldg(xi + s) * (1.0f / 32768.0f) : __ldg(xf + s);
      v[k] = smp;
    }
#pragma unroll
    for (int k = 0; k < NLD; ++k) {
      const int i = tid + k * MEL_THREADS;
      if (i < MEL_SPAN) xs[i] = v[k];
    }
  }
  for (int i = tid; i < 400; i += MEL_THREADS) { const double2 t = twiddle[i]; tw[i] = make_double2(t.x, -t.y); }
  for (int i = tid; i < MEL_F; i += MEL_THREADS) pw[i * MEL_BINS] = 0.f;   // bin 0 is never formed (zero filter weight)
  __syncthreads();

  // ---- stage A: thread = (frame, n2): 16-point DFT over n1 of the real sequence x[25 n1 + n2], outputs k1 = 0..8
  for (int item = tid; item < MEL_F * 25; item += MEL_THREADS) {
    const int f = item / 25, n2 = item - f * 25;
    const float* fr = xs + f * 160;
    double e0[4], e2[4];                                                  // E_r[0], E_r[2] (real);  E_r[1] = (d0[r], -d1[r]), E_r[3] = conj
    double d0[4], d1[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {                                         // real 4-point DFT of x[25 (4m + r) + n2], m = 0..3
      const double x0 = (double)__fmul_rn(fr[25 * r + n2], window[25 * r + n2]);
      const double x1 = (double)__fmul_rn(fr[25 * (r + 4) + n2], window[25 * (r + 4) + n2]);
      const double x2 = (double)__fmul_rn(fr[25 * (r + 8) + n2], window[25 * (r + 8) + n2]);
      const double x3 = (double)__fmul_rn(fr[25 * (r + 12) + n2], window[25 * (r + 12) + n2]);
      const double sa = x0 + x2, sb = x1 + x3;
      e0[r] = sa + sb; e2[r] = sa - sb; d0[r] = x0 - x2; d1[r] = x1 - x3;
    }
    double2* dst = Tb + (f * MEL_K1) * 25 + n2;
#pragma unroll
    for (int k1 = 0; k1 < MEL_K1; ++k1) {
      const int q = k1 & 3;
      double2 acc;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const double2 er = q == 0 ? make_double2(e0[r], 0.0) : q == 2 ? make_double2(e2[r], 0.0)
                         : q == 1 ? make_double2(d0[r], -d1[r]) : make_double2(d0[r], d1[r]);
        if (r == 0) acc = er;
        else acc = cfma(er, tw[(25 * r * k1) % 400], acc);                 // W16^(r k1) = W400^(25 r k1)
      }
      dst[k1 * 25] = cmul(acc, tw[n2 * k1]);
    }
  }
  __syncthreads();
  double2 w5[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) w5[j] = tw[80 * j];
  // ---- stage B1: thread = (frame, k1, b): 5-point DFT over a of T[5a + b], times W25^(b c), in place
  for (int item = tid; item < MEL_F * MEL_K1 * 5; item += MEL_THREADS) {
    const int b = item % 5, fk = item / 5;
    double2* base = Tb + fk * 25 + b;
    double2 in[5], out[5];
#pragma unroll
    for (int a = 0; a < 5; ++a) in[a] = base[5 * a];
    dft5(in, w5, out);
#pragma unroll
    for (int c = 0; c < 5; ++c) base[5 * c] = c == 0 ? out[0] : cmul(out[c], tw[16 * b * c]);   // 16 b c <= 256
  }
  __syncthreads();
  // ---- stage B2: thread = (frame, k1, c): 5-point DFT over b of Z[b][c] -> bins k1 + 16 (c + 5 d), |X|^2
  for (int item = tid; item < MEL_F * MEL_K1 * 5; item += MEL_THREADS) {
    const int c = item % 5, fk = item / 5;
    const int k1 = fk % MEL_K1, f = fk / MEL_K1;
    const double2* base = Tb + fk * 25 + 5 * c;
    double2 in[5], out[5];
#pragma unroll
    for (int b = 0; b < 5; ++b) in[b] = base[b];
    dft5(in, w5, out);
#pragma unroll
    for (int d = 0; d < 5; ++d) {
      const int k = k1 + 16 * (c + 5 * d);
      const float p = (float)(out[d].x * out[d].x + out[d].y * out[d].y);
      if (k >= 1 && k < MEL_BINS) pw[f * MEL_BINS + k] = p;
      else if (k > 200 && k1 >= 1 && k1 <= 7) pw[f * MEL_BINS + 400 - k] = p;       // |X[400 - k]|^2 = |X[k]|^2 (real input)
    }
  }
  __syncthreads();

  // mel projection: an item = one filter x one half of the CTA's frames, so a weight is fetched once for MEL_F / 2 frames and
  // the frames give independent accumulation chains
  float lmax = -INFINITY;
  constexpr int FH = MEL_F / 2;
  for (int i = tid; i < 2 * n_mels; i += MEL_THREADS) {
    const int half = i >= n_mels ? 1 : 0, m = i - half * n_mels;
    const int b0 = fb_start[m], o0 = fb_off[m], cnt = fb_off[m + 1] - o0;
    const float* p = pw + (half * FH) * MEL_BINS + b0;
    float acc[FH];
#pragma unroll
    for (int f = 0; f < FH; ++f) acc[f] = 0.f;
    for (int j = 0; j < cnt; ++j) {
      const float w = fb_w[o0 + j];
#pragma unroll
      for (int f = 0; f < FH; ++f) acc[f] = fmaf(w, p[f * MEL_BINS + j], acc[f]);
    }
#pragma unroll
    for (int f = 0; f < FH; ++f) {
      const int t = t0 + half * FH + f;
      if (t >= n_frames) continue;
      const float v = log10f(fmaxf(acc[f], 1e-10f));
      lmax = fmaxf(lmax, v);
      if (t < n_store) logspec[((long long)clip * frames_alloc + t) * n_mels + m] = v;
    }
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) s_max[tid >> 5] = lmax;
  __syncthreads();
  if (tid == 0) {
    float m = s_max[0];
    for (int i = 1; i < MEL_THREADS / 32; ++i) m = fmaxf(m, s_max[i]);
    if (m > -INFINITY) atomic_max_float(clip_max + clip, m);
  }
}

__global__ void fill_kernel(float* p, float v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// clamp_scope 1: every clip uses the maximum over the whole batch
__global__ void share_max_kernel(float* clip_max, int B) {
  __shared__ float s[32];
  float m = -INFINITY;
  for (int i = threadIdx.x; i < B; i += blockDim.x) m = fmaxf(m, clip_max[i]);
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : -INFINITY;
    m = warp_max(m);
    s[0] = m;
  }
  __syncthreads();
  m = s[0];
  for (int i = threadIdx.x; i < B; i += blockDim.x) clip_max[i] = m;
}

cudaError_t launch_share_max(float* clip_max, int B, cudaStream_t st) {
  share_max_kernel<<<1, 256, 0, st>>>(clip_max, B);
  return cudaGetLastError();
}

// Normalise and lay out.  mode 0: fp32 channel-major [B, n_mels, n_store] (the reference's layout)
//                         mode 1: fp32 time-major [B, n_store, n_mels] (input of the conv1 im2col, fp32 mode)
//                         mode 2: bf16 time-major [B, n_store, n_mels] (bf16 mode)
// grid: (ceil(n_store/32), B), block (32, 8)
__global__ void mel_norm_kernel(const float* __restrict__ logspec, const float* __restrict__ clip_max, int n_store,
                                int frames_alloc, int n_mels, int mode, void* __restrict__ out) {
  __shared__ float tile[32][129];
  const int clip = blockIdx.y;
  const int t0 = blockIdx.x * 32;
  const float floor_v = clip_max[clip] - 8.0f;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int t = t0 + r;
    for (int m = threadIdx.x; m < n_mels; m += 32) {
      float v = -INFINITY;
      if (t < n_store) {
        v = logspec[((long long)clip * frames_alloc + t) * n_mels + m];
        v = (fmaxf(v, floor_v) + 4.0f) / 4.0f;
      }
      tile[r][m] = v;
    }
  }
  __syncthreads();
  if (mode == 0) {
    float* o = reinterpret_cast<float*>(out) + (long long)clip * n_mels * n_store;
    for (int m = threadIdx.y; m < n_mels; m += blockDim.y) {
      const int t = t0 + threadIdx.x;
      if (t < n_store) o[(long long)m * n_store + t] = tile[threadIdx.x][m];
    }
  } else {
    const long long rows = n_store;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      const int t = t0 + r;
      if (t >= n_store) continue;
      const long long off = ((long long)clip * rows + t) * n_mels;
      for (int m = threadIdx.x; m < n_mels; m += 32) {
        if (mode == 1) reinterpret_cast<float*>(out)[off + m] = tile[r][m];
        else reinterpret_cast<__nv_bfloat16*>(out)[off + m] = __float2bfloat16_rn(tile[r][m]);
      }
    }
  }
}

// channel-major fp32 mel [B, n_mels, T] (what callers of Whisper.encoder pass) -> time-major [B, T, n_mels]
// grid: (ceil(T/32), B), block (32, 8)
__global__ void mel_to_timemajor_kernel(const float* __restrict__ mel, int T, int n_mels, int mode, void* __restrict__ out) {
  __shared__ float tile[32][129];
  const int clip = blockIdx.y;
  const int t0 = blockIdx.x * 32;
  const float* src = mel + (long long)clip * n_mels * T;
  for (int m = threadIdx.y; m < n_mels; m += blockDim.y) {
    const int t = t0 + threadIdx.x;
    tile[threadIdx.x][m] = (t < T) ? src[(long long)m * T + t] : 0.f;
  }
  __syncthreads();
  const long long rows = T;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int t = t0 + r;
    if (t >= T) continue;
    const long long off = ((long long)clip * rows + t) * n_mels;
    for (int m = threadIdx.x; m < n_mels; m += 32) {
      if (mode == 1) reinterpret_cast<float*>(out)[off + m] = tile[r][m];
      else reinterpret_cast<__nv_bfloat16*>(out)[off + m] = __float2bfloat16_rn(tile[r][m]);
    }
  }
}

// ------------------------------------------------------------------------------------------ host launchers
size_t mel_power_smem_bytes() {
  return sizeof(double2) * (400 + MEL_F * MEL_K1 * 25) + sizeof(float) * (MEL_SPAN + MEL_F * MEL_BINS);
}

cudaError_t launch_mel_power(const MelTables& tb, const void* pcm, bool pcm_i16, long long clip_stride, const int* n_valid_arr,
                             int n_valid_all, int n_pad, int B, int n_frames, int n_store, int frames_alloc,
                             float* logspec, float* clip_max, cudaStream_t st) {
  static unsigned long long attr_mask = 0;
  const size_t smem = mel_power_smem_bytes();
  if (cudaError_t e = opt_in_smem(mel_power_kernel, (int)smem, attr_mask); e != cudaSuccess) return e;
  fill_kernel<<<(B + 127) / 128, 128, 0, st>>>(clip_max, -INFINITY, B);
  dim3 grid((n_frames + MEL_F - 1) / MEL_F, B);
  mel_power_kernel<<<grid, MEL_THREADS, smem, st>>>(pcm, pcm_i16 ? 1 : 0, clip_stride, n_valid_arr, n_valid_all, n_pad, n_frames, n_store,
                                                    frames_alloc, tb.n_mels, tb.twiddle, tb.window, tb.fb_start,
                                                    tb.fb_off, tb.fb_w, logspec, clip_max);
  return cudaGetLastError();
}

cudaError_t launch_mel_norm(const float* logspec, const float* clip_max, int B, int n_store, int frames_alloc,
                            int n_mels, int mode, void* out, cudaStream_t st) {
  dim3 grid((n_store + 31) / 32, B), block(32, 8);
  mel_norm_kernel<<<grid, block, 0, st>>>(logspec, clip_max, n_store, frames_alloc, n_mels, mode, out);
  return cudaGetLastError();
}

cudaError_t launch_mel_to_timemajor(const float* mel, int B, int T, int n_mels, int mode, void* out, cudaStream_t st) {
  dim3 grid((T + 31) / 32, B), block(32, 8);
  mel_to_timemajor_kernel<<<grid, block, 0, st>>>(mel, T, n_mels, mode, out);
  return cudaGetLastError();
}

}  // namespace wat
