// Fused log-mel front end (reference: package/whisper-at/whisper_at/audio.py:110-157).
//
//   phase 1  mel_power_kernel : framing (centre / reflect, hop 160) -> hann window -> 400-point real DFT
//            -> |X|^2 -> sparse slaney mel projection -> log10(max(.,1e-10)), plus the per-clip maximum.
//   phase 2  mel_norm_kernel  : max(x, clipmax - 8), (x + 4) / 4, written in the layout the consumer wants.
//
// The DFT is evaluated directly (no FFT) with fp64 accumulation over the twice-folded frame.  First fold (real input):
//   Re X[k] = x[0] + (-1)^k x[200] + sum_{n=1..199} E[n] cos(2 pi n k / 400),   E[n] = x[n] + x[400-n]
//   Im X[k] =                      - sum_{n=1..199} O[n] sin(2 pi n k / 400),   O[n] = x[n] - x[400-n]
// second fold (n <-> 200-n, cos/sin pick up (-1)^k): for n = 1..99 only
//   sum E cos = E[100] cos(pi k/2) + sum_n (E[n] + (-1)^k E[200-n]) cos(.),  sum O sin = O[100] sin(pi k/2) + sum_n (O[n] - (-1)^k O[200-n]) sin(.)
// so every bin needs 2 x 99 DFMA instead of 2 x 199; even and odd bins read their own folded sequences.
// so the result is the exact transform of the fp32 windowed frame (the reference's own fp32 FFT is
// ~1e-5 away from it in log-mel units; SURVEY.md §7 "mel tolerance is tight").  Only bins 1..199 are
// evaluated: columns 0 and 200 of the slaney filterbank are zero.
#include "common.cuh"
#include "kernels.h"

namespace wat {

constexpr int MEL_F = 20;                          // frames per CTA
constexpr int MEL_THREADS = 256;
constexpr int MEL_SPAN = MEL_F * 160 + 240;        // samples a CTA touches
constexpr int MEL_BINS = 200;                      // bins 0..199 (200 has zero weight)

struct __align__(16) EO { double e, o; };

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  // monotone int encoding: works for mixed signs
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// grid: (ceil(n_frames/MEL_F), B).  logspec: [B, frames_alloc, n_mels] (frames >= n_store only feed the max)
__global__ void __launch_bounds__(MEL_THREADS, 2)
mel_power_kernel(const void* __restrict__ pcm, int pcm_i16, long long clip_stride, const int* __restrict__ n_valid_arr,
                 int n_valid_all, int n_pad, int n_frames, int n_store, int frames_alloc, int n_mels,
                 const double2* __restrict__ twiddle,      // [400] (cos, sin)(2 pi j / 400)
                 const float* __restrict__ window,         // [400] periodic hann
                 const int* __restrict__ fb_start,         // [n_mels] first bin of each filter
                 const int* __restrict__ fb_off,           // [n_mels+1] offsets into fb_w
                 const float* __restrict__ fb_w,           // packed non-zero weights
                 float* __restrict__ logspec, float* __restrict__ clip_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EO* eo = reinterpret_cast<EO*>(smem_raw);                               // [2 parities][MEL_F][100]
  float* xs = reinterpret_cast<float*>(eo + 2 * MEL_F * 100);             // [MEL_SPAN]
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
  for (int i = tid; i < MEL_SPAN; i += MEL_THREADS) {
    int s = s0 + i;
    if (s < 0) s = -s;                                                    // reflect about sample 0
    if (s >= n_total) s = 2 * (n_total - 1) - s;                          // reflect about the padded end
    float smp = 0.f;                                                      // appended `padding` samples are zero
    if (s >= 0 && s < n_valid) smp = pcm_i16 ? (float)__ldg(xi + s) * (1.0f / 32768.0f) : __ldg(xf + s);
    xs[i] = smp;
  }
  __syncthreads();
  // fold twice.  eo[(par * MEL_F + f) * 100 + n], n = 1..99: (E[n] +/- E[200-n], O[n] -/+ O[200-n]) for bins of parity par;
  // slot n = 0 keeps the per-frame specials: par 0 -> (x[0] + x[200], E[100]);  par 1 -> (x[0] - x[200], O[100])
  for (int i = tid; i < MEL_F * 100; i += MEL_THREADS) {
    const int f = i / 100, n = i - f * 100;
    const float* fr = xs + f * 160;
    auto xw = [&](int j) { return (double)__fmul_rn(fr[j], window[j]); };
    EO ev, od;
    if (n == 0) {
      const double x0 = xw(0), x200 = xw(200), a = xw(100), b = xw(300);
      ev.e = x0 + x200; ev.o = a + b;                               // E[100]
      od.e = x0 - x200; od.o = a - b;                               // O[100]
    } else {
      const double a = xw(n), b = xw(400 - n), c = xw(200 - n), d = xw(200 + n);
      const double En = a + b, On = a - b, Em = c + d, Om = c - d;  // m = 200 - n
      ev.e = En + Em; ev.o = On - Om;
      od.e = En - Em; od.o = On + Om;
    }
    eo[(0 * MEL_F + f) * 100 + n] = ev;
    eo[(1 * MEL_F + f) * 100 + n] = od;
  }
  __syncthreads();

  // DFT: thread = 4 bins of one parity x 4 frames.  (25 even + 25 odd bin groups) x 5 frame groups = 250 threads.
  if (tid < 250) {
    const int bg = tid % 50, fg = tid / 50;
    const int par = bg >= 25;
    const int kbase = 8 * (bg - 25 * par) + par;                   // bins kbase, kbase+2, kbase+4, kbase+6
    double re[4][4], im[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) { re[a][b] = 0.0; im[a][b] = 0.0; }
    // twiddles w = e^(2 pi i n k / 400) by an fp64 rotation per step instead of a table gather: a gather over bins that
    // are 8 apart puts every lane of a warp on the same shared-memory banks (25-way conflicts made this kernel 6x slower
    // than its DFMA count); 99 rotations accumulate ~1e-14 of error, far below the fp32 inputs.
    double2 w[4], r[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) { r[a] = twiddle[kbase + 2 * a]; w[a] = r[a]; }
    const EO* base = eo + (par * MEL_F + fg * 4) * 100;
    for (int n = 1; n < 100; ++n) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const EO v = base[b * 100 + n];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          re[a][b] = fma(v.e, w[a].x, re[a][b]);
          im[a][b] = fma(v.o, w[a].y, im[a][b]);
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const double wx = w[a].x * r[a].x - w[a].y * r[a].y;
        w[a].y = fma(w[a].x, r[a].y, w[a].y * r[a].x);
        w[a].x = wx;
      }
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const EO sp = base[b * 100];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int k = kbase + 2 * a;
        double r, i2;
        if (!par) {                                                 // even k = 2m: cos(pi k/2) = (-1)^m, sin = 0
          r = re[a][b] + sp.e + (((k >> 1) & 1) ? -sp.o : sp.o);
          i2 = im[a][b];
        } else {                                                    // odd k: cos = 0, sin(pi k/2) = (-1)^((k-1)/2)
          r = re[a][b] + sp.e;
          i2 = im[a][b] + ((((k - 1) >> 1) & 1) ? -sp.o : sp.o);
        }
        const double p = r * r + i2 * i2;
        if (k < MEL_BINS) pw[(fg * 4 + b) * MEL_BINS + k] = (float)p;
      }
    }
  }
  __syncthreads();

  float lmax = -INFINITY;
  for (int i = tid; i < MEL_F * n_mels; i += MEL_THREADS) {
    const int f = i / n_mels, m = i - f * n_mels;
    const int t = t0 + f;
    if (t >= n_frames) continue;
    const int b0 = fb_start[m], o0 = fb_off[m], cnt = fb_off[m + 1] - o0;
    const float* p = pw + f * MEL_BINS + b0;
    float acc = 0.f;
    for (int j = 0; j < cnt; ++j) acc = fmaf(fb_w[o0 + j], p[j], acc);
    const float v = log10f(fmaxf(acc, 1e-10f));
    lmax = fmaxf(lmax, v);
    if (t < n_store) logspec[((long long)clip * frames_alloc + t) * n_mels + m] = v;
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
  return sizeof(EO) * 2 * MEL_F * 100 + sizeof(float) * (MEL_SPAN + MEL_F * MEL_BINS);
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
