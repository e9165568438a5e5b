// fp32 SIMT kernels: the fp32 correctness mode of the encoder/head GEMMs and attention, plus the
// memory-bound glue every mode uses (LayerNorm, pooling, TL-TR window regroup, short-sequence attention).
// References: package/whisper-at/whisper_at/model.py:29-31 (LayerNorm fp32), :92-107 (attention),
// :171-174 (20x average pooling), :351-379 (ATModel.forward).
#include "common.cuh"
#include "kernels.h"

namespace wat {

// ------------------------------------------------------------------------------------------ fp32 GEMM
constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmF32 g) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Ws[SG_BK][SG_BN + 4];
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const float* A = g.A;
  const int tid = threadIdx.x;
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += SG_BK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), wv = av;
    if (m0 + lr < g.M && k0 + lk < g.K) av = *reinterpret_cast<const float4*>(A + (long long)(m0 + lr) * g.lda + k0 + lk);
    if (n0 + lr < g.N && k0 + lk < g.K) wv = *reinterpret_cast<const float4*>(g.W + (long long)(n0 + lr) * g.K + k0 + lk);
    __syncthreads();
    As[lk + 0][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
    Ws[lk + 0][lr] = wv.x; Ws[lk + 1][lr] = wv.y; Ws[lk + 2][lr] = wv.z; Ws[lk + 3][lr] = wv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[n];
      if (g.act == 1) v = gelu_erf(v);
      if (g.R) v += g.R[(long long)(g.r_mod > 0 ? m % g.r_mod : m) * g.ldr + n];
      g.C[(long long)m * g.ldc + n] = v;
    }
  }
}

cudaError_t launch_gemm_f32(const GemmF32& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  if ((g.K & 3) || (g.lda & 3)) return cudaErrorInvalidValue;
  dim3 grid((g.N + SG_BN - 1) / SG_BN, (g.M + SG_BM - 1) / SG_BM, 1);
  gemm_f32_kernel<<<grid, 256, 0, st>>>(g);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ LayerNorm
// rows of the residual stream are fp32 (fp32 mode) or fp16 (bf16 mode): 4 elements at a time
__device__ __forceinline__ float4 ldrow4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldrow4(const __half* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  return f16x4_to_f32(u.x, u.y);
}
__device__ __forceinline__ void strow4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void strow4(__half* p, float4 v) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_f16_sat(v.x, v.y), pack_f16_sat(v.z, v.w));
}

template <typename OutT>
__device__ __forceinline__ void store4(OutT* p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  uint2 u = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  *reinterpret_cast<uint2*>(p) = u;
}

template <typename OutT>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int M, int D, OutT* __restrict__ out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (long long)row * D;
  float s = 0.f;
  for (int e = lane * 4; e < D; e += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + e);
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
  for (int e = lane * 4; e < D; e += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + e);
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-5f);
  OutT* o = out + (long long)row * D;
  for (int e = lane * 4; e < D; e += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + e);
    const float4 gm = *reinterpret_cast<const float4*>(gamma + e);
    const float4 bt = *reinterpret_cast<const float4*>(beta + e);
    float4 y;
    y.x = (v.x - mean) * rstd * gm.x + bt.x;
    y.y = (v.y - mean) * rstd * gm.y + bt.y;
    y.z = (v.z - mean) * rstd * gm.z + bt.z;
    y.w = (v.w - mean) * rstd * gm.w + bt.w;
    store4<OutT>(o + e, y);
  }
}

// same, with the row held in registers (one global read; D a multiple of 128, <= 1280): every load of a row is in
// flight before the first reduction starts
template <typename OutT, typename InT = float>
__global__ void __launch_bounds__(256) layernorm_reg_kernel(const InT* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int M, int D, OutT* __restrict__ out) {
  // rows are walked from the END: the producer GEMM wrote the residual stream front to back, so its last row blocks are
  // the ones still in L2; and the rows normalised last (the first ones) are what the next GEMM reads first
  const int row = (gridDim.x - 1 - blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const int n4 = D >> 7;
  const InT* xr = x + (long long)row * D;
  float4 v[10];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 10; ++i)
    if (i < n4) {
      v[i] = ldrow4(xr + (i * 32 + lane) * 4);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  const float inv_d = 1.0f / (float)D;
  const float mean = warp_sum(s) * inv_d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 10; ++i)
    if (i < n4) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  const float rstd = rsqrtf(warp_sum(q) * inv_d + 1e-5f);
  OutT* o = out + (long long)row * D;
#pragma unroll
  for (int i = 0; i < 10; ++i)
    if (i < n4) {
      const int c = (i * 32 + lane) * 4;
      const float4 gm = *reinterpret_cast<const float4*>(gamma + c);
      const float4 bt = *reinterpret_cast<const float4*>(beta + c);
      float4 y;
      y.x = (v[i].x - mean) * rstd * gm.x + bt.x;
      y.y = (v[i].y - mean) * rstd * gm.y + bt.y;
      y.z = (v[i].z - mean) * rstd * gm.z + bt.z;
      y.w = (v[i].w - mean) * rstd * gm.w + bt.w;
      store4<OutT>(o + c, y);
    }
}

// x rows in fp16 (the bf16 mode's residual stream) -> fp32 LayerNorm output (ln_post for the ASR hand-off); D % 128 == 0, <= 1280
cudaError_t launch_layernorm_f16in(const void* x, const float* gamma, const float* beta, int M, int D, float* out, cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  if ((D & 127) || D > 1280) return cudaErrorInvalidValue;
  layernorm_reg_kernel<float, __half><<<(M + 7) / 8, 256, 0, st>>>((const __half*)x, gamma, beta, M, D, out);
  return cudaGetLastError();
}

cudaError_t launch_layernorm(const float* x, const float* gamma, const float* beta, int M, int D, void* out,
                             bool out_bf16, cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  if (D & 3) return cudaErrorInvalidValue;
  const int grid = (M + 7) / 8;
  if ((D & 127) == 0 && D <= 1280) {
    if (out_bf16) layernorm_reg_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, gamma, beta, M, D, (__nv_bfloat16*)out);
    else layernorm_reg_kernel<float><<<grid, 256, 0, st>>>(x, gamma, beta, M, D, (float*)out);
    return cudaGetLastError();
  }
  if (out_bf16) layernorm_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(x, gamma, beta, M, D, (__nv_bfloat16*)out);
  else layernorm_kernel<float><<<grid, 256, 0, st>>>(x, gamma, beta, M, D, (float*)out);
  return cudaGetLastError();
}

// LayerNorm of 20 consecutive rows (one pooling window of the encoder) + the 20x average pool of the SAME rows
// (model.py:171-174: the pooled state of layer l is the mean of x after block l, i.e. of the input of block l+1's
// first LayerNorm), so the fp32 residual stream is read once for both.  One CTA of 10 warps per window; a warp
// normalises rows w and w + 10 one after the other with the row held in registers (like layernorm_kernel, which runs
// at the HBM roofline) and adds them into its own smem partial of the pool; the partials are summed in a fixed order.
template <typename OutT>
__global__ void __launch_bounds__(320, 3) layernorm_pool20_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, int D, OutT* __restrict__ out,
                                                               float* __restrict__ pooled, int layer, int L, int P) {
  extern __shared__ float ln_part[];                      // [10 warps][D]
  const int g = gridDim.x - 1 - blockIdx.x;               // window index = b * P + w, walked from the end (see layernorm_reg_kernel)
  const int b = g / P, w = g - b * P;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n4 = D >> 7;                                  // float4 per lane per row (D is a multiple of 128, <= 1280)
  float* mine = ln_part + warp * D;
  const float inv_d = 1.0f / (float)D;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const long long row = (long long)g * 20 + warp + half * 10;
    const float* xr = x + row * D;
    float4 v[10];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i)
      if (i < n4) {
        v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i)
      if (i < n4) {
        const float a = v[i].x - mean, bb = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + bb * bb) + (cc * cc + d * d);
      }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + 1e-5f);
    OutT* o = out + row * D;
#pragma unroll
    for (int i = 0; i < 10; ++i)
      if (i < n4) {
        const int c = (i * 32 + lane) * 4;
        const float4 gm = *reinterpret_cast<const float4*>(gamma + c);
        const float4 bt = *reinterpret_cast<const float4*>(beta + c);
        float4 y;
        y.x = (v[i].x - mean) * rstd * gm.x + bt.x;
        y.y = (v[i].y - mean) * rstd * gm.y + bt.y;
        y.z = (v[i].z - mean) * rstd * gm.z + bt.z;
        y.w = (v[i].w - mean) * rstd * gm.w + bt.w;
        store4<OutT>(o + c, y);
        float4* pm = reinterpret_cast<float4*>(mine + c);
        if (half == 0) *pm = v[i];
        else { const float4 p = *pm; *pm = make_float4(p.x + v[i].x, p.y + v[i].y, p.z + v[i].z, p.w + v[i].w); }
      }
  }
  __syncthreads();
  for (int c = threadIdx.x * 4; c < D; c += 1280) {
    float4 a = *reinterpret_cast<const float4*>(ln_part + c);
#pragma unroll
    for (int k = 1; k < 10; ++k) {
      const float4 p = *reinterpret_cast<const float4*>(ln_part + k * D + c);
      a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    }
    *reinterpret_cast<float4*>(pooled + (((long long)b * L + layer) * P + w) * D + c) =
        make_float4(a.x * 0.05f, a.y * 0.05f, a.z * 0.05f, a.w * 0.05f);
  }
}

cudaError_t launch_layernorm_pool20(const float* x, const float* gamma, const float* beta, int B, int T, int D, void* out,
                                    bool out_bf16, float* pooled, int layer, int L, cudaStream_t st) {
  if ((D & 127) || D > 1280 || T % 20) return cudaErrorInvalidValue;
  const int P = T / 20;
  const int grid = B * P, block = 320;
  const size_t smem = (size_t)10 * D * sizeof(float);
  static unsigned long long attr_mask_h = 0, attr_mask_f = 0;
  if (cudaError_t e = opt_in_smem(layernorm_pool20_kernel<__nv_bfloat16>, 10 * 1280 * 4, attr_mask_h); e != cudaSuccess) return e;
  if (cudaError_t e = opt_in_smem(layernorm_pool20_kernel<float>, 10 * 1280 * 4, attr_mask_f); e != cudaSuccess) return e;
  if (out_bf16) layernorm_pool20_kernel<__nv_bfloat16><<<grid, block, smem, st>>>(x, gamma, beta, D, (__nv_bfloat16*)out, pooled, layer, L, P);
  else layernorm_pool20_kernel<float><<<grid, block, smem, st>>>(x, gamma, beta, D, (float*)out, pooled, layer, L, P);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ fp32 attention, hd 64
constexpr int AF_KT = 32;

__global__ void __launch_bounds__(128) attn_f32_hd64_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                            const float* __restrict__ v, long long ld,
                                                            float* __restrict__ out, long long ldo, int T, float c_log2) {
  __shared__ __align__(16) float Ks[AF_KT][64];
  __shared__ __align__(16) float Vs[AF_KT][64];
  const int h = blockIdx.y;
  const long long base = (long long)blockIdx.z * T;
  const int tid = threadIdx.x;
  const int t = blockIdx.x * 128 + tid;
  const bool valid = t < T;
  float qr[64], o[64];
#pragma unroll
  for (int e = 0; e < 64; e += 4) {
    float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) v4 = *reinterpret_cast<const float4*>(q + (base + t) * ld + h * 64 + e);
    qr[e] = v4.x * c_log2; qr[e + 1] = v4.y * c_log2; qr[e + 2] = v4.z * c_log2; qr[e + 3] = v4.w * c_log2;
    o[e] = o[e + 1] = o[e + 2] = o[e + 3] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int kt0 = 0; kt0 < T; kt0 += AF_KT) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 128;
      const int r = idx >> 4, c4 = (idx & 15) * 4;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (kt0 + r < T) {
        kv = *reinterpret_cast<const float4*>(k + (base + kt0 + r) * ld + h * 64 + c4);
        vv = *reinterpret_cast<const float4*>(v + (base + kt0 + r) * ld + h * 64 + c4);
      }
      *reinterpret_cast<float4*>(&Ks[r][c4]) = kv;
      *reinterpret_cast<float4*>(&Vs[r][c4]) = vv;
    }
    __syncthreads();
    float s[AF_KT];
    float mt = -INFINITY;
#pragma unroll
    for (int j = 0; j < AF_KT; ++j) {
      float a = 0.f;
#pragma unroll
      for (int e = 0; e < 64; e += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(&Ks[j][e]);
        a = fmaf(qr[e], k4.x, a); a = fmaf(qr[e + 1], k4.y, a); a = fmaf(qr[e + 2], k4.z, a); a = fmaf(qr[e + 3], k4.w, a);
      }
      s[j] = (kt0 + j < T) ? a : -INFINITY;
      mt = fmaxf(mt, s[j]);
    }
    const float m_new = fmaxf(m, mt);
    const float alpha = exp2f(m - m_new);
    l *= alpha;
#pragma unroll
    for (int e = 0; e < 64; ++e) o[e] *= alpha;
#pragma unroll
    for (int j = 0; j < AF_KT; ++j) {
      const float p = exp2f(s[j] - m_new);
      l += p;
#pragma unroll
      for (int e = 0; e < 64; e += 4) {
        const float4 v4 = *reinterpret_cast<const float4*>(&Vs[j][e]);
        o[e] = fmaf(p, v4.x, o[e]); o[e + 1] = fmaf(p, v4.y, o[e + 1]);
        o[e + 2] = fmaf(p, v4.z, o[e + 2]); o[e + 3] = fmaf(p, v4.w, o[e + 3]);
      }
    }
    m = m_new;
  }
  if (valid) {
    const float inv = 1.0f / l;
    float* op = out + (base + t) * ldo + h * 64;
#pragma unroll
    for (int e = 0; e < 64; e += 4)
      *reinterpret_cast<float4*>(op + e) = make_float4(o[e] * inv, o[e + 1] * inv, o[e + 2] * inv, o[e + 3] * inv);
  }
}

cudaError_t launch_attn_f32_hd64(const float* q, const float* k, const float* v, long long ld, float* out,
                                 long long ldo, int n_seq, int T, int n_head, cudaStream_t st) {
  dim3 grid((T + 127) / 128, n_head, n_seq);
  const float c = 0.125f * 1.4426950408889634f;            // (64^-0.25)^2 * log2(e)
  attn_f32_hd64_kernel<<<grid, 128, 0, st>>>(q, k, v, ld, out, ldo, T, c);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ short-sequence attention
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) attn_small_kernel(const TIn* __restrict__ qkv, TOut* __restrict__ out, int T,
                                                         int n_head, int hd, float scale2) {
  extern __shared__ float S[];                            // [T][T+1]
  const int seq = blockIdx.x, h = blockIdx.y;
  const int D = n_head * hd;
  const long long row0 = (long long)seq * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ldS = T + 1;
  for (int p = warp; p < T * T; p += 8) {
    const int i = p / T, j = p - i * T;
    const TIn* qp = qkv + (row0 + i) * 3 * D + h * hd;
    const TIn* kp = qkv + (row0 + j) * 3 * D + D + h * hd;
    float a = 0.f;
    for (int e = lane * 4; e < hd; e += 128) {
      const float4 q4 = load4(qp + e), k4 = load4(kp + e);
      a = fmaf(q4.x, k4.x, a); a = fmaf(q4.y, k4.y, a); a = fmaf(q4.z, k4.z, a); a = fmaf(q4.w, k4.w, a);
    }
    a = warp_sum(a);
    if (lane == 0) S[i * ldS + j] = a * scale2;
  }
  __syncthreads();
  for (int i = warp; i < T; i += 8) {
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) mx = fmaxf(mx, S[i * ldS + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float p = expf(S[i * ldS + j] - mx);
      S[i * ldS + j] = p;
      sum += p;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < T; j += 32) S[i * ldS + j] *= inv;
  }
  __syncthreads();
  const int hd4 = hd >> 2;
  for (int idx = threadIdx.x; idx < T * hd4; idx += 256) {
    const int i = idx / hd4, e = (idx - i * hd4) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < T; ++j) {
      const float p = S[i * ldS + j];
      const float4 v4 = load4(qkv + (row0 + j) * 3 * D + 2 * D + h * hd + e);
      acc.x = fmaf(p, v4.x, acc.x); acc.y = fmaf(p, v4.y, acc.y); acc.z = fmaf(p, v4.z, acc.z); acc.w = fmaf(p, v4.w, acc.w);
    }
    store4<TOut>(out + (row0 + i) * D + h * hd + e, acc);
  }
}

// bf16 fast path of the head attention for T <= 32 (decision windows of <= 32 pooled frames; <= 32 encoder layers) and
// hd % 16 == 0.  One CTA per (GROUP of sequences, head): G = 32 / T consecutive sequences (their rows are contiguous) share one
// 32 x 32 score tile, whose off-diagonal blocks are masked - at at_time_res = 2 (T = 5) that is 6 sequences per CTA instead
// of a 97%-empty tile each.
//   1. S = Q K^T on mma.sync m16n8k16 (bf16 in, fp32 accumulate; the contraction split over the 8 warps, fragments loaded
//      straight from global memory - every Q / K element is used exactly once), partial tiles summed through smem in a
//      fixed order;
//   2. row softmax in fp32 over the row's own sequence (warp per row);
//   3. O = P V in fp32 FMAs: thread = 4 columns x 16 rows, V read once per row half, P broadcast from smem (zero outside
//      the row's sequence).
// Per-sequence results do not depend on G or on the sequence's position in its group: the masked terms are exact zeros and
// every row sums its own keys in the same order.
// (The tcgen05 path is reserved for the encoder: these tiles are at most 32 x 32 and the whole head attention is ~1% of a step.)
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// PACK = false: one sequence per CTA (T > 16), every bound is T itself
template <bool PACK>
__global__ void __launch_bounds__(256) attn_small_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                             __nv_bfloat16* __restrict__ out, int T, int n_head, int hd,
                                                             float scale2, int n_seq, int G) {
  __shared__ float Sp[8][32][33];                         // per-warp partial scores; Sp[0] is reused for P
  const int seq0 = PACK ? blockIdx.x * G : blockIdx.x, h = blockIdx.y;
  const int R = PACK ? min(G, n_seq - seq0) * T : T;       // rows (= keys) of this tile
  const int D = n_head * hd;
  const long long row0 = (long long)seq0 * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const __nv_bfloat16* qb = qkv + row0 * 3 * D + h * hd;
  const __nv_bfloat16* kb = qb + D;
  const __nv_bfloat16* vb = qb + 2 * D;
  {
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll 2
    for (int ks = warp; ks < (hd >> 4); ks += 8) {
      const int k0 = ks * 16 + tig * 2;
      uint32_t a[2][4], b[4][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r0 = mt * 16 + g, r1 = r0 + 8;
        const uint32_t* p0 = reinterpret_cast<const uint32_t*>(qb + (long long)r0 * 3 * D + k0);
        const uint32_t* p1 = reinterpret_cast<const uint32_t*>(qb + (long long)r1 * 3 * D + k0);
        a[mt][0] = r0 < R ? p0[0] : 0u;
        a[mt][1] = r1 < R ? p1[0] : 0u;
        a[mt][2] = r0 < R ? p0[4] : 0u;                   // columns + 8
        a[mt][3] = r1 < R ? p1[4] : 0u;
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int n = nt * 8 + g;
        const uint32_t* pk = reinterpret_cast<const uint32_t*>(kb + (long long)n * 3 * D + k0);
        b[nt][0] = n < R ? pk[0] : 0u;
        b[nt][1] = n < R ? pk[4] : 0u;
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[mt][nt], a[mt], b[nt]);
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int r = mt * 16 + g, c = nt * 8 + tig * 2;
        Sp[warp][r][c] = acc[mt][nt][0];
        Sp[warp][r][c + 1] = acc[mt][nt][1];
        Sp[warp][r + 8][c] = acc[mt][nt][2];
        Sp[warp][r + 8][c + 1] = acc[mt][nt][3];
      }
  }
  __syncthreads();
  for (int i = warp; i < R; i += 8) {                     // row i is only ever touched by this warp from here on
    const int j0 = PACK ? (i / T) * T : 0;                 // keys of row i's own sequence: [j0, j0 + T)
    const bool mine = lane >= j0 && lane < j0 + T;
    float sc = -INFINITY;
    if (mine) {
      float a = Sp[0][i][lane];
#pragma unroll
      for (int k = 1; k < 8; ++k) a += Sp[k][i][lane];
      sc = a * scale2;
    }
    const float mx = warp_max(sc);
    const float p = mine ? expf(sc - mx) : 0.f;
    const float inv = 1.0f / warp_sum(p);
    Sp[0][i][lane] = p * inv;                              // exact zeros outside the sequence (and for lanes >= R)
  }
  __syncthreads();
  const int ncg = hd >> 2, halves = (R + 15) >> 4;
  for (int idx = threadIdx.x; idx < ncg * halves; idx += 256) {
    const int half = idx / ncg, e = (idx - half * ncg) * 4;
    const int rbase = half * 16;
    // keys that any of the 16 rows of this half can see: the sequences overlapping [rbase, rbase + 16)
    const int jlo = PACK ? (rbase / T) * T : 0, jhi = PACK ? min(R, ((min(rbase + 15, R - 1)) / T + 1) * T) : T;
    float4 acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 5
    for (int j = jlo; j < jhi; ++j) {
      const float4 v4 = load4(vb + (long long)j * 3 * D + e);
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const float p = Sp[0][rbase + r][j];
        acc[r].x = fmaf(p, v4.x, acc[r].x); acc[r].y = fmaf(p, v4.y, acc[r].y);
        acc[r].z = fmaf(p, v4.z, acc[r].z); acc[r].w = fmaf(p, v4.w, acc[r].w);
      }
    }
#pragma unroll
    for (int r = 0; r < 16; ++r)
      if (rbase + r < R) store4<__nv_bfloat16>(out + (row0 + rbase + r) * D + h * hd + e, acc[r]);
  }
}

cudaError_t launch_attn_small(const void* qkv, bool in_bf16, void* out, bool out_bf16, int n_seq, int T, int n_head,
                              int hd, cudaStream_t st) {
  if (n_seq <= 0) return cudaSuccess;
  if (T > 128 || (hd & 3)) return cudaErrorInvalidValue;
  dim3 grid(n_seq, n_head);
  const size_t smem = sizeof(float) * T * (T + 1);
  const float scale2 = 1.0f / sqrtf((float)hd);           // (hd^-0.25)^2
  if (smem > 48 * 1024) {
    static unsigned long long attr_mask_f = 0, attr_mask_h = 0;
    if (cudaError_t e = opt_in_smem(attn_small_kernel<float, float>, 80 * 1024, attr_mask_f); e != cudaSuccess) return e;
    if (cudaError_t e = opt_in_smem(attn_small_kernel<__nv_bfloat16, __nv_bfloat16>, 80 * 1024, attr_mask_h); e != cudaSuccess) return e;
  }
  if (in_bf16 && out_bf16 && T <= 32 && (hd & 15) == 0) {
    const int G = 32 / T;                                  // sequences per 32 x 32 score tile
    dim3 grid_g((n_seq + G - 1) / G, n_head);
    if (G > 1) attn_small_mma_kernel<true><<<grid_g, 256, 0, st>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, T, n_head, hd, scale2, n_seq, G);
    else attn_small_mma_kernel<false><<<grid_g, 256, 0, st>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, T, n_head, hd, scale2, n_seq, 1);
    return cudaGetLastError();
  }
  if (!in_bf16 && !out_bf16)
    attn_small_kernel<float, float><<<grid, 256, smem, st>>>((const float*)qkv, (float*)out, T, n_head, hd, scale2);
  else if (in_bf16 && out_bf16)
    attn_small_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, smem, st>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, T, n_head, hd, scale2);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ means / pooling / regroup
// The row producers of the TL-TR head that are not GEMMs (window regroup, layer reduction, group mean) can also leave what a
// LayerNorm folded into the next GEMM needs (gemm_tc.cu): the row's bf16 copy and its (sum, sum of squares) as ONE slice.
// One 128-thread CTA per output row; RowTail collects the thread's share and reduces it in a fixed order.
struct RowTail {
  float s = 0.f, q = 0.f;
  __device__ __forceinline__ void put(float4 v, int e, __nv_bfloat16* xb_row) {
    s += (v.x + v.y) + (v.z + v.w);
    q = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, q))));
    if (xb_row) *reinterpret_cast<uint2*>(xb_row + e) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
  __device__ __forceinline__ void finish(float2* stats_row) {      // every thread of the 128-thread CTA calls it
    __shared__ float2 part[4];
    const float ws = warp_sum(s), wq = warp_sum(q);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = make_float2(ws, wq);
    __syncthreads();
    if (threadIdx.x == 0)
      *stats_row = make_float2((part[0].x + part[1].x) + (part[2].x + part[3].x), (part[0].y + part[1].y) + (part[2].y + part[3].y));
  }
};

template <typename InT, typename OutT>
__global__ void __launch_bounds__(128) group_mean_kernel(const InT* __restrict__ x, int win, int D, OutT* __restrict__ out,
                                                         long long out_stride, __nv_bfloat16* __restrict__ xb, float2* __restrict__ stats) {
  const int g = blockIdx.x;
  const float inv = 1.0f / (float)win;
  RowTail tail;
  for (int e = threadIdx.x * 4; e < D; e += blockDim.x * 4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < win; ++r) {
      const float4 v = ldrow4(x + ((long long)g * win + r) * D + e);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    a = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
    strow4(out + (long long)g * out_stride + e, a);
    if (stats) tail.put(a, e, xb ? xb + (long long)g * D : nullptr);
  }
  if (stats) tail.finish(stats + g);
}

cudaError_t launch_group_mean(const void* x, bool in_f16, int n_groups, int win, int D, void* out, bool out_f16, long long out_stride,
                              cudaStream_t st, __nv_bfloat16* xb, float* stats) {
  if (n_groups <= 0) return cudaSuccess;
  float2* s2 = reinterpret_cast<float2*>(stats);
  if (!in_f16 && !out_f16) group_mean_kernel<float, float><<<n_groups, 128, 0, st>>>((const float*)x, win, D, (float*)out, out_stride, xb, s2);
  else if (in_f16 && out_f16) group_mean_kernel<__half, __half><<<n_groups, 128, 0, st>>>((const __half*)x, win, D, (__half*)out, out_stride, xb, s2);
  else if (in_f16) group_mean_kernel<__half, float><<<n_groups, 128, 0, st>>>((const __half*)x, win, D, (float*)out, out_stride, xb, s2);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

__global__ void pool20_kernel(const float* __restrict__ x, int T, int D, int layer, int L, float* __restrict__ pooled) {
  const int P = T / 20;
  const int g = blockIdx.x;                                // b * P + w
  const int b = g / P, w = g - b * P;
  const float* src = x + ((long long)b * T + (long long)w * 20) * D;
  float* dst = pooled + (((long long)b * L + layer) * P + w) * D;
  for (int e = threadIdx.x * 4; e < D; e += blockDim.x * 4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int r = 0; r < 20; ++r) {
      const float4 v = *reinterpret_cast<const float4*>(src + (long long)r * D + e);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    *reinterpret_cast<float4*>(dst + e) = make_float4(a.x * 0.05f, a.y * 0.05f, a.z * 0.05f, a.w * 0.05f);
  }
}

cudaError_t launch_pool20(const float* x, int B, int T, int D, int layer, int L, float* pooled, cudaStream_t st) {
  pool20_kernel<<<B * (T / 20), 128, 0, st>>>(x, T, D, layer, L, pooled);
  return cudaGetLastError();
}

// thread = 8 columns of one pooling window: 20 rows of 16-byte loads, all in flight before the first add
__global__ void __launch_bounds__(160) pool20_bf16_kernel(const __nv_bfloat16* __restrict__ xb, int T, int D, int layer, int L,
                                                          float* __restrict__ pooled) {
  const int P = T / 20;
  const int g = gridDim.x - 1 - blockIdx.x;                // walked from the end: the producer GEMM wrote the last rows last (L2)
  const int b = g / P, w = g - b * P;
  const __nv_bfloat16* src = xb + ((long long)b * T + (long long)w * 20) * D;
  float* dst = pooled + (((long long)b * L + layer) * P + w) * D;
  for (int e = threadIdx.x * 8; e < D; e += blockDim.x * 8) {
    uint4 v[20];
#pragma unroll
    for (int r = 0; r < 20; ++r) v[r] = *reinterpret_cast<const uint4*>(src + (long long)r * D + e);
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 0.f;
#pragma unroll
    for (int r = 0; r < 20; ++r) {
      const uint32_t u[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        a[2 * k] += __uint_as_float(u[k] << 16);             // bf16 -> fp32 is a shift
        a[2 * k + 1] += __uint_as_float(u[k] & 0xffff0000u);
      }
    }
    *reinterpret_cast<float4*>(dst + e) = make_float4(a[0] * 0.05f, a[1] * 0.05f, a[2] * 0.05f, a[3] * 0.05f);
    *reinterpret_cast<float4*>(dst + e + 4) = make_float4(a[4] * 0.05f, a[5] * 0.05f, a[6] * 0.05f, a[7] * 0.05f);
  }
}

cudaError_t launch_pool20_bf16(const __nv_bfloat16* xb, int B, int T, int D, int layer, int L, float* pooled, cudaStream_t st) {
  if ((D & 7) || T % 20) return cudaErrorInvalidValue;
  pool20_bf16_kernel<<<B * (T / 20), 160, 0, st>>>(xb, T, D, layer, L, pooled);
  return cudaGetLastError();
}

template <typename OutT>
__global__ void __launch_bounds__(128) head_gather_kernel(const float* __restrict__ pooled, int L, int Tp_total, int t_start, int Tp, int dw,
                                   int S, int D, OutT* __restrict__ out, __nv_bfloat16* __restrict__ xb, float2* __restrict__ stats) {
  // out row = ((b*S + s)*L + l)*dw + tau
  long long r = blockIdx.x;
  const int tau = (int)(r % dw); r /= dw;
  const int l = (int)(r % L); r /= L;
  const int s = (int)(r % S);
  const int b = (int)(r / S);
  const int t = s * dw + tau;
  OutT* o = out + (long long)blockIdx.x * D;
  __nv_bfloat16* xr = xb ? xb + (long long)blockIdx.x * D : nullptr;
  RowTail tail;
  if (t < Tp) {
    const float* src = pooled + (((long long)b * L + l) * Tp_total + t_start + t) * D;
    for (int e = threadIdx.x * 4; e < D; e += blockDim.x * 4) {
      const float4 v = *reinterpret_cast<const float4*>(src + e);
      strow4(o + e, v);
      if (stats) tail.put(v, e, xr);
    }
  } else {                                                        // zero rows of a ragged last window: LN of zeros = beta
    for (int e = threadIdx.x * 4; e < D; e += blockDim.x * 4) {
      strow4(o + e, make_float4(0.f, 0.f, 0.f, 0.f));
      if (stats) tail.put(make_float4(0.f, 0.f, 0.f, 0.f), e, xr);
    }
  }
  if (stats) tail.finish(stats + blockIdx.x);
}

cudaError_t launch_head_gather(const float* pooled, int B, int L, int Tp_total, int t_start, int Tp, int dw, int S, int D,
                               void* out, bool out_f16, cudaStream_t st, __nv_bfloat16* xb, float* stats) {
  const long long rows = (long long)B * S * L * dw;
  if (rows <= 0) return cudaSuccess;
  if (out_f16) head_gather_kernel<__half><<<(unsigned)rows, 128, 0, st>>>(pooled, L, Tp_total, t_start, Tp, dw, S, D, (__half*)out, xb, reinterpret_cast<float2*>(stats));
  else head_gather_kernel<float><<<(unsigned)rows, 128, 0, st>>>(pooled, L, Tp_total, t_start, Tp, dw, S, D, (float*)out, xb, reinterpret_cast<float2*>(stats));
  return cudaGetLastError();
}

// Layer reduction of the baseline heads (models.py:113-167): mean over layers, last layer, or the learned layer weights
// divided by their sum; rows of a ragged last window are zero like head_gather's.
template <typename OutT>
__global__ void __launch_bounds__(128) head_layer_reduce_kernel(const float* __restrict__ pooled, int L, int Tp_total, int t_start, int Tp, int dw,
                                         int S, int D, int kind, const float* __restrict__ w, OutT* __restrict__ out,
                                         __nv_bfloat16* __restrict__ xb, float2* __restrict__ stats) {
  long long r = blockIdx.x;
  const int tau = (int)(r % dw); r /= dw;
  const int s = (int)(r % S);
  const int b = (int)(r / S);
  const int t = s * dw + tau;
  OutT* o = out + (long long)blockIdx.x * D;
  __nv_bfloat16* xr = xb ? xb + (long long)blockIdx.x * D : nullptr;
  RowTail tail;
  if (t >= Tp) {
    for (int e = threadIdx.x * 4; e < D; e += blockDim.x * 4) {
      strow4(o + e, make_float4(0.f, 0.f, 0.f, 0.f));
      if (stats) tail.put(make_float4(0.f, 0.f, 0.f, 0.f), e, xr);
    }
    if (stats) tail.finish(stats + blockIdx.x);
    return;
  }
  const float* src = pooled + ((long long)b * L * Tp_total + t_start + t) * D;
  const long long lstride = (long long)Tp_total * D;
  float wsum = 0.f;
  if (kind == 2) for (int l = 0; l < L; ++l) wsum += w[l];
  for (int e = threadIdx.x * 4; e < D; e += blockDim.x * 4) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kind == 1) {
      a = *reinterpret_cast<const float4*>(src + (L - 1) * lstride + e);
    } else {
      for (int l = 0; l < L; ++l) {
        const float4 v = *reinterpret_cast<const float4*>(src + l * lstride + e);
        const float c = kind == 2 ? w[l] : 1.0f;
        a.x = fmaf(c, v.x, a.x); a.y = fmaf(c, v.y, a.y); a.z = fmaf(c, v.z, a.z); a.w = fmaf(c, v.w, a.w);
      }
      const float inv = kind == 2 ? 1.0f / wsum : 1.0f / (float)L;
      a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
    }
    strow4(o + e, a);
    if (stats) tail.put(a, e, xr);
  }
  if (stats) tail.finish(stats + blockIdx.x);
}

cudaError_t launch_head_layer_reduce(const float* pooled, int B, int L, int Tp_total, int t_start, int Tp, int dw, int S, int D,
                                     int kind, const float* w, void* out, bool out_f16, cudaStream_t st, __nv_bfloat16* xb, float* stats) {
  const long long rows = (long long)B * S * dw;
  if (rows <= 0) return cudaSuccess;
  if (kind == 2 && !w) return cudaErrorInvalidValue;
  if (out_f16) head_layer_reduce_kernel<__half><<<(unsigned)rows, 128, 0, st>>>(pooled, L, Tp_total, t_start, Tp, dw, S, D, kind, w, (__half*)out, xb,
                                                                                  reinterpret_cast<float2*>(stats));
  else head_layer_reduce_kernel<float><<<(unsigned)rows, 128, 0, st>>>(pooled, L, Tp_total, t_start, Tp, dw, S, D, kind, w, (float*)out, xb,
                                                                      reinterpret_cast<float2*>(stats));
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ im2col (k=3, pad=1)
__global__ void im2col_k3_kernel(const uint4* __restrict__ src, int Tin, int Cv, int stride, int Tout, uint4* __restrict__ out) {
  const long long row = blockIdx.x;
  const int b = (int)(row / Tout), j = (int)(row - (long long)b * Tout);
  for (int i = threadIdx.x; i < 3 * Cv; i += blockDim.x) {
    const int k = i / Cv, c = i - k * Cv;
    const int t = j * stride + k - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (t >= 0 && t < Tin) v = src[((long long)b * Tin + t) * Cv + c];
    out[row * 3 * Cv + i] = v;
  }
}

cudaError_t launch_im2col_k3(const void* src, bool bf16, int B, int Tin, int C, int stride, int Tout, void* out,
                             cudaStream_t st) {
  const int per16 = bf16 ? 8 : 4;
  if (C % per16) return cudaErrorInvalidValue;
  im2col_k3_kernel<<<(unsigned)((long long)B * Tout), 128, 0, st>>>((const uint4*)src, Tin, C / per16, stride, Tout, (uint4*)out);
  return cudaGetLastError();
}

// LayerNorm folded into the consuming GEMM (gemm_tc.cu): W' = bf16(W diag(gamma)), colsum[n] = sum_k W'[n,k] (of the ROUNDED
// values: it cancels against the same products in the GEMM), bias_out[n] = bias[n] + sum_k W[n,k] beta[k].  Warp per row n.
__global__ void __launch_bounds__(256) fold_ln_weights_kernel(const float* __restrict__ W, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ bias, int N, int K,
                                                              __nv_bfloat16* __restrict__ Wout, float* __restrict__ colsum,
                                                              float* __restrict__ bias_out) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  float cs = 0.f, bs = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = W[(long long)n * K + k];
    const __nv_bfloat16 wg = __float2bfloat16_rn(w * gamma[k]);
    Wout[(long long)n * K + k] = wg;
    cs += __bfloat162float(wg);
    bs = fmaf(w, beta[k], bs);
  }
  cs = warp_sum(cs);
  bs = warp_sum(bs);
  if (lane == 0) { colsum[n] = cs; bias_out[n] = (bias ? bias[n] : 0.f) + bs; }
}

cudaError_t launch_fold_ln_weights(const float* W, const float* gamma, const float* beta, const float* bias, int N, int K,
                                   __nv_bfloat16* Wout, float* colsum, float* bias_out, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  fold_ln_weights_kernel<<<(N + 7) / 8, 256, 0, st>>>(W, gamma, beta, bias, N, K, Wout, colsum, bias_out);
  return cudaGetLastError();
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += step) dst[i] = __float2bfloat16_rn(src[i]);
}

cudaError_t launch_f32_to_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  long long blocks = (n + 255) / 256;
  if (blocks > 65535) blocks = 65535;
  f32_to_bf16_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, n);
  return cudaGetLastError();
}

}  // namespace wat
