// bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands fed by TMA).
//
//   C[M, N] = epilogue( A[M, K] . W[N, K]^T )        A, W bf16 row-major (K contiguous)
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer    : A tile 128x64 + W tile BNx64 per stage, SWIZZLE_128B, mbarrier complete_tx
//   warp 1      MMA issuer      : one thread issues 4 x tcgen05.mma (M128, N=BN, K16) per stage; tcgen05.commit
//                                 releases the smem stage / publishes the accumulator
//   warps 2..9  epilogue        : tcgen05.ld (lane = output row; two warps per TMEM lane quadrant, each half of
//                                 the columns), bias / GELU / residual / layout, global stores
// Two accumulator buffers in TMEM (2 x BN columns) let the epilogue of tile i overlap the MMAs of tile i+1.
// Tiles are ordered n-fastest so the CTAs that share an A row-block run together and A is read from HBM once.
//
// Used for every dense contraction of the path (reference call sites: model.py:36 F.linear via Linear,
// model.py:47 conv via im2col): QKV, attention out-proj, MLP fc1/fc2, conv stem, TL-TR head linears.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace wat {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_THREADS = 320;                 // TMA warp, MMA warp, 8 epilogue warps
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;

template <int BN>
struct TcCfg {
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 3 : 5;
  static constexpr int TMEM_COLS = 2 * BN;                       // 512 or 256: powers of two
  static constexpr int WARP_STAGE_BYTES = 7168;                  // per epilogue warp: 32 x 36 float transpose tile + 32 x 8 float2 row-statistics slots
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 8 * WARP_STAGE_BYTES + 256 /*barriers*/;
};

struct GemmTcDev {
  const float* bias;
  void* C; long long ldc;
  const float* R; long long ldr; int r_mod;
  int M, N, K, act, epi;
  int tma_store;           // bf16 outputs leave through smem + TMA store (CTA-pair kernel)
  int res_tma;             // residual epilogue in row layout: residual box in by TMA load, x (fp16) and xb (bf16) boxes out by TMA store
  __nv_bfloat16* vt; int seq_T; int seq_Tpad; int n_head; int D;
  float q_scale;           // QKV epilogue: q columns [0, D) are multiplied by it (0 = off)
  int c_f16, r_f16;        // fp32 epilogues: C is stored / R is read as fp16 (the bf16 mode's residual stream)
  // fp32 epilogues as PRODUCER of the next LayerNorm (see kernels.h GemmTc): bf16 copy, per-slice row statistics
  __nv_bfloat16* xb; long long ldxb;
  float* stats; int stats_np;
  // LayerNorm of the A rows folded into this GEMM (CONSUMER)
  const float* ln_stats; int ln_np; const float* ln_colsum;
  long long* trace;        // optional clock trace of cluster 0 (test hook)
};

// bias / activation / residual / layout for 32 consecutive columns [n0, n0+32) of output row `row`
// `stage` (optional, warp-private 32 x 36 floats): fp32 outputs are transposed through it so that global accesses are
// 128-byte coalesced (8 lanes x 16 B per row) instead of one 16 B access per row.
// `bias_chunk`: the 32 bias values of this chunk (global, or the per-tile copy the warp prefetched into smem).
// the warp's 32 x 32 fp32 residual block in the coalesced (transposed) access pattern of the fp32 epilogue
__device__ __forceinline__ void load_residual_chunk(const GemmTcDev& g, int row_base, int n0, int lane, float4 (&res)[8]) {
  const int cc = (lane & 7) * 4;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int grow = row_base + it * 4 + (lane >> 3);
    res[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grow < g.M && n0 < g.N) {
      const long long rrow = g.r_mod > 0 ? (grow % g.r_mod) : grow;
      if (g.r_f16) {                                              // 4 halves: the raw bits travel in .x / .y and are converted at use
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(g.R) + rrow * g.ldr + n0 + cc);
        res[it].x = __uint_as_float(u.x); res[it].y = __uint_as_float(u.y);
      } else {
        res[it] = *reinterpret_cast<const float4*>(g.R + rrow * g.ldr + n0 + cc);
      }
    }
  }
}

// ---- LayerNorm folded into the consuming GEMM.  With W' = W diag(gamma) (bf16), cs[n] = sum_k W'[n,k] and
// b'[n] = b[n] + sum_k W[n,k] beta[k] (prepared once, launch_fold_ln_weights):
//     LN(x) W^T + b  =  rstd (x W'^T - mean cs) + b'
// so the GEMM reads the un-normalised bf16 copy of x its producer's epilogue wrote, and the row's (mean, rstd) come from the
// per-slice (sum, sum of squares) partials the same epilogue left in `stats` (summed here in a fixed order).
__device__ __forceinline__ void ln_row_scalars(const GemmTcDev& g, int row, bool row_ok, float& mean, float& rstd) {
  mean = 0.f; rstd = 0.f;
  if (!row_ok) return;
  const float2* sp = reinterpret_cast<const float2*>(g.ln_stats) + (long long)row * g.ln_np;
  float s = 0.f, q = 0.f;
  for (int p = 0; p < g.ln_np; ++p) { const float2 v = sp[p]; s += v.x; q += v.y; }
  const float inv_k = 1.0f / (float)g.K;
  mean = s * inv_k;
  rstd = rsqrtf(fmaxf(q * inv_k - mean * mean, 0.f) + 1e-5f);
}
// r[i] <- rstd (r[i] - mean cs[i]) + bias[i] for the 32 columns of a chunk (thread = row); bias is applied here, not later
__device__ __forceinline__ void ln_fold_chunk(uint32_t (&r)[32], float mean, float rstd, const float* cs, const float* bias) {
  const float2 nm = make_float2(-mean, -mean), rs = make_float2(rstd, rstd);
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 c = *reinterpret_cast<const float4*>(cs + i);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) b = *reinterpret_cast<const float4*>(bias + i);
    float2 t0 = __ffma2_rn(nm, make_float2(c.x, c.y), make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
    float2 t1 = __ffma2_rn(nm, make_float2(c.z, c.w), make_float2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
    t0 = __ffma2_rn(rs, t0, make_float2(b.x, b.y));
    t1 = __ffma2_rn(rs, t1, make_float2(b.z, b.w));
    r[i] = __float_as_uint(t0.x); r[i + 1] = __float_as_uint(t0.y); r[i + 2] = __float_as_uint(t1.x); r[i + 3] = __float_as_uint(t1.y);
  }
}

// ---- transposed phase of the fp32 epilogues.  After the smem transpose lane l holds, for it = 0..7, columns (l%8)*4..+3 of
// row it*4 + l/8 of the warp's 32 x 32 chunk: a warp-level access covers 4 rows x 128 contiguous bytes.
template <int NIT>
__device__ __forceinline__ void load_residual_rows(const GemmTcDev& g, int row_base, int n0, int lane, int it0, float4 (&res)[NIT]) {
  const int cc = (lane & 7) * 4;
#pragma unroll
  for (int k = 0; k < NIT; ++k) {
    const int grow = row_base + (it0 + k) * 4 + (lane >> 3);
    res[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grow < g.M && n0 < g.N) {
      const long long rrow = g.r_mod > 0 ? (grow % g.r_mod) : grow;
      if (g.r_f16) {
        const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(g.R) + rrow * g.ldr + n0 + cc);
        res[k].x = __uint_as_float(u.x); res[k].y = __uint_as_float(u.y);
      } else {
        res[k] = *reinterpret_cast<const float4*>(g.R + rrow * g.ldr + n0 + cc);
      }
    }
  }
}
// rows it0 .. it0+NIT-1 (per lane) of the chunk: stage (+ residual) -> C, bf16 copy, row statistics.  nxt_row_base >= 0: `res` is
// refilled with the residual block of the chunk at (nxt_row_base, nxt_n0) as soon as it has been consumed, i.e. BEFORE this
// chunk's stores are issued (the residual stream is updated in place, but a chunk's own columns are only read before they are
// written).
template <int NIT>
__device__ __forceinline__ void epi_rows_generic(const GemmTcDev& g, const float* stage, float2* rowacc, int row_base, int n0, int lane, int it0,
                                         float4 (&res)[NIT], bool add_res, int nxt_row_base = -1, int nxt_n0 = 0) {
  const int cc = (lane & 7) * 4;
  float4 t[NIT];
#pragma unroll
  for (int k = 0; k < NIT; ++k) {
    t[k] = *reinterpret_cast<const float4*>(stage + ((it0 + k) * 4 + (lane >> 3)) * 36 + cc);
    if (add_res) {
      const float4 rv = g.r_f16 ? f16x4_to_f32(__float_as_uint(res[k].x), __float_as_uint(res[k].y)) : res[k];
      t[k].x += rv.x; t[k].y += rv.y; t[k].z += rv.z; t[k].w += rv.w;
    }
  }
  if (nxt_row_base >= 0) load_residual_rows<NIT>(g, nxt_row_base, nxt_n0, lane, 0, res);
#pragma unroll
  for (int k = 0; k < NIT; ++k) {
    const int grow = row_base + (it0 + k) * 4 + (lane >> 3);
    if (grow < g.M) {
      if (g.c_f16)
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(g.C) + (long long)grow * g.ldc + n0 + cc) =
            make_uint2(pack_f16_sat(t[k].x, t[k].y), pack_f16_sat(t[k].z, t[k].w));
      else
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.C) + (long long)grow * g.ldc + n0 + cc) = t[k];
    }
  }
  // ---- this output feeds a LayerNorm: what the (folded) consumer needs is produced here, while the values are in registers
  if (g.xb) {                                                     // bf16 copy of the rows: the next GEMM's A operand
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const int grow = row_base + (it0 + k) * 4 + (lane >> 3);
      if (grow < g.M)
        *reinterpret_cast<uint2*>(g.xb + (long long)grow * g.ldxb + n0 + cc) = make_uint2(pack_bf16(t[k].x, t[k].y), pack_bf16(t[k].z, t[k].w));
    }
  }
  if (rowacc) {
    // per-row (sum, sum of squares) of this warp's column slice.  Each lane owns one slot per row it touches (row it*4 +
    // lane/8, slot lane%8: 32 rows x 8 slots of float2, warp-private smem, conflict-free) and adds its 4 columns there;
    // the 8 slots of a row are summed in a fixed order when the slice is complete - no shuffles on the way.
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      float2* a = rowacc + (it0 + k) * 32 + lane;
      float2 acc = *a;
      acc.x += (t[k].x + t[k].y) + (t[k].z + t[k].w);
      acc.y = fmaf(t[k].x, t[k].x, fmaf(t[k].y, t[k].y, fmaf(t[k].z, t[k].z, fmaf(t[k].w, t[k].w, acc.y))));
      *a = acc;
    }
  }
}
// The same for the case the bf16 encoder runs ~all the time - C in fp16 with its bf16 copy and the row statistics, every row of
// the chunk inside the matrix: no per-row bounds checks, one base pointer per stream advanced by a constant stride (the generic
// version above spends ~600 of its ~900 instructions per chunk on 64-bit address arithmetic, predicates and flag tests).
template <int NIT, bool RF16>
__device__ __forceinline__ void load_residual_rows_fast(const GemmTcDev& g, int row_base, int n0, int lane, int it0, float4 (&res)[NIT]) {
  const long long e0 = (long long)(row_base + it0 * 4 + (lane >> 3)) * g.ldr + n0 + (lane & 7) * 4;
  const long long step = 4 * g.ldr;
  if (RF16) {
    const __half* rp = reinterpret_cast<const __half*>(g.R) + e0;
#pragma unroll
    for (int k = 0; k < NIT; ++k) {
      const uint2 u = *reinterpret_cast<const uint2*>(rp + k * step);
      res[k].x = __uint_as_float(u.x); res[k].y = __uint_as_float(u.y);
    }
  } else {
    const float* rp = g.R + e0;
#pragma unroll
    for (int k = 0; k < NIT; ++k) res[k] = *reinterpret_cast<const float4*>(rp + k * step);
  }
}
template <int NIT, bool RF16>
__device__ __forceinline__ void epi_rows_fast(const GemmTcDev& g, const float* stage, float2* rowacc, int row_base, int n0, int lane, int it0,
                                              float4 (&res)[NIT], bool add_res, int nxt_row_base, int nxt_n0) {
  const int cc = (lane & 7) * 4, lrow = it0 * 4 + (lane >> 3);
  float4 t[NIT];
  const float* sp = stage + lrow * 36 + cc;
#pragma unroll
  for (int k = 0; k < NIT; ++k) {
    t[k] = *reinterpret_cast<const float4*>(sp + k * 4 * 36);
    if (add_res) {
      const float4 rv = RF16 ? f16x4_to_f32(__float_as_uint(res[k].x), __float_as_uint(res[k].y)) : res[k];
      t[k].x += rv.x; t[k].y += rv.y; t[k].z += rv.z; t[k].w += rv.w;
    }
  }
  if (nxt_row_base >= 0) {
    if (g.r_mod == 0 && nxt_row_base + 32 <= g.M && nxt_n0 < g.N) load_residual_rows_fast<NIT, RF16>(g, nxt_row_base, nxt_n0, lane, 0, res);
    else load_residual_rows<NIT>(g, nxt_row_base, nxt_n0, lane, 0, res);
  }
  {
    __half* cp = reinterpret_cast<__half*>(g.C) + (long long)(row_base + lrow) * g.ldc + n0 + cc;
    const long long step = 4 * g.ldc;
#pragma unroll
    for (int k = 0; k < NIT; ++k) *reinterpret_cast<uint2*>(cp + k * step) = make_uint2(pack_f16_sat(t[k].x, t[k].y), pack_f16_sat(t[k].z, t[k].w));
  }
  {
    __nv_bfloat16* xp = g.xb + (long long)(row_base + lrow) * g.ldxb + n0 + cc;
    const long long step = 4 * g.ldxb;
#pragma unroll
    for (int k = 0; k < NIT; ++k) *reinterpret_cast<uint2*>(xp + k * step) = make_uint2(pack_bf16(t[k].x, t[k].y), pack_bf16(t[k].z, t[k].w));
  }
  float2* a = rowacc + it0 * 32 + lane;
#pragma unroll
  for (int k = 0; k < NIT; ++k) {
    float2 acc = a[k * 32];
    acc.x += (t[k].x + t[k].y) + (t[k].z + t[k].w);
    acc.y = fmaf(t[k].x, t[k].x, fmaf(t[k].y, t[k].y, fmaf(t[k].z, t[k].z, fmaf(t[k].w, t[k].w, acc.y))));
    a[k * 32] = acc;
  }
}
template <int NIT>
__device__ __forceinline__ void epi_rows(const GemmTcDev& g, const float* stage, float2* rowacc, int row_base, int n0, int lane, int it0,
                                         float4 (&res)[NIT], bool add_res, int nxt_row_base = -1, int nxt_n0 = 0) {
  if (g.c_f16 && g.xb && rowacc && row_base + 32 <= g.M && (!add_res || g.r_mod == 0)) {           // warp-uniform
    if (g.r_f16) epi_rows_fast<NIT, true>(g, stage, rowacc, row_base, n0, lane, it0, res, add_res, nxt_row_base, nxt_n0);
    else epi_rows_fast<NIT, false>(g, stage, rowacc, row_base, n0, lane, it0, res, add_res, nxt_row_base, nxt_n0);
  } else {
    epi_rows_generic<NIT>(g, stage, rowacc, row_base, n0, lane, it0, res, add_res, nxt_row_base, nxt_n0);
  }
}
// `res` (optional): the residual block of THIS chunk, loaded by the caller ahead of time; it is refilled with the block at
// (nxt_row_base, nxt_n0) before this chunk's stores are issued, so a residual read is always one chunk ahead of its use
// (the residual stream is updated in place, but a chunk's own columns are only read before they are written).
__device__ __forceinline__ void tc_epilogue_chunk(const GemmTcDev& g, int row, bool row_ok, int n0, const uint32_t (&r)[32], int vb, int vtok,
                                                  const float* bias_chunk, float* stage = nullptr, int lane = 0, long long* tr2 = nullptr,
                                                  float4 (*res_io)[8] = nullptr, int nxt_row_base = -1, int nxt_n0 = 0,
                                                  float2* rowacc = nullptr) {
  if (n0 >= g.N) return;
  if (stage != nullptr && (g.epi == TC_EPI_F32_RES || g.epi == TC_EPI_F32)) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    if (bias_chunk) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias_chunk + i);
        v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
      }
    }
    if (tr2) tr2[0] = clock64();
    if (g.act == 1) {
#pragma unroll
      for (int i = 0; i < 32; i += 8) gelu_erf_poly8(v + i);
    }
    if (tr2) tr2[1] = clock64();
#pragma unroll
    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(stage + lane * 36 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    __syncwarp();
    const int row_base = row - lane;
    if (res_io) {
      // 8-warp kernel: the residual block of this chunk was loaded one chunk ahead (registers); refill it for the next chunk
      // before this chunk's stores are issued
      epi_rows<8>(g, stage, rowacc, row_base, n0, lane, 0, *res_io, true, nxt_row_base, nxt_n0);
    } else {
      // 16-warp kernel (96 registers): two halves of four rows-per-lane, residual loaded just in time - the other warps hide it
#pragma unroll 1
      for (int it0 = 0; it0 < 8; it0 += 4) {
        float4 res4[4];
        if (g.epi == TC_EPI_F32_RES) load_residual_rows<4>(g, row_base, n0, lane, it0, res4);
        epi_rows<4>(g, stage, rowacc, row_base, n0, lane, it0, res4, g.epi == TC_EPI_F32_RES);
      }
    }
    if (tr2) tr2[2] = clock64();
    __syncwarp();
    return;
  }
  if (!row_ok) return;
  if (g.epi == TC_EPI_BF16 || (g.epi == TC_EPI_QKV && n0 < 2 * g.D)) {
    // 8 columns at a time: bias, GELU, pack, one 16-byte store (keeps the live set small: this path also runs in the
    // 16-epilogue-warp kernel, which has ~96 registers per thread)
    __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(g.C) + (long long)row * g.ldc + n0;
    const float qs = (g.epi == TC_EPI_QKV && g.q_scale != 0.f && n0 < g.D) ? g.q_scale : 1.0f;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[i + e]);
      if (bias_chunk) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias_chunk + i);
        const float4 b1 = *reinterpret_cast<const float4*>(bias_chunk + i + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (g.act == 1) gelu_erf_poly8(v);
      if (qs != 1.0f) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= qs;
      }
      *reinterpret_cast<uint4*>(cp + i) =
          make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
    return;
  }
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  if (bias_chunk) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias_chunk + i);
      v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
    }
  }
  if (g.act == 1) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) gelu_erf_poly8(v + i);
  }
  if (g.epi == TC_EPI_QKV) {
    const int vc = n0 - 2 * g.D;                         // h * 64 + e ; a 32-chunk never straddles a head
    const int h = vc >> 6, e0 = vc & 63;
    __nv_bfloat16* vp = g.vt + (((long long)vb * g.n_head + h) * VT_ROWS + e0) * g.seq_Tpad + vtok;
#pragma unroll
    for (int i = 0; i < 32; ++i) vp[(long long)i * g.seq_Tpad] = __float2bfloat16_rn(v[i]);
  } else {                                                       // fp32 output without a transpose tile (16-epilogue-warp kernel only)
    if (g.epi == TC_EPI_F32_RES) {
      const long long rr = g.r_mod > 0 ? (row % g.r_mod) : row;
      const float* rp = g.R + rr * g.ldr + n0;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 r4 = *reinterpret_cast<const float4*>(rp + i);
        v[i] += r4.x; v[i + 1] += r4.y; v[i + 2] += r4.z; v[i + 3] += r4.w;
      }
    }
    float* cp = reinterpret_cast<float*>(g.C) + (long long)row * g.ldc + n0;
#pragma unroll
    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(cp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
}

// bf16 output chunk (32 rows x 32 columns) through a warp-private smem box and ONE TMA store (SASS UTMASTG) instead of four
// 16-byte stores per thread: a per-thread store instruction touches 32 different rows, i.e. 32 half-used sectors per request.
// `stage`: 2 KB, 1024-byte aligned, laid out as the 64-byte-swizzled box the tensor map describes (16-byte chunk c of row r
// sits at chunk c ^ ((r >> 1) & 3)), which also makes the warp's 16-byte smem writes bank-conflict free.
__device__ __forceinline__ void tc_epilogue_chunk_bf16_tma(const GemmTcDev& g, const CUtensorMap* tmC, int row_base, int n0, const uint32_t (&r)[32],
                                                           const float* bias_chunk, uint8_t* stage, int lane) {
  if (n0 >= g.N) return;
  if (lane == 0) tma_store_wait_read();                           // the previous box of this warp has left smem
  __syncwarp();
  uint8_t* my_row = stage + lane * 64;
  const int sw = (lane >> 1) & 3;
  const float qs = (g.epi == TC_EPI_QKV && g.q_scale != 0.f && n0 < g.D) ? g.q_scale : 1.0f;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[i + e]);
    if (bias_chunk) {
      const float4 b0 = *reinterpret_cast<const float4*>(bias_chunk + i);
      const float4 b1 = *reinterpret_cast<const float4*>(bias_chunk + i + 4);
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (g.act == 1) gelu_erf_poly8(v);
    if (qs != 1.0f) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= qs;
    }
    *reinterpret_cast<uint4*>(my_row + (((i >> 3) ^ sw) << 4)) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tmC, stage, n0, row_base);
    tma_store_commit();
  }
}

// Residual epilogue in ROW layout (CTA-pair kernel, fp16 residual stream): thread = output row from the TMEM load to the end.
// The chunk's 32 x 32 fp16 residual box arrives by a TMA load issued TWO chunks earlier (three warp-private boxes and
// mbarriers in turn), the sum acc + bias + residual is formed in fp32, its (sum, sum of squares) stay in the thread's registers
// (a row's statistics need no other lane), the fp16 row goes back into the residual's own box, the bf16 copy into a fourth
// box, and each box leaves by ONE TMA store.  Against the transposing epilogue above this drops the fp32 transpose through
// smem, the per-lane statistics slots and every per-row address and predicate (the tensor maps clip rows >= M): ~180 instead
// of ~500 instructions per chunk.
// Hand-over of the boxes: the only wait on the bulk stores sits AFTER the chunk's arithmetic, where the previous chunk's boxes
// have long been read out; it frees the bf16 box for this chunk and box (cc + 2) % 3 = (cc - 1) % 3 for the load of chunk
// cc + 2, which is issued right behind this chunk's stores - a full chunk ahead of its use.
// `wst`: 8 KB = residual boxes 0..2 | bf16 box (2 KB each, 64-byte-swizzled like the bf16 store box); `rbar`: 3 mbarriers.
__device__ __forceinline__ void tc_epilogue_chunk_res_tma(const CUtensorMap* tmX, const CUtensorMap* tmXB, const CUtensorMap* tmR,
                                                          uint8_t* wst, uint64_t* rbar, uint32_t cc, int row_base, int n0,
                                                          const uint32_t (&r)[32], const float* bias_chunk, int lane, int nxt_row_base,
                                                          int nxt_n0, float2& s2, float2& q2) {
  const uint32_t bi = cc % 3;
  uint8_t* rin = wst + bi * 2048;
  uint8_t* xbb = wst + 3 * 2048;
  float4 bv[8];                                                   // the chunk's 32 bias values (L1): requested before the box is waited for
#pragma unroll
  for (int i = 0; i < 8; ++i) bv[i] = bias_chunk ? __ldg(reinterpret_cast<const float4*>(bias_chunk) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  mbar_wait(&rbar[bi], (cc / 3) & 1);                             // this chunk's residual box has landed
  const int sw = (lane >> 1) & 3;
  uint8_t* my = rin + lane * 64;
  uint4 xbo[4];
  uint4 rv4[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) rv4[c] = *reinterpret_cast<const uint4*>(my + ((c ^ sw) << 4));   // the row's 32 halves, all loads in flight
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t rw[4] = {rv4[c].x, rv4[c].y, rv4[c].z, rv4[c].w};
    float2 v[4];                                                  // packed fp32 pairs: FADD2 / FFMA2 halve the instruction count
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = make_float2(__uint_as_float(r[c * 8 + 2 * k]), __uint_as_float(r[c * 8 + 2 * k + 1]));
    {
      const float4 b0 = bv[2 * c], b1 = bv[2 * c + 1];
      v[0] = __fadd2_rn(v[0], make_float2(b0.x, b0.y)); v[1] = __fadd2_rn(v[1], make_float2(b0.z, b0.w));
      v[2] = __fadd2_rn(v[2], make_float2(b1.x, b1.y)); v[3] = __fadd2_rn(v[3], make_float2(b1.z, b1.w));
    }
    uint32_t xo[4], xb2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = __fadd2_rn(v[k], __half22float2(*reinterpret_cast<const __half2*>(&rw[k])));
      s2 = __fadd2_rn(s2, v[k]);
      q2 = __ffma2_rn(v[k], v[k], q2);
      xo[k] = pack_f16_sat(v[k].x, v[k].y);
      xb2[k] = pack_bf16(v[k].x, v[k].y);
    }
    // a lane only ever touches its own row of a box: the fp16 result overwrites the residual it was read from
    *reinterpret_cast<uint4*>(my + ((c ^ sw) << 4)) = make_uint4(xo[0], xo[1], xo[2], xo[3]);
    xbo[c] = make_uint4(xb2[0], xb2[1], xb2[2], xb2[3]);
  }
  if (lane == 0) tma_store_wait_read();                           // chunk cc - 1's boxes (its x box, the bf16 box) have left smem
  __syncwarp();
  uint8_t* myb = xbb + lane * 64;
#pragma unroll
  for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(myb + ((c ^ sw) << 4)) = xbo[c];
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tmX, rin, n0, row_base);
    tma_store_2d(tmXB, xbb, n0, row_base);
    tma_store_commit();
    if (nxt_row_base >= 0) {                                      // the residual box of chunk cc + 2
      const uint32_t nb = (cc + 2) % 3;
      mbar_expect_tx(&rbar[nb], 2048);
      tma_load_2d(wst + nb * 2048, tmR, &rbar[nb], nxt_n0, nxt_row_base);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTcDev g) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* epi_stage_area = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage_area + 8 * Cfg::WARP_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tmem_full = bars + 2 * Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (g.M + TC_BM - 1) / TC_BM;
  const int n_tiles = (g.N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int k_blocks = (g.K + TC_BK - 1) / TC_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 256); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, Cfg::TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Producer and MMA warps run converged (all 32 lanes follow the control flow) and only the elected lane issues:
  // inside a divergent single-lane region the compiler wraps every TMA / tcgen05.mma in an ELECT + R2UR.BROADCAST loop.
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_blk = tile % n_tiles, m_blk = tile / n_tiles;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + TC_A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * TC_BK, m_blk * TC_BM);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * TC_BK, n_blk * BN);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(TC_BM, BN);
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait_spin(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t a_desc = make_smem_desc_sw128(sa);
          const uint64_t b_desc = make_smem_desc_sw128(sa + TC_A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);   // +32 B per K=16 step
          umma_commit(&empty_bar[stage]);
          if (kb == k_blocks - 1) umma_commit(&tmem_full[as]);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;                                     // TMEM lane quadrant this warp may read
    const int chalf = (warp - 2) >> 2;                          // which half of the tile's columns
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int n_blk = tile % n_tiles, m_blk = tile / n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int row = m_blk * TC_BM + q * 32 + lane;
      const bool row_ok = row < g.M;
      float* my_stage = reinterpret_cast<float*>(epi_stage_area + (warp - 2) * Cfg::WARP_STAGE_BYTES);
      float2* rowacc = g.stats ? reinterpret_cast<float2*>(my_stage + 32 * 36) : nullptr;
      if (rowacc) {
#pragma unroll
        for (int k = 0; k < 8; ++k) rowacc[k * 32 + lane] = make_float2(0.f, 0.f);
        __syncwarp();
      }
      float ln_mean = 0.f, ln_rstd = 0.f;
      if (g.ln_stats) ln_row_scalars(g, row, row_ok, ln_mean, ln_rstd);        // while the tile's MMAs are still running
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      int vb = 0, vtok = 0;
      if (g.epi == TC_EPI_QKV && row_ok) { vb = row / g.seq_T; vtok = row - vb * g.seq_T; }
      const bool f32_out = g.epi == TC_EPI_F32_RES || g.epi == TC_EPI_F32;
#pragma unroll 1
      for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_row + c0, r);
        tc_wait_ld();
        if (c0 + 32 == (chalf + 1) * (BN / 2)) { tc_fence_before(); mbar_arrive(&tmem_empty[as]); }
        const int n0 = n_blk * BN + c0;
        const float* bias_chunk = g.bias ? g.bias + n0 : nullptr;
        if (g.ln_stats && n0 < g.N) { ln_fold_chunk(r, ln_mean, ln_rstd, g.ln_colsum + n0, bias_chunk); bias_chunk = nullptr; }
        tc_epilogue_chunk(g, row, row_ok, n0, r, vb, vtok, bias_chunk, f32_out ? my_stage : nullptr, lane, nullptr, nullptr, -1, 0, rowacc);
      }
      if (rowacc) {                                               // this warp's slice of the row statistics: partial p = 2 n_blk + chalf
        __syncwarp();
        if (row_ok && n_blk * BN + chalf * (BN / 2) < g.N) {
          float2 a = rowacc[lane * 8];                            // this thread's row: slots lane*8 .. lane*8+7
#pragma unroll
          for (int k = 1; k < 8; ++k) { const float2 b = rowacc[lane * 8 + k]; a.x += b.x; a.y += b.y; }
          reinterpret_cast<float2*>(g.stats)[(long long)row * g.stats_np + n_blk * 2 + chalf] = a;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}


// ------------------------------------------------------------------------------------------ CTA-pair variant
// 256 x 256 output tile per cluster of two CTAs (tcgen05 cta_group::2, MMA M = 256): each CTA stages its own 128 A
// rows and HALF of the W tile (128 of the 256 N rows), so a k-block costs 32 KB of L2->smem traffic per SM instead
// of 48 KB for the same 128x256x64 MACs per SM; 5 pipeline stages + the epilogue staging fit.  The leader CTA (rank 0) owns the
// `full` barriers (both CTAs' TMA bytes are signalled there) and issues every MMA; tcgen05.commit multicasts the
// stage-free / accumulator-ready arrivals to both CTAs; each CTA's epilogue warps drain their own 128 TMEM lanes.
constexpr int TC2_STAGE_BYTES = 2 * TC_A_BYTES;                  // A 128x64 + W-half 128x64
// EW = epilogue warps: 8 (two per TMEM lane quadrant, 128 columns each; fp32 outputs use a per-warp transpose tile) or
// 16 (four per quadrant, 64 columns each) for the bf16 + GELU epilogue, whose per-tile latency with 8 warps (~13k
// cycles) exceeds the tile's 10k MMA cycles at K = 1280.
// (A 16-warp / 3-stage variant of the fp32 residual epilogue was measured for the attention out-projection, which is bound by
// its epilogue's memory traffic - 12 bytes per output element in 128-byte pieces spread over 128 rows - and was slower.)
template <int EW, int STAGES> struct Tc2Cfg {
  static constexpr int THREADS = 64 + 32 * EW;
  static constexpr int COLS = 256 / (EW / 4);                    // columns per epilogue warp
  static constexpr bool F32_STAGE = EW == 8;                     // has the fp32 transpose tile + row-statistics slots
  // per-warp epilogue staging, 1024-byte aligned: the fp32 transpose tile (32 x 36 floats; 8-warp kernel only), which doubles
  // as the 2 KB bf16 box of the TMA-store epilogue
  static constexpr int WARP_STAGE_BYTES = F32_STAGE ? 7168 : 2048;   // (+ 32 x 8 float2 row-statistics slots)
  static constexpr int STAGE_AREA = EW * WARP_STAGE_BYTES;
  static constexpr int BIAS_AREA = 2 * EW * COLS * 4;            // this warp's bias slice of the current tile + its LN column sums
  // res_tma epilogue (8 warps): the stage and bias areas together are 8 x 8 KB = four 2 KB boxes per warp; its 8 x 3 mbarriers follow
  // the pipeline barriers
  static constexpr int SMEM_BYTES = STAGES * TC2_STAGE_BYTES + 1024 + STAGE_AREA + BIAS_AREA + 384;
};

// RES_TMA: the instantiation for the TMA-box residual epilogue (8 epilogue warps); every other epilogue is compiled out of it, and it
// out of theirs - sharing one kernel cost the QKV epilogue 10% through register spills.
template <int EW, int TC2_STAGES, bool RES_TMA>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Tc2Cfg<EW, TC2_STAGES>::THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmXB, const GemmTcDev g) {
  constexpr int BN = 256;
  using Cfg2 = Tc2Cfg<EW, TC2_STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* epi_stage_area = smem + TC2_STAGES * TC2_STAGE_BYTES;                       // 1024-byte aligned
  float* epi_bias_area = reinterpret_cast<float*>(epi_stage_area + Cfg2::STAGE_AREA);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage_area + Cfg2::STAGE_AREA + Cfg2::BIAS_AREA);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + TC2_STAGES;
  uint64_t* tmem_full = bars + 2 * TC2_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  const int m_pairs = (g.M + 2 * TC_BM - 1) / (2 * TC_BM);
  const int n_tiles = (g.N + BN - 1) / BN;
  const int total_tiles = m_pairs * n_tiles;
  const int k_blocks = (g.K + TC_BK - 1) / TC_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (g.tma_store) tma_prefetch_desc(&tmC);
    for (int s = 0; s < TC2_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 2 * EW); }   // EW warps x 2 CTAs
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
  constexpr bool res_tma = RES_TMA;
  static_assert(!RES_TMA || (EW == 8 && Cfg2::STAGE_AREA + Cfg2::BIAS_AREA == 8 * 8192), "res_tma: 8 epilogue warps, four 2 KB boxes each");
  uint64_t* res_bars = bars + 16;                                  // 8 warps x 3 (res_tma only)
  if (res_tma && warp >= 2 && lane == 0) {                         // warp-private barriers of the residual boxes (tc_epilogue_chunk_res_tma)
    uint64_t* rbar = res_bars + (warp - 2) * 3;
    mbar_init(&rbar[0], 1);
    mbar_init(&rbar[1], 1);
    mbar_init(&rbar[2], 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmXB);
    tma_prefetch_desc(&tmC);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
      const int n_blk = tile % n_tiles, m_pair = tile / n_tiles;
      const int row0 = m_pair * 2 * TC_BM + (int)rank * TC_BM;
      const int col0 = n_blk * BN + (int)rank * (BN / 2);
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * TC2_STAGE_BYTES;
          uint8_t* sb = sa + TC_A_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * TC2_STAGE_BYTES);
          const uint32_t bar_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
          tma_load_2d_pair(sa, &tmA, bar_leader, kb * TC_BK, row0);
          tma_load_2d_pair(sb, &tmB, bar_leader, kb * TC_BK, col0);
        }
        __syncwarp();
        if (++stage == TC2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {                                              // warp-uniform: the leader CTA issues every MMA
      constexpr uint32_t idesc = make_idesc_bf16(2 * TC_BM, BN);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        // Two pipeline stages per iteration: a barrier check costs ~100-170 cycles even when the phase is already
        // complete, comparable to the 512 tensor cycles of one k-block, so the two try_waits are issued back to back
        // (their latencies overlap) and 8 MMAs + 2 commits follow in one burst.
        int kb = 0;
        for (; kb + 1 < k_blocks; kb += 2) {
          int stage1 = stage + 1; uint32_t phase1 = phase;
          if (stage1 == TC2_STAGES) { stage1 = 0; phase1 ^= 1; }
          const bool ok0 = mbar_try_wait_nohint(&full_bar[stage], phase);
          const bool ok1 = mbar_try_wait_nohint(&full_bar[stage1], phase1);
          if (!ok0) mbar_wait_spin(&full_bar[stage], phase);
          if (!ok1) mbar_wait_spin(&full_bar[stage1], phase1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa0 = smem_u32(smem + stage * TC2_STAGE_BYTES);
            const uint32_t sa1 = smem_u32(smem + stage1 * TC2_STAGE_BYTES);
            const uint64_t a0 = make_smem_desc_sw128(sa0), b0 = make_smem_desc_sw128(sa0 + TC_A_BYTES);
            const uint64_t a1 = make_smem_desc_sw128(sa1), b1 = make_smem_desc_sw128(sa1 + TC_A_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) umma_bf16_ss_pair(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, (kb | k) != 0);
            umma_commit_pair(&empty_bar[stage], 3);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) umma_bf16_ss_pair(d_tmem, a1 + 2 * k, b1 + 2 * k, idesc, 1);
            umma_commit_pair(&empty_bar[stage1], 3);
            if (kb + 2 == k_blocks) umma_commit_pair(&tmem_full[as], 3);
          }
          __syncwarp();
          stage = stage1; phase = phase1;
          if (++stage == TC2_STAGES) { stage = 0; phase ^= 1; }
        }
        if (kb < k_blocks) {                                      // odd tail
          mbar_wait_spin(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * TC2_STAGE_BYTES);
            const uint64_t a_desc = make_smem_desc_sw128(sa);
            const uint64_t b_desc = make_smem_desc_sw128(sa + TC_A_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)
              umma_bf16_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            umma_commit_pair(&empty_bar[stage], 3);
            umma_commit_pair(&tmem_full[as], 3);
          }
          __syncwarp();
          if (++stage == TC2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int chalf = (warp - 2) >> 2;                             // which column slice of the tile (COLS wide)
    constexpr int COLS = Cfg2::COLS;
    int it = 0;
    // fp32 residual epilogue (8 warps): the residual block of a chunk is loaded one chunk (or one tile) ahead of its use
    const bool res_ahead = !RES_TMA && EW == 8 && g.epi == TC_EPI_F32_RES;
    float4 res[8];
    uint8_t* wst = epi_stage_area + (warp - 2) * 8192;            // res_tma: this warp's four boxes (spans the stage and bias areas)
    uint64_t* rbar = res_bars + (warp - 2) * 3;
    uint32_t cc = 0;                                               // chunks this warp has processed (residual box / barrier phase)
    // coordinates of this warp's k-th chunk (4 chunks per tile, tiles cluster_id, cluster_id + n_clusters, ...); false past the end
    auto chunk_at = [&](int k, int& rb, int& cn0) {
      const int t = cluster_id + (k >> 2) * n_clusters;
      if (t >= total_tiles) return false;
      rb = (t / n_tiles) * 2 * TC_BM + (int)rank * TC_BM + q * 32;
      cn0 = (t % n_tiles) * BN + chalf * COLS + (k & 3) * 32;
      return true;
    };
    if (res_tma && lane == 0) {
      for (int k = 0; k < 2; ++k) {
        int rb, cn0;
        if (chunk_at(k, rb, cn0)) { mbar_expect_tx(&rbar[k], 2048); tma_load_2d(wst + k * 2048, &tmR, &rbar[k], cn0, rb); }
      }
    }
    if (res_ahead && cluster_id < total_tiles)
      load_residual_chunk(g, (cluster_id / n_tiles) * 2 * TC_BM + (int)rank * TC_BM + q * 32, (cluster_id % n_tiles) * BN + chalf * COLS, lane, res);
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
      const int n_blk = tile % n_tiles, m_pair = tile / n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      // this warp's 128 bias values -> smem while the MMAs of the tile are still running
      float* my_stage = reinterpret_cast<float*>(epi_stage_area + (warp - 2) * Cfg2::WARP_STAGE_BYTES);
      float* bias_s = epi_bias_area + (warp - 2) * 2 * COLS;
      float* cs_s = bias_s + COLS;
      if (!res_tma && lane * 4 < COLS) {                           // (res_tma reads its bias from global: the area holds its boxes)
        const int nb = n_blk * BN + chalf * COLS + lane * 4;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), c4 = b4;
        if (g.bias && nb < g.N) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + nb));
        if (g.ln_stats && nb < g.N) c4 = __ldg(reinterpret_cast<const float4*>(g.ln_colsum + nb));
        *reinterpret_cast<float4*>(bias_s + lane * 4) = b4;
        *reinterpret_cast<float4*>(cs_s + lane * 4) = c4;
      }
      float2* rowacc = (Cfg2::F32_STAGE && g.stats && !res_tma) ? reinterpret_cast<float2*>(my_stage + 32 * 36) : nullptr;
      int nxt_tile_rb = -1, nxt_tile_n0 = 0;                      // res_tma: first chunk of this warp's next tile
      if (res_tma && tile + n_clusters < total_tiles) {
        const int nt = tile + n_clusters, nm = nt / n_tiles;
        nxt_tile_rb = nm * 2 * TC_BM + (int)rank * TC_BM + q * 32;
        nxt_tile_n0 = (nt - nm * n_tiles) * BN + chalf * COLS;
      }
      float2 st_s2 = make_float2(0.f, 0.f), st_q2 = st_s2;        // res_tma: (sum, sum of squares) of this thread's row over this warp's column slice, as even / odd column partials
      if (rowacc) {
#pragma unroll
        for (int k = 0; k < 8; ++k) rowacc[k * 32 + lane] = make_float2(0.f, 0.f);
      }
      __syncwarp();
      const int row = m_pair * 2 * TC_BM + (int)rank * TC_BM + q * 32 + lane;
      const bool row_ok = row < g.M;
      float ln_mean = 0.f, ln_rstd = 0.f;
      if (!RES_TMA && g.ln_stats) ln_row_scalars(g, row, row_ok, ln_mean, ln_rstd);        // while the tile's MMAs are still running
      long long* tr = (g.trace && blockIdx.x == 0 && threadIdx.x == 64 && it < 24) ? g.trace + it * 8 : nullptr;
      if (tr) tr[0] = clock64();
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      if (tr) tr[1] = clock64();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      int vb = 0, vtok = 0;
      if (!RES_TMA && g.epi == TC_EPI_QKV && row_ok) { vb = row / g.seq_T; vtok = row - vb * g.seq_T; }
#pragma unroll 1
      for (int c0 = chalf * COLS; c0 < (chalf + 1) * COLS; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_row + c0, r);
        tc_wait_ld();
        if (c0 + 32 == (chalf + 1) * COLS) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[as]), 0));
        }
        long long* tr2 = (tr && c0 == chalf * COLS + 32 && it >= 8 && it < 16) ? g.trace + 24 * 8 + (it - 8) * 4 : nullptr;
        if (tr2) tr2[3] = clock64();
        const float* bias_chunk = g.bias ? bias_s + (c0 - chalf * COLS) : nullptr;
        if (!RES_TMA && g.ln_stats && n_blk * BN + c0 < g.N) { ln_fold_chunk(r, ln_mean, ln_rstd, cs_s + (c0 - chalf * COLS), bias_chunk); bias_chunk = nullptr; }
        if constexpr (RES_TMA) {
          // the chunk two ahead: same tile for the first two chunks, else the next tile's first two
          int rb2 = -1, n02 = 0;
          const int kc = (c0 - chalf * COLS) >> 5;
          if (kc < 2) { rb2 = row - lane; n02 = n_blk * BN + c0 + 64; }
          else if (nxt_tile_rb >= 0) { rb2 = nxt_tile_rb; n02 = nxt_tile_n0 + (kc - 2) * 32; }
          tc_epilogue_chunk_res_tma(&tmC, &tmXB, &tmR, wst, rbar, cc, row - lane, n_blk * BN + c0, r,
                                    g.bias ? g.bias + n_blk * BN + c0 : nullptr, lane, rb2, n02, st_s2, st_q2);
          ++cc;
        } else if (res_ahead) {
          int nrb = -1, nn0 = 0;
          if (c0 + 32 < (chalf + 1) * COLS) { nrb = row - lane; nn0 = n_blk * BN + c0 + 32; }
          else if (tile + n_clusters < total_tiles) {
            const int nt = tile + n_clusters;
            nrb = (nt / n_tiles) * 2 * TC_BM + (int)rank * TC_BM + q * 32;
            nn0 = (nt % n_tiles) * BN + chalf * COLS;
          }
          tc_epilogue_chunk(g, row, row_ok, n_blk * BN + c0, r, vb, vtok, bias_chunk, my_stage, lane, tr2, &res, nrb, nn0, rowacc);
        } else if (g.tma_store && (g.epi == TC_EPI_BF16 || n_blk * BN + c0 < 2 * g.D)) {
          tc_epilogue_chunk_bf16_tma(g, &tmC, row - lane, n_blk * BN + c0, r, bias_chunk, reinterpret_cast<uint8_t*>(my_stage), lane);
        } else {
          tc_epilogue_chunk(g, row, row_ok, n_blk * BN + c0, r, vb, vtok, bias_chunk, Cfg2::F32_STAGE ? my_stage : nullptr, lane, tr2, nullptr, -1, 0, rowacc);
        }
        if (tr) tr[2 + (c0 - chalf * COLS) / 32] = clock64();
      }
      if (res_tma && row_ok)
        reinterpret_cast<float2*>(g.stats)[(long long)row * g.stats_np + n_blk * (BN / COLS) + chalf] = make_float2(st_s2.x + st_s2.y, st_q2.x + st_q2.y);
      if (rowacc) {                                               // this warp's slice of the row statistics: partial p = n_blk (BN / COLS) + chalf
        __syncwarp();
        if (row_ok && n_blk * BN + chalf * COLS < g.N) {
          float2 a = rowacc[lane * 8];                            // this thread's row: slots lane*8 .. lane*8+7
#pragma unroll
          for (int k = 1; k < 8; ++k) { const float2 b = rowacc[lane * 8 + k]; a.x += b.x; a.y += b.y; }
          reinterpret_cast<float2*>(g.stats)[(long long)row * g.stats_np + n_blk * (BN / COLS) + chalf] = a;
        }
      }
    }
  }

  if ((g.tma_store || res_tma) && warp >= 2 && lane == 0) tma_store_wait_all();     // bulk stores read this CTA's smem: drain before exit
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------ host side
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2D bf16 row-major [rows, cols] with row stride ld (elements); box = 64 cols x box_rows, 128B swizzle
static bool make_map_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 16-bit [rows, cols] with row stride ld (elements): box = 32 cols x 32 rows, 64-byte swizzle (the epilogues' TMA stores and the
// residual box load; rows >= `rows` are clipped on store and zero-filled on load)
static bool make_map_store(CUtensorMap* m, void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, bool f16 = false) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int res_tma_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("WAT_GEMM_RES_TMA");
    mode = e ? (atoi(e) != 0) : 1;
  }
  return mode;
}

static int tma_store_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("WAT_GEMM_TMA_STORE");
    mode = e ? (atoi(e) != 0) : 1;
  }
  return mode;
}

template <int BN>
static cudaError_t launch_tc(const GemmTc& g, int num_sms, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  static unsigned long long attr_mask = 0;
  if (cudaError_t e = opt_in_smem(gemm_tc_kernel<BN>, Cfg::SMEM_BYTES, attr_mask); e != cudaSuccess) return e;
  CUtensorMap tmA, tmB;
  if (!make_map_2d(&tmA, g.A, g.M, g.K, g.lda, TC_BM)) return cudaErrorInvalidValue;
  if (!make_map_2d(&tmB, g.W, g.N, g.K, g.K, BN)) return cudaErrorInvalidValue;
  GemmTcDev d;
  d.bias = g.bias; d.C = g.C; d.ldc = g.ldc; d.R = g.R; d.ldr = g.ldr; d.r_mod = g.r_mod;
  d.M = g.M; d.N = g.N; d.K = g.K; d.act = g.act; d.epi = g.epi; d.tma_store = 0; d.res_tma = 0;
  d.vt = g.vt; d.seq_T = g.seq_T; d.seq_Tpad = g.seq_Tpad; d.n_head = g.n_head; d.D = g.N / 3; d.q_scale = g.q_scale; d.c_f16 = g.c_f16; d.r_f16 = g.r_f16;
  d.trace = nullptr;
  d.xb = g.xb; d.ldxb = g.ldxb; d.stats = g.stats; d.stats_np = g.stats_np;
  d.ln_stats = g.ln_stats; d.ln_np = g.ln_np; d.ln_colsum = g.ln_colsum;
  const int m_tiles = (g.M + TC_BM - 1) / TC_BM, n_tiles = (g.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles;
  const int grid = total < num_sms ? total : num_sms;
  gemm_tc_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, d);
  return cudaGetLastError();
}


template <int EW, int STAGES>
static cudaError_t launch_tc2(const GemmTc& g, int num_sms, cudaStream_t st) {
  using Cfg2 = Tc2Cfg<EW, STAGES>;
  static unsigned long long attr_mask = 0, attr_mask_res = 0;
  if (cudaError_t e = opt_in_smem(gemm_tc2_kernel<EW, STAGES, false>, Cfg2::SMEM_BYTES, attr_mask); e != cudaSuccess) return e;
  if constexpr (EW == 8)
    if (cudaError_t e = opt_in_smem(gemm_tc2_kernel<EW, STAGES, true>, Cfg2::SMEM_BYTES, attr_mask_res); e != cudaSuccess) return e;
  CUtensorMap tmA, tmB;
  if (!make_map_2d(&tmA, g.A, g.M, g.K, g.lda, TC_BM)) return cudaErrorInvalidValue;
  if (!make_map_2d(&tmB, g.W, g.N, g.K, g.K, 128)) return cudaErrorInvalidValue;      // each CTA loads half of the 256 W rows
  GemmTcDev d;
  d.bias = g.bias; d.C = g.C; d.ldc = g.ldc; d.R = g.R; d.ldr = g.ldr; d.r_mod = g.r_mod;
  d.M = g.M; d.N = g.N; d.K = g.K; d.act = g.act; d.epi = g.epi;
  d.vt = g.vt; d.seq_T = g.seq_T; d.seq_Tpad = g.seq_Tpad; d.n_head = g.n_head; d.D = g.N / 3; d.q_scale = g.q_scale; d.c_f16 = g.c_f16; d.r_f16 = g.r_f16;
  d.trace = g.trace;
  d.xb = g.xb; d.ldxb = g.ldxb; d.stats = g.stats; d.stats_np = g.stats_np;
  d.ln_stats = g.ln_stats; d.ln_np = g.ln_np; d.ln_colsum = g.ln_colsum;
  // bf16 outputs (fc1, the q | k part of QKV) are written by TMA from a smem box; needs a 16-byte aligned, 16-byte pitched C
  CUtensorMap tmC = tmA;
  d.tma_store = 0;
  if (tma_store_mode() && (g.epi == TC_EPI_BF16 || g.epi == TC_EPI_QKV) && (g.ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(g.C) & 15) == 0) {
    const uint64_t c_cols = g.epi == TC_EPI_QKV ? 2 * (uint64_t)(g.N / 3) : (uint64_t)g.N;
    if (!make_map_store(&tmC, g.C, g.M, c_cols, g.ldc)) return cudaErrorInvalidValue;
    d.tma_store = 1;
  }
  // fp16 residual update that also produces the next LayerNorm's inputs (out-proj, fc2): row-layout epilogue through TMA boxes
  CUtensorMap tmR = tmA, tmXB = tmA;
  d.res_tma = 0;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (EW == 8 && res_tma_mode() && g.epi == TC_EPI_F32_RES && g.c_f16 && g.r_f16 && g.R && g.xb && g.stats && g.r_mod == 0 && !g.ln_stats &&
      (g.ldc & 7) == 0 && (g.ldr & 7) == 0 && (g.ldxb & 7) == 0 && al16(g.C) && al16(g.R) && al16(g.xb)) {
    if (!make_map_store(&tmC, g.C, g.M, g.N, g.ldc, true)) return cudaErrorInvalidValue;
    if (!make_map_store(&tmR, const_cast<float*>(g.R), g.M, g.N, g.ldr, true)) return cudaErrorInvalidValue;
    if (!make_map_store(&tmXB, g.xb, g.M, g.N, g.ldxb, false)) return cudaErrorInvalidValue;
    d.res_tma = 1;
  }
  const int total = ((g.M + 255) / 256) * ((g.N + 255) / 256);
  int clusters = num_sms / 2;
  if (clusters > total) clusters = total;
  if constexpr (EW == 8) {
    if (d.res_tma) {
      gemm_tc2_kernel<EW, STAGES, true><<<2 * clusters, Cfg2::THREADS, Cfg2::SMEM_BYTES, st>>>(tmA, tmB, tmC, tmR, tmXB, d);
      return cudaGetLastError();
    }
  }
  gemm_tc2_kernel<EW, STAGES, false><<<2 * clusters, Cfg2::THREADS, Cfg2::SMEM_BYTES, st>>>(tmA, tmB, tmC, tmR, tmXB, d);
  return cudaGetLastError();
}

// 0: 1-CTA kernel, 1: CTA-pair kernel when the shape allows (N % 256 == 0).  WAT_GEMM_PAIR=0/1 overrides.
static int pair_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("WAT_GEMM_PAIR");
    mode = e ? (atoi(e) != 0) : 1;
  }
  return mode;
}

static bool use_pair_kernel(int M, int N, int force_pair) {
  return N % 256 == 0 && (force_pair > 0 || (force_pair == 0 && pair_mode() && M >= 256));
}
// number of per-row statistics slices a launch with GemmTc::stats set writes: one per epilogue warp column slice of the
// kernel that will run (two per N tile)
int gemm_tc_stats_slices(int M, int N, int K, int epi, int force_pair) {
  (void)K; (void)epi;
  if (use_pair_kernel(M, N, force_pair) || N % 256 == 0) return 2 * (N / 256);
  return 2 * (N / 128);
}

cudaError_t launch_gemm_tc(const GemmTc& g, int num_sms, cudaStream_t st) {
  if (g.M <= 0) return cudaSuccess;
  const bool f32_out = g.epi == TC_EPI_F32_RES || g.epi == TC_EPI_F32;
  if ((g.xb || g.stats) && !f32_out) return cudaErrorInvalidValue;                     // producer extras live in the fp32 epilogues
  if (g.stats && g.stats_np != gemm_tc_stats_slices(g.M, g.N, g.K, g.epi, g.force_pair)) return cudaErrorInvalidValue;
  if (g.xb && (g.ldxb & 3)) return cudaErrorInvalidValue;
  if (g.ln_stats && (!g.ln_colsum || g.ln_np <= 0)) return cudaErrorInvalidValue;
  if ((g.K & 7) || (g.lda & 7) || (g.N & 31) || (g.ldc & 7)) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.W) & 15)) return cudaErrorInvalidValue;
  if (g.epi == TC_EPI_QKV && ((g.N % 3) || ((g.N / 3) & 63) || g.seq_T <= 0)) return cudaErrorInvalidValue;
  if (use_pair_kernel(g.M, g.N, g.force_pair)) {
    static const int wide = getenv("WAT_GEMM_EW16") ? atoi(getenv("WAT_GEMM_EW16")) : 1;
    if (wide && g.epi == TC_EPI_BF16 && g.act == 1) return launch_tc2<16, 5>(g, num_sms, st);     // GELU epilogue
    return launch_tc2<8, 5>(g, num_sms, st);
  }
  if (g.N % 256 == 0) return launch_tc<256>(g, num_sms, st);
  if (g.N % 128 == 0) return launch_tc<128>(g, num_sms, st);
  return cudaErrorInvalidValue;
}

}  // namespace wat
