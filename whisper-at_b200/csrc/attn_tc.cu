// Non-causal multi-head self-attention for the encoder (head_dim 64, T = 1500) on tcgen05 tensor cores.
// Reference semantics: model.py:92-107 — softmax((q*s)(k*s)^T) v with s = 64^-0.25, softmax in fp32; the
// [B, H, T, T] score tensor the reference materialises is never written here.
//
// One CTA per (clip, head, 128-query tile); two CTAs are resident per SM (256 TMEM columns, 96 KB smem each) so
// one CTA's softmax overlaps the other's MMAs.
//   warp 0      TMA producer : Q tile once; K tile [128 keys x 64] double-buffered; V^T tile [64 x 128 keys]
//   warp 1      MMA issuer   : S = Q K^T (M128 N128 K64) into TMEM; O += P V (M128 N64 K128), P from smem
//   warps 2..5  softmax      : thread = query row = TMEM lane: online softmax on S (exp2, fp32), P -> bf16 into the
//                              128B-swizzled smem tile the MMA reads, O rescale in TMEM, final O / l -> bf16
// K tail: 1500 = 11*128 + 92; TMA zero-fills rows >= T and those keys are masked to -inf before the softmax.
#include "common.cuh"
#include "kernels.h"

namespace wat {

constexpr int AT_THREADS = 192;
constexpr int AT_Q_BYTES = 128 * 64 * 2;      // 16 KB
constexpr int AT_K_BYTES = 128 * 64 * 2;      // 16 KB per stage, 2 stages
constexpr int AT_V_BYTES = 64 * 128 * 2;      // 16 KB (two 64x64 K-blocks of 8 KB)
constexpr int AT_P_BYTES = 128 * 128 * 2;     // 32 KB (two 128x64 K-blocks of 16 KB)
constexpr int AT_SMEM = AT_Q_BYTES + 2 * AT_K_BYTES + AT_V_BYTES + AT_P_BYTES + 1024 + 128;
constexpr int AT_TMEM_COLS = 256;             // S: cols [0,128), O: cols [128,192)

// One KV tile of the online softmax for one query row (thread == TMEM lane).
//   pass 1: row max of S (two chunks in flight);  pass 2: P = 2^(S*c - m) -> bf16 into the swizzled smem tile.
// The reference max m_used is only moved when the true max exceeds it by more than 2^8 (lazy rescale, as in
// FlashAttention-4): P stays <= 256, the O accumulator in TMEM is rescaled only on those rare steps, and the final
// O / l is unchanged.  MASK handles the ragged last tile (keys >= nvalid are excluded).
template <bool MASK>
__device__ __forceinline__ void softmax_tile(uint32_t tS, uint32_t tO, uint8_t* p_row, int sw, float c_log2, int nvalid,
                                             uint64_t* prev_pv, uint32_t prev_par, float& m_used, float& l) {
  uint32_t a[32], b[32];
  float mx = -INFINITY;
  tmem_ld32(tS, a);
  tmem_ld32(tS + 32, b);
  tc_wait_ld();
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (!MASK || i < nvalid) mx = fmaxf(mx, __uint_as_float(a[i]));
    if (!MASK || 32 + i < nvalid) mx = fmaxf(mx, __uint_as_float(b[i]));
  }
  tmem_ld32(tS + 64, a);
  tmem_ld32(tS + 96, b);
  tc_wait_ld();
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (!MASK || 64 + i < nvalid) mx = fmaxf(mx, __uint_as_float(a[i]));
    if (!MASK || 96 + i < nvalid) mx = fmaxf(mx, __uint_as_float(b[i]));
  }
  mx *= c_log2;
  // PV(j-1) must have finished before O is rescaled or the P tile is overwritten (it ran under the max pass above)
  const bool have_o = prev_pv != nullptr;
  if (have_o) { mbar_wait_spin(prev_pv, prev_par); tc_fence_after(); }
  const bool need = mx > m_used + 8.0f;
  if (__any_sync(0xffffffffu, need)) {
    const float m_new = need ? mx : m_used;
    const float alpha = ex2_approx(m_used - m_new);              // 1 for rows that keep their reference
    l *= alpha;
    m_used = m_new;
    if (have_o) {
      tmem_ld32(tO, a);
      tmem_ld32(tO + 32, b);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        a[i] = __float_as_uint(__uint_as_float(a[i]) * alpha);
        b[i] = __float_as_uint(__uint_as_float(b[i]) * alpha);
      }
      tmem_st32(tO, a);
      tmem_st32(tO + 32, b);
      tc_wait_st();
    }
  }
  const float neg_m = -m_used;
  float lsum = 0.f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    tmem_ld32(tS + half * 64, a);
    tmem_ld32(tS + half * 64 + 32, b);
    tc_wait_ld();
    uint8_t* blk = p_row + half * 16384;                          // K-block (64 keys) of the P tile
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
      float p[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int i = g4 * 8 + e;                                 // column inside this half
        const float sv = __uint_as_float(i < 32 ? a[i & 31] : b[i & 31]);
        float pv = ex2_approx(fmaf(sv, c_log2, neg_m));
        if (MASK && half * 64 + i >= nvalid) pv = 0.f;
        p[e] = pv;
        lsum += pv;
      }
      *reinterpret_cast<uint4*>(blk + ((g4 ^ sw) << 4)) =
          make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
    }
  }
  l += lsum;
}

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmVT,
               __nv_bfloat16* __restrict__ out, int T, int D, int n_head, int q_tiles, float c_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_Q_BYTES;
  uint8_t* sV = sK + 2 * AT_K_BYTES;
  uint8_t* sP = sV + AT_V_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + AT_P_BYTES);
  uint64_t* q_full = bars;          // 1
  uint64_t* k_full = bars + 1;      // 2
  uint64_t* k_empty = bars + 3;     // 2
  uint64_t* v_full = bars + 5;      // 1
  uint64_t* v_empty = bars + 6;     // 1
  uint64_t* s_full = bars + 7;      // 1
  uint64_t* p_full = bars + 8;      // 1 (128 arrivals)
  uint64_t* pv_done = bars + 9;     // 1: PV(j) finished (O and the P tile may be touched again); the last one = O final
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % q_tiles;
  const int bh = blockIdx.x / q_tiles;
  const int h = bh % n_head, b = bh / n_head;
  const int n_kv = (T + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQK);
    tma_prefetch_desc(&tmVT);
    mbar_init(q_full, 1);
    mbar_init(&k_full[0], 1); mbar_init(&k_full[1], 1);
    mbar_init(&k_empty[0], 1); mbar_init(&k_empty[1], 1);
    mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, AT_TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, AT_Q_BYTES);
      tma_load_3d(sQ, &tmQK, q_full, h * 64, qt * 128, b);
      for (int j = 0; j < n_kv; ++j) {
        const int ks = j & 1;
        const uint32_t kph = (j >> 1) & 1;
        mbar_wait(&k_empty[ks], kph ^ 1);
        mbar_expect_tx(&k_full[ks], AT_K_BYTES);
        tma_load_3d(sK + ks * AT_K_BYTES, &tmQK, &k_full[ks], D + h * 64, j * 128, b);
        mbar_wait(v_empty, (j & 1) ^ 1);
        mbar_expect_tx(v_full, AT_V_BYTES);
        tma_load_2d(sV, &tmVT, v_full, j * 128, bh * 64);
        tma_load_2d(sV + AT_V_BYTES / 2, &tmVT, v_full, j * 128 + 64, bh * 64);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64);
      const uint64_t dQ = make_smem_desc_sw128(smem_u32(sQ));
      const uint64_t dP = make_smem_desc_sw128(smem_u32(sP));
      const uint64_t dV = make_smem_desc_sw128(smem_u32(sV));
      mbar_wait(q_full, 0);
      // S(0) = Q K(0)^T
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      {
        const uint64_t dK = make_smem_desc_sw128(smem_u32(sK));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_S, dQ + 2 * k, dK + 2 * k, idesc_s, k != 0);
        umma_commit(s_full);
        umma_commit(&k_empty[0]);
      }
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait_spin(p_full, j & 1);                            // softmax(j) has consumed S(j) and written P(j)
        tc_fence_after();
        if (j + 1 < n_kv) {                                       // S(j+1) first: the next softmax starts while PV(j) runs
          const int ks = (j + 1) & 1;
          mbar_wait(&k_full[ks], ((j + 1) >> 1) & 1);
          tc_fence_after();
          const uint64_t dK = make_smem_desc_sw128(smem_u32(sK + ks * AT_K_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_S, dQ + 2 * k, dK + 2 * k, idesc_s, k != 0);
          umma_commit(s_full);
          umma_commit(&k_empty[ks]);
        }
        mbar_wait(v_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int kb = k >> 2, kk = k & 3;
          umma_bf16_ss(tmem_O, dP + kb * (16384 >> 4) + 2 * kk, dV + kb * (8192 >> 4) + 2 * kk, idesc_o, (j | k) != 0);
        }
        umma_commit(v_empty);
        umma_commit(pv_done);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;                                  // query row in tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    float m_used = -INFINITY, l = 0.f;                            // reference max (log2 units, may lag by <= 8) and row sum
    uint8_t* p_row = sP + r * 128;
    const int sw = r & 7;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait_spin(s_full, j & 1);
      tc_fence_after();
      const int nvalid = T - j * 128;
      uint64_t* prev_pv = j > 0 ? pv_done : nullptr;
      const uint32_t prev_par = (j - 1) & 1;
      if (nvalid >= 128) softmax_tile<false>(tmem_S + lane_off, tmem_O + lane_off, p_row, sw, c_log2, 128, prev_pv, prev_par, m_used, l);
      else softmax_tile<true>(tmem_S + lane_off, tmem_O + lane_off, p_row, sw, c_log2, nvalid, prev_pv, prev_par, m_used, l);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    mbar_wait_spin(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    const int tq = qt * 128 + r;
    const float inv = 1.0f / l;
    __nv_bfloat16* op = out + ((long long)b * T + tq) * D + h * 64;
    uint32_t v0[32], v1[32];
    tmem_ld32(tmem_O + lane_off, v0);
    tmem_ld32(tmem_O + lane_off + 32, v1);
    tc_wait_ld();
    if (tq < T) {
#pragma unroll
      for (int i = 0; i < 32; i += 8)
        *reinterpret_cast<uint4*>(op + i) =
            make_uint4(pack_bf16(__uint_as_float(v0[i]) * inv, __uint_as_float(v0[i + 1]) * inv),
                       pack_bf16(__uint_as_float(v0[i + 2]) * inv, __uint_as_float(v0[i + 3]) * inv),
                       pack_bf16(__uint_as_float(v0[i + 4]) * inv, __uint_as_float(v0[i + 5]) * inv),
                       pack_bf16(__uint_as_float(v0[i + 6]) * inv, __uint_as_float(v0[i + 7]) * inv));
#pragma unroll
      for (int i = 0; i < 32; i += 8)
        *reinterpret_cast<uint4*>(op + 32 + i) =
            make_uint4(pack_bf16(__uint_as_float(v1[i]) * inv, __uint_as_float(v1[i + 1]) * inv),
                       pack_bf16(__uint_as_float(v1[i + 2]) * inv, __uint_as_float(v1[i + 3]) * inv),
                       pack_bf16(__uint_as_float(v1[i + 4]) * inv, __uint_as_float(v1[i + 5]) * inv),
                       pack_bf16(__uint_as_float(v1[i + 6]) * inv, __uint_as_float(v1[i + 7]) * inv));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, AT_TMEM_COLS); }
}

cudaError_t launch_attn_tc(const __nv_bfloat16* qk, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int T, int Tpad,
                           int n_head, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return cudaErrorInvalidValue;
  const int D = n_head * 64;
  CUtensorMap tmQK, tmVT;
  {
    cuuint64_t dims[3] = {(cuuint64_t)2 * D, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)2 * D * 2, (cuuint64_t)T * 2 * D * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&tmQK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(qk), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Tpad, (cuuint64_t)B * n_head * 64};
    cuuint64_t strides[1] = {(cuuint64_t)Tpad * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    if (enc(&tmVT, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(vt), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  const int q_tiles = (T + 127) / 128;
  const float c_log2 = 0.125f * 1.4426950408889634f;
  attn_tc_kernel<<<B * n_head * q_tiles, AT_THREADS, AT_SMEM, st>>>(tmQK, tmVT, out, T, D, n_head, q_tiles, c_log2);
  return cudaGetLastError();
}

}  // namespace wat
