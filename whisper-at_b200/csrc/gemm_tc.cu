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
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;                       // 512 or 256: powers of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct GemmTcDev {
  const float* bias;
  void* C; long long ldc;
  const float* R; long long ldr; int r_mod;
  int M, N, K, act, epi;
  __nv_bfloat16* vt; int seq_T; int seq_Tpad; int n_head; int D;
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
      res[it] = *reinterpret_cast<const float4*>(g.R + rrow * g.ldr + n0 + cc);
    }
  }
}

// `res` (optional): the residual block of THIS chunk, loaded by the caller ahead of time; it is refilled with the block at
// (nxt_row_base, nxt_n0) before this chunk's stores are issued, so a residual read is always one chunk ahead of its use
// (the residual stream is updated in place, but a chunk's own columns are only read before they are written).
__device__ __forceinline__ void tc_epilogue_chunk(const GemmTcDev& g, int row, bool row_ok, int n0, const uint32_t (&r)[32], int vb, int vtok,
                                                  const float* bias_chunk, float* stage = nullptr, int lane = 0, long long* tr2 = nullptr,
                                                  float4 (*res_io)[8] = nullptr, int nxt_row_base = -1, int nxt_n0 = 0) {
  if (n0 >= g.N) return;
  if (stage != nullptr && (g.epi == TC_EPI_F32_RES || g.epi == TC_EPI_F32)) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    if (g.bias) {
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
    const int cc = (lane & 7) * 4;
    // all residual loads first (R aliases C for the in-place residual stream, so the compiler would otherwise
    // serialise load -> add -> store per row and expose one memory latency per iteration)
    float4 res_local[8];
    float4 (&res)[8] = res_io ? *res_io : res_local;
    if (g.epi == TC_EPI_F32_RES && !res_io) load_residual_chunk(g, row_base, n0, lane, res);
    float4 t[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      t[it] = *reinterpret_cast<const float4*>(stage + (it * 4 + (lane >> 3)) * 36 + cc);
      if (g.epi == TC_EPI_F32_RES) { t[it].x += res[it].x; t[it].y += res[it].y; t[it].z += res[it].z; t[it].w += res[it].w; }
    }
    if (res_io && nxt_row_base >= 0) load_residual_chunk(g, nxt_row_base, nxt_n0, lane, res);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int grow = row_base + it * 4 + (lane >> 3);
      if (grow < g.M) *reinterpret_cast<float4*>(reinterpret_cast<float*>(g.C) + (long long)grow * g.ldc + n0 + cc) = t[it];
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
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[i + e]);
      if (g.bias) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias_chunk + i);
        const float4 b1 = *reinterpret_cast<const float4*>(bias_chunk + i + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (g.act == 1) gelu_erf_poly8(v);
      *reinterpret_cast<uint4*>(cp + i) =
          make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
    return;
  }
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  if (g.bias) {
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
  if (false) {
  } else if (g.epi == TC_EPI_QKV) {
    const int vc = n0 - 2 * g.D;                         // h * 64 + e ; a 32-chunk never straddles a head
    const int h = vc >> 6, e0 = vc & 63;
    __nv_bfloat16* vp = g.vt + (((long long)vb * g.n_head + h) * VT_ROWS + e0) * g.seq_Tpad + vtok;
#pragma unroll
    for (int i = 0; i < 32; ++i) vp[(long long)i * g.seq_Tpad] = __float2bfloat16_rn(v[i]);
  } else {
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

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTcDev g) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
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
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const int row = m_blk * TC_BM + q * 32 + lane;
      const bool row_ok = row < g.M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      int vb = 0, vtok = 0;
      if (g.epi == TC_EPI_QKV && row_ok) { vb = row / g.seq_T; vtok = row - vb * g.seq_T; }
#pragma unroll 1
      for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_row + c0, r);
        tc_wait_ld();
        if (c0 + 32 == (chalf + 1) * (BN / 2)) { tc_fence_before(); mbar_arrive(&tmem_empty[as]); }
        tc_epilogue_chunk(g, row, row_ok, n_blk * BN + c0, r, vb, vtok, g.bias ? g.bias + n_blk * BN + c0 : nullptr);
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
constexpr int TC2_STAGES = 5;
constexpr int TC2_STAGE_BYTES = 2 * TC_A_BYTES;                  // A 128x64 + W-half 128x64
// EW = epilogue warps: 8 (two per TMEM lane quadrant, 128 columns each; fp32 outputs use a per-warp transpose tile) or
// 16 (four per quadrant, 64 columns each) for the bf16 + GELU epilogue, whose per-tile latency with 8 warps (~13k
// cycles) exceeds the tile's 10k MMA cycles at K = 1280.
template <int EW> struct Tc2Cfg {
  static constexpr int THREADS = 64 + 32 * EW;
  static constexpr int COLS = 256 / (EW / 4);                    // columns per epilogue warp
  static constexpr int WARP_FLOATS = (EW == 8 ? 32 * 36 : 0) + COLS;   // transpose tile (8-warp kernel only) + bias slice
  static constexpr int SMEM_BYTES = TC2_STAGES * TC2_STAGE_BYTES + 1024 + 256 + EW * WARP_FLOATS * 4;
};

template <int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Tc2Cfg<EW>::THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmTcDev g) {
  constexpr int BN = 256;
  using Cfg2 = Tc2Cfg<EW>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC2_STAGES * TC2_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + TC2_STAGES;
  uint64_t* tmem_full = bars + 2 * TC2_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + TC2_STAGES * TC2_STAGE_BYTES + 256);

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
    for (int s = 0; s < TC2_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 2 * EW); }   // EW warps x 2 CTAs
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc_pair(tmem_slot, 512); tmem_relinquish_pair(); }
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
    const bool res_ahead = EW == 8 && g.epi == TC_EPI_F32_RES;
    float4 res[8];
    if (res_ahead && cluster_id < total_tiles)
      load_residual_chunk(g, (cluster_id / n_tiles) * 2 * TC_BM + (int)rank * TC_BM + q * 32, (cluster_id % n_tiles) * BN + chalf * COLS, lane, res);
    for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, ++it) {
      const int n_blk = tile % n_tiles, m_pair = tile / n_tiles;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      // this warp's 128 bias values -> smem while the MMAs of the tile are still running
      float* my_stage = epi_stage + (warp - 2) * Cfg2::WARP_FLOATS;
      float* bias_s = my_stage + (EW == 8 ? 32 * 36 : 0);
      if (g.bias && lane * 4 < COLS) {
        const int nb = n_blk * BN + chalf * COLS + lane * 4;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nb < g.N) b4 = __ldg(reinterpret_cast<const float4*>(g.bias + nb));
        *reinterpret_cast<float4*>(bias_s + lane * 4) = b4;
      }
      __syncwarp();
      long long* tr = (g.trace && blockIdx.x == 0 && threadIdx.x == 64 && it < 24) ? g.trace + it * 8 : nullptr;
      if (tr) tr[0] = clock64();
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      if (tr) tr[1] = clock64();
      const int row = m_pair * 2 * TC_BM + (int)rank * TC_BM + q * 32 + lane;
      const bool row_ok = row < g.M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
      int vb = 0, vtok = 0;
      if (g.epi == TC_EPI_QKV && row_ok) { vb = row / g.seq_T; vtok = row - vb * g.seq_T; }
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
        if (res_ahead) {
          int nrb = -1, nn0 = 0;
          if (c0 + 32 < (chalf + 1) * COLS) { nrb = row - lane; nn0 = n_blk * BN + c0 + 32; }
          else if (tile + n_clusters < total_tiles) {
            const int nt = tile + n_clusters;
            nrb = (nt / n_tiles) * 2 * TC_BM + (int)rank * TC_BM + q * 32;
            nn0 = (nt % n_tiles) * BN + chalf * COLS;
          }
          tc_epilogue_chunk(g, row, row_ok, n_blk * BN + c0, r, vb, vtok, g.bias ? bias_s + (c0 - chalf * COLS) : nullptr, my_stage, lane, tr2, &res, nrb, nn0);
        } else {
          tc_epilogue_chunk(g, row, row_ok, n_blk * BN + c0, r, vb, vtok, g.bias ? bias_s + (c0 - chalf * COLS) : nullptr, EW == 8 ? my_stage : nullptr, lane, tr2);
        }
        if (tr) tr[2 + (c0 - chalf * COLS) / 32] = clock64();
      }
    }
  }

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
  d.M = g.M; d.N = g.N; d.K = g.K; d.act = g.act; d.epi = g.epi;
  d.vt = g.vt; d.seq_T = g.seq_T; d.seq_Tpad = g.seq_Tpad; d.n_head = g.n_head; d.D = g.N / 3;
  d.trace = nullptr;
  const int m_tiles = (g.M + TC_BM - 1) / TC_BM, n_tiles = (g.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles;
  const int grid = total < num_sms ? total : num_sms;
  gemm_tc_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, d);
  return cudaGetLastError();
}


template <int EW>
static cudaError_t launch_tc2(const GemmTc& g, int num_sms, cudaStream_t st) {
  using Cfg2 = Tc2Cfg<EW>;
  static unsigned long long attr_mask = 0;
  if (cudaError_t e = opt_in_smem(gemm_tc2_kernel<EW>, Cfg2::SMEM_BYTES, attr_mask); e != cudaSuccess) return e;
  CUtensorMap tmA, tmB;
  if (!make_map_2d(&tmA, g.A, g.M, g.K, g.lda, TC_BM)) return cudaErrorInvalidValue;
  if (!make_map_2d(&tmB, g.W, g.N, g.K, g.K, 128)) return cudaErrorInvalidValue;      // each CTA loads half of the 256 W rows
  GemmTcDev d;
  d.bias = g.bias; d.C = g.C; d.ldc = g.ldc; d.R = g.R; d.ldr = g.ldr; d.r_mod = g.r_mod;
  d.M = g.M; d.N = g.N; d.K = g.K; d.act = g.act; d.epi = g.epi;
  d.vt = g.vt; d.seq_T = g.seq_T; d.seq_Tpad = g.seq_Tpad; d.n_head = g.n_head; d.D = g.N / 3;
  d.trace = g.trace;
  const int total = ((g.M + 255) / 256) * ((g.N + 255) / 256);
  int clusters = num_sms / 2;
  if (clusters > total) clusters = total;
  gemm_tc2_kernel<EW><<<2 * clusters, Cfg2::THREADS, Cfg2::SMEM_BYTES, st>>>(tmA, tmB, d);
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

cudaError_t launch_gemm_tc(const GemmTc& g, int num_sms, cudaStream_t st) {
  if (g.M <= 0) return cudaSuccess;
  if ((g.K & 7) || (g.lda & 7) || (g.N & 31) || (g.ldc & 7)) return cudaErrorInvalidValue;
  if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.W) & 15)) return cudaErrorInvalidValue;
  if (g.epi == TC_EPI_QKV && ((g.N % 3) || ((g.N / 3) & 63) || g.seq_T <= 0)) return cudaErrorInvalidValue;
  if (g.N % 256 == 0 && (g.force_pair > 0 || (g.force_pair == 0 && pair_mode() && g.M >= 256))) {
    static const int wide = getenv("WAT_GEMM_EW16") ? atoi(getenv("WAT_GEMM_EW16")) : 1;
    if (wide && g.epi == TC_EPI_BF16 && g.act == 1) return launch_tc2<16>(g, num_sms, st);     // GELU epilogue
    return launch_tc2<8>(g, num_sms, st);
  }
  if (g.N % 256 == 0) return launch_tc<256>(g, num_sms, st);
  if (g.N % 128 == 0) return launch_tc<128>(g, num_sms, st);
  return cudaErrorInvalidValue;
}

}  // namespace wat
