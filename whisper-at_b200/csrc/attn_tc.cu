// Non-causal multi-head self-attention for the encoder (head_dim 64, T = 1500) on tcgen05 tensor cores.
// Reference semantics: model.py:92-107 — softmax((q*s)(k*s)^T) v with s = 64^-0.25, softmax in fp32; the
// [B, H, T, T] score tensor the reference materialises is never written here.
//
// One CTA per (clip, head, 128-query tile), KV tiles of 64 keys, five warps, THREE CTAs resident per SM (57 KB smem,
// 128 + 32 TMEM columns and a 128-register budget each: registers are allocated per SM sub-partition, 15 warps -> 4 per
// sub-partition).  The exponentials (MUFU.EX2, ~16 per clock per SM) are the binding resource; three co-resident CTAs
// keep that pipe fed while any one of them waits on the tensor core.
//   warp 0      control : TMA producer AND MMA issuer in one converged warp (the elected lane issues).  Per step j:
//                         S(j+1) = Q K(j+1)^T (M128 N64 K64, both operands in smem) as soon as the softmax threads hold
//                         S(j) in registers; prefetch K(j+2) (3 stages) and V^T(j+1) (2 stages); O += P(j) V(j) with the
//                         A operand P(j) read from TENSOR MEMORY (TS-form tcgen05.mma) and V^T(j) from smem.
//   warps 1..4  softmax : thread = query row = TMEM lane.  One tcgen05.ld of the 64 scores (kept in registers), row max,
//                         P = 2^(S*c - m) with one ex2.approx per element (packed FFMA2 / FADD2 around it), bf16 pairs
//                         kept in the registers S frees and stored to TMEM with one tcgen05.st once PV(j-1) has read
//                         the previous P: the exponentials of step j overlap PV(j-1).  Lazy rescale of O, final O / l.
// Shared-memory bandwidth was the hidden limit of the earlier versions (P written to and read back from smem, Q re-read
// for every 64 keys): an SS-form M128 N64 K16 MMA reads 6 KB of operands for 32 cycles of math.  P in TMEM removes a
// third of that traffic and the generic->async proxy fence.
// K tail: 1500 = 23*64 + 28; TMA zero-fills rows >= T and those keys are excluded in the last tile only.
//
// Two passes.  The exponentials bind this kernel (head_dim 64: 128 MACs per exponential), and with them the softmax warps'
// instruction count.  Pass 0 therefore runs the softmax WITHOUT a running maximum: q arrives pre-multiplied by
// 64^-0.5 log2(e) (QKV epilogue), so P = 2^S directly - no row max, no subtraction, no rescale of O, no pv_done hand-shake for
// it.  Every quantity is floating point (P bf16 with an 8-bit exponent, O and l fp32), so a common factor 2^-m is immaterial as
// long as nothing leaves the exponent range: a row is accepted iff its sum l ends in [2^-100, 2^100] (that also catches inf
// and NaN).  If any row of the tile fails, the CTA re-initialises its barriers and repeats the tile with the classic
// online softmax (pass 1: running max with lazy rescale), whose result is the one written.  Scores of +-69 (natural units)
// before the max is subtracted are far outside what Whisper produces, so pass 1 is the exception; it is what the unit tests
// with inflated weights exercise.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace wat {

// optional per-phase clock trace of one CTA (test hook wat_dbg_attention tc=3): slot = step * 8 + k
#define AT_TRACE(base, step, k) do { if (TRACE && tr) tr[(base) + (step) * 8 + (k)] = clock64(); } while (0)

constexpr int AT_KV = 64;                     // keys per step
constexpr int AT_KSTAGE = 2;                  // K pipeline stages
constexpr int AT_NSTAGE = 2;                  // V^T pipeline stages
constexpr int AT_Q_BYTES = 128 * 64 * 2;      // 16 KB
constexpr int AT_K_BYTES = AT_KV * 64 * 2;    // 8 KB per stage
constexpr int AT_V_BYTES = 64 * AT_KV * 2;    // 8 KB per stage
constexpr int AT_SMEM = AT_Q_BYTES + AT_KSTAGE * AT_K_BYTES + AT_NSTAGE * AT_V_BYTES + 1024 + 256;
constexpr int AT_SMS = 148;                   // B200; only used to guess which tile follows a CTA on its SM slot
constexpr int AT_POLY_PAIR = 0;                // which pair (0,2,4,6; -1 = none) of every 8 scores takes the polynomial 2^x
constexpr float AT_LAZY_LOG2 = 24.0f;         // rescale O only when a row max grows by more than 2^24
constexpr int AT_TMEM_COLS = 128;             // S: [0,64)  O: [64,128)
constexpr int AT_TMEM_P_COLS = 32;            // P as packed bf16 pairs, a second allocation: 3 x (128 + 32) <= 512 columns per SM

struct AttnBars {
  uint64_t q_full;
  uint64_t k_full[AT_KSTAGE];   // "stage empty" barriers are not needed: the control warp learns that an MMA has read its
  uint64_t v_full[AT_NSTAGE];   // operands from s_free / p_full, which the softmax threads only signal after that MMA's result
  uint64_t s_full;          // S(j) written by Q K(j)^T
  uint64_t s_free;          // the 128 softmax threads hold S(j) in registers: the MMA warp may overwrite S
  uint64_t p_full;          // P(j) stored to TMEM by the 128 softmax threads
  uint64_t pv_done;         // PV(j) finished: the P tile is reusable, O is stable
  uint32_t tmem_slot, tmem_slot_p;
};

// One KV tile of the online softmax for one query row (thread == TMEM lane).  The reference max m_used is only moved
// when the true max exceeds it by more than 2^24 (lazy rescale; P is bf16 and O / l are fp32, so a P of up to 2^24 loses
// nothing: every term carries the same 2^-m_used factor and it cancels in O / l): the O accumulator in TMEM is
// rescaled only on those rare steps, and the final O / l is unchanged.  MASK handles the ragged last tile.
template <bool MASK, bool TRACE>
__device__ __forceinline__ void softmax_tile(uint32_t tS, uint32_t tO, uint32_t tP, float c_log2, int nvalid,
                                             int j, int n_kv, AttnBars* bars, float& m_used, float& l, bool& s_next,
                                             long long* tr) {
  uint32_t a[32], b[32];
  tmem_ld32(tS, a);
  tmem_ld32(tS + 32, b);
  tc_wait_ld();
  tc_fence_before();
  mbar_arrive(&bars->s_free);                                     // S is in registers: Q K(j+1)^T may overwrite the buffer
  AT_TRACE(0, j, 2);
  // four independent max chains (a single chain of 64 dependent FMNMX costs ~4 cycles each)
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    if (!MASK || i < nvalid) mx0 = fmaxf(mx0, __uint_as_float(a[i]));
    if (!MASK || i + 1 < nvalid) mx1 = fmaxf(mx1, __uint_as_float(a[i + 1]));
    if (!MASK || 32 + i < nvalid) mx2 = fmaxf(mx2, __uint_as_float(b[i]));
    if (!MASK || 33 + i < nvalid) mx3 = fmaxf(mx3, __uint_as_float(b[i + 1]));
  }
  float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
  mx *= c_log2;
  const bool need = mx > m_used + AT_LAZY_LOG2;
  if (__any_sync(0xffffffffu, need)) {
    if (j > 0) { mbar_wait_spin(&bars->pv_done, (j - 1) & 1); tc_fence_after(); }   // a rescale touches O: PV(j-1) must be done
    if (TRACE && tr) tr[j * 8 + 7] = 1;
    const float m_new = need ? mx : m_used;
    const float alpha = ex2_approx(m_used - m_new);              // 1 for rows that keep their reference
    l *= alpha;
    m_used = m_new;
    if (j > 0) {
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[32];
        tmem_ld32(tO + hh * 32, o);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st32(tO + hh * 32, o);
      }
      tc_wait_st();
    }
  }
  AT_TRACE(0, j, 3);
  const float2 c2 = make_float2(c_log2, c_log2), nm2 = make_float2(-m_used, -m_used);
  float2 ls0 = make_float2(0.f, 0.f), ls1 = ls0, ls2 = ls0, ls3 = ls0;
  uint32_t pk[32];                                                // P(j) as bf16 pairs; takes over the registers S(j) frees
#pragma unroll
  for (int g4 = 0; g4 < 8; ++g4) {
    float p[8];
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      const int i = g4 * 8 + e;
      const float s0 = __uint_as_float(i < 32 ? a[i & 31] : b[i & 31]);
      const float s1 = __uint_as_float(i < 32 ? a[(i + 1) & 31] : b[(i + 1) & 31]);
      const float2 x = __ffma2_rn(make_float2(s0, s1), c2, nm2);  // one FFMA2 for two elements
      if (e == AT_POLY_PAIR) {                                    // one pair in four on the FMA pipe: MUFU is the binding unit
        const float2 y = ex2_poly2(x);
        p[e] = y.x;
        p[e + 1] = y.y;
      } else {
        p[e] = ex2_approx(x.x);
        p[e + 1] = ex2_approx(x.y);
      }
      if (MASK && i >= nvalid) p[e] = 0.f;
      if (MASK && i + 1 >= nvalid) p[e + 1] = 0.f;
    }
    ls0 = __fadd2_rn(ls0, make_float2(p[0], p[1]));               // packed row-sum partials
    ls1 = __fadd2_rn(ls1, make_float2(p[2], p[3]));
    ls2 = __fadd2_rn(ls2, make_float2(p[4], p[5]));
    ls3 = __fadd2_rn(ls3, make_float2(p[6], p[7]));
    pk[g4 * 4 + 0] = pack_bf16(p[0], p[1]);
    pk[g4 * 4 + 1] = pack_bf16(p[2], p[3]);
    pk[g4 * 4 + 2] = pack_bf16(p[4], p[5]);
    pk[g4 * 4 + 3] = pack_bf16(p[6], p[7]);
  }
  AT_TRACE(0, j, 4);
  // the single P tile in TMEM is read by PV(j-1): only the store waits for it, the exponentials above overlapped it
  // (the check of S(j+1) is issued with it: an mbarrier check costs ~150 cycles even when the phase is complete)
  {
    const bool okp = j > 0 ? mbar_test_wait(&bars->pv_done, (j - 1) & 1) : true;
    s_next = j + 1 < n_kv ? mbar_test_wait(&bars->s_full, (j + 1) & 1) : true;
    if (!okp) mbar_wait_spin(&bars->pv_done, (j - 1) & 1);
    tc_fence_after();
  }
  tmem_st32(tP, pk);
  tc_wait_st();
  const float2 t = __fadd2_rn(__fadd2_rn(ls0, ls1), __fadd2_rn(ls2, ls3));
  l += t.x + t.y;
}

// Pass 0: P = 2^S with no reference maximum (see the header).  POLY selects which pairs of every 8 scores take the FMA-pipe
// polynomial instead of MUFU.EX2: bit p of POLY = pair p (scores 2p, 2p+1) of the group; the pattern alternates between the low
// and the high nibble on odd groups so that e.g. 0x31 gives 1 of 4 and 2 of 4 pairs in turn (37.5%).
template <bool MASK, bool TRACE, int POLY>
__device__ __forceinline__ void softmax_tile_fast(uint32_t tS, uint32_t tP, int nvalid, int j, int n_kv, AttnBars* bars, float& l,
                                                  bool& s_next, long long* tr) {
  constexpr int NC = AT_KV;
  uint32_t a[NC];
  {
    uint32_t lo[32], hi[32];
    tmem_ld32(tS, lo);
    tmem_ld32(tS + 32, hi);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) { a[i] = lo[i]; a[32 + i] = hi[i]; }
  }
  tc_fence_before();
  mbar_arrive(&bars->s_free);                                     // S is in registers: Q K(j+1)^T may overwrite the buffer
  AT_TRACE(0, j, 2);
  float2 ls0 = make_float2(0.f, 0.f), ls1 = ls0, ls2 = ls0, ls3 = ls0;
  uint32_t pk[NC / 2];
#pragma unroll
  for (int g4 = 0; g4 < NC / 8; ++g4) {
    float p[8];
    const int pat = (g4 & 1) ? (POLY >> 4) & 15 : POLY & 15;
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      const int i = g4 * 8 + e;
      const float2 x = make_float2(__uint_as_float(a[i]), __uint_as_float(a[i + 1]));
      if ((pat >> (e >> 1)) & 1) {
        const float2 y = ex2_poly2_clamped(x);
        p[e] = y.x;
        p[e + 1] = y.y;
      } else {
        p[e] = ex2_approx(x.x);
        p[e + 1] = ex2_approx(x.y);
      }
      if (MASK && i >= nvalid) p[e] = 0.f;
      if (MASK && i + 1 >= nvalid) p[e + 1] = 0.f;
    }
    ls0 = __fadd2_rn(ls0, make_float2(p[0], p[1]));
    ls1 = __fadd2_rn(ls1, make_float2(p[2], p[3]));
    ls2 = __fadd2_rn(ls2, make_float2(p[4], p[5]));
    ls3 = __fadd2_rn(ls3, make_float2(p[6], p[7]));
    pk[g4 * 4 + 0] = pack_bf16(p[0], p[1]);
    pk[g4 * 4 + 1] = pack_bf16(p[2], p[3]);
    pk[g4 * 4 + 2] = pack_bf16(p[4], p[5]);
    pk[g4 * 4 + 3] = pack_bf16(p[6], p[7]);
  }
  AT_TRACE(0, j, 4);
  {
    const bool okp = j > 0 ? mbar_test_wait(&bars->pv_done, (j - 1) & 1) : true;
    s_next = j + 1 < n_kv ? mbar_test_wait(&bars->s_full, (j + 1) & 1) : true;
    if (!okp) mbar_wait_spin(&bars->pv_done, (j - 1) & 1);
    tc_fence_after();
  }
  tmem_st32(tP, pk);
  tc_wait_st();
  const float2 t = __fadd2_rn(__fadd2_rn(ls0, ls1), __fadd2_rn(ls2, ls3));
  l += t.x + t.y;
}

template <bool TRACE, int POLY>
__global__ void __launch_bounds__(160, 3)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmVT, __nv_bfloat16* __restrict__ out, int T, int D, int n_head,
               int q_tiles, float c_log2, int first_pass, long long* __restrict__ trace, unsigned int* __restrict__ n_repeat) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_Q_BYTES;
  uint8_t* sV = sK + AT_KSTAGE * AT_K_BYTES;
  AttnBars* bars = reinterpret_cast<AttnBars*>(sV + AT_NSTAGE * AT_V_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % q_tiles;
  const int bh = blockIdx.x / q_tiles;
  const int h = bh % n_head, b = bh / n_head;
  const int n_kv = (T + AT_KV - 1) / AT_KV;

  // warps 0..3 are the softmax warps (TMEM lane quadrant = warp), the LAST warp is the control warp (placing it first made
  // no difference)
  constexpr int ctrl = 4, first_sm = 0;
  if (warp == ctrl && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmVT);
  }
  if (warp == first_sm) {
    tmem_alloc(&bars->tmem_slot, AT_TMEM_COLS);
    tmem_alloc(&bars->tmem_slot_p, AT_TMEM_P_COLS);
    tmem_relinquish();
  }
  __shared__ int s_bad;

#pragma unroll 1
  for (int pass = first_pass; pass < 2; ++pass) {
  const bool fast = pass == 0;
  if (warp == ctrl && lane == 0) {
    if (pass > first_pass) {                                      // second pass of this tile: every phase of pass 0 has completed
      mbar_inval(&bars->q_full);
      for (int s = 0; s < AT_KSTAGE; ++s) mbar_inval(&bars->k_full[s]);
      for (int s = 0; s < AT_NSTAGE; ++s) mbar_inval(&bars->v_full[s]);
      mbar_inval(&bars->s_full); mbar_inval(&bars->p_full); mbar_inval(&bars->pv_done); mbar_inval(&bars->s_free);
      if (n_repeat) atomicAdd(n_repeat, 1u);
    }
    s_bad = 0;
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < AT_KSTAGE; ++s) mbar_init(&bars->k_full[s], 1);
    for (int s = 0; s < AT_NSTAGE; ++s) mbar_init(&bars->v_full[s], 1);
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->p_full, 128);                                 // one arrival per softmax thread (= query row)
    mbar_init(&bars->pv_done, 1);
    mbar_init(&bars->s_free, 128);
    fence_mbar_init();
    // first loads right away: their latency (the Q tile is always a first touch) overlaps the TMEM allocation and the CTA
    // barrier below; nobody else touches these barriers before that barrier
    mbar_expect_tx(&bars->q_full, AT_Q_BYTES);
    tma_load_3d(sQ, &tmQ, &bars->q_full, h * 64, qt * 128, b);
    mbar_expect_tx(&bars->k_full[0], AT_K_BYTES);
    tma_load_3d(sK, &tmK, &bars->k_full[0], D + h * 64, 0, b);
    mbar_expect_tx(&bars->v_full[0], AT_V_BYTES);
    tma_load_2d(sV, &tmVT, &bars->v_full[0], 0, bh * VT_ROWS);
    if (n_kv > 1) {
      mbar_expect_tx(&bars->k_full[1], AT_K_BYTES);
      tma_load_3d(sK + AT_K_BYTES, &tmK, &bars->k_full[1], D + h * 64, AT_KV, b);
    }
    // and the first boxes of the tile that will follow this CTA on its SM slot (three CTAs per SM) go to L2
    const int nxt = blockIdx.x + 3 * AT_SMS;
    if (pass == first_pass && nxt < (int)gridDim.x) {
      const int nqt = nxt % q_tiles, nbh = nxt / q_tiles, nh = nbh % n_head, nbb = nbh / n_head;
      tma_prefetch_l2_3d(&tmQ, nh * 64, nqt * 128, nbb);
      tma_prefetch_l2_3d(&tmK, D + nh * 64, 0, nbb);
      tma_prefetch_l2_3d(&tmK, D + nh * 64, AT_KV, nbb);
      tma_prefetch_l2_2d(&tmVT, 0, nbh * VT_ROWS);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_slot;
  const uint32_t tmem_O = tmem_base + 64;
  const uint32_t tmem_P = bars->tmem_slot_p;

  if (warp == ctrl) {
    // Control warp: TMA producer and MMA issuer in one converged warp (only the elected lane issues).  Five warps per CTA
    // keep three CTAs on an SM at <= 4 warps per sub-partition, i.e. a 128-register budget for the softmax warps.
    // K(j+2) and V^T(j+1) are requested in the two issue blocks of step j (see the loop), K(0), K(1), V^T(0) and Q above.
    constexpr uint32_t idesc = make_idesc_bf16(128, 64);           // both MMAs are M128 N64
    const uint64_t dQ = make_smem_desc_sw128(smem_u32(sQ));
    long long* tr = (trace && blockIdx.x == gridDim.x / 2 && lane == 0) ? trace : nullptr;
    mbar_wait_spin(&bars->q_full, 0);
    mbar_wait_spin(&bars->k_full[0], 0);
    tc_fence_after();
    if (elect_one()) {
      const uint64_t dK = make_smem_desc_sw128(smem_u32(sK));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base, dQ + 2 * k, dK + 2 * k, idesc, k != 0);
      umma_commit(&bars->s_full);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      AT_TRACE(512, j, 0);
      // Two hand-overs per step, each one barrier-check group (issued back to back with the non-blocking test_wait: a
      // check costs ~150 cycles even when its phase is complete) and one issue block:
      //  1. S(j) sits in registers (s_free)  -> Q K(j+1)^T, and K(j+2) is requested into the stage Q K(j)^T has read
      //     (it has: the softmax threads signalled s_free(j) after loading its result);
      //  2. P(j) is in TMEM (p_full)         -> P(j) V(j), and V^T(j+1) is requested into the stage PV(j-1) has read
      //     (it has: the softmax threads stored P(j) only after pv_done(j-1)).
      if (j + 1 < n_kv) {
        const uint32_t phk = ((j + 1) >> 1) & 1;
        const bool okk = mbar_test_wait(&bars->k_full[st ^ 1], phk);
        const bool oks = mbar_test_wait(&bars->s_free, j & 1);
        if (!okk) mbar_wait_spin(&bars->k_full[st ^ 1], phk);
        AT_TRACE(512, j, 5);
        if (!oks) mbar_wait_spin(&bars->s_free, j & 1);
        AT_TRACE(512, j, 6);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t dK = make_smem_desc_sw128(smem_u32(sK + (st ^ 1) * AT_K_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base, dQ + 2 * k, dK + 2 * k, idesc, k != 0);
          umma_commit(&bars->s_full);
          AT_TRACE(512, j, 7);
          if (j + 2 < n_kv) {
            mbar_expect_tx(&bars->k_full[st], AT_K_BYTES);
            tma_load_3d(sK + st * AT_K_BYTES, &tmK, &bars->k_full[st], D + h * 64, (j + 2) * AT_KV, b);
          }
        }
        __syncwarp();
      }
      AT_TRACE(512, j, 1);
      {
        const bool okp = mbar_test_wait(&bars->p_full, j & 1);
        const bool okv = mbar_test_wait(&bars->v_full[st], ph);
        if (!okp) mbar_wait_spin(&bars->p_full, j & 1);
        AT_TRACE(512, j, 2);
        if (!okv) mbar_wait_spin(&bars->v_full[st], ph);
      }
      AT_TRACE(512, j, 3);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dV = make_smem_desc_sw128(smem_u32(sV + st * AT_V_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem_O, tmem_P + 8 * k, dV + 2 * k, idesc, (j | k) != 0);
        umma_commit(&bars->pv_done);
        if (j + 1 < n_kv) {
          mbar_expect_tx(&bars->v_full[st ^ 1], AT_V_BYTES);
          tma_load_2d(sV + (st ^ 1) * AT_V_BYTES, &tmVT, &bars->v_full[st ^ 1], (j + 1) * AT_KV, bh * VT_ROWS);
        }
      }
      __syncwarp();
      AT_TRACE(512, j, 4);
    }
  } else {
    const int q = warp;
    const int r = q * 32 + lane;                                  // query row in tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    float m_used = -INFINITY, l = 0.f;                            // reference max (log2 units, may lag by <= 24), row sum
    long long* tr = (trace && blockIdx.x == gridDim.x / 2 && warp == first_sm && lane == 0) ? trace : nullptr;
    bool s_ready = false;                                         // S(j) already seen complete by the previous step
    for (int j = 0; j < n_kv; ++j) {
      AT_TRACE(0, j, 0);
      if (!s_ready) mbar_wait_spin(&bars->s_full, j & 1);
      tc_fence_after();
      AT_TRACE(0, j, 1);
      const int nvalid = T - j * AT_KV;
      const uint32_t tS = tmem_base + lane_off;
      if (fast) {
        if (nvalid >= AT_KV) softmax_tile_fast<false, TRACE, POLY>(tS, tmem_P + lane_off, AT_KV, j, n_kv, bars, l, s_ready, tr);
        else softmax_tile_fast<true, TRACE, POLY>(tS, tmem_P + lane_off, nvalid, j, n_kv, bars, l, s_ready, tr);
      } else {
        if (nvalid >= AT_KV) softmax_tile<false, TRACE>(tS, tmem_O + lane_off, tmem_P + lane_off, c_log2, AT_KV, j, n_kv, bars, m_used, l, s_ready, tr);
        else softmax_tile<true, TRACE>(tS, tmem_O + lane_off, tmem_P + lane_off, c_log2, nvalid, j, n_kv, bars, m_used, l, s_ready, tr);
      }
      AT_TRACE(0, j, 5);
      tc_fence_before();
      mbar_arrive(&bars->p_full);
      AT_TRACE(0, j, 6);
    }
    mbar_wait_spin(&bars->pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    // pass 0 is accepted iff every row sum stayed inside the exponent range (inf and NaN fail the comparisons too)
    if (fast && !(l >= 7.888609052e-31f && l <= 1.267650600e30f)) s_bad = 1;
    named_bar_sync(1, 128);                                        // the four softmax warps: s_bad is final
    if (!(fast && s_bad)) {
      // O / l
      const int tq = qt * 128 + r;
      constexpr int ncol = 64, c0 = 0;
      __nv_bfloat16* op = out + ((long long)b * T + tq) * D + h * 64 + c0;
      const float inv = 1.0f / l;
#pragma unroll 1
      for (int cc = 0; cc < ncol; cc += 32) {
        uint32_t v0[32];
        tmem_ld32(tmem_O + lane_off + c0 + cc, v0);
        tc_wait_ld();
        if (tq < T) {
#pragma unroll
          for (int i = 0; i < 32; i += 8)
            *reinterpret_cast<uint4*>(op + cc + i) =
                make_uint4(pack_bf16(__uint_as_float(v0[i]) * inv, __uint_as_float(v0[i + 1]) * inv),
                           pack_bf16(__uint_as_float(v0[i + 2]) * inv, __uint_as_float(v0[i + 3]) * inv),
                           pack_bf16(__uint_as_float(v0[i + 4]) * inv, __uint_as_float(v0[i + 5]) * inv),
                           pack_bf16(__uint_as_float(v0[i + 6]) * inv, __uint_as_float(v0[i + 7]) * inv));
        }
      }
    }
  }
  // end of the pass: the control warp joins the softmax warps; a rejected pass 0 is repeated as pass 1
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (!(fast && s_bad)) break;
  __syncthreads();                                                // everyone has read s_bad before it is cleared
  }

  tc_fence_before();
  __syncthreads();
  if (warp == first_sm) { tc_fence_after(); tmem_dealloc(bars->tmem_slot_p, AT_TMEM_P_COLS); tmem_dealloc(bars->tmem_slot, AT_TMEM_COLS); }
}

static bool make_map_bf16(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                          const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// POLY patterns of softmax_tile_fast (share of the exponentials evaluated on the FMA pipe): 0x11 = 25%, 0x31 = 37.5%, 0x33 = 50%
template <int POLY>
static cudaError_t launch_attn_variant(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmVT, __nv_bfloat16* out, int grid,
                                       int T, int D, int n_head, int q_tiles, float c_log2, int first_pass, long long* trace,
                                       unsigned int* n_repeat, cudaStream_t st) {
  static unsigned long long attr_mask = 0, attr_mask_tr = 0;
  if (cudaError_t e = opt_in_smem(attn_tc_kernel<false, POLY>, AT_SMEM, attr_mask); e != cudaSuccess) return e;
  if (cudaError_t e = opt_in_smem(attn_tc_kernel<true, POLY>, AT_SMEM, attr_mask_tr); e != cudaSuccess) return e;
  constexpr int threads = 160;
  if (trace) attn_tc_kernel<true, POLY><<<grid, threads, AT_SMEM, st>>>(tmQ, tmK, tmVT, out, T, D, n_head, q_tiles, c_log2, first_pass, trace, n_repeat);
  else attn_tc_kernel<false, POLY><<<grid, threads, AT_SMEM, st>>>(tmQ, tmK, tmVT, out, T, D, n_head, q_tiles, c_log2, first_pass, nullptr, n_repeat);
  return cudaGetLastError();
}

// q_prescaled: the q half of qk already carries the factor 64^-0.5 log2(e) (GemmTc::q_scale in the QKV epilogue); only then can
// the max-free first pass run.  n_repeat (optional, device): incremented once per tile that had to be repeated with pass 1.
cudaError_t launch_attn_tc(const __nv_bfloat16* qk, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int T, int Tpad,
                           int n_head, bool q_prescaled, cudaStream_t st, long long* trace, unsigned int* n_repeat) {
  const int D = n_head * 64;
  if (Tpad % AT_KV || Tpad < ((T + AT_KV - 1) / AT_KV) * AT_KV) return cudaErrorInvalidValue;
  CUtensorMap tmQ, tmK, tmVT;
  {
    cuuint64_t dims[3] = {(cuuint64_t)2 * D, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)2 * D * 2, (cuuint64_t)T * 2 * D * 2};
    cuuint32_t boxq[3] = {64, 128, 1}, boxk[3] = {64, AT_KV, 1};
    if (!make_map_bf16(&tmQ, qk, 3, dims, strides, boxq)) return cudaErrorInvalidValue;
    if (!make_map_bf16(&tmK, qk, 3, dims, strides, boxk)) return cudaErrorInvalidValue;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Tpad, (cuuint64_t)B * n_head * VT_ROWS};
    cuuint64_t strides[1] = {(cuuint64_t)Tpad * 2};
    cuuint32_t box[2] = {AT_KV, 64};
    if (!make_map_bf16(&tmVT, vt, 2, dims, strides, box)) return cudaErrorInvalidValue;
  }
  const int q_tiles = (T + 127) / 128;
  const int grid = B * n_head * q_tiles;
  const float c_log2 = q_prescaled ? 1.0f : AT_QSCALE;
  static const int fast_on = getenv("WAT_ATTN_FAST") ? atoi(getenv("WAT_ATTN_FAST")) : 1;
  static const int poly = getenv("WAT_ATTN_POLY") ? atoi(getenv("WAT_ATTN_POLY")) : 1;
  const int first_pass = (fast_on && q_prescaled) ? 0 : 1;
  // Measured without gain on top of this kernel (round 2): loading S from TMEM in two halves with the second load in flight
  // during the first half's exponentials, and asking for PV(j-1)'s completion at the start of the step (79.6-81.8 ms vs 81.0
  // on the same box) - the per-step latencies are hidden by the three co-resident CTAs, not exposed.
  // One elected mbarrier arrival per softmax warp (barrier counts 4 instead of 128; both arrivals follow a warp-collective
  // tcgen05.wait) was measured too: 81.5-82.3 ms against 78.8-79.0 ms on the same box - the __syncwarp and the divergent
  // branch cost more than the 32 per-thread arrivals they replace.
  // Two threads per query row in the max-free pass (9 warps per CTA, each softmax thread 32 of the 64 keys of a step) was built
  // and measured: 94 ms per step against 81 ms for one thread per row on the same box - more warps per step cost more in
  // hand-overs and registers (72 per thread) than the shorter per-thread exponential phase gains.  Removed again.
#define WAT_ATTN_LAUNCH(P) launch_attn_variant<P>(tmQ, tmK, tmVT, out, grid, T, D, n_head, q_tiles, c_log2, first_pass, trace, n_repeat, st)
  if (poly == 0) return WAT_ATTN_LAUNCH(0x11);
  if (poly == 2) return WAT_ATTN_LAUNCH(0x33);
  return WAT_ATTN_LAUNCH(0x31);
#undef WAT_ATTN_LAUNCH
}

}  // namespace wat
