// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM wrappers (inline PTX),
// small math helpers.  No CUTLASS/CuTe: descriptor encodings are written out by hand below.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace wat {

// ------------------------------------------------------------------------------------------
// watchdog: every mbarrier wait is bounded (wall time, %globaltimer) so a protocol bug traps instead of
// hanging the GPU (a hung box is a strike on the shared pool).
#ifndef WAT_WATCHDOG_NS
#define WAT_WATCHDOG_NS 4000000000ull
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint
// expires) instead of polling, so a waiting producer / MMA thread does not steal issue slots from the math warps
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
static __device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity) {
  const uint64_t t0 = global_timer_ns();
  while (!mbar_try_wait(bar, parity)) {
    if (global_timer_ns() - t0 > WAT_WATCHDOG_NS) {
      printf("wat: mbarrier watchdog block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
// non-blocking phase check (try_wait suspends the warp for a hardware time limit when the phase is still open; test_wait
// returns at once) - used to issue several checks back to back before falling into a wait
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// latency-critical hand-offs: poll without the suspend hint (wake-up from the hinted sleep costs ~a microsecond)
__device__ __forceinline__ bool mbar_try_wait_nohint(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// named barrier among `count` threads of the CTA (id 1..15; 0 is __syncthreads): ~30 cycles, far cheaper than an mbarrier round trip
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  uint32_t polls = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait_nohint(bar, parity)) {
    if ((++polls & 0xFFFFu) == 0) {
      const uint64_t t = global_timer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > WAT_WATCHDOG_NS) {
        printf("wat: mbarrier watchdog (spin) block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
        __trap();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// pull a tensor-map box into L2 only (no smem destination, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
               : "memory");
}
// TMA store of a smem box to global memory (bulk-group completion); rows / columns outside the tensor map are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the issuing thread's bulk stores have finished READING their smem source (it may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... all but the most recently committed group
__device__ __forceinline__ void tma_store_wait_read_but1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (M = 128 rows = TMEM lanes, K packed two bf16 per 32-bit column,
// i.e. 8 columns per K = 16 step) is read from tensor memory, so only B costs shared-memory bandwidth
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------ CTA pair (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on an mbarrier of another CTA of the cluster.  Relaxed: the hand-overs it is used for (TMEM accumulator
// drained) are ordered by tcgen05.fence::before_thread_sync / after_thread_sync; a release at cluster scope compiles to
// MEMBAR.ALL.GPU + ERRBAR, which makes the arriving thread wait for all of its earlier global stores (the epilogue's
// output) - ncu showed 17% of the GELU GEMM's stall samples on those three instructions.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are signalled on an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 MMA across the CTA pair: rows 0..127 from / into the leader CTA, rows 128..255 the peer; each CTA holds
// its own A rows and half of B's N rows at the SAME smem offsets.  Issued by one thread of the leader CTA.
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (multicast) on the mbarrier at this smem offset in every CTA of `cta_mask` when the MMAs issued so far finish
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// instruction descriptor, kind::f16: bf16 x bf16 -> f32, A and B both K-major
// bits: [4,6) c_format=1(F32) | [7,10) a_format=1(BF16) | [10,13) b_format=1(BF16) | 15 a_major=0 | 16 b_major=0
//       [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// shared-memory matrix descriptor for a K-major tile stored as rows of 128 bytes (64 bf16) with the
// 128B swizzle (what TMA SWIZZLE_128B writes): 8-row groups are 1024 B apart (SBO), LBO unused.
// bits: [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SW128)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread i gets lane i's row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
               : "memory");
}

// ------------------------------------------------------------------------------------------ math
// 2^x for two values on the FMA pipe instead of MUFU (Cody-Waite: n = round(x), 2^f by a degree-3 minimax polynomial on
// [-0.5, 0.5], max relative error 7.5e-5 - far below the bf16 rounding of the P tile it feeds), exponent added as an integer.
// x is clamped to >= -126; callers guarantee x <= ~100.
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);    // 1.5 * 2^23: the sum's low mantissa bits hold round(x)
  const float2 r = __fadd2_rn(x, magic);
  const float2 nf = __fadd2_rn(r, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = __fadd2_rn(x, make_float2(-nf.x, -nf.y));
  float2 p = __ffma2_rn(f, make_float2(0.0551716685f, 0.0551716685f), make_float2(0.242611125f, 0.242611125f));
  p = __ffma2_rn(p, f, make_float2(0.693260968f, 0.693260968f));
  p = __ffma2_rn(p, f, make_float2(0.999928057f, 0.999928057f));
  return make_float2(__uint_as_float(__float_as_uint(p.x) + (__float_as_uint(r.x) << 23)),
                     __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(r.y) << 23)));
}
// The same for the attention fast path, where x is NOT bounded by a running max: x is first clamped to [-126, 128] by one
// saturating FFMA per value (u = sat((x + 126) / 254), x_c = 254 u - 126), so 2^x_c never wraps the exponent field: x >= 128
// gives inf / NaN (exponent 255) and x = 127 a finite 2^127, both of which the caller detects through the row sum.
__device__ __forceinline__ float2 ex2_poly2_clamped(float2 x) {
  const float ux = __saturatef(fmaf(x.x, 1.0f / 254.0f, 126.0f / 254.0f));
  const float uy = __saturatef(fmaf(x.y, 1.0f / 254.0f, 126.0f / 254.0f));
  const float2 u = make_float2(ux, uy), k254 = make_float2(254.0f, 254.0f);
  const float2 r = __ffma2_rn(u, k254, make_float2(12582912.0f - 126.0f, 12582912.0f - 126.0f));    // x_c + 1.5 * 2^23: low bits = round(x_c)
  const float2 n126 = __fadd2_rn(r, make_float2(-(12582912.0f - 126.0f), -(12582912.0f - 126.0f)));  // round(x_c) + 126
  const float2 f = __ffma2_rn(u, k254, make_float2(-n126.x, -n126.y));                                // x_c - round(x_c)
  float2 p = __ffma2_rn(f, make_float2(0.0551716685f, 0.0551716685f), make_float2(0.242611125f, 0.242611125f));
  p = __ffma2_rn(p, f, make_float2(0.693260968f, 0.693260968f));
  p = __ffma2_rn(p, f, make_float2(0.999928057f, 0.999928057f));
  return make_float2(__uint_as_float(__float_as_uint(p.x) + (__float_as_uint(r.x) << 23)),
                     __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(r.y) << 23)));
}
__device__ __forceinline__ float ex2_approx(float x) {           // single MUFU.EX2 (flush-to-zero, no range fix-up)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// GELU (exact-erf form, model.py nn.GELU) of 8 values with NO special-function unit.  With t = clamp(x / (3.3 sqrt 2), -1, 1)
// (one saturating FFMA, s = sat(x c + 1/2), t = 2 s - 1) and w = 2 t^2 - 1:   0.5 (1 + erf(x / sqrt 2)) = 0.5 + t G(w),
// G a degree-10 polynomial (|z| = 3.3 is where 1 - erf = 3e-6).  Max |GELU error| 6e-6 in fp32 Horner form - well inside
// the bf16 rounding of the tile it produces.  8.5 instructions per value, all on the FMA pipe: the fc1 epilogue has to
// emit 6.4 GELUs per clock per SM to keep up with the MMA, and the rcp + ex2 form needed 80% of the MUFU pipe for that.
__device__ __forceinline__ void gelu_erf_poly8(float* v) {
  float2 x[4], t[4], w[4], g[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) x[k] = make_float2(v[2 * k], v[2 * k + 1]);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 s = make_float2(__saturatef(fmaf(x[k].x, 0.10713739108887084f, 0.5f)),
                                 __saturatef(fmaf(x[k].y, 0.10713739108887084f, 0.5f)));
    t[k] = __ffma2_rn(s, make_float2(2.0f, 2.0f), make_float2(-1.0f, -1.0f));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = __ffma2_rn(__fmul2_rn(t[k], t[k]), make_float2(2.0f, 2.0f), make_float2(-1.0f, -1.0f));
#pragma unroll
  for (int k = 0; k < 4; ++k) g[k] = __ffma2_rn(w[k], make_float2(0.00599917769f, 0.00599917769f), make_float2(-0.0156302564f, -0.0156302564f));
#define WAT_GELU_STEP(c)                                                                   \
  _Pragma("unroll") for (int k = 0; k < 4; ++k) g[k] = __ffma2_rn(g[k], w[k], make_float2(c, c));
  WAT_GELU_STEP(0.0184577368f)
  WAT_GELU_STEP(-0.0294251293f)
  WAT_GELU_STEP(0.0602968112f)
  WAT_GELU_STEP(-0.0977825522f)
  WAT_GELU_STEP(0.138803765f)
  WAT_GELU_STEP(-0.188849464f)
  WAT_GELU_STEP(0.2509166f)
  WAT_GELU_STEP(-0.349210799f)
  WAT_GELU_STEP(0.706423819f)
#undef WAT_GELU_STEP
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 o = __fmul2_rn(x[k], __ffma2_rn(g[k], t[k], make_float2(0.5f, 0.5f)));
    v[2 * k] = o.x;
    v[2 * k + 1] = o.y;
  }
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// fp16 storage of the residual stream (bf16 mode): 4 values <-> 8 bytes; conversion to fp16 saturates instead of producing inf
__device__ __forceinline__ uint32_t pack_f16_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float4 f16x4_to_f32(uint32_t a, uint32_t b) {
  const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&a));
  const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&b));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace wat
