// PTX wrappers shared by the tcgen05 kernels (sm_100a): mbarrier, bulk-async copy, TMEM, UMMA descriptors.
#pragma once
#include "common.cuh"

namespace mrinr {

constexpr long long kWatchdogCycles = 4000000000LL;   // ~2 s: a stuck pipeline traps instead of hanging the GPU

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase flips (or the hint expires)
// instead of spinning.  Spinning matters here: a quarter-rate MUFU.SIN leaves only ~2.5 free issue slots per 8 cycles
// on its SM sub-partition (tools/mufu_bench.cu), so every instruction a waiting warp issues delays the epilogue.
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok;
}
static __device__ __noinline__ __attribute__((noreturn)) void mbar_timeout(int32_t* errflag, int code) {
  if (errflag) atomicExch(errflag, code);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int32_t* errflag, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kWatchdogCycles) mbar_timeout(errflag, code);
  }
}
// wait used by the single-thread roles (MMA issuer, producer, forwarder): backs off with nanosleep between probes so
// that the spinning warp does not take issue slots from the two epilogue warps of its SM sub-partition
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, int32_t* errflag, int code,
                                                  unsigned ns = 64) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(ns);
    if (clock64() - t0 > kWatchdogCycles) mbar_timeout(errflag, code);
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// the same copy, delivered to the same shared-memory offset (and mbarrier offset) of every CTA in `mask`
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// completion of this CTA's prior MMAs -> one arrival on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, M=128 N=256 K=16, f16/bf16 in, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version
  return d;
}
// instruction descriptor (kind::f16): D fp32, A/B f16 (0) or bf16 (1), both K-major, M=128, N=256
__host__ __device__ constexpr uint32_t make_idesc(int bf16, int M, int N) {
  return (1u << 4) | ((uint32_t)bf16 << 7) | ((uint32_t)bf16 << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  if (BF16) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
  } else {
    const __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// hidden activation: Sine (modulated_siren.py:54) or Morlet (:80).  MUFU.SIN after the 1/2pi scaling
// keeps |err| <= max(2^-21, |x| 2^-23); hidden pre-activations are O(1).
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// fp32-accurate sine without the library's slow path (a call that forces every live register to be spilled around
// it): three-term Cody-Waite reduction by pi/2 + degree-9 / degree-8 minimax polynomials.  |error| <= ~1 ulp for
// |x| <= 1e4 (the output pre-activation is O(1)); beyond that the reduction loses bits gradually instead of
// switching to Payne-Hanek.
__device__ __forceinline__ float sin_accurate(float x) {
  const float qf = rintf(x * 0.636619772367581343f);
  const int k = (int)qf;
  float r = fmaf(qf, -1.57079601287841796875f, x);
  r = fmaf(qf, -3.13916473151482683e-7f, r);
  r = fmaf(qf, -5.39030253e-15f, r);
  const float r2 = r * r;
  float ps = fmaf(r2, 2.60831598e-6f, -1.98106907e-4f);
  ps = fmaf(ps, r2, 8.33307858e-3f);
  ps = fmaf(ps, r2, -1.66666597e-1f);
  const float sn = fmaf(ps * r2, r, r);
  float pc = fmaf(r2, 2.44331568e-5f, -1.38873163e-3f);
  pc = fmaf(pc, r2, 4.16666418e-2f);
  pc = fmaf(pc, r2, -0.5f);
  const float cs = fmaf(pc, r2, 1.0f);
  const float v = (k & 1) ? cs : sn;
  return (k & 2) ? -v : v;
}
template <int ACT, bool W0ONE>
__device__ __forceinline__ float act_fast(float x, float w0) {
  const float s = __sinf(W0ONE ? x : w0 * x);
  if (ACT == MRINR_ACT_MORLET) return s * fast_ex2(x * x * -0.72134752044448170368f);   // exp(-x^2/2)
  return s;
}


// ---- cluster / cta_group::2 variants -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster.  Default (.release.cta)
// semantics on purpose: `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR and cost 23 % of all issue slots
// (profiles/r01_siren_v3a.md).  The data being published was made visible to the async proxy of the writing SM by
// fence.proxy.async before this arrive, and its only consumer is that SM's own tensor core.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_n(uint32_t bar, uint32_t cta, uint32_t count) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra], %2;\n\t}" ::"r"(bar), "r"(cta), "r"(count)
      : "memory");
}
// (plain acquire.cta wait: an acquire at cluster scope would invalidate this SM's L1 on every probe)
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int32_t* errflag, int code) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > kWatchdogCycles) mbar_timeout(errflag, code);
  }
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// completion of all prior MMAs of this thread -> arrive on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
// D[tmem, both CTAs] (+)= A[256 x 16: 128 rows per CTA] * B[256 x 16: 128 rows per CTA]^T, issued by the leader CTA
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// One lane of a converged warp (elect.sync): the tcgen05 issue pattern that lets the compiler keep descriptors in
// uniform registers.  Issuing from a divergent `if (lane == 0)` region instead makes ptxas wrap every UTCHMMA in an
// ELECT / R2UR waterfall loop (~25 instructions, ~200 cycles per MMA: measured, profiles/r01_siren_v4a.md).
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(p));
  return p != 0;
}
// descriptor = {lo, hi}: only the low word (start address, LBO) changes between K steps
__device__ __forceinline__ void umma_f16_pair_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi,
                                                   uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3fffu) | (((lbo_bytes >> 4) & 0x3fffu) << 16);
}
__device__ __forceinline__ uint32_t smem_desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14);   // version 1, no swizzle, base offset 0
}

// ---- order-pinned epilogue arithmetic -----------------------------------------------------------------------
// The epilogue is MUFU-bound (one sine per activation, 4 lanes/clk per SM sub-partition = one warp instruction every
// 8 cycles).  Left to itself ptxas batches the 32 MUFU.SIN of a chunk back to back and the dependent multiplies /
// packs after them, so a warp alternates between "only MUFU" and "no MUFU" stretches and the unit idles ~30 % of the
// time with two warps per sub-partition.  These volatile wrappers pin the program order so that the source can
// interleave: sines of group g, then the multiplies and packs of group g-1.
__device__ __forceinline__ float vsin(float x) {
  float y;
#ifdef MRINR_POWER_NO_SIN      // tools/power_split.py only: wrong results, the special-function unit stays idle
  asm volatile("mul.f32 %0, %1, 0f3F000000;" : "=f"(y) : "f"(x));
#else
  asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
#endif
  return y;
}
// Sine of a feature whose modulation is m: with MRINR_V5_SKIP_DEAD (measurement builds only, `make skipdead`) the
// special-function instruction is predicated off when m == 0 -- the product sin(x) * 0 is 0 either way.
__device__ __forceinline__ float vsin_live(float x, float m) {
#ifdef MRINR_V5_SKIP_DEAD
  // a warp vote makes the branch uniform (ptxas speculates a lane-predicated sin.approx: MUFU.SIN + select); a tile's
  // rows share their patch's modulation, so the vote is unanimous except in the two-patch remainder tiles
  float y = 0.f;
  if (__any_sync(0xffffffffu, m != 0.f)) y = vsin(x);
  return y;
#else
  (void)m;
  return vsin(x);
#endif
}
__device__ __forceinline__ float vmul(float a, float b) {
  float y;
  asm volatile("mul.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b));
  return y;
}
template <bool BF16>
__device__ __forceinline__ uint32_t vpack2(float lo, float hi) {
  uint32_t y;
  if (BF16) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
  else      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
  return y;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// ---- packed fp32 pairs (FMUL2 / FFMA2 / FADD2: one instruction, two lanes of work) ---------------------------
// The arithmetic wrappers are volatile (part of the order-pinned epilogue, see above); the pack / unpack moves are not,
// so that the compiler can keep constants and register pairs wherever it likes.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t vmul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm volatile("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t vadd2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm volatile("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t vsub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm volatile("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t vfma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// Morlet envelope exp(-x^2/2) (modulated_siren.py:80) for two pre-activations, on the FMA pipe: the special-function
// unit already evaluates one sine per activation and is the epilogue's bottleneck (16 results/clk/SM); a second MUFU
// op per activation (ex2.approx) doubled the hidden phases.  2^y with y = -x^2 log2(e)/2: round-to-integer by the
// magic-number trick, degree-4 polynomial of the fraction in [-0.5, 0.5] (relative error 2.6e-6; coefficients fitted
// for minimal relative error, checked in tests/test_host_logic.py), exponent added with an integer shift-add.
// x^2 is clamped at 174 (|x| > 13.2 -> ~1e-38 instead of a wrapped exponent).
#define GAUSS_ASM asm volatile
__device__ __forceinline__ uint64_t gmul2(uint64_t a, uint64_t b) {
  uint64_t r;
  GAUSS_ASM("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t gsub2(uint64_t a, uint64_t b) {
  uint64_t r;
  GAUSS_ASM("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t gfma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  GAUSS_ASM("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t gauss2(float x0, float x1) {
  const uint64_t X = pk2(x0, x1);
  float t0, t1;
  upk2(gmul2(X, X), t0, t1);
  t0 = fminf(t0, 174.0f);
  t1 = fminf(t1, 174.0f);
  const uint64_t T = pk2(t0, t1);
  const uint64_t NEGC = 0xBF38AA3BBF38AA3BULL;                     // -log2(e)/2, twice
  const uint64_t Z = gfma2(T, NEGC, 0x4B4000004B400000ULL);        // + 1.5 * 2^23: round(y) in the low mantissa bits
  const uint64_t NNEG = gsub2(0x4B4000004B400000ULL, Z);           // -n, n = round(y)
  const uint64_t F = gfma2(T, NEGC, NNEG);                         // y - n  in [-0.5, 0.5]
  uint64_t Pp = gfma2(0x3C1CCBEA3C1CCBEAULL, F, 0x3D650A203D650A20ULL);
  Pp = gfma2(Pp, F, 0x3E76036D3E76036DULL);
  Pp = gfma2(Pp, F, 0x3F31706E3F31706EULL);
  Pp = gfma2(Pp, F, 0x3F7FFFF43F7FFFF4ULL);
  float p0, p1, z0, z1;
  upk2(Pp, p0, p1);
  upk2(Z, z0, z1);
  const float g0 = __int_as_float(__float_as_int(p0) + (__float_as_int(z0) << 23));
  const float g1 = __int_as_float(__float_as_int(p1) + (__float_as_int(z1) << 23));
  return pk2(g0, g1);
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

struct SirenTcParams {
  const float* table0;      // [C,256]
  const uint16_t* table16;  // [C,256]  the same table rounded to the operand format (fp16 / bf16)
  const uint16_t* w16;      // [(L-1)][32][256][8]
  const uint16_t* w16p;     // [(L-1)][2][32][128][8]   (cta_group::2 path)
  const uint16_t* w16q;     // [(L-1)][2][34][128][8]   (cta_group::2 path, bias carried by a 17th K step)
  const uint16_t* w16x3;    // [(L-1)][2 (hi pass, lo pass)][2][34][128][8]   fp16x3 mode: W = hi + lo
  const float* layer0;      // [3][256]  W_0[:,0], W_0[:,1], b_0
  const float* grid;        // [C][2]
  float w0_initial;
  const float* bias;        // [L][256]
  const float* last_w;      // [256]
  const float* last_b;      // [1]
  const float* mods;        // [L,B,256]
  const int32_t* idx;       // compacted patch list or null
  const int32_t* nactive;   // device scalar or null
  float* out;               // [B,C]
  int32_t* errflag;
  long long B;
  int C, L;
  float w0;
};

}  // namespace mrinr
