// Fused modulated-SIREN synthesis kernel, version 5: coordinate-block tiles.
//
// What the v4 timeline said (profiles/r01_siren.md, tools/timeline.py): per cluster iteration (2 tiles per CTA) the
// epilogue warps spend 2 x ~4100 cycles in the layer-0 phases, 6 x ~2650 in the hidden phases and 2 x ~3300 in the
// output phases = ~33 600 cycles, against 17 408 cycles of tensor work.  The special-function unit (one MUFU.SIN per
// activation, 2048 cycles per tile and layer) is the wall, and layer 0 alone is 20 % of it -- although
// sin(w0_initial (W_0 g + b_0)) does not depend on the patch at all (modulated_siren.py:448: every patch sees the
// same grid).  A polynomial sine on the FMA pipe does not help (tools/epi_bench.cu: <= 10 %).
//
// This version removes layer 0 from the special-function unit:
//   * A tile is no longer 128 consecutive rows of the flat [patch, coordinate] stream but one COORDINATE BLOCK of one
//     patch: block tau covers coordinates 128 tau .. 128 tau + 127 (tau < C / 128); the remaining C % 128 coordinates
//     of up to two patches share one tile (C = 576: 4 full blocks + one 2 x 64 block per pair of patches).
//   * Every cluster owns a contiguous range of patches and walks it block-major: all its tiles of block 0, then block
//     1, ...  So a thread (one tile row, 128 or 64 columns) sees the same coordinate for hundreds of consecutive tiles and keeps
//     its slice of the pre-computed layer-0 table act_0(W_0 g_c + b_0) (MrinrPacked::d_table0) in REGISTERS as packed
//     16-bit pairs.  The layer-0 phase is then: unpack, multiply by the modulation, pack, store -- no sine.
//     The table is re-read from global memory only when the block changes (C/128 + 1 times per sub-block of 128
//     patches).
//   * A tile belongs to one patch, so its modulation vectors are warp-uniform.  The producer thread streams them into
//     a small shared-memory ring with cp.async.bulk (1 KB per patch and layer) and the epilogue reads them with
//     broadcast LDS.128 -- no per-row global loads, no L1 misses on the critical path (v4: 19 % long-scoreboard stalls).
//
// Everything else is v4: CTA pairs (tcgen05 cta_group::2), two tile slots per CTA (ping-pong), bias through a 17th
// K step, weights through a 5-slab ring, MMAs issued by a converged warp with elect.sync.
//
// Precision modes (template parameter PREC):
//   MRINR_PREC_FP16 / MRINR_PREC_BF16: one MMA per product, 16-bit operands, fp32 accumulation -- the fast path.
//   MRINR_PREC_FP16X3: fp32-class accuracy on the tensor cores for weights / modulations large enough that an 11-bit
//     significand no longer meets the 1e-3 bound (W x 2 with dense modulations).  Activations and weights
//     are split into hi = rn16(v), lo = rn16(v - hi) and a product is three MMAs, A_hi W_hi + A_lo W_hi + A_hi W_lo
//     (the dropped lo x lo term is 2^-22 relative), as dense_tc.cu does for the modulator.  The second A operand takes
//     the shared memory of the second tile slot, so a CTA has ONE tile in flight (no ping-pong: the tensor core waits
//     for the epilogue) and the weights stream through the ring twice per layer (hi pass incl. the bias step, lo
//     pass).  The layer-0 table is read from global memory in fp32 instead of living in registers as 16-bit pairs.
#include "tc_ptx.cuh"
#include "siren_sched.h"

namespace mrinr {
namespace v5 {

constexpr int kH = 256;
constexpr int kSlabBytes = 16384;    // K=64 x N=128 (this CTA's half of the output rows) x 2 B
constexpr int kBiasSlabBytes = 4096; // K=16 x N=128 x 2 B
constexpr int kNumSlabs = 5;         // 4 weight slabs + the bias step
constexpr int kLayerBytes = 4 * kSlabBytes + kBiasSlabBytes;   // per (layer, rank) in the packed array
#ifndef MRINR_V5_EPI_WARPS
#define MRINR_V5_EPI_WARPS 8
#endif
constexpr int kEpiWarps = MRINR_V5_EPI_WARPS;  // 8 or 16: 2 or 4 epilogue warps per SM sub-partition
constexpr int kColGroups = kEpiWarps / 4;      // a warp owns rows 32(w&3)..+31 and columns [kCols*(w>>2), +kCols)
constexpr int kCols = 256 / kColGroups;        // 128 or 64 columns per thread per phase
constexpr int kPairs = kCols / 32;             // pairs of 16-column chunks per phase
constexpr int kThreads = kEpiWarps * 32 + 128;  // + one warpgroup: MMA issuer, producer, two idle warps
// setmaxnreg works on whole warpgroups: the epilogue warpgroups grow to kRegsEpilogue registers and the helper
// warpgroup shrinks to kRegsOther (per sub-partition: 2 x 232 + 40 <= 512).  At the launch cap of 168 the compiler
// spilled the tile-walk state, and with 226 KB of shared memory there is next to no L1 left: every reload was an
// L2 round trip (~300 cycles) on the critical path between phases (tools/timeline.py).
// The pool setmaxnreg redistributes is what the CTA was launched with (threads x launch registers), so
//   8 epilogue warps: 384 x 168 = 64512 = 256 x 232 + 128 x 40;  16: 640 x 96 = 61440 = 512 x 112 + 128 x 32
// -- a request beyond the pool never completes (the 16-warp build with 112 / 40 hung for exactly this reason).
#ifndef MRINR_V5_REGS_EPI
#define MRINR_V5_REGS_EPI (MRINR_V5_EPI_WARPS == 8 ? 232 : 112)
#endif
#ifndef MRINR_V5_REGS_OTHER
#define MRINR_V5_REGS_OTHER (MRINR_V5_EPI_WARPS == 8 ? 40 : 32)
#endif
constexpr int kRegsEpilogue = MRINR_V5_REGS_EPI, kRegsOther = MRINR_V5_REGS_OTHER;
static_assert(kEpiWarps * 32 * kRegsEpilogue + 128 * kRegsOther <= (kEpiWarps == 8 ? 168 : 96) * kThreads,
              "setmaxnreg targets exceed the CTA's register pool");
// Register budget: the register file is split per SM sub-partition (16 384 each) and the 10 (18) warps of the CTA
// land 3 (5) on some sub-partition, so the cap is 168 (96) registers per thread -- not 65536 / kThreads.
constexpr int kMaxLayers = 16;
constexpr int kTmemCols = 512;
constexpr int kModStages = 4;                  // modulation ring depth per slot (power of two)

constexpr int kOffA = 0;                                          // [2 slots][64 KB]
constexpr int kOffOnes = 2 * 65536;                               // [2 kc][128][8] constant "ones" K step
constexpr int kOffW = kOffOnes + 4096;                            // 4 x 16 KB + 4 KB
constexpr int kOffMods = kOffW + kLayerBytes;                     // [kModStages][2 slots][kMaxSub][256] f32
constexpr int kOffLastW = kOffMods + kModStages * 2 * kMaxSub * kH * 4;   // [256] f32
constexpr int kOffPart = kOffLastW + kH * 4;                      // [2 slots][3 column groups][128] f32 partial dots
constexpr int kOffBar = kOffPart + 2 * 3 * kTileM * 4;
constexpr int kNumBars = 48;
constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
static_assert(kSmemBytes <= 232448, "shared memory budget");

constexpr int kBarWFull = 0;     // [5] local, transaction based
constexpr int kBarWPeer = 5;     // [5] leader: the peer's slab has landed
constexpr int kBarWEmpty = 10;   // [5] both CTAs, via multicast commit
constexpr int kBarAFull = 15;    // [2] leader: one arrival per epilogue warp of both CTAs: operand of slot s complete
constexpr int kBarAccFull = 17;  // [2] both CTAs, via multicast commit
constexpr int kBarModFull = 19;  // [2 slots][kModStages] local, transaction based
constexpr int kBarModEmpty = 27; // [2 slots][kModStages] local, one arrival per epilogue warp

#ifdef MRINR_TIMELINE
// development aid: per-phase timestamps of one lane per selected warp (see tools/timeline.py); compiled out by
// default.  Timestamps are kept in a per-thread local array and written out once at the end (no atomics in the loop).
__device__ long long g_timeline[8192];
__device__ int g_timeline_n;
#define TL_DECL long long tl_buf[160]; int tl_tag[160]; int tl_n = 0;
#define TL(tag)                                                                          \
  do {                                                                                   \
    if (lane == 0 && blockIdx.x == 0 && (warp == 0 || warp == kEpiWarps) && tl_n < 160) { \
      tl_buf[tl_n] = clock64(); tl_tag[tl_n] = (tag); ++tl_n;                             \
    }                                                                                    \
  } while (0)
#define TL_FLUSH                                                                          \
  do {                                                                                   \
    if (lane == 0 && blockIdx.x == 0 && (warp == 0 || warp == kEpiWarps)) {               \
      const int base = atomicAdd(&g_timeline_n, tl_n);                                   \
      for (int _i = 0; _i < tl_n && base + _i < 2048; ++_i) {                             \
        g_timeline[(base + _i) * 4 + 0] = warp;                                          \
        g_timeline[(base + _i) * 4 + 1] = tl_tag[_i];                                    \
        g_timeline[(base + _i) * 4 + 2] = tl_buf[_i];                                    \
        g_timeline[(base + _i) * 4 + 3] = 0;                                             \
      }                                                                                  \
    }                                                                                    \
  } while (0)
#else
#define TL_DECL
#define TL(tag) do { } while (0)
#define TL_FLUSH do { } while (0)
#endif

// Packed 16-bit pair -> two fp32 values.  volatile on purpose: the table registers are loop-invariant, and without it
// the compiler hoists all 128 conversions out of the tile loop and keeps the fp32 copies in local memory.
template <bool BF16>
__device__ __forceinline__ void unpack2(uint32_t v, float& lo, float& hi) {
  if (BF16) {
    asm volatile("shl.b32 %0, %2, 16;\n\tand.b32 %1, %2, 0xffff0000;" : "=f"(lo), "=f"(hi) : "r"(v));
  } else {
    asm volatile(
        "{\n\t.reg .f16 l, h;\n\t"
        "mov.b32 {l, h}, %2;\n\t"
        "cvt.f32.f16 %0, l;\n\t"
        "cvt.f32.f16 %1, h;\n\t}"
        : "=f"(lo), "=f"(hi)
        : "r"(v));
  }
}

// Morlet envelope exp(-x^2/2) of two activations of group g (0..3) of a 16-activation chunk.  The special-function unit
// (one MUFU.SIN per activation already) and the FMA pipe are both close to their limits in the Morlet epilogue, so the
// envelope is split between them: groups whose bit is set in kMorletFmaMask use the FMA-pipe evaluation (gauss2),
// the others ex2.approx on the special-function unit.  Measured, sustained TFLOP/s of the kernel on dense modulations
// (profiles/r02_morlet_variants.txt): all MUFU (round 1) 792-805, all FMA 773-780, groups 0 and 2 on the FMA pipe
// 820-831 (0xA the same; 0x1 802, 0x7 804).  Also measured and NOT better: free instruction scheduling of the envelope
// (831), an Estrin polynomial (811), and a software-pipelined all-FMA version -- four packed chains advanced stage by
// stage with the next block's MUFU.SIN between the stages -- 767: the packed FMAs are throughput-, not latency-bound
// (an FFMA2 occupies the 32-lane FMA pipe for two issue cycles; ~10 packed operations per activation pair).
#ifndef MRINR_MORLET_FMA_MASK
#define MRINR_MORLET_FMA_MASK 0x5
#endif
constexpr int kMorletFmaMask = MRINR_MORLET_FMA_MASK;
__device__ __forceinline__ uint64_t morlet_env2(int g, float x0, float x1) {
  if ((kMorletFmaMask >> g) & 1) return gauss2(x0, x1);
  float t0, t1;
  upk2(vmul2(pk2(x0, x1), pk2(x0, x1)), t0, t1);
  float e0, e1;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0 * -0.72134752044448170368f));
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1 * -0.72134752044448170368f));
  return pk2(e0, e1);
}

template <int ACT, int PREC, bool W0ONE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) siren_tc5_kernel(const SirenTcParams P) {
  constexpr bool BF16 = (PREC == MRINR_PREC_BF16);
  constexpr bool X3 = (PREC == MRINR_PREC_FP16X3);   // split operands, three MMAs per product, one tile slot
  constexpr int kSlots = X3 ? 1 : 2;                 // tiles in flight per CTA
  constexpr int kPasses = X3 ? 2 : 1;                // trips of a layer's weights through the slab ring
  constexpr int kTpi = 2 * kSlots;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int C = P.C, L = P.L;
  const uint32_t rank = cluster_ctarank();

  float* s_mods = reinterpret_cast<float*>(smem + kOffMods);
  float* s_lastw = reinterpret_cast<float*>(smem + kOffLastW);
  float* s_part = reinterpret_cast<float*>(smem + kOffPart);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  const uint32_t sA = smem_u32(smem + kOffA);
  const uint32_t sOnes = smem_u32(smem + kOffOnes);
  const uint32_t sW = smem_u32(smem + kOffW);
  const uint32_t sMods = smem_u32(smem + kOffMods);
  const uint32_t bar0 = smem_u32(s_bar);
  auto bar = [bar0](int i) -> uint32_t { return bar0 + 8u * (uint32_t)i; };

  const long long n_act = P.nactive ? (long long)*P.nactive : P.B;
  const Sched S = make_sched(n_act, C, blockIdx.x >> 1, gridDim.x >> 1, kTpi);
  const size_t layer_stride = (size_t)P.B * kH;

  // ---- one-time setup ----
  for (int i = tid; i < kH; i += kThreads) s_lastw[i] = P.last_w[i];
  for (int i = tid; i < 4096 / 16; i += kThreads) {
    // ones block: K slots 0 and 1 of every row are 1.0 (they meet the bias hi / lo rows of B), the rest 0
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (i < kTileM) v.x = BF16 ? 0x3f803f80u : 0x3c003c00u;
    reinterpret_cast<uint4*>(smem + kOffOnes)[i] = v;
  }
  if (tid == 0) {
    for (int s = 0; s < kNumSlabs; ++s) {
      mbar_init(bar(kBarWFull + s), 1);
      mbar_init(bar(kBarWPeer + s), 1);
      mbar_init(bar(kBarWEmpty + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(kBarAFull + s), 2 * kEpiWarps);
      mbar_init(bar(kBarAccFull + s), 1);
    }
    for (int s = 0; s < 2 * kModStages; ++s) {
      mbar_init(bar(kBarModFull + s), 1);
      mbar_init(bar(kBarModEmpty + s), kEpiWarps);
    }
    fence_barrier_init();
  }
  fence_proxy_async();         // the ones block is read by the tensor core (async proxy)
  if (warp == kEpiWarps) tmem_alloc_pair(smem_u32(s_tmem), kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  TL_DECL

  if (warp < kEpiWarps) {
    // =========================== epilogue warps (both CTAs, identical) ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpilogue));
    const int q = warp & 3;
    const int cg = warp >> 2;          // column group
    const int t = q * 32 + lane;       // tile row == TMEM lane
    const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const float last_b = P.last_b ? *P.last_b : 0.f;
    const float w_mine = P.last_w[tid & (kH - 1)];   // output-layer weight of the column this thread pre-scales

    // per coordinate block (changes C/128 + 1 times per launch)
    int cur_type = -1;
    int c_row = 0;                     // this row's coordinate within the patch
    int sub_row = 0;                   // which of the tile's patches this row belongs to (remainder block only)
    int sub_prev = 0;                  // sub_row of the previous iteration's tiles (their output phase runs late)
    bool row_live = false;             // false: padding row of a remainder tile
    uint32_t T2[X3 ? 1 : kCols / 2];   // layer-0 table of (c_row, this thread's columns), packed pairs (not in X3)
    const float* trow32 = P.table0;    // X3: the fp32 table row of c_row, read per tile
    float* out0 = nullptr; float* out1 = nullptr;      // output element of this row for the tile in slot 0 / 1

    auto mod_stage = [&](int slot, uint32_t use) -> float* {
      return s_mods + (((use & (kModStages - 1)) * 2 + slot) * kMaxSub) * kH;
    };
    auto mod_full_bar = [&](int slot, uint32_t use) -> uint32_t {
      return bar(kBarModFull + slot * kModStages + (int)(use & (kModStages - 1)));
    };
    // non-blocking probe, issued early so that its latency (~100 cycles) hides behind other work; the modulation
    // vectors are requested several phases ahead, so the probe almost always succeeds
    auto peek_mods = [&](int slot, uint32_t use) -> uint32_t {
      return mbar_try_wait(mod_full_bar(slot, use), (use / kModStages) & 1u);
    };
    auto wait_mods = [&](int slot, uint32_t use) {
      mbar_wait(mod_full_bar(slot, use), (use / kModStages) & 1u, P.errflag, 6);
    };
    // this warp is done with the phase: its part of A[slot] is written (publish_a) and its modulation reads are over
    auto finish_phase = [&](int slot, uint32_t use, bool publish_a) {
      if (publish_a) fence_proxy_async();        // generic-proxy writes of A -> visible to the tensor core's reads
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(kBarModEmpty + slot * kModStages + (int)(use & (kModStages - 1))));
        if (publish_a) mbar_arrive_cluster(bar(kBarAFull + slot), 0);
      }
    };
    auto load4x4 = [&](const float* mp, int hc, float4 (&m)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) m[i] = *reinterpret_cast<const float4*>(mp + hc * 16 + i * 4);
    };
    // Four activations h = act(x) * mod of group g -> operand entries: pk = rn16(h); X3 also lo = rn16(h - pk).
    auto emit4 = [&](int g, float a0, float a1, float a2, float a3, uint32_t (&pk)[8], uint32_t (&lo)[8]) {
      pk[g * 2 + 0] = vpack2<BF16>(a0, a1);
      pk[g * 2 + 1] = vpack2<BF16>(a2, a3);
      if (X3) {
        float h0, h1, h2, h3;
        unpack2<false>(pk[g * 2 + 0], h0, h1);
        unpack2<false>(pk[g * 2 + 1], h2, h3);
        lo[g * 2 + 0] = pack2<false>(a0 - h0, a1 - h1);
        lo[g * 2 + 1] = pack2<false>(a2 - h2, a3 - h3);
      }
    };
    // 16 activations x = pre-activation (bias included) -> h = act(x) * mod, packed to 8 x (2 x 16 bit) (+ 8 residuals).
    // Order-pinned pipeline: sines of group g (special-function unit), then the FMA-pipe work of group g-1.
    auto act16_pack = [&](const uint32_t (&v)[16], const float4 (&m)[4], uint32_t (&pk)[8], uint32_t (&lo)[8]) {
      float s[16];
      auto finish_group = [&](int g) {
        const float4 mm = m[g];
        if (ACT == MRINR_ACT_SINE) {
          emit4(g, vmul(s[g * 4 + 0], mm.x), vmul(s[g * 4 + 1], mm.y), vmul(s[g * 4 + 2], mm.z), vmul(s[g * 4 + 3], mm.w),
                pk, lo);
        } else {
          // Morlet (modulated_siren.py:80): sin(w0 x) exp(-x^2/2).  One MUFU.SIN per activation as for sine; the
          // envelope runs on the FMA pipe as packed pairs (gauss2) underneath the next group's sines.
          const uint64_t h01 = vmul2(morlet_env2(g, __uint_as_float(v[g * 4 + 0]), __uint_as_float(v[g * 4 + 1])),
                                     vmul2(pk2(s[g * 4 + 0], s[g * 4 + 1]), pk2(mm.x, mm.y)));
          const uint64_t h23 = vmul2(morlet_env2(g, __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3])),
                                     vmul2(pk2(s[g * 4 + 2], s[g * 4 + 3]), pk2(mm.z, mm.w)));
          float a0, a1, a2, a3;
          upk2(h01, a0, a1);
          upk2(h23, a2, a3);
          emit4(g, a0, a1, a2, a3, pk, lo);
        }
      };
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float mg[4] = {m[g].x, m[g].y, m[g].z, m[g].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float x = __uint_as_float(v[g * 4 + i]);
          // (X3 too: a full-precision sine was measured there -- same error, the 22-bit operand split is the mode's
          // floor, not sin.approx -- and cost 23 % of its throughput: profiles/r02_precision_table.txt)
          s[g * 4 + i] = vsin_live(W0ONE ? x : P.w0 * x, mg[i]);
        }
        if (g > 0) finish_group(g - 1);
      }
      finish_group(3);
    };
    // operand entries of columns cg*kCols + hc*16 .. +15 of row t: pk into slot's A buffer; X3: lo into the second one
    auto store16 = [&](int slot, int hc, const uint32_t (&pk)[8], const uint32_t (&lo)[8]) {
      uint8_t* base = smem + kOffA + slot * 65536 + (cg * (kCols / 8) + hc * 2) * 2048 + t * 16;
#ifdef MRINR_POWER_NO_STS      // tools/power_split.py only: wrong results, the operand stores never execute (what an
      if (P.w0 != 12345.f) {   // A operand that does not go through shared memory could save at most)
        asm volatile("" ::"r"(pk[0] ^ pk[1] ^ pk[2] ^ pk[3] ^ pk[4] ^ pk[5] ^ pk[6] ^ pk[7]));
        return;
      }
#endif
#ifdef MRINR_POWER_CONST_A     // tools/power_split.py only: the stores execute, but with constant data (separates the
      {                        // energy of the stores from the energy of operand bits toggling in the tensor core)
        asm volatile("" ::"r"(pk[0] ^ pk[1] ^ pk[2] ^ pk[3] ^ pk[4] ^ pk[5] ^ pk[6] ^ pk[7]));
        const uint4 c = make_uint4(0x2e662e66u, 0x2e662e66u, 0x2e662e66u, 0x2e662e66u);      // 0.1 in fp16
        *reinterpret_cast<uint4*>(base) = c;
        *reinterpret_cast<uint4*>(base + 2048) = c;
        return;
      }
#endif
      *reinterpret_cast<uint4*>(base) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(base + 2048) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      if (X3) {
        *reinterpret_cast<uint4*>(base + 65536) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(base + 65536 + 2048) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
    };

    // output phase of a finished tile: y = sin(w0 (h_{L-1} . w_last + b_last)), h_{L-1} = act(D) * mod_{L-1}.
    // All rows of a tile share the modulation vector, so the epilogue threads first turn the ring stage into
    // mod_{L-1} * w_last in place (one or two elements per thread); a row then needs one FFMA per activation.
    auto final_phase = [&](int slot, float* outp, uint32_t ev, uint32_t use) {
      const uint32_t tcol = taddr_row + (uint32_t)slot * 256u + (uint32_t)(cg * kCols);
      TL(1000 + slot);
      wait_mods(slot, use);
      float* ring = mod_stage(slot, use);
#pragma unroll
      for (int j = tid; j < kMaxSub * kH; j += kEpiWarps * 32) ring[j] *= w_mine;
      // these generic-proxy writes must be ordered before the producer's next cp.async.bulk into this stage; fencing
      // here (not at the end of the phase) hides the fence behind the barrier and the accumulator wait
      fence_proxy_async();
      named_bar_sync(1, kEpiWarps * 32);
      const float* mp = ring + sub_prev * kH + cg * kCols;
      mbar_wait(bar(kBarAccFull + slot), ev & 1u, P.errflag, 5);
      TL(1010 + slot);
      tc_fence_after();
      uint32_t va[16], vb[16];
      tmem_ld16(tcol, va);
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
      uint64_t d01 = 0ull, d23 = 0ull;             // Morlet: packed pairs of partial dots
      auto dot16 = [&](const uint32_t (&v)[16], int hc) {
        float4 mw[4];
        load4x4(mp, hc, mw);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float x0 = __uint_as_float(v[g * 4 + 0]), x1 = __uint_as_float(v[g * 4 + 1]);
          const float x2 = __uint_as_float(v[g * 4 + 2]), x3 = __uint_as_float(v[g * 4 + 3]);
          const float h0 = vsin_live(W0ONE ? x0 : P.w0 * x0, mw[g].x), h1 = vsin_live(W0ONE ? x1 : P.w0 * x1, mw[g].y);
          const float h2 = vsin_live(W0ONE ? x2 : P.w0 * x2, mw[g].z), h3 = vsin_live(W0ONE ? x3 : P.w0 * x3, mw[g].w);
          if (ACT == MRINR_ACT_SINE) {
            d0 = fmaf(h0, mw[g].x, d0);
            d1 = fmaf(h1, mw[g].y, d1);
            d2 = fmaf(h2, mw[g].z, d2);
            d3 = fmaf(h3, mw[g].w, d3);
          } else {
            d01 = vfma2(vmul2(morlet_env2(g, x0, x1), pk2(h0, h1)), pk2(mw[g].x, mw[g].y), d01);
            d23 = vfma2(vmul2(morlet_env2(g, x2, x3), pk2(h2, h3)), pk2(mw[g].z, mw[g].w), d23);
          }
        }
      };
#pragma unroll 1
      for (int hp = 0; hp < kPairs; ++hp) {
        tmem_ld_wait();
        tmem_ld16(tcol + (uint32_t)(hp * 2 + 1) * 16u, vb);
        dot16(va, hp * 2);
        tmem_ld_wait();
        if (hp < kPairs - 1) tmem_ld16(tcol + (uint32_t)(hp * 2 + 2) * 16u, va);
        dot16(vb, hp * 2 + 1);
      }
      TL(1020 + slot);
      tc_fence_before();
      finish_phase(slot, use, false);
      if (ACT != MRINR_ACT_SINE) {
        upk2(d01, d0, d1);
        upk2(d23, d2, d3);
      }
      float dot = (d0 + d1) + (d2 + d3);
      // combine the column groups of a row: groups 1.. hand their partial dot to group 0's warp of the same quarter
      float* part = s_part + slot * 3 * kTileM;
      if (cg != 0) {
        part[(cg - 1) * kTileM + t] = dot;
        asm volatile("bar.arrive %0, %1;" ::"r"(2 + q), "r"(32 * kColGroups) : "memory");
      } else {
        asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "r"(32 * kColGroups) : "memory");
#pragma unroll
        for (int gq = 0; gq < kColGroups - 1; ++gq) dot += part[gq * kTileM + t];
        // output layer: always sine, never modulated (modulated_siren.py:211-213, :233)
        if (outp != nullptr) *outp = sin_accurate(P.w0 * (dot + last_b));
      }
      TL(1030 + slot);
    };

    Walk w;
    w.set_block(S);
    for (long long it = 0; it <= S.total; ++it) {
      const bool last = (it == S.total);                           // extra pass: only the pending output phases
      const uint32_t use0 = (uint32_t)it * (uint32_t)L;           // modulation-ring sequence number of layer 0
      const uint32_t ev0 = (uint32_t)it * (uint32_t)(L - 1);      // accumulator event of this iteration's layer 1
      sub_prev = sub_row;
      if (!last && w.type != cur_type) {
        // ---- new coordinate block: this row's coordinate and its slice of the layer-0 table ----
        cur_type = w.type;
        if (cur_type < S.n_full) {
          c_row = cur_type * kTileM + t;
          sub_row = 0;
          row_live = true;
        } else {
          const int sub = t / S.rem;
          row_live = sub < S.ksub;
          sub_row = row_live ? sub : 0;
          c_row = S.n_full * kTileM + (row_live ? t - sub * S.rem : 0);
        }
        trow32 = P.table0 + (size_t)c_row * kH + cg * kCols;
        if (!X3) {
          const float4* trow = reinterpret_cast<const float4*>(trow32);
#pragma unroll
          for (int i = 0; i < kCols / 4; ++i) {
            const float4 v = __ldg(trow + i);
            T2[i * 2 + 0] = pack2<BF16>(v.x, v.y);
            T2[i * 2 + 1] = pack2<BF16>(v.z, v.w);
          }
        }
      }
      // ---- per slot: finish the previous tile of the slot, then write the layer-0 operand of the new one ----
#pragma unroll 1
      for (int slot = 0; slot < kSlots; ++slot) {
        TL(910 + slot);
        const uint32_t ok0 = last ? 1u : peek_mods(slot, use0);
        TL(920 + slot);
        if (it > 0) final_phase(slot, slot ? out1 : out0, ev0 - 1u, use0 - 1u);
        if (last) continue;
        {
          // which patch does this row belong to?  (a phantom tile / padding row has no output)
          const int ti = w.j * kTpi + slot * 2 + (int)rank;
          const int pl = cur_type < S.n_full ? ti : ti * S.ksub + sub_row;
          float* o = nullptr;
          if (row_live && pl < w.nps) {
            const long long patch = P.idx ? (long long)P.idx[S.pa + w.base + pl] : S.pa + w.base + pl;
            o = P.out + patch * C + c_row;
          }
          if (slot) out1 = o; else out0 = o;
        }
        // layer 0 (modulated_siren.py:154-156 with dim_in = 2): h = table[c] * mod_0
        TL(2000 + slot);
        if (!ok0) wait_mods(slot, use0);
        const float* mp = mod_stage(slot, use0) + sub_row * kH + cg * kCols;
        TL(2005 + slot);
        if (!X3) {
#pragma unroll
          for (int hc = 0; hc < kCols / 16; ++hc) {
            float4 m[4];
            load4x4(mp, hc, m);
            uint32_t pk[8], lo[8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float a0, a1, a2, a3;
              unpack2<BF16>(T2[hc * 8 + g * 2 + 0], a0, a1);
              unpack2<BF16>(T2[hc * 8 + g * 2 + 1], a2, a3);
              pk[g * 2 + 0] = pack2<BF16>(a0 * m[g].x, a1 * m[g].y);
              pk[g * 2 + 1] = pack2<BF16>(a2 * m[g].z, a3 * m[g].w);
            }
            store16(slot, hc, pk, lo);
          }
        } else {
          // X3: the table in full precision, straight from global memory (L2-resident: 590 KB shared by all CTAs)
#pragma unroll 2
          for (int hc = 0; hc < kCols / 16; ++hc) {
            float4 m[4];
            load4x4(mp, hc, m);
            uint32_t pk[8], lo[8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 tv = __ldg(reinterpret_cast<const float4*>(trow32 + hc * 16 + g * 4));
              emit4(g, tv.x * m[g].x, tv.y * m[g].y, tv.z * m[g].z, tv.w * m[g].w, pk, lo);
            }
            store16(slot, hc, pk, lo);
          }
        }
        finish_phase(slot, use0, true);
        TL(2010 + slot);
      }
      if (last) break;

      // ---- hidden layers 1 .. L-2, alternating slots: the other slot's MMAs run underneath ----
      for (int l = 1; l <= L - 2; ++l) {
        const uint32_t ev = ev0 + (uint32_t)(l - 1);
        const uint32_t use = use0 + (uint32_t)l;
#pragma unroll 1
        for (int slot = 0; slot < kSlots; ++slot) {
          const uint32_t tcol = taddr_row + (uint32_t)slot * 256u + (uint32_t)(cg * kCols);
          TL(3000 + l * 10 + slot);
          const uint32_t okm = peek_mods(slot, use);
          mbar_wait(bar(kBarAccFull + slot), ev & 1u, P.errflag, 4);
          if (!okm) wait_mods(slot, use);
          const float* mp = mod_stage(slot, use) + sub_row * kH + cg * kCols;
          TL(4000 + l * 10 + slot);
          tc_fence_after();
          uint32_t va[16], vb[16];
          tmem_ld16(tcol, va);
#pragma unroll 1
          for (int hp = 0; hp < kPairs; ++hp) {
            float4 m[4];
            uint32_t pk[8], lo[8];
            tmem_ld_wait();
            tmem_ld16(tcol + (uint32_t)(hp * 2 + 1) * 16u, vb);     // next chunk lands while this one is processed
            load4x4(mp, hp * 2, m);
            act16_pack(va, m, pk, lo);
            store16(slot, hp * 2, pk, lo);
            tmem_ld_wait();
            if (hp < kPairs - 1) tmem_ld16(tcol + (uint32_t)(hp * 2 + 2) * 16u, va);
            load4x4(mp, hp * 2 + 1, m);
            act16_pack(vb, m, pk, lo);
            store16(slot, hp * 2 + 1, pk, lo);
          }
          tc_fence_before();
          finish_phase(slot, use, true);
          TL(5000 + l * 10 + slot);
        }
      }
      TL(900);
      w.next(S);
      TL(901);
    }
  } else if (warp == kEpiWarps) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
    if (rank == 0) {
      // =========================== MMA issuer (leader CTA) ===========================
      // The whole warp runs the loop converged; one elected lane issues the tcgen05 instructions, so that all
      // descriptor arithmetic stays warp-uniform (uniform registers, no per-instruction waterfall).
      const uint32_t idesc = make_idesc(BF16 ? 1 : 0, 2 * kTileM, kH);
      const uint32_t desc_hi = smem_desc_hi(128);
      const uint32_t a_lo0 = smem_desc_lo(sA, 2048);
      const uint32_t ones_lo = smem_desc_lo(sOnes, 2048);
      const uint32_t b_lo0 = smem_desc_lo(sW, 2048);
      uint32_t ev = 0;
      for (long long it = 0; it < S.total; ++it) {
        for (int l = 1; l < L; ++l, ++ev) {
          // not unrolled over the slots: the issuer shares its sub-partition's instruction cache with two epilogue
          // warps, and 12 KB of straight-line issue code evicted their loops at every phase change
#pragma unroll 1
          for (int slot = 0; slot < kSlots; ++slot) {
            const uint32_t a_lo = a_lo0 + (uint32_t)slot * (65536u >> 4);
            const uint32_t a2_lo = a_lo0 + (65536u >> 4);      // X3: the residual operand A_lo lives in the second buffer
            const uint32_t d_tmem = tmem_base + (uint32_t)slot * 256u;
            TL(6000 + l * 10 + slot);
            mbar_wait_backoff(bar(kBarAFull + slot), ev & 1u, P.errflag, 1, 32);
            TL(7000 + l * 10 + slot);
#pragma unroll 1
            for (int pass = 0; pass < kPasses; ++pass) {
              const uint32_t wev = ev * (uint32_t)kPasses + (uint32_t)pass;     // trips of the slab ring so far
#pragma unroll
              for (int s = 0; s < kNumSlabs; ++s) {
                if (slot == 0) {
                  mbar_wait_backoff(bar(kBarWFull + s), wev & 1u, P.errflag, 2, 32);
                  mbar_wait_backoff(bar(kBarWPeer + s), wev & 1u, P.errflag, 8, 32);
                }
                tc_fence_after();
                if (elect_one()) {
                  const uint32_t b_lo = b_lo0 + (uint32_t)s * (kSlabBytes >> 4);
#ifndef MRINR_POWER_NO_MMA     // tools/power_split.py only: wrong results, the tensor core stays idle
                  if (s < 4) {
                    if (pass == 0) {
#pragma unroll
                      for (int kk = 0; kk < 4; ++kk)           // A (X3: A_hi) x W (X3: W_hi)
                        umma_f16_pair_lohi(d_tmem, a_lo + (uint32_t)(s * 4 + kk) * 256u, b_lo + (uint32_t)kk * 256u,
                                           desc_hi, idesc, (s | kk) != 0 ? 1u : 0u);
                      if (X3) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)         // A_lo x W_hi
                          umma_f16_pair_lohi(d_tmem, a2_lo + (uint32_t)(s * 4 + kk) * 256u, b_lo + (uint32_t)kk * 256u,
                                             desc_hi, idesc, 1u);
                      }
                    } else {
#pragma unroll
                      for (int kk = 0; kk < 4; ++kk)           // A_hi x W_lo
                        umma_f16_pair_lohi(d_tmem, a_lo + (uint32_t)(s * 4 + kk) * 256u, b_lo + (uint32_t)kk * 256u,
                                           desc_hi, idesc, 1u);
                    }
                  } else if (pass == 0) {
#ifndef MRINR_POWER_NO_BIAS    // tools/power_split.py only: what the 17th K step (bias through the MMA) costs
                    umma_f16_pair_lohi(d_tmem, ones_lo, b_lo, desc_hi, idesc, 1u);      // + bias
#endif
                  }
#endif
                  // The last slot is the last user of a slab: hand every slab back as soon as its MMAs are issued, so
                  // that the next weights stream in underneath the remaining MMAs (one commit per slab either way;
                  // with a single hand-back at the end of the layer the ~1500-cycle L2 round trip of the reload sat
                  // between MMA(l, slot 1) and MMA(l+1, slot 0))
                  if (slot == kSlots - 1) umma_commit_pair(bar(kBarWEmpty + s), 3);
                }
                __syncwarp();
              }
            }
            if (elect_one()) umma_commit_pair(bar(kBarAccFull + slot), 3);
            __syncwarp();
            TL(8000 + l * 10 + slot);
          }
        }
      }
    } else if (lane == 0) {
      // =========================== forwarder (peer CTA): my slab has landed ===========================
      const long long n_trips = S.total * (long long)(L - 1) * kPasses;     // trips of the slab ring
      for (long long wev = 0; wev < n_trips; ++wev) {
#pragma unroll 1
        for (int s = 0; s < kNumSlabs; ++s) {
          mbar_wait_backoff(bar(kBarWFull + s), (uint32_t)wev & 1u, P.errflag, 9);
          mbar_arrive_cluster(bar(kBarWPeer + s), 0);
        }
      }
    }
    __syncwarp();
  } else if (warp > kEpiWarps + 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));      // idle warps of the helper warpgroup
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
    // ============ producer (both CTAs): own half of every weight slab + the modulation vectors of my tiles ============
    // Order per iteration: for every layer l >= 1: mods(l) [+ mods(0) of the next iteration after l = L-2], weights(l).
    // The modulations of a phase are always requested before the producer can block on a weight slab that
    // (transitively) waits for that phase.
    // The whole warp runs the loops converged and one elected lane issues the copies (as in the MMA issuer): from
    // `if (lane == 0)` every cp.async.bulk was wrapped in an ELECT / R2UR waterfall and the unrolled producer was
    // 48 KB of straight-line code streaming through the SM's instruction caches once per iteration.
    {
      uint32_t ev = 0;
      Walk w;
      w.set_block(S);
      // modulation vectors of the two slots' tiles of the walk position ww, layer l, ring sequence number itx * L + l
      auto issue_mods_of = [&](long long itx, int l, const Walk& ww) {
        const uint32_t use = (uint32_t)itx * (uint32_t)L + (uint32_t)l;
        const uint32_t stage = use & (kModStages - 1);
        const bool full = ww.type < S.n_full;
        const int ns = full ? 1 : S.ksub;
#pragma unroll 1
        for (int slot = 0; slot < kSlots; ++slot) {
          const uint32_t full_bar = bar(kBarModFull + slot * kModStages + stage);
          mbar_wait_backoff(bar(kBarModEmpty + slot * kModStages + stage), ((use / kModStages) & 1u) ^ 1u, P.errflag, 10);
          const int ti = ww.j * kTpi + slot * 2 + (int)rank;
#pragma unroll 1
          for (int sidx = 0; sidx < ns; ++sidx) {
            int pl = full ? ti : ti * S.ksub + sidx;
            if (pl >= ww.nps) pl = ww.nps - 1;                   // phantom tile / missing patch: any valid vector
            const long long patch = P.idx ? (long long)P.idx[S.pa + ww.base + pl] : S.pa + ww.base + pl;
            if (elect_one()) {
              if (sidx == 0) mbar_expect_tx(full_bar, (uint32_t)ns * kH * 4u);
              bulk_g2s(sMods + (uint32_t)(((stage * 2 + slot) * kMaxSub + sidx) * kH * 4),
                       P.mods + (size_t)l * layer_stride + (size_t)patch * kH, kH * 4, full_bar);
            }
            __syncwarp();
          }
        }
      };
      // The layer-0 vectors of iteration it+1 are requested in the middle of iteration it (after layer L-2's request:
      // their ring stage is the one an early layer of iteration it has released by then).
      if (S.total > 0) issue_mods_of(0, 0, w);
      const int l_early = L - 2 >= 1 ? L - 2 : 1;
#pragma unroll 1
      for (long long it = 0; it < S.total; ++it) {
        Walk wn = w;
        wn.next(S);
#pragma unroll 1
        for (int l = 1; l < L; ++l, ++ev) {
          issue_mods_of(it, l, w);
          if (l == l_early && it + 1 < S.total) issue_mods_of(it + 1, 0, wn);
#pragma unroll 1
          for (int pass = 0; pass < kPasses; ++pass) {
            // X3: [(L-1)][hi pass, lo pass][rank]; otherwise [(L-1)][rank]
            const uint8_t* src = reinterpret_cast<const uint8_t*>(X3 ? P.w16x3 : P.w16q) +
                                 (((size_t)(l - 1) * kPasses + pass) * 2 + rank) * kLayerBytes;
            const uint32_t wev = ev * (uint32_t)kPasses + (uint32_t)pass;
#pragma unroll 1
            for (int sl = 0; sl < kNumSlabs; ++sl) {
              const uint32_t bytes = sl < 4 ? kSlabBytes : kBiasSlabBytes;
              mbar_wait_backoff(bar(kBarWEmpty + sl), (wev & 1u) ^ 1u, P.errflag, 7);
              if (elect_one()) {
                mbar_expect_tx(bar(kBarWFull + sl), bytes);
                bulk_g2s(sW + sl * kSlabBytes, src + (size_t)sl * kSlabBytes, bytes, bar(kBarWFull + sl));
              }
              __syncwarp();
            }
          }
        }
        w = wn;
      }
    }
    __syncwarp();
  }

  TL_FLUSH;
  // ---- teardown: both CTAs must be done before the pair's TMEM is released ----
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (warp == kEpiWarps) tmem_dealloc_pair(tmem_base, kTmemCols);
}

template <int ACT, int PREC, bool W0ONE>
static int launch_one(const SirenTcParams& P, int grid, cudaStream_t st) {
  MRINR_SMEM_OPT_IN((siren_tc5_kernel<ACT, PREC, W0ONE>), kSmemBytes);
  siren_tc5_kernel<ACT, PREC, W0ONE><<<grid, kThreads, kSmemBytes, st>>>(P);
  count_launch();
  return check_launch("siren_tc5");
}

template <int ACT, int PREC>
static int launch_w0(const SirenTcParams& P, int grid, bool w0one, cudaStream_t st) {
  return w0one ? launch_one<ACT, PREC, true>(P, grid, st) : launch_one<ACT, PREC, false>(P, grid, st);
}

int launch_siren_tc_v5(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                       int64_t B, float* d_out, cudaStream_t st) {
  MRINR_REQUIRE(p->H == kH && p->L <= kMaxLayers && p->L >= 3 && p->C >= kTileM, MRINR_E_UNSUPPORTED,
                "siren_tc5: the tensor-core path needs dim_hidden 256, 3 <= num_layers <= 16 and siren_patch_size^2 >= "
                "128 (got H=%d L=%d C=%d); use MRINR_PREC_FP32", p->H, p->L, p->C);
  const bool x3 = (p->precision == MRINR_PREC_FP16X3);
  MRINR_REQUIRE(x3 ? p->d_net_w16x3 != nullptr : p->d_net_w16q != nullptr, MRINR_E_ARG,
                "siren_tc5: the weights were not packed for precision mode %d", p->precision);
  SirenTcParams P;
  P.table0 = p->d_table0; P.table16 = p->d_table16; P.w16 = p->d_net_w16; P.w16p = p->d_net_w16p;
  P.w16q = p->d_net_w16q; P.w16x3 = p->d_net_w16x3;
  P.layer0 = p->d_layer0; P.grid = p->d_grid; P.w0_initial = p->w0_initial;
  P.bias = p->d_net_bias; P.last_w = p->d_last_w;
  P.last_b = p->d_last_b; P.mods = d_mods; P.idx = d_idx; P.nactive = d_nactive; P.out = d_out;
  P.errflag = p->d_errflag; P.B = B; P.C = p->C; P.L = p->L; P.w0 = p->w0;
  // one cluster per SM pair, but never more clusters than there are groups of tiles-per-iteration patches (a cluster
  // iteration processes 4 tiles -- 2 in the fp16x3 mode -- of the same coordinate block)
  const long long tpi = x3 ? 2 : 4;
  long long clusters = p->num_sms / 2;
  if (p->synth_clusters > 0 && p->synth_clusters < clusters) clusters = p->synth_clusters;
  const long long groups = (B + tpi - 1) / tpi;
  if (clusters > groups) clusters = groups;
  if (clusters < 1) clusters = 1;
  const int grid = (int)(clusters * 2);
  const bool w0one = (p->w0 == 1.0f);
  const bool morlet = (p->activation == MRINR_ACT_MORLET);
  switch (p->precision) {
    case MRINR_PREC_FP16:
      return morlet ? launch_w0<MRINR_ACT_MORLET, MRINR_PREC_FP16>(P, grid, w0one, st)
                    : launch_w0<MRINR_ACT_SINE, MRINR_PREC_FP16>(P, grid, w0one, st);
    case MRINR_PREC_BF16:
      return morlet ? launch_w0<MRINR_ACT_MORLET, MRINR_PREC_BF16>(P, grid, w0one, st)
                    : launch_w0<MRINR_ACT_SINE, MRINR_PREC_BF16>(P, grid, w0one, st);
    case MRINR_PREC_FP16X3:
      return morlet ? launch_w0<MRINR_ACT_MORLET, MRINR_PREC_FP16X3>(P, grid, w0one, st)
                    : launch_w0<MRINR_ACT_SINE, MRINR_PREC_FP16X3>(P, grid, w0one, st);
    default:
      break;
  }
  set_error("siren_tc5: precision mode %d is not a tensor-core mode", p->precision);
  return MRINR_E_ARG;
}

}  // namespace v5

#ifndef MRINR_LAB
// The product library ships exactly one tensor-core synthesis kernel (the lab library, `make lab`, routes this call
// through lab/siren_dispatch.cu to the retired variants instead).
int launch_siren_tc(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                    int64_t B, float* d_out, cudaStream_t st) {
  return v5::launch_siren_tc_v5(p, d_mods, d_idx, d_nactive, B, d_out, st);
}
#endif
}  // namespace mrinr

#ifdef MRINR_TIMELINE
extern "C" __attribute__((visibility("default"))) int mrinr_debug_timeline(long long* host_out, int max_entries) {
  int n = 0;
  cudaMemcpyFromSymbol(&n, mrinr::v5::g_timeline_n, sizeof(int));
  if (n > max_entries) n = max_entries;
  if (n > 2048) n = 2048;
  cudaMemcpyFromSymbol(host_out, mrinr::v5::g_timeline, (size_t)n * 4 * sizeof(long long));
  int zero = 0;
  cudaMemcpyToSymbol(mrinr::v5::g_timeline_n, &zero, sizeof(int));
  return n;
}
#endif
