// Fused modulated-SIREN synthesis kernel, version 6: layer 0 on its own warps.
//
// v5 (siren_tc5.cu) took layer 0 off the special-function unit: a tile is one coordinate block of one patch, so
// h_0 = table[c] * mod_0 with a patch-independent table.  Its timeline (profiles/r01_siren.md) still shows the eight
// epilogue warps spending 2 x ~1250 of ~27 200 cycles per iteration on that multiply, and 64 of their 168 registers
// on the table.  But the layer-0 phase needs neither the special-function unit nor TMEM, and the slot's A operand
// is free exactly when the slot's output phase starts -- so it can run BESIDE the output phase on other warps:
//
//   warps 0-7   epilogue: output phase of the previous tile, hidden layers 1..L-2 (MUFU-bound)      [v5, minus layer 0]
//   warps 8-11  layer 0:  when MMA(L-1) of the slot's previous tile retires, read the tile's rows of the 16-bit table
//               from global memory (L2-resident, 64 KB per tile, 64-byte row segments), multiply by mod_0 from the
//               modulation ring, write the A operand of layer 1.  FMA-pipe and LSU work that fills issue slots the
//               MUFU-bound warps leave empty.
//   warp 12     MMA issuer (leader CTA) / forwarder (peer)            warp 13  weight + modulation producer
//
// MEASURED RESULT (profiles/r01_siren.md): correct (same parity tests as v5), but SLOWER than v5 -- 977-1015 vs 1143
// TFLOP/s.  With the layer-0 warps' work compiled out the iteration drops to ~24 000 cycles as planned, but their
// HMUL2 / STS.128 stream running beside the MUFU-bound output phase stretches that phase from ~2400 to 3900-5900
// cycles (MIO / shared-memory port contention), which costs more than the 2 x 1250 cycles it hides.  Kept as an
// opt-in variant (MRINR_TC_VARIANT=6) and as the record of the experiment; v5 is the default.
//
// The table no longer lives in registers (it would need 128 per thread on 4 warps); without it the epilogue warps fit
// the 128-register cap of a 14-warp CTA.  The schedule (coordinate-block tiles, sub-blocks of 128 patches), the
// modulation ring, CTA pairs, ping-pong slots, the bias K step and the weight ring are those of v5.
#include "tc_ptx.cuh"

namespace mrinr {
namespace v6 {

constexpr int kH = 256;
constexpr int kTileM = 128;
constexpr int kSlabBytes = 16384;    // K=64 x N=128 (this CTA's half of the output rows) x 2 B
constexpr int kBiasSlabBytes = 4096; // K=16 x N=128 x 2 B
constexpr int kNumSlabs = 5;         // 4 weight slabs + the bias step
constexpr int kLayerBytes = 4 * kSlabBytes + kBiasSlabBytes;   // per (layer, rank) in the packed array
#ifndef MRINR_V6_EPI_WARPS
#define MRINR_V6_EPI_WARPS 8
#endif
constexpr int kEpiWarps = MRINR_V6_EPI_WARPS;  // 8 or 16: 2 or 4 epilogue warps per SM sub-partition
constexpr int kColGroups = kEpiWarps / 4;      // a warp owns rows 32(w&3)..+31 and columns [kCols*(w>>2), +kCols)
constexpr int kCols = 256 / kColGroups;        // 128 or 64 columns per thread per phase
constexpr int kPairs = kCols / 32;             // pairs of 16-column chunks per phase
constexpr int kL0Warps = 4;                    // layer-0 warps (one per SM sub-partition)
constexpr int kWarpMma = kEpiWarps + kL0Warps; // MMA issuer / forwarder
constexpr int kThreads = (kEpiWarps + kL0Warps) * 32 + 64;
// Register budget: the register file is split per SM sub-partition (16 384 each) and the 10 (18) warps of the CTA
// land 4 on some sub-partition, so the cap is 128 registers per thread -- not 65536 / kThreads.
constexpr int kMaxLayers = 16;
constexpr int kTmemCols = 512;
constexpr int kMaxSub = 2;                     // patches sharing a remainder tile
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
constexpr int kBarModEmpty = 27; // [2 slots][kModStages] local, kEpiWarps arrivals (layer-0 warps arrive with count 2)
constexpr int kBarAFree = 35;    // [2] both CTAs, via multicast commit: the last layer's MMAs of the slot's tile have retired
constexpr int kBarL0Done = 37;   // [2] local, one arrival per layer-0 warp: A[slot] holds the layer-1 operand

#ifdef MRINR_TIMELINE_V6
// development aid: per-phase timestamps of one lane per selected warp (see tools/timeline.py); compiled out by
// default.  Timestamps are kept in a per-thread local array and written out once at the end (no atomics in the loop).
__device__ long long g_timeline[8192];
__device__ int g_timeline_n;
#define TL_DECL long long tl_buf[160]; int tl_tag[160]; int tl_n = 0;
#define TL(tag)                                                                          \
  do {                                                                                   \
    if (lane == 0 && blockIdx.x == 0 && (warp == 0 || warp == kEpiWarps || warp == kWarpMma) && tl_n < 160) { \
      tl_buf[tl_n] = clock64(); tl_tag[tl_n] = (tag); ++tl_n;                             \
    }                                                                                    \
  } while (0)
#define TL_FLUSH                                                                          \
  do {                                                                                   \
    if (lane == 0 && blockIdx.x == 0 && (warp == 0 || warp == kEpiWarps || warp == kWarpMma)) {               \
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

// The schedule of one cluster; every role derives it from the same inputs, so nothing is communicated.
// The cluster's patches are walked in sub-blocks of kSubBlock patches, block-major inside a sub-block: the modulation
// vectors of a sub-block (5 KB per patch, read once per coordinate block) then stay in L2 between their
// C/128 + 1 uses (74 clusters x 128 patches x 5 KB = 47 MB) instead of streaming from HBM every time.
constexpr int kSubBlock = 128;
struct Sched {
  long long pa;          // this cluster's patches: compacted indices [pa, pa + np)
  int np;
  int n_full, rem, ksub; // C = 128 n_full + rem; ksub = patches per remainder tile (0 if rem == 0)
  int n_types;           // coordinate blocks per patch: n_full (+ 1 if rem)
  long long total;       // cluster iterations (4 tiles each)
};
__device__ __forceinline__ int iters_rem(const Sched& s, int n) {   // remainder-block iterations for n patches
  return s.rem ? ((n + s.ksub - 1) / s.ksub + 3) / 4 : 0;
}
__device__ __forceinline__ Sched make_sched(long long n_act, int C, long long cluster_id, long long n_clusters) {
  Sched s;
  s.pa = n_act * cluster_id / n_clusters;
  s.np = (int)(n_act * (cluster_id + 1) / n_clusters - s.pa);
  s.n_full = C / kTileM;
  s.rem = C - s.n_full * kTileM;
  s.ksub = s.rem ? (kTileM / s.rem < kMaxSub ? kTileM / s.rem : kMaxSub) : 0;
  s.n_types = s.n_full + (s.rem ? 1 : 0);
  const int blocks = s.np / kSubBlock, tail = s.np - blocks * kSubBlock;
  s.total = (long long)blocks * (s.n_full * (kSubBlock / 4) + iters_rem(s, kSubBlock));
  if (tail) s.total += s.n_full * ((tail + 3) / 4) + iters_rem(s, tail);
  return s;
}
struct Walk {          // (sub-block, coordinate block, iteration within the block), advanced without divisions
  int type = 0;
  int j = 0;
  int base = 0;        // first patch of the sub-block, relative to Sched::pa
  int nps = 0;         // patches in the sub-block
  int itf = 0, itr = 0;
  __device__ __forceinline__ void set_block(const Sched& s) {
    nps = s.np - base < kSubBlock ? s.np - base : kSubBlock;
    itf = (nps + 3) / 4;
    itr = iters_rem(s, nps);
  }
  __device__ __forceinline__ void next(const Sched& s) {
    if (++j == (type < s.n_full ? itf : itr)) {
      j = 0;
      if (++type == s.n_types) { type = 0; base += kSubBlock; set_block(s); }
    }
  }
};

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

// product of two packed 16-bit pairs in the operand format
template <bool BF16>
__device__ __forceinline__ uint32_t mul2_16(uint32_t a, uint32_t b) {
  uint32_t r;
  if (BF16) asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  else      asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

template <int ACT, bool BF16, bool W0ONE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) siren_tc6_kernel(const SirenTcParams P) {
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
  const Sched S = make_sched(n_act, C, blockIdx.x >> 1, gridDim.x >> 1);
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
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(kBarAFree + s), 1);
      mbar_init(bar(kBarL0Done + s), kL0Warps);
    }
    for (int s = 0; s < 2 * kModStages; ++s) {
      mbar_init(bar(kBarModFull + s), 1);
      mbar_init(bar(kBarModEmpty + s), kEpiWarps);
    }
    fence_barrier_init();
  }
  fence_proxy_async();         // the ones block is read by the tensor core (async proxy)
  if (warp == kWarpMma) tmem_alloc_pair(smem_u32(s_tmem), kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  TL_DECL

  if (warp < kEpiWarps) {
    // =========================== epilogue warps (both CTAs, identical) ===========================
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
      fence_proxy_async();
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
    // 16 activations x = pre-activation (bias included) -> h = act(x) * mod, packed to 8 x (2 x 16 bit).
    // Sine uses the order-pinned pipeline (sines of group g, then multiply + pack of group g-1).
    auto act16_pack = [&](const uint32_t (&v)[16], const float4 (&m)[4], uint32_t (&pk)[8]) {
      if (ACT == MRINR_ACT_SINE) {
        float s[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float x = __uint_as_float(v[g * 4 + i]);
            s[g * 4 + i] = vsin(W0ONE ? x : P.w0 * x);
          }
          if (g > 0) {
            const float4 mm = m[g - 1];
            pk[(g - 1) * 2 + 0] = vpack2<BF16>(vmul(s[g * 4 - 4], mm.x), vmul(s[g * 4 - 3], mm.y));
            pk[(g - 1) * 2 + 1] = vpack2<BF16>(vmul(s[g * 4 - 2], mm.z), vmul(s[g * 4 - 1], mm.w));
          }
        }
        pk[6] = vpack2<BF16>(vmul(s[12], m[3].x), vmul(s[13], m[3].y));
        pk[7] = vpack2<BF16>(vmul(s[14], m[3].z), vmul(s[15], m[3].w));
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          pk[g * 2 + 0] = pack2<BF16>(act_fast<ACT, W0ONE>(__uint_as_float(v[g * 4 + 0]), P.w0) * m[g].x,
                                      act_fast<ACT, W0ONE>(__uint_as_float(v[g * 4 + 1]), P.w0) * m[g].y);
          pk[g * 2 + 1] = pack2<BF16>(act_fast<ACT, W0ONE>(__uint_as_float(v[g * 4 + 2]), P.w0) * m[g].z,
                                      act_fast<ACT, W0ONE>(__uint_as_float(v[g * 4 + 3]), P.w0) * m[g].w);
        }
      }
    };
    auto store16 = [&](int slot, int hc, const uint32_t (&pk)[8]) {   // columns cg*kCols + hc*16 .. +15 of row t
      uint8_t* base = smem + kOffA + slot * 65536 + (cg * (kCols / 8) + hc * 2) * 2048 + t * 16;
      *reinterpret_cast<uint4*>(base) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(base + 2048) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
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
      named_bar_sync(1, kEpiWarps * 32);
      const float* mp = ring + sub_prev * kH + cg * kCols;
      mbar_wait(bar(kBarAccFull + slot), ev & 1u, P.errflag, 5);
      TL(1010 + slot);
      tc_fence_after();
      uint32_t va[16], vb[16];
      tmem_ld16(tcol, va);
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
      auto dot16 = [&](const uint32_t (&v)[16], int hc) {
        float4 mw[4];
        load4x4(mp, hc, mw);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float x0 = __uint_as_float(v[g * 4 + 0]), x1 = __uint_as_float(v[g * 4 + 1]);
          const float x2 = __uint_as_float(v[g * 4 + 2]), x3 = __uint_as_float(v[g * 4 + 3]);
          float h0, h1, h2, h3;
          if (ACT == MRINR_ACT_SINE) {
            h0 = vsin(W0ONE ? x0 : P.w0 * x0); h1 = vsin(W0ONE ? x1 : P.w0 * x1);
            h2 = vsin(W0ONE ? x2 : P.w0 * x2); h3 = vsin(W0ONE ? x3 : P.w0 * x3);
          } else {
            h0 = act_fast<ACT, W0ONE>(x0, P.w0); h1 = act_fast<ACT, W0ONE>(x1, P.w0);
            h2 = act_fast<ACT, W0ONE>(x2, P.w0); h3 = act_fast<ACT, W0ONE>(x3, P.w0);
          }
          d0 = fmaf(h0, mw[g].x, d0);
          d1 = fmaf(h1, mw[g].y, d1);
          d2 = fmaf(h2, mw[g].z, d2);
          d3 = fmaf(h3, mw[g].w, d3);
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
      tc_fence_before();
      TL(1015 + slot);
      finish_phase(slot, use, false);
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
      }
      // ---- per slot: finish the previous tile of the slot (its layer-0 successor is written by the layer-0 warps) ----
#pragma unroll 1
      for (int slot = 0; slot < 2; ++slot) {
        if (it > 0) final_phase(slot, slot ? out1 : out0, ev0 - 1u, use0 - 1u);
        if (last) continue;
        {
          // which patch does this row belong to?  (a phantom tile / padding row has no output)
          const int ti = w.j * 4 + slot * 2 + (int)rank;
          const int pl = cur_type < S.n_full ? ti : ti * S.ksub + sub_row;
          float* o = nullptr;
          if (row_live && pl < w.nps) {
            const long long patch = P.idx ? (long long)P.idx[S.pa + w.base + pl] : S.pa + w.base + pl;
            o = P.out + patch * C + c_row;
          }
          if (slot) out1 = o; else out0 = o;
        }
        // Layer 1 of the new tile may start when (a) the layer-0 warps have written A[slot] and (b) this warp no
        // longer reads the slot's accumulator (the output phase above): MMA(1) overwrites it.  So the epilogue
        // warps, not the layer-0 warps, give the MMA issuer its a_full arrivals for layer 1.
        TL(1020 + slot);
        mbar_wait(bar(kBarL0Done + slot), (uint32_t)it & 1u, P.errflag, 13);
        TL(1030 + slot);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(bar(kBarAFull + slot), 0);
      }
      if (last) break;

      // ---- hidden layers 1 .. L-2, alternating slots: the other slot's MMAs run underneath ----
      for (int l = 1; l <= L - 2; ++l) {
        const uint32_t ev = ev0 + (uint32_t)(l - 1);
        const uint32_t use = use0 + (uint32_t)l;
#pragma unroll 1
        for (int slot = 0; slot < 2; ++slot) {
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
            uint32_t pk[8];
            tmem_ld_wait();
            tmem_ld16(tcol + (uint32_t)(hp * 2 + 1) * 16u, vb);     // next chunk lands while this one is processed
            load4x4(mp, hp * 2, m);
            act16_pack(va, m, pk);
            store16(slot, hp * 2, pk);
            tmem_ld_wait();
            if (hp < kPairs - 1) tmem_ld16(tcol + (uint32_t)(hp * 2 + 2) * 16u, va);
            load4x4(mp, hp * 2 + 1, m);
            act16_pack(vb, m, pk);
            store16(slot, hp * 2 + 1, pk);
          }
          tc_fence_before();
          finish_phase(slot, use, true);
          TL(5000 + l * 10 + slot);
        }
      }
      w.next(S);
    }
  } else if (warp < kWarpMma) {
    // =========================== layer-0 warps (both CTAs) ===========================
    // A warp instruction covers 8 rows x 4 K-chunks (64-byte row segments in global memory, 4 x 128 contiguous bytes
    // in the operand layout).  Warp lw owns K chunks 8 lw .. 8 lw + 7 of all 128 rows = 32 sixteen-byte pieces per
    // lane, fetched in 4 batches of 8 (two batches in flight; the first two are requested before any barrier wait:
    // the table does not depend on them).  For a full coordinate block every address is base + compile-time offset.
    const int lw = warp - kEpiWarps;
    const int r8 = lane & 7, kq = lane >> 3;
    const uint4* tab = reinterpret_cast<const uint4*>(P.table16);       // [C][32] 16-byte chunks
    Walk w;
    w.set_block(S);
    for (long long it = 0; it < S.total; ++it, w.next(S)) {
      const uint32_t use0 = (uint32_t)it * (uint32_t)L;
      const bool full = w.type < S.n_full;
#pragma unroll 1
      for (int slot = 0; slot < 2; ++slot) {
        const float* ring = s_mods + (((use0 & (kModStages - 1)) * 2 + slot) * kMaxSub) * kH;
        uint8_t* a_base = smem + kOffA + slot * 65536 + r8 * 16;
        auto wait_inputs = [&]() {
          if (it > 0) mbar_wait(bar(kBarAFree + slot), (uint32_t)(it - 1) & 1u, P.errflag, 11);   // A[slot] is free
          mbar_wait(bar(kBarModFull + slot * kModStages + (int)(use0 & (kModStages - 1))), (use0 / kModStages) & 1u,
                    P.errflag, 12);
        };
        if (full) {
          // batch b (0..3): K chunk kc = 8 lw + 4 (b >> 1) + kq, rows 8 rg + r8 with rg = 8 (b & 1) .. + 7.
          // The multiply runs on packed 16-bit pairs (HMUL2 / HMUL2.BF16): 6 instructions per 16-byte piece instead
          // of 22 with fp32 products; the extra rounding (table, modulation and product each rounded to the operand
          // format) stays inside the 1e-3 budget (tests: 4.4e-4 -> measured below).
          const uint4* t0 = tab + (size_t)(w.type * kTileM + r8) * (kH / 8) + lw * 8 + kq;
          uint4 ta[8], tb[8], tc[8];
          auto load8 = [&](int b, uint4 (&tv)[8]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) tv[i] = __ldg(t0 + ((b & 1) * 8 + i) * 8 * (kH / 8) + (b >> 1) * 4);
          };
          auto proc8 = [&](int b, const uint4 (&tv)[8]) {
            const int kc = lw * 8 + (b >> 1) * 4 + kq;
            const float4 m0 = *reinterpret_cast<const float4*>(ring + kc * 8);
            const float4 m1 = *reinterpret_cast<const float4*>(ring + kc * 8 + 4);
            const uint32_t p0 = pack2<BF16>(m0.x, m0.y), p1 = pack2<BF16>(m0.z, m0.w);
            const uint32_t p2 = pack2<BF16>(m1.x, m1.y), p3 = pack2<BF16>(m1.z, m1.w);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<uint4*>(a_base + kc * 2048 + ((b & 1) * 8 + i) * 128) =
                  make_uint4(mul2_16<BF16>(tv[i].x, p0), mul2_16<BF16>(tv[i].y, p1), mul2_16<BF16>(tv[i].z, p2),
                             mul2_16<BF16>(tv[i].w, p3));
          };
          TL(9000 + slot);
          load8(0, ta);
          load8(1, tb);
          load8(2, tc);
          wait_inputs();
          TL(9100 + slot);
#ifndef MRINR_V6_NOL0WORK
          proc8(0, ta);
          load8(3, ta);
          proc8(1, tb);
          proc8(2, tc);
          proc8(3, ta);
#endif
          TL(9200 + slot);
        } else {
          // remainder block: rows of up to kMaxSub patches, padding rows clamped to coordinate 0 of the block
          wait_inputs();
#pragma unroll 1
          for (int p = 0; p < 32; ++p) {
            const int rg = p & 15, kc = lw * 8 + (p >> 4) * 4 + kq;
            const int row = rg * 8 + r8;
            const int sb = row / S.rem;
            const bool live = sb < S.ksub;
            const int sub = live ? sb : 0;
            const int c = S.n_full * kTileM + (live ? row - sb * S.rem : 0);
            const uint4 tv = __ldg(tab + (size_t)c * (kH / 8) + kc);
            const float4 m0 = *reinterpret_cast<const float4*>(ring + sub * kH + kc * 8);
            const float4 m1 = *reinterpret_cast<const float4*>(ring + sub * kH + kc * 8 + 4);
            *reinterpret_cast<uint4*>(a_base + kc * 2048 + rg * 128) =
                make_uint4(mul2_16<BF16>(tv.x, pack2<BF16>(m0.x, m0.y)), mul2_16<BF16>(tv.y, pack2<BF16>(m0.z, m0.w)),
                           mul2_16<BF16>(tv.z, pack2<BF16>(m1.x, m1.y)), mul2_16<BF16>(tv.w, pack2<BF16>(m1.z, m1.w)));
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_n(bar(kBarModEmpty + slot * kModStages + (int)(use0 & (kModStages - 1))), 2);   // stands for 2 warps
          mbar_arrive(bar(kBarL0Done + slot));          // the epilogue warps forward it to the MMA issuer (see there)
        }
      }
    }
  } else if (warp == kWarpMma) {
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
#pragma unroll
          for (int slot = 0; slot < 2; ++slot) {
            const uint32_t a_lo = a_lo0 + (uint32_t)slot * (65536u >> 4);
            const uint32_t d_tmem = tmem_base + (uint32_t)slot * 256u;
            TL(6000 + l * 10 + slot);
            mbar_wait_backoff(bar(kBarAFull + slot), ev & 1u, P.errflag, 1, 32);
            TL(7000 + l * 10 + slot);
#pragma unroll
            for (int s = 0; s < kNumSlabs; ++s) {
              if (slot == 0) {
                mbar_wait_backoff(bar(kBarWFull + s), ev & 1u, P.errflag, 2, 32);
                mbar_wait_backoff(bar(kBarWPeer + s), ev & 1u, P.errflag, 8, 32);
              }
              tc_fence_after();
              if (elect_one()) {
                const uint32_t b_lo = b_lo0 + (uint32_t)s * (kSlabBytes >> 4);
                if (s < 4) {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    umma_f16_pair_lohi(d_tmem, a_lo + (uint32_t)(s * 4 + kk) * 256u, b_lo + (uint32_t)kk * 256u, desc_hi,
                                       idesc, (s | kk) != 0 ? 1u : 0u);
                } else {
                  umma_f16_pair_lohi(d_tmem, ones_lo, b_lo, desc_hi, idesc, 1u);      // + bias
                }
                if (slot == 1 && s == kNumSlabs - 1) {
                  // both slots have consumed this layer's slabs: hand all of them back with one commit each
#pragma unroll
                  for (int r = 0; r < kNumSlabs; ++r) umma_commit_pair(bar(kBarWEmpty + r), 3);
                }
              }
              __syncwarp();
            }
            if (elect_one()) {
              umma_commit_pair(bar(kBarAccFull + slot), 3);
              if (l == L - 1) umma_commit_pair(bar(kBarAFree + slot), 3);    // A[slot] may be rewritten (layer-0 warps)
            }
            __syncwarp();
            TL(8000 + l * 10 + slot);
          }
        }
      }
    } else if (lane == 0) {
      // =========================== forwarder (peer CTA): my slab has landed ===========================
      uint32_t ev = 0;
      for (long long it = 0; it < S.total; ++it) {
        for (int l = 1; l < L; ++l, ++ev) {
#pragma unroll 1
          for (int s = 0; s < kNumSlabs; ++s) {
            mbar_wait_backoff(bar(kBarWFull + s), ev & 1u, P.errflag, 9);
            mbar_arrive_cluster(bar(kBarWPeer + s), 0);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ============ producer (both CTAs): own half of every weight slab + the modulation vectors of my tiles ============
    // Order per iteration: mods(0), then for every layer l >= 1: mods(l), weights(l).  The modulations of a phase are
    // always requested before the producer can block on a weight slab that (transitively) waits for that phase.
    if (lane == 0) {
      uint32_t ev = 0;
      Walk w;
      w.set_block(S);
      for (long long it = 0; it < S.total; ++it, w.next(S)) {
        const bool full = w.type < S.n_full;
        const int nsub = full ? 1 : S.ksub;
        long long patch[2][kMaxSub];
#pragma unroll
        for (int slot = 0; slot < 2; ++slot) {
          const int ti = w.j * 4 + slot * 2 + (int)rank;
#pragma unroll
          for (int s = 0; s < kMaxSub; ++s) {
            int pl = full ? ti : ti * S.ksub + s;
            if (pl >= w.nps) pl = w.nps - 1;                     // phantom tile / missing patch: any valid vector
            patch[slot][s] = P.idx ? (long long)P.idx[S.pa + w.base + pl] : S.pa + w.base + pl;
          }
        }
        auto issue_mods = [&](int l) {
          const uint32_t use = (uint32_t)it * (uint32_t)L + (uint32_t)l;
          const uint32_t stage = use & (kModStages - 1);
#pragma unroll
          for (int slot = 0; slot < 2; ++slot) {
            const uint32_t full_bar = bar(kBarModFull + slot * kModStages + stage);
            mbar_wait_backoff(bar(kBarModEmpty + slot * kModStages + stage), ((use / kModStages) & 1u) ^ 1u, P.errflag, 10);
            mbar_expect_tx(full_bar, (uint32_t)nsub * kH * 4u);
            for (int s = 0; s < nsub; ++s)
              bulk_g2s(sMods + (uint32_t)(((stage * 2 + slot) * kMaxSub + s) * kH * 4),
                       P.mods + (size_t)l * layer_stride + (size_t)patch[slot][s] * kH, kH * 4, full_bar);
          }
        };
        issue_mods(0);
        for (int l = 1; l < L; ++l, ++ev) {
          issue_mods(l);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(P.w16q) + ((size_t)(l - 1) * 2 + rank) * kLayerBytes;
#pragma unroll 1
          for (int s = 0; s < kNumSlabs; ++s) {
            const uint32_t bytes = s < 4 ? kSlabBytes : kBiasSlabBytes;
            mbar_wait_backoff(bar(kBarWEmpty + s), (ev & 1u) ^ 1u, P.errflag, 7);
            mbar_expect_tx(bar(kBarWFull + s), bytes);
            bulk_g2s(sW + s * kSlabBytes, src + (size_t)s * kSlabBytes, bytes, bar(kBarWFull + s));
          }
        }
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
  if (warp == kWarpMma) tmem_dealloc_pair(tmem_base, kTmemCols);
}

template <int ACT, bool BF16, bool W0ONE>
static int launch_one(const SirenTcParams& P, int grid, cudaStream_t st) {
  MRINR_SMEM_OPT_IN((siren_tc6_kernel<ACT, BF16, W0ONE>), kSmemBytes);
  siren_tc6_kernel<ACT, BF16, W0ONE><<<grid, kThreads, kSmemBytes, st>>>(P);
  count_launch();
  return check_launch("siren_tc6");
}

int launch_siren_tc_v6(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                       int64_t B, float* d_out, cudaStream_t st) {
  MRINR_REQUIRE(p->H == kH && p->L <= kMaxLayers && p->L >= 3 && p->C >= kTileM, MRINR_E_UNSUPPORTED,
                "siren_tc6: unsupported configuration (H=%d L=%d C=%d)", p->H, p->L, p->C);
  SirenTcParams P;
  P.table0 = p->d_table0; P.table16 = p->d_table16; P.w16 = p->d_net_w16; P.w16p = p->d_net_w16p; P.w16q = p->d_net_w16q;
  P.layer0 = p->d_layer0; P.grid = p->d_grid; P.w0_initial = p->w0_initial;
  P.bias = p->d_net_bias; P.last_w = p->d_last_w;
  P.last_b = p->d_last_b; P.mods = d_mods; P.idx = d_idx; P.nactive = d_nactive; P.out = d_out;
  P.errflag = p->d_errflag; P.B = B; P.C = p->C; P.L = p->L; P.w0 = p->w0;
  // one cluster per SM pair, but never more clusters than there are groups of 4 patches (a cluster iteration
  // processes 4 tiles of the same coordinate block)
  long long clusters = p->num_sms / 2;
  const long long groups = (B + 3) / 4;
  if (clusters > groups) clusters = groups;
  if (clusters < 1) clusters = 1;
  const int grid = (int)(clusters * 2);
  const bool w0one = (p->w0 == 1.0f);
  const bool bf16 = (p->precision == MRINR_PREC_BF16);
  const bool morlet = (p->activation == MRINR_ACT_MORLET);
#define MRINR_TC_CASE(A, Bf, W) return launch_one<A, Bf, W>(P, grid, st)
  if (!morlet) {
    if (!bf16) { if (w0one) MRINR_TC_CASE(MRINR_ACT_SINE, false, true); else MRINR_TC_CASE(MRINR_ACT_SINE, false, false); }
    else       { if (w0one) MRINR_TC_CASE(MRINR_ACT_SINE, true, true);  else MRINR_TC_CASE(MRINR_ACT_SINE, true, false); }
  } else {
    if (!bf16) { if (w0one) MRINR_TC_CASE(MRINR_ACT_MORLET, false, true); else MRINR_TC_CASE(MRINR_ACT_MORLET, false, false); }
    else       { if (w0one) MRINR_TC_CASE(MRINR_ACT_MORLET, true, true);  else MRINR_TC_CASE(MRINR_ACT_MORLET, true, false); }
  }
#undef MRINR_TC_CASE
}

}  // namespace v6
}  // namespace mrinr

#ifdef MRINR_TIMELINE_V6
extern "C" __attribute__((visibility("default"))) int mrinr_debug_timeline(long long* host_out, int max_entries) {
  int n = 0;
  cudaMemcpyFromSymbol(&n, mrinr::v6::g_timeline_n, sizeof(int));
  if (n > max_entries) n = max_entries;
  if (n > 2048) n = 2048;
  cudaMemcpyFromSymbol(host_out, mrinr::v6::g_timeline, (size_t)n * 4 * sizeof(long long));
  int zero = 0;
  cudaMemcpyToSymbol(mrinr::v6::g_timeline_n, &zero, sizeof(int));
  return n;
}
#endif
