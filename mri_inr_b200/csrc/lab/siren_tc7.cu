// Fused modulated-SIREN synthesis kernel, version 7: one epilogue warpgroup per tile slot.
//
// v5's timeline (profiles/r01_siren.md): the eight epilogue warps walk ten phases per cluster iteration in lockstep,
// and every phase change costs ~300-600 cycles of latency (barrier round trips, fence.proxy.async, refilling the
// TMEM-load / MUFU pipeline) during which the SM sub-partitions issue nothing: ~4400 of ~27 200 cycles.  Running
// other work BESIDE a MUFU-bound phase does not help (v6: the issue port is the shared resource), but running it
// INSTEAD of idling does.  So the warps are split by tile slot:
//
//   warpgroup 0 (warps 0-3)   all phases of slot 0: output phase of the previous tile, layer 0, hidden layers
//   warpgroup 1 (warps 4-7)   the same for slot 1
//   warpgroup 2 (warps 8-11)  warp 8: MMA issuer / forwarder, warp 9: weight + modulation producer, 10-11: idle
//
// A slot's chain is strictly serial (phase -> MMA of the next layer -> phase ...), so while warpgroup 0 waits for its
// MMA, warpgroup 1 owns the sub-partitions' issue ports, and vice versa: the two chains interleave by themselves and
// phase-change latency of one hides behind the arithmetic of the other.  A thread now owns one row and all 256
// columns, so the layer-0 table needs 128 registers; warpgroup 2 gives its registers away with setmaxnreg (40 each),
// the epilogue warpgroups grow to 232 (3 warps per sub-partition: 232 + 232 + 40 <= 512).  No partial-dot exchange
// between column groups is needed any more.  Weight slabs are handed back per slab (not per layer) because slot 1's
// MMAs now trail slot 0's by half a period.  Everything else (coordinate-block tiles, sub-blocks, modulation ring,
// CTA pairs, bias K step) is v5.
#include "tc_ptx.cuh"

namespace mrinr {
namespace v7 {

constexpr int kH = 256;
constexpr int kTileM = 128;
constexpr int kSlabBytes = 16384;    // K=64 x N=128 (this CTA's half of the output rows) x 2 B
constexpr int kBiasSlabBytes = 4096; // K=16 x N=128 x 2 B
constexpr int kNumSlabs = 5;         // 4 weight slabs + the bias step
constexpr int kLayerBytes = 4 * kSlabBytes + kBiasSlabBytes;   // per (layer, rank) in the packed array
constexpr int kEpiWarps = 8;                   // two warpgroups: warps 0-3 own slot 0, warps 4-7 own slot 1
constexpr int kGroupThreads = 128;             // threads of one epilogue warpgroup
constexpr int kCols = 256;                     // a thread owns one row (TMEM lane) and all columns
constexpr int kPairs = kCols / 32;             // pairs of 16-column chunks per phase
constexpr int kThreads = 384;                  // 3 warpgroups
constexpr int kWarpMma = 8, kWarpProducer = 9;
constexpr int kRegsEpilogue = 232, kRegsOther = 40;   // setmaxnreg targets
constexpr int kMaxLayers = 16;
constexpr int kTmemCols = 512;
constexpr int kMaxSub = 2;                     // patches sharing a remainder tile
constexpr int kModStages = 4;                  // modulation ring depth per slot (power of two)

constexpr int kOffA = 0;                                          // [2 slots][64 KB]
constexpr int kOffOnes = 2 * 65536;                               // [2 kc][128][8] constant "ones" K step
constexpr int kOffW = kOffOnes + 4096;                            // 4 x 16 KB + 4 KB
constexpr int kOffMods = kOffW + kLayerBytes;                     // [kModStages][2 slots][kMaxSub][256] f32
constexpr int kOffLastW = kOffMods + kModStages * 2 * kMaxSub * kH * 4;   // [256] f32
constexpr int kOffBar = kOffLastW + kH * 4;
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

#ifdef MRINR_TIMELINE_V7
// development aid: per-phase timestamps of one lane per selected warp (see tools/timeline.py); compiled out by
// default.  Timestamps are kept in a per-thread local array and written out once at the end (no atomics in the loop).
__device__ long long g_timeline[8192];
__device__ int g_timeline_n;
#define TL_DECL long long tl_buf[160]; int tl_tag[160]; int tl_n = 0;
#define TL(tag)                                                                          \
  do {                                                                                   \
    if (lane == 0 && blockIdx.x == 0 && (warp == 0 || warp == 4 || warp == kWarpMma) && tl_n < 160) { \
      tl_buf[tl_n] = clock64(); tl_tag[tl_n] = (tag); ++tl_n;                             \
    }                                                                                    \
  } while (0)
#define TL_FLUSH                                                                          \
  do {                                                                                   \
    if (lane == 0 && blockIdx.x == 0 && (warp == 0 || warp == 4 || warp == kWarpMma)) {               \
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

template <int ACT, bool BF16, bool W0ONE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) siren_tc7_kernel(const SirenTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int C = P.C, L = P.L;
  const uint32_t rank = cluster_ctarank();

  float* s_mods = reinterpret_cast<float*>(smem + kOffMods);
  float* s_lastw = reinterpret_cast<float*>(smem + kOffLastW);
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
      mbar_init(bar(kBarAFull + s), 2 * 4);        // the slot's warpgroup in both CTAs
      mbar_init(bar(kBarAccFull + s), 1);
    }
    for (int s = 0; s < 2 * kModStages; ++s) {
      mbar_init(bar(kBarModFull + s), 1);
      mbar_init(bar(kBarModEmpty + s), 4);
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
    // =========================== epilogue warpgroups (both CTAs): warpgroup g owns tile slot g ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpilogue));
    const int slot = warp >> 2;
    const int q = warp & 3;
    const int t = q * 32 + lane;       // tile row == TMEM lane
    const int gt = tid & (kGroupThreads - 1);
    const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t tcol = taddr_row + (uint32_t)slot * 256u;
    const float last_b = P.last_b ? *P.last_b : 0.f;
    uint8_t* a_row = smem + kOffA + slot * 65536 + t * 16;         // this row's 16-byte chunk of K chunk 0

    // per coordinate block (changes C/128 + 1 times per sub-block of patches)
    int cur_type = -1;
    int c_row = 0;                     // this row's coordinate within the patch
    int sub_row = 0;                   // which of the tile's patches this row belongs to (remainder block only)
    int sub_prev = 0;                  // sub_row of the previous tile (its output phase runs late)
    bool row_live = false;             // false: padding row of a remainder tile
    uint32_t T2[kCols / 2];            // layer-0 table of (c_row, all columns), packed pairs: 128 registers
    float* outp = nullptr;             // output element of this row for the tile in flight

    auto mod_stage = [&](uint32_t use) -> float* {
      return s_mods + (((use & (kModStages - 1)) * 2 + slot) * kMaxSub) * kH;
    };
    auto mod_full_bar = [&](uint32_t use) -> uint32_t {
      return bar(kBarModFull + slot * kModStages + (int)(use & (kModStages - 1)));
    };
    auto peek_mods = [&](uint32_t use) -> uint32_t { return mbar_try_wait(mod_full_bar(use), (use / kModStages) & 1u); };
    auto wait_mods = [&](uint32_t use) { mbar_wait(mod_full_bar(use), (use / kModStages) & 1u, P.errflag, 6); };
    // this warp is done with the phase: its part of A[slot] is written (publish_a) and its modulation reads are over
    auto finish_phase = [&](uint32_t use, bool publish_a) {
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
    // 16 activations x = pre-activation (bias included) -> h = act(x) * mod, packed to 8 x (2 x 16 bit).
    // Sine uses the order-pinned pipeline (sines of group g, then multiply + pack of group g-1).
    auto act16_pack = [&](const uint32_t (&v)[16], const float4 (&m)[4], uint32_t (&pk)[8]) {
      if (ACT == MRINR_ACT_SINE) {
        float sv[16];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float x = __uint_as_float(v[g * 4 + i]);
            sv[g * 4 + i] = vsin(W0ONE ? x : P.w0 * x);
          }
          if (g > 0) {
            const float4 mm = m[g - 1];
            pk[(g - 1) * 2 + 0] = vpack2<BF16>(vmul(sv[g * 4 - 4], mm.x), vmul(sv[g * 4 - 3], mm.y));
            pk[(g - 1) * 2 + 1] = vpack2<BF16>(vmul(sv[g * 4 - 2], mm.z), vmul(sv[g * 4 - 1], mm.w));
          }
        }
        pk[6] = vpack2<BF16>(vmul(sv[12], m[3].x), vmul(sv[13], m[3].y));
        pk[7] = vpack2<BF16>(vmul(sv[14], m[3].z), vmul(sv[15], m[3].w));
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
    auto store16 = [&](int hc, const uint32_t (&pk)[8]) {       // columns hc*16 .. +15 of row t
      uint8_t* base = a_row + hc * 2 * 2048;
      *reinterpret_cast<uint4*>(base) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(base + 2048) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    };

    // output phase of a finished tile: y = sin(w0 (h_{L-1} . w_last + b_last)), h_{L-1} = act(D) * mod_{L-1}.
    // All rows of a tile share the modulation vector, so the warpgroup first turns the ring stage into
    // mod_{L-1} * w_last in place (four elements per thread); a row then needs one FFMA per activation.
    auto final_phase = [&](uint32_t ev, uint32_t use) {
      TL(1000 + slot);
      wait_mods(use);
      float* ring = mod_stage(use);
#pragma unroll
      for (int j = gt; j < kMaxSub * kH; j += kGroupThreads) ring[j] *= s_lastw[j & (kH - 1)];
      // these generic-proxy writes must be ordered before the producer's next cp.async.bulk into this stage
      fence_proxy_async();
      named_bar_sync(1 + slot, kGroupThreads);
      const float* mp = ring + sub_prev * kH;
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
      finish_phase(use, false);
      // output layer: always sine, never modulated (modulated_siren.py:211-213, :233)
      if (outp != nullptr) *outp = sin_accurate(P.w0 * ((d0 + d1) + (d2 + d3) + last_b));
    };

    Walk w;
    w.set_block(S);
    for (long long it = 0; it <= S.total; ++it) {
      const bool last = (it == S.total);                           // extra pass: only the pending output phase
      const uint32_t use0 = (uint32_t)it * (uint32_t)L;           // modulation-ring sequence number of layer 0
      const uint32_t ev0 = (uint32_t)it * (uint32_t)(L - 1);      // accumulator event of this iteration's layer 1
      sub_prev = sub_row;
      if (!last && w.type != cur_type) {
        // ---- new coordinate block: this row's coordinate and its row of the layer-0 table ----
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
        const float4* trow = reinterpret_cast<const float4*>(P.table0 + (size_t)c_row * kH);
#pragma unroll
        for (int i = 0; i < kCols / 4; ++i) {
          const float4 v = __ldg(trow + i);
          T2[i * 2 + 0] = pack2<BF16>(v.x, v.y);
          T2[i * 2 + 1] = pack2<BF16>(v.z, v.w);
        }
      }
      const uint32_t ok0 = last ? 1u : peek_mods(use0);
      if (it > 0) final_phase(ev0 - 1u, use0 - 1u);
      if (last) break;
      {
        // which patch does this row belong to?  (a phantom tile / padding row has no output)
        const int ti = w.j * 4 + slot * 2 + (int)rank;
        const int pl = cur_type < S.n_full ? ti : ti * S.ksub + sub_row;
        outp = nullptr;
        if (row_live && pl < w.nps) {
          const long long patch = P.idx ? (long long)P.idx[S.pa + w.base + pl] : S.pa + w.base + pl;
          outp = P.out + patch * C + c_row;
        }
      }
      // ---- layer 0 (modulated_siren.py:154-156 with dim_in = 2): h = table[c] * mod_0 ----
      {
        TL(2000 + slot);
        if (!ok0) wait_mods(use0);
        const float* mp = mod_stage(use0) + sub_row * kH;
        TL(2005 + slot);
#pragma unroll
        for (int hc = 0; hc < kCols / 16; ++hc) {
          float4 m[4];
          load4x4(mp, hc, m);
          uint32_t pk[8];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float a0, a1, a2, a3;
            unpack2<BF16>(T2[hc * 8 + g * 2 + 0], a0, a1);
            unpack2<BF16>(T2[hc * 8 + g * 2 + 1], a2, a3);
            pk[g * 2 + 0] = pack2<BF16>(a0 * m[g].x, a1 * m[g].y);
            pk[g * 2 + 1] = pack2<BF16>(a2 * m[g].z, a3 * m[g].w);
          }
          store16(hc, pk);
        }
        finish_phase(use0, true);
        TL(2010 + slot);
      }
      // ---- hidden layers 1 .. L-2: the other warpgroup's phases fill this one's waits ----
      for (int l = 1; l <= L - 2; ++l) {
        const uint32_t ev = ev0 + (uint32_t)(l - 1);
        const uint32_t use = use0 + (uint32_t)l;
        TL(3000 + l * 10 + slot);
        const uint32_t okm = peek_mods(use);
        mbar_wait(bar(kBarAccFull + slot), ev & 1u, P.errflag, 4);
        if (!okm) wait_mods(use);
        const float* mp = mod_stage(use) + sub_row * kH;
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
          store16(hp * 2, pk);
          tmem_ld_wait();
          if (hp < kPairs - 1) tmem_ld16(tcol + (uint32_t)(hp * 2 + 2) * 16u, va);
          load4x4(mp, hp * 2 + 1, m);
          act16_pack(vb, m, pk);
          store16(hp * 2 + 1, pk);
        }
        tc_fence_before();
        finish_phase(use, true);
        TL(5000 + l * 10 + slot);
      }
      w.next(S);
    }
  } else {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
   if (warp == kWarpMma) {
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
                // slot 1 trails slot 0 by half a period: hand every slab back as soon as slot 1 has used it, so that
                // the next layer's weights are in place when slot 0 comes round again
                if (slot == 1) umma_commit_pair(bar(kBarWEmpty + s), 3);
              }
              __syncwarp();
            }
            if (elect_one()) umma_commit_pair(bar(kBarAccFull + slot), 3);
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
   } else if (warp == kWarpProducer) {
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
  MRINR_SMEM_OPT_IN((siren_tc7_kernel<ACT, BF16, W0ONE>), kSmemBytes);
  siren_tc7_kernel<ACT, BF16, W0ONE><<<grid, kThreads, kSmemBytes, st>>>(P);
  count_launch();
  return check_launch("siren_tc7");
}

int launch_siren_tc_v7(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                       int64_t B, float* d_out, cudaStream_t st) {
  MRINR_REQUIRE(p->H == kH && p->L <= kMaxLayers && p->L >= 3 && p->C >= kTileM, MRINR_E_UNSUPPORTED,
                "siren_tc7: unsupported configuration (H=%d L=%d C=%d)", p->H, p->L, p->C);
  SirenTcParams P;
  P.table0 = p->d_table0; P.w16 = p->d_net_w16; P.w16p = p->d_net_w16p; P.w16q = p->d_net_w16q;
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

}  // namespace v7
}  // namespace mrinr

#ifdef MRINR_TIMELINE_V7
extern "C" __attribute__((visibility("default"))) int mrinr_debug_timeline(long long* host_out, int max_entries) {
  int n = 0;
  cudaMemcpyFromSymbol(&n, mrinr::v7::g_timeline_n, sizeof(int));
  if (n > max_entries) n = max_entries;
  if (n > 2048) n = 2048;
  cudaMemcpyFromSymbol(host_out, mrinr::v7::g_timeline, (size_t)n * 4 * sizeof(long long));
  int zero = 0;
  cudaMemcpyToSymbol(mrinr::v7::g_timeline_n, &zero, sizeof(int));
  return n;
}
#endif
