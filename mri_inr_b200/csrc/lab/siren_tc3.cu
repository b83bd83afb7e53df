// Fused modulated-SIREN synthesis kernel, version 3: CTA pairs (tcgen05 cta_group::2).
//
// Measurement that motivates it (profiles/r01_siren_v2.md): the single-CTA kernels stream 4 x 128 KB of weights
// plus a 128 KB slice of the layer-0 table per 128-row tile and were bound by L2->SM delivery (5.6 TB/s, 670 KB per
// tile).  Here two CTAs of a cluster (one TPC) execute each MMA together as M=256: every CTA keeps its own 128-row
// tile (A operand, TMEM accumulators, epilogue) but only HALF of each layer's weights (its 128 output rows), and
// the tensor cores read the other half from the peer's shared memory -- weight traffic per tile halves.  Layer 0 is
// evaluated in the kernel (2 FFMA + sine per element) instead of being read from the table: no table traffic.
//
// Schedule = version 2 (siren_tc2.cu): two TMEM accumulators, chunk chasing (32-column a_ready barriers), 8 epilogue
// warps per CTA, deferred output epilogue.  Cross-CTA signalling:
//   a_ready[g], acc_free   : live in the leader CTA (rank 0); one arrival per warp, sent with mapa + mbarrier.arrive.shared::cluster
//   w_full[s]              : local transaction barrier of each CTA; the peer forwards "my half has landed" to the
//                            leader's w_peer[s] (the peer's otherwise idle MMA warp does the forwarding)
//   w_empty[s], acc_full[b]: tcgen05.commit.cta_group::2 multicast to both CTAs
// Warp roles per CTA (320 threads): warps 0-7 epilogue; warp 8 lane 0: MMA issuer (leader) / forwarder (peer);
// warp 9 lane 0: weight producer (cp.async.bulk of 16 KB half-slabs).
#include "tc_ptx.cuh"

namespace mrinr {
namespace v3 {

constexpr int kH = 256;
constexpr int kTileM = 128;
constexpr int kSlabBytes = 16384;   // K=64 x N=128 (this CTA's half) x 2 B
constexpr int kNumSlabs = 4;
constexpr int kEpiWarps = 8;
constexpr int kThreads = kEpiWarps * 32 + 64;
constexpr int kMaxLayers = 16;
constexpr int kTmemCols = 512;

constexpr int kOffA = 0;
constexpr int kOffW = 65536;
constexpr int kOffBias = kOffW + kNumSlabs * kSlabBytes;          // [16][256] f32
constexpr int kOffL0 = kOffBias + kMaxLayers * kH * 4;            // [3][256] f32: W0[:,0], W0[:,1], b0
constexpr int kOffLastW = kOffL0 + 3 * kH * 4;                    // [256] f32
constexpr int kOffPart = kOffLastW + kH * 4;                      // [2][128] f32
constexpr int kOffBar = kOffPart + 2 * kTileM * 4;
constexpr int kOffTmemPtr = kOffBar + 32 * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;

constexpr int kBarWFull = 0;     // [4] local, transaction based
constexpr int kBarWPeer = 4;     // [4] leader: the peer's half-slab has landed (1 arrival)
constexpr int kBarWEmpty = 8;    // [4] both CTAs, via multicast commit
constexpr int kBarAReady = 12;   // [8] leader: 8 warp arrivals (4 warps of the column half x 2 CTAs)
constexpr int kBarAccFull = 20;  // [2] both CTAs, via multicast commit
constexpr int kBarAccFree = 22;  // leader: 16 warp arrivals

struct RowInfo {
  const float* mod_base;
  float* out;
};

template <int ACT, bool BF16, bool W0ONE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) siren_tc3_kernel(const SirenTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int C = P.C, L = P.L;
  const uint32_t rank = cluster_ctarank();

  float* s_bias = reinterpret_cast<float*>(smem + kOffBias);
  float* s_l0 = reinterpret_cast<float*>(smem + kOffL0);
  float* s_lastw = reinterpret_cast<float*>(smem + kOffLastW);
  float* s_part = reinterpret_cast<float*>(smem + kOffPart);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  const uint32_t sA = smem_u32(smem + kOffA);
  const uint32_t sW = smem_u32(smem + kOffW);
  const uint32_t bar0 = smem_u32(s_bar);
  auto bar = [bar0](int i) -> uint32_t { return bar0 + 8u * (uint32_t)i; };

  const long long n_act = P.nactive ? (long long)*P.nactive : P.B;
  const long long total_rows = n_act * C;
  const long long n_tiles = (total_rows + kTileM - 1) / kTileM;
  const long long n_pairs = (n_tiles + 1) / 2;
  const long long cluster_id = blockIdx.x >> 1;
  const long long n_clusters = gridDim.x >> 1;
  const size_t layer_stride = (size_t)P.B * kH;

  // ---- one-time setup ----
  for (int i = tid; i < L * kH; i += kThreads) s_bias[i] = P.bias[i];
  for (int i = tid; i < 3 * kH; i += kThreads) s_l0[i] = P.layer0[i];
  for (int i = tid; i < kH; i += kThreads) s_lastw[i] = P.last_w[i];
  if (tid == 0) {
    for (int s = 0; s < kNumSlabs; ++s) {
      mbar_init(bar(kBarWFull + s), 1);
      mbar_init(bar(kBarWPeer + s), 1);
      mbar_init(bar(kBarWEmpty + s), 1);
    }
    for (int g = 0; g < 8; ++g) mbar_init(bar(kBarAReady + g), 8);
    mbar_init(bar(kBarAccFull + 0), 1);
    mbar_init(bar(kBarAccFull + 1), 1);
    mbar_init(bar(kBarAccFree), 2 * kEpiWarps);
    fence_barrier_init();
  }
  if (warp == kEpiWarps) tmem_alloc_pair(smem_u32(s_tmem), kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers are initialised and its TMEM allocated before anyone signals
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp < kEpiWarps) {
    // =========================== epilogue warps (both CTAs, identical) ===========================
    const int q = warp & 3;
    const int half = warp >> 2;
    const int t = q * 32 + lane;
    const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const float last_b = P.last_b ? *P.last_b : 0.f;

    uint32_t ev = 0;
    bool have_prev = false;
    uint32_t prev_ev = 0;
    RowInfo prev_row{nullptr, nullptr};
    uint32_t tile_iter = 0;

    // publish chunk g of the next layer's operand: every lane has fenced its writes; one arrival per warp
    auto publish = [&](int g) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar(kBarAReady + g), 0);
    };

    auto final_epilogue = [&](const RowInfo& ri, uint32_t e, uint32_t it) {
      const uint32_t acc_col = (e & 1u) * 256u + (uint32_t)half * 128u;
      const float* bias_l = s_bias + (L - 1) * kH + half * 128;
      const float* mod_l = ri.mod_base + (size_t)(L - 1) * layer_stride + half * 128;
      const float* lw = s_lastw + half * 128;
      float dot = 0.f;
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        uint32_t v[32];
        tmem_ld32(taddr_row + acc_col + j * 32, v);
        float4 m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = __ldg(reinterpret_cast<const float4*>(mod_l + j * 32) + i);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = *reinterpret_cast<const float4*>(bias_l + j * 32 + i * 4);
          const float4 w = *reinterpret_cast<const float4*>(lw + j * 32 + i * 4);
          dot = fmaf(act_fast<ACT, W0ONE>(__uint_as_float(v[i * 4 + 0]) + b.x, P.w0) * m[i].x, w.x, dot);
          dot = fmaf(act_fast<ACT, W0ONE>(__uint_as_float(v[i * 4 + 1]) + b.y, P.w0) * m[i].y, w.y, dot);
          dot = fmaf(act_fast<ACT, W0ONE>(__uint_as_float(v[i * 4 + 2]) + b.z, P.w0) * m[i].z, w.z, dot);
          dot = fmaf(act_fast<ACT, W0ONE>(__uint_as_float(v[i * 4 + 3]) + b.w, P.w0) * m[i].w, w.w, dot);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar(kBarAccFree), 0);
      float* part = s_part + (it & 1u) * kTileM;
      if (half == 1) {
        part[t] = dot;
        asm volatile("bar.arrive %0, 64;" ::"r"(2 + q) : "memory");
      } else {
        asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");
        // output layer: always sine, never modulated (modulated_siren.py:211-213, :233)
        if (ri.out != nullptr) *ri.out = sinf(P.w0 * (dot + part[t] + last_b));
      }
    };

    for (long long pair = cluster_id; pair < n_pairs; pair += n_clusters, ++tile_iter) {
      const long long tile = pair * 2 + rank;          // may be a phantom tile (>= n_tiles): all rows invalid
      const long long R0 = tile * kTileM;
      const long long pc0 = R0 / C;
      const int c0 = (int)(R0 - pc0 * C);
      const int boundary = C - c0;
      RowInfo my{nullptr, nullptr};
      float g0, g1;
      {
        long long pc = pc0 + (t >= boundary ? 1 : 0);
        if (pc >= n_act) pc = n_act - 1;
        const long long patch = P.idx ? (long long)P.idx[pc] : pc;
        my.mod_base = P.mods + (size_t)patch * kH;
        int c = c0 + t;
        if (c >= C) c -= C;
        if (R0 + t < total_rows) my.out = P.out + patch * C + c;
        const float2 g = __ldg(reinterpret_cast<const float2*>(P.grid) + c);
        g0 = g.x;
        g1 = g.y;
      }

      if (have_prev) {   // A is free once the previous tile's last MMA event has completed
        mbar_wait(bar(kBarAccFull + (prev_ev & 1u)), (prev_ev >> 1) & 1u, P.errflag, 3);
        tc_fence_after();
      }

      // ---- layer 0 (modulated_siren.py:154-156 with dim_in = 2): h = act(w0_initial * (W0 g + b0)) * mod_0 ----
      {
        const float* mod_l = my.mod_base + half * 128;
        prefetch_l1(mod_l + layer_stride + (lane & 3) * 32);     // warm L1 with layer 1's modulation half-row
        const float* wa = s_l0 + half * 128;
        const float* wb = s_l0 + kH + half * 128;
        const float* wc = s_l0 + 2 * kH + half * 128;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          float4 m[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) m[i] = __ldg(reinterpret_cast<const float4*>(mod_l + j * 32) + i);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(wa + j * 32 + i * 4);
            const float4 b = *reinterpret_cast<const float4*>(wb + j * 32 + i * 4);
            const float4 cc = *reinterpret_cast<const float4*>(wc + j * 32 + i * 4);
            const float h0 = act_fast<ACT, false>(fmaf(g1, b.x, fmaf(g0, a.x, cc.x)), P.w0_initial) * m[i].x;
            const float h1 = act_fast<ACT, false>(fmaf(g1, b.y, fmaf(g0, a.y, cc.y)), P.w0_initial) * m[i].y;
            const float h2 = act_fast<ACT, false>(fmaf(g1, b.z, fmaf(g0, a.z, cc.z)), P.w0_initial) * m[i].z;
            const float h3 = act_fast<ACT, false>(fmaf(g1, b.w, fmaf(g0, a.w, cc.w)), P.w0_initial) * m[i].w;
            pk[i * 2 + 0] = pack2<BF16>(h0, h1);
            pk[i * 2 + 1] = pack2<BF16>(h2, h3);
          }
          const int kc0 = (half * 4 + j) * 4;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq)
            *reinterpret_cast<uint4*>(smem + kOffA + (kc0 + qq) * 2048 + t * 16) =
                make_uint4(pk[qq * 4 + 0], pk[qq * 4 + 1], pk[qq * 4 + 2], pk[qq * 4 + 3]);
          publish(half * 4 + j);
        }
      }

      if (have_prev) final_epilogue(prev_row, prev_ev, tile_iter - 1);

      for (int l = 1; l <= L - 2; ++l) {
        const uint32_t e = ev + (uint32_t)(l - 1);
        const uint32_t acc_col = (e & 1u) * 256u + (uint32_t)half * 128u;
        const float* bias_l = s_bias + l * kH + half * 128;
        const float* mod_l = my.mod_base + (size_t)l * layer_stride + half * 128;
        prefetch_l1(mod_l + layer_stride + (lane & 3) * 32);     // next layer's half-row (l+1 <= L-1 exists)
        mbar_wait(bar(kBarAccFull + (e & 1u)), (e >> 1) & 1u, P.errflag, 4);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          uint32_t v[32];
          tmem_ld32(taddr_row + acc_col + j * 32, v);
          float4 m[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) m[i] = __ldg(reinterpret_cast<const float4*>(mod_l + j * 32) + i);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b = *reinterpret_cast<const float4*>(bias_l + j * 32 + i * 4);
            const float h0 = act_fast<ACT, W0ONE>(__uint_as_float(v[i * 4 + 0]) + b.x, P.w0) * m[i].x;
            const float h1 = act_fast<ACT, W0ONE>(__uint_as_float(v[i * 4 + 1]) + b.y, P.w0) * m[i].y;
            const float h2 = act_fast<ACT, W0ONE>(__uint_as_float(v[i * 4 + 2]) + b.z, P.w0) * m[i].z;
            const float h3 = act_fast<ACT, W0ONE>(__uint_as_float(v[i * 4 + 3]) + b.w, P.w0) * m[i].w;
            pk[i * 2 + 0] = pack2<BF16>(h0, h1);
            pk[i * 2 + 1] = pack2<BF16>(h2, h3);
          }
          const int kc0 = (half * 4 + j) * 4;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq)
            *reinterpret_cast<uint4*>(smem + kOffA + (kc0 + qq) * 2048 + t * 16) =
                make_uint4(pk[qq * 4 + 0], pk[qq * 4 + 1], pk[qq * 4 + 2], pk[qq * 4 + 3]);
          tc_fence_before();
          publish(half * 4 + j);
        }
      }
      prev_ev = ev + (uint32_t)(L - 2);
      prev_row = my;
      have_prev = true;
      ev += (uint32_t)(L - 1);
    }
    if (have_prev) {
      mbar_wait(bar(kBarAccFull + (prev_ev & 1u)), (prev_ev >> 1) & 1u, P.errflag, 5);
      tc_fence_after();
      final_epilogue(prev_row, prev_ev, tile_iter - 1);
    }
  } else if (warp == kEpiWarps) {
    if (lane == 0 && rank == 0) {
      // =========================== MMA issuer (leader CTA) ===========================
      const uint32_t idesc = make_idesc(BF16 ? 1 : 0, 2 * kTileM, kH);
      const uint64_t adesc0 = make_smem_desc(sA, 2048, 128);
      uint32_t e = 0, n_free = 0;
      const uint32_t per_tile = (uint32_t)(L - 1);
      for (long long pair = cluster_id; pair < n_pairs; pair += n_clusters) {
        for (int l = 1; l < L; ++l, ++e) {
          if (e >= 2 && ((e - 2) % per_tile) == per_tile - 1) {
            mbar_wait_cluster(bar(kBarAccFree), n_free & 1u, P.errflag, 6);
            ++n_free;
          }
          const uint32_t d_tmem = tmem_base + (e & 1u) * 256u;
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            const int g = (i >> 1) + (i & 1) * 4;
            const int s = g >> 1;
            mbar_wait_cluster(bar(kBarAReady + g), e & 1u, P.errflag, 1);
            if ((g & 1) == 0) {
              mbar_wait(bar(kBarWFull + s), e & 1u, P.errflag, 2);
              mbar_wait_cluster(bar(kBarWPeer + s), e & 1u, P.errflag, 8);
            }
            tc_fence_after();
            const uint64_t bdesc0 = make_smem_desc(sW + s * kSlabBytes, 2048, 128);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const int k = g * 2 + kk;
              const uint64_t ad = adesc0 + (uint64_t)((k * 4096) >> 4);
              const uint64_t bd = bdesc0 + (uint64_t)(((k & 3) * 4096) >> 4);
              umma_f16_pair(d_tmem, ad, bd, idesc, (i | kk) != 0 ? 1u : 0u);
            }
            if (g & 1) umma_commit_pair(bar(kBarWEmpty + s), 3);
          }
          umma_commit_pair(bar(kBarAccFull + (e & 1u)), 3);
        }
      }
    } else if (lane == 0) {
      // =========================== forwarder (peer CTA): my half-slab has landed ===========================
      uint32_t e = 0;
      for (long long pair = cluster_id; pair < n_pairs; pair += n_clusters) {
        for (int l = 1; l < L; ++l, ++e) {
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            const int s = (i >> 1) + (i & 1) * 2;
            mbar_wait(bar(kBarWFull + s), e & 1u, P.errflag, 9);
            mbar_arrive_cluster(bar(kBarWPeer + s), 0);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== weight producer (both CTAs: own half of every slab) ===========================
    if (lane == 0) {
      uint32_t e = 0;
      for (long long pair = cluster_id; pair < n_pairs; pair += n_clusters) {
        for (int l = 1; l < L; ++l, ++e) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(P.w16p) + (size_t)(l - 1) * kH * kH * 2 +
                               (size_t)rank * (kH / 2) * kH * 2;
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            const int s = (i >> 1) + (i & 1) * 2;
            mbar_wait(bar(kBarWEmpty + s), (e & 1u) ^ 1u, P.errflag, 7);
            mbar_expect_tx(bar(kBarWFull + s), kSlabBytes);
            bulk_g2s(sW + s * kSlabBytes, src + (size_t)s * kSlabBytes, kSlabBytes, bar(kBarWFull + s));
          }
        }
      }
    }
    __syncwarp();
  }

  // ---- teardown: both CTAs must be done before the pair's TMEM is released ----
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (warp == kEpiWarps) tmem_dealloc_pair(tmem_base, kTmemCols);
}

template <int ACT, bool BF16, bool W0ONE>
static int launch_one(const SirenTcParams& P, int grid, cudaStream_t st) {
  MRINR_SMEM_OPT_IN((siren_tc3_kernel<ACT, BF16, W0ONE>), kSmemBytes);
  siren_tc3_kernel<ACT, BF16, W0ONE><<<grid, kThreads, kSmemBytes, st>>>(P);
  count_launch();
  return check_launch("siren_tc3");
}

int launch_siren_tc_v3(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                       int64_t B, float* d_out, cudaStream_t st) {
  MRINR_REQUIRE(p->H == kH && p->L <= kMaxLayers && p->L >= 3 && p->C >= kTileM, MRINR_E_UNSUPPORTED,
                "siren_tc3: unsupported configuration (H=%d L=%d C=%d)", p->H, p->L, p->C);
  SirenTcParams P;
  P.table0 = p->d_table0; P.w16 = p->d_net_w16; P.w16p = p->d_net_w16p; P.layer0 = p->d_layer0; P.grid = p->d_grid;
  P.w0_initial = p->w0_initial;
  P.bias = p->d_net_bias; P.last_w = p->d_last_w;
  P.last_b = p->d_last_b; P.mods = d_mods; P.idx = d_idx; P.nactive = d_nactive; P.out = d_out;
  P.errflag = p->d_errflag; P.B = B; P.C = p->C; P.L = p->L; P.w0 = p->w0;
  const long long n_tiles = (B * p->C + kTileM - 1) / kTileM;
  const long long n_pairs = (n_tiles + 1) / 2;
  long long clusters = p->num_sms / 2;
  if (clusters > n_pairs) clusters = n_pairs;
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

}  // namespace v3
}  // namespace mrinr
