// Fused modulated-SIREN synthesis kernel, version 2: the epilogue of layer l and the MMAs of layer l+1 overlap.
//
// Same math, operand layouts and weight ring as siren_tc.cu (see its header); what changes is the schedule:
//
//  * two TMEM accumulators (2 x 256 columns): MMA event e (one per tile-layer) writes acc[e & 1] while the
//    epilogue of event e-1 is still reading acc[(e-1) & 1];
//  * chunk chasing: the epilogue publishes the next layer's A operand in 32-column chunks (a_ready[0..7]); the
//    MMA thread issues the two K=16 steps of a chunk as soon as that chunk has landed, so the tensor pipe
//    trails the epilogue by one chunk instead of one layer;
//  * 8 epilogue warps: warp w owns TMEM lane quarter (w & 3) = rows 32(w&3)..+31 and column half (w >> 2), so
//    the two halves of every row are processed concurrently and every SM sub-partition has two epilogue warps;
//  * the last hidden layer's epilogue (fp32 dot with w_last + output sine) of tile i is deferred until after the
//    layer-0 operand of tile i+1 has been produced, so the tensor pipe is already busy with tile i+1 while the
//    output of tile i is finished (acc_free guards the accumulator it still reads);
//  * modulation vectors are read straight from global memory (warp-uniform 16-byte loads, L1 resident: 2 KB per
//    tile-layer) instead of being staged through shared memory.
//
// Warp roles (320 threads): warps 0-7 epilogue, warp 8 lane 0 issues the MMAs, warp 9 lane 0 streams weight slabs.
#include "tc_ptx.cuh"

#include <cstdlib>

namespace mrinr {
namespace v1 {
int launch_siren_tc_v1(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                       int64_t B, float* d_out, cudaStream_t st);
}
namespace v2 {

constexpr int kH = 256;
constexpr int kTileM = 128;
constexpr int kSlabBytes = 32768;   // K=64 x N=256 x 2 B
constexpr int kNumSlabs = 4;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = kEpiThreads + 64;
constexpr int kMaxLayers = 16;
constexpr int kTmemCols = 512;

constexpr int kOffA = 0;
constexpr int kOffW = 65536;
constexpr int kOffBias = kOffW + kNumSlabs * kSlabBytes;          // [16][256] f32
constexpr int kOffLastW = kOffBias + kMaxLayers * kH * 4;         // [256] f32
constexpr int kOffPart = kOffLastW + kH * 4;                      // [2][128] f32 partial dots of column half 1
constexpr int kOffBar = kOffPart + 2 * kTileM * 4;                // barriers
constexpr int kOffTmemPtr = kOffBar + 32 * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;

// barrier indices (8 bytes each)
constexpr int kBarWFull = 0;     // [4]
constexpr int kBarWEmpty = 4;    // [4]
constexpr int kBarAReady = 8;    // [8]  128 arrivals each (the 4 warps of one column half)
constexpr int kBarAccFull = 16;  // [2]
constexpr int kBarAccFree = 18;  // 256 arrivals: deferred final epilogue finished reading its accumulator

struct RowInfo {
  const float* mod_base;   // mods + patch*256 (layer 0); layer l adds l*B*256
  float* out;              // &out[patch*C + c] or nullptr for a padding row
};

template <int ACT, bool BF16, bool W0ONE>
__global__ void __launch_bounds__(kThreads, 1) siren_tc2_kernel(const SirenTcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int C = P.C, L = P.L;

  float* s_bias = reinterpret_cast<float*>(smem + kOffBias);
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
  const size_t layer_stride = (size_t)P.B * kH;   // floats between mods of consecutive layers

  // ---- one-time setup ----
  for (int i = tid; i < L * kH; i += kThreads) s_bias[i] = P.bias[i];
  for (int i = tid; i < kH; i += kThreads) s_lastw[i] = P.last_w[i];
  if (tid == 0) {
    for (int s = 0; s < kNumSlabs; ++s) {
      mbar_init(bar(kBarWFull + s), 1);
      mbar_init(bar(kBarWEmpty + s), 1);
    }
    for (int g = 0; g < 8; ++g) mbar_init(bar(kBarAReady + g), kEpiThreads / 2);
    mbar_init(bar(kBarAccFull + 0), 1);
    mbar_init(bar(kBarAccFull + 1), 1);
    mbar_init(bar(kBarAccFree), kEpiThreads);
    fence_barrier_init();
  }
  if (warp == kEpiWarps) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp < kEpiWarps) {
    // =========================== epilogue warps ===========================
    const int q = warp & 3;            // TMEM lane quarter
    const int half = warp >> 2;        // column half: columns [128*half, 128*half+128)
    const int t = q * 32 + lane;       // row within the tile == TMEM lane
    const uint32_t taddr_row = tmem_base + ((uint32_t)(q * 32) << 16);
    const float last_b = P.last_b ? *P.last_b : 0.f;
    const int rsub = lane & 7, kq = lane >> 3;   // layer-0 mapping: (row in group of 8, one of 4 k-slabs)

    uint32_t ev = 0;                   // event id of this tile's layer 1
    bool have_prev = false;
    uint32_t prev_ev = 0;              // event id of the previous tile's last hidden layer
    RowInfo prev_row{nullptr, nullptr};
    uint32_t tile_iter = 0;

    // final epilogue of a finished tile: dot with w_last over this thread's 128 columns, pair-combine, sine
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
      mbar_arrive(bar(kBarAccFree));
      // combine the two column halves of a row: warp q+4 hands its partial to warp q
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

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_iter) {
      const long long R0 = tile * kTileM;
      const long long pc0 = R0 / C;
      const int c0 = (int)(R0 - pc0 * C);
      const int boundary = C - c0;                       // rows >= boundary belong to compact patch pc0+1
      auto patch_of = [&](int row) -> long long {        // original patch index of a tile row (clamped for padding)
        long long pc = pc0 + (row >= boundary ? 1 : 0);
        if (pc >= n_act) pc = n_act - 1;
        return P.idx ? (long long)P.idx[pc] : pc;
      };
      RowInfo my{nullptr, nullptr};
      {
        const long long patch = patch_of(t);
        my.mod_base = P.mods + (size_t)patch * kH;
        int c = c0 + t;
        if (c >= C) c -= C;
        if (R0 + t < total_rows) my.out = P.out + patch * C + c;
      }

      // the A tile is free once the previous tile's last MMA event has completed
      if (have_prev) {
        mbar_wait(bar(kBarAccFull + (prev_ev & 1u)), (prev_ev >> 1) & 1u, P.errflag, 3);
        tc_fence_after();
      }

      // ---- layer 0: A[:, 128*half ..] = f16(table0[c] * mod_0), chunk by chunk ----
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const int g = half * 4 + j;            // 32-column chunk == two K=16 steps of the next layer
        const int kc = g * 4 + kq;             // this lane's 8-column slab
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int row = q * 32 + g8 * 8 + rsub;
          int c = c0 + row;
          if (c >= C) c -= C;
          const float4* trow = reinterpret_cast<const float4*>(P.table0 + (size_t)c * kH + kc * 8);
          const float4* mrow = reinterpret_cast<const float4*>(P.mods + (size_t)patch_of(row) * kH + kc * 8);
          const float4 a0 = __ldg(trow), a1 = __ldg(trow + 1);
          const float4 m0 = __ldg(mrow), m1 = __ldg(mrow + 1);
          uint4 pk;
          pk.x = pack2<BF16>(a0.x * m0.x, a0.y * m0.y);
          pk.y = pack2<BF16>(a0.z * m0.z, a0.w * m0.w);
          pk.z = pack2<BF16>(a1.x * m1.x, a1.y * m1.y);
          pk.w = pack2<BF16>(a1.z * m1.z, a1.w * m1.w);
          *reinterpret_cast<uint4*>(smem + kOffA + kc * 2048 + row * 16) = pk;
        }
        fence_proxy_async();
        mbar_arrive(bar(kBarAReady + g));
      }

      // ---- deferred output of the previous tile (the tensor pipe is already working on this tile) ----
      if (have_prev) final_epilogue(prev_row, prev_ev, tile_iter - 1);

      // ---- hidden layers 1 .. L-2: epilogue publishes the next layer's operand chunk by chunk ----
      for (int l = 1; l <= L - 2; ++l) {
        const uint32_t e = ev + (uint32_t)(l - 1);
        const uint32_t acc_col = (e & 1u) * 256u + (uint32_t)half * 128u;
        const float* bias_l = s_bias + l * kH + half * 128;
        const float* mod_l = my.mod_base + (size_t)l * layer_stride + half * 128;
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
          fence_proxy_async();
          mbar_arrive(bar(kBarAReady + half * 4 + j));
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
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BF16 ? 1 : 0, kTileM, kH);
      const uint64_t adesc0 = make_smem_desc(sA, 2048, 128);
      uint32_t e = 0, n_free = 0;
      const uint32_t per_tile = (uint32_t)(L - 1);
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 1; l < L; ++l, ++e) {
          // acc[e&1] was last read by event e-2; if that was a tile's deferred final epilogue, wait for it
          if (e >= 2 && ((e - 2) % per_tile) == per_tile - 1) {
            mbar_wait(bar(kBarAccFree), n_free & 1u, P.errflag, 6);
            ++n_free;
          }
          const uint32_t d_tmem = tmem_base + (e & 1u) * 256u;
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            const int g = (i >> 1) + (i & 1) * 4;            // chunk order 0,4,1,5,2,6,3,7 (both halves advance together)
            const int s = g >> 1;                            // weight slab (K=64) holding this chunk's two K steps
            mbar_wait(bar(kBarAReady + g), e & 1u, P.errflag, 1);
            if ((g & 1) == 0) mbar_wait(bar(kBarWFull + s), e & 1u, P.errflag, 2);
            tc_fence_after();
            const uint64_t bdesc0 = make_smem_desc(sW + s * kSlabBytes, 4096, 128);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const int k = g * 2 + kk;                      // K=16 step within the layer
              const uint64_t ad = adesc0 + (uint64_t)((k * 4096) >> 4);
              const uint64_t bd = bdesc0 + (uint64_t)(((k & 3) * 8192) >> 4);
              umma_f16(d_tmem, ad, bd, idesc, (i | kk) != 0 ? 1u : 0u);
            }
            if (g & 1) umma_commit(bar(kBarWEmpty + s));     // slab fully consumed once these MMAs retire
          }
          umma_commit(bar(kBarAccFull + (e & 1u)));
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== weight producer ===========================
    if (lane == 0) {
      uint32_t e = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 1; l < L; ++l, ++e) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(P.w16) + (size_t)(l - 1) * kH * kH * 2;
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            const int s = (i >> 1) + (i & 1) * 2;            // slab order 0,2,1,3 = order of first use
            mbar_wait(bar(kBarWEmpty + s), (e & 1u) ^ 1u, P.errflag, 7);
            mbar_expect_tx(bar(kBarWFull + s), kSlabBytes);
            bulk_g2s(sW + s * kSlabBytes, src + (size_t)s * kSlabBytes, kSlabBytes, bar(kBarWFull + s));
          }
        }
      }
    }
    __syncwarp();
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kEpiWarps) tmem_dealloc(tmem_base, kTmemCols);
}

template <int ACT, bool BF16, bool W0ONE>
static int launch_one(const SirenTcParams& P, int grid, cudaStream_t st) {
  MRINR_SMEM_OPT_IN((siren_tc2_kernel<ACT, BF16, W0ONE>), kSmemBytes);
  siren_tc2_kernel<ACT, BF16, W0ONE><<<grid, kThreads, kSmemBytes, st>>>(P);
  count_launch();
  return check_launch("siren_tc2");
}

int launch_siren_tc_v2(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                       int64_t B, float* d_out, cudaStream_t st) {
  MRINR_REQUIRE(p->H == kH && p->L <= kMaxLayers && p->L >= 2 && p->C >= kTileM, MRINR_E_UNSUPPORTED,
                "siren_tc2: unsupported configuration (H=%d L=%d C=%d)", p->H, p->L, p->C);
  SirenTcParams P;
  P.table0 = p->d_table0; P.w16 = p->d_net_w16; P.bias = p->d_net_bias; P.last_w = p->d_last_w;
  P.last_b = p->d_last_b; P.mods = d_mods; P.idx = d_idx; P.nactive = d_nactive; P.out = d_out;
  P.errflag = p->d_errflag; P.B = B; P.C = p->C; P.L = p->L; P.w0 = p->w0;
  const long long n_tiles = (B * p->C + kTileM - 1) / kTileM;
  int grid = p->num_sms;
  if ((long long)grid > n_tiles) grid = (int)n_tiles;
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

}  // namespace v2

}  // namespace mrinr
