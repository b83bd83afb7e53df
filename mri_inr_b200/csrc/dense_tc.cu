// Dense layers of the patch encoder and the modulator on the tensor cores, at fp32-level accuracy.
//
//   C[M,N] = act( [A1 | A2] . W^T + bias ),   A1 [M,K1], A2 [M,K2] fp32 row-major (A2 optional),  W [N, K1+K2]
//
// covers Conv2d(32,64,8) on an 8x8 map (a [B,2048] x [2048,64] product), Linear(64,Z) (siren_encoder.py:503-512) and
// every Modulator layer, relu(A_l cat(h_{l-1}, z) + c_l) (modulated_siren.py:325-343; the hidden columns of A_l come
// first, :341, so A1 = h_{l-1}, A2 = z).  One launch per layer: the intermediates are a few KB per patch and HBM is
// idle next to the synthesis kernel, so nothing is gained by keeping them on chip.
//
// Precision: tcgen05 kind::f16 has an 11-bit significand.  Every fp32 operand x is split into two fp16 numbers,
// hi = rn16(x), lo = rn16(x - hi) (22 significant bits), and the product is formed from three MMAs,
//   A.B ~= A_hi.B_hi + A_hi.B_lo + A_lo.B_hi        (the dropped lo.lo term is 2^-22 relative),
// accumulated in fp32 in TMEM: ~1e-6 relative, i.e. the noise level of an fp32 FFMA loop with a different summation
// order.  tests/test_gpu_parity.py holds the result to 1e-5 against the reference.
// Operand range: hi and lo are fp16, so the split is exact to 2^-22 only for 6.1e-5 <= |x| < 65504 (below that the
// absolute floor is the fp16 subnormal spacing 6e-8, above it hi overflows).  The reference's inputs are images
// normalised to [0, 1] (normalize_scan, preprocessing.py:127-137) and its weights / activations are O(1).
//
// Per CTA: one 128-row tile, all N columns (N = 64, 128 or 256), K in slabs of 32 through a 2-stage ring (96 KB for
// N = 256), so that TWO CTAs are resident per SM: one CTA's prologue / epilogue (all latency) hides behind the
// other's MMAs.  (With one 192 KB CTA per SM the tensor pipe was 18 % active, profiles/r01_ncu_dense_split.txt.)
//   warps 0-3  A producers: coalesced loads of the fp32 rows (two slabs ahead, in registers), split, write both halves in
//              the UMMA K-major core-matrix layout; afterwards the epilogue (tcgen05.ld, bias, activation, store)
//   warp 4     MMA issuer (converged warp + elect.sync); owns the TMEM allocation
//   warp 5     weight producer: one cp.async.bulk per slab (hi and lo halves are contiguous in the packed array)
// Clusters of kCluster CTAs share the weights: every CTA fetches 1/kCluster of a slab and MULTICASTS it into the same
// shared-memory offset of all CTAs of the cluster (cp.async.bulk ... .multicast::cluster), so the L2 -> SM weight
// traffic per row tile drops by kCluster (the kernel is bound by L2 delivery: 64 KB of weights per 32 KB of operand
// rows without it).  A stage may only be refilled when the MMAs of ALL CTAs have consumed it, so the tensor core's
// completion (tcgen05.commit) is multicast too: the `empty` barrier of every CTA counts kCluster arrivals.
#include "tc_ptx.cuh"

namespace mrinr {
namespace dense {

constexpr int kBM = 128;
constexpr int kSlabK = 32;
constexpr int kKc = kSlabK / 8;    // 16-byte K chunks per slab
constexpr int kStages = 2;
constexpr int kThreads = 192;
#ifndef MRINR_DENSE_CLUSTER
#define MRINR_DENSE_CLUSTER 2
#endif
constexpr int kCluster = MRINR_DENSE_CLUSTER;
// A operand half (hi or lo) of a stage: [4 kc][128 rows][8] fp16, K chunks kALbo bytes apart.  The 64 bytes of padding
// per chunk spread a producer warp's store instruction (4 rows x 4 K chunks x 16 bytes) over all 32 banks: with chunks
// exactly 2 KB apart the four K chunks of a row fell on the same banks (4-way conflict, ncu: 75 % of the kernel's
// shared-memory wavefronts were conflict replays on a shared-memory port that also feeds the tensor core).
constexpr int kALbo = kBM * 16 + 64;
constexpr int kAHalfBytes = kKc * kALbo;

template <int N>
struct Cfg {
  static constexpr int kWHalfBytes = N * kSlabK * 2;   // [4 kc][N][8] fp16
  static constexpr int kStageBytes = 2 * kAHalfBytes + 2 * kWHalfBytes;
  static constexpr int kOffBias = kStages * kStageBytes;          // [N] f32
  static constexpr int kOffBar = kOffBias + N * 4;
  static constexpr int kSmemBytes = kOffBar + 64 + 16;
  static constexpr int kTmemCols = N < 32 ? 32 : N;
};

struct DenseParams {
  const float* a1; long long lda1; int K1;
  const float* a2; long long lda2; int K2;
  const uint16_t* w;      // packed by pack_split_kernel
  const float* bias;      // [N] or null
  float* c; long long ldc;
  long long M;
  int act;                // 0 none, 1 relu, 2 leaky relu
  float slope;
  int32_t* errflag;
};

// W [N, K] fp32 row-major (nn.Linear / flattened Conv2d weight) -> [K/32 slabs][hi, lo][4 kc][N][8] fp16
__global__ void pack_split_kernel(const float* __restrict__ w, int N, int K, uint16_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * K) return;
  const int e = (int)(i & 7);
  const int n = (int)((i >> 3) % N);
  const int kc = (int)(((i >> 3) / N) % kKc);
  const int slab = (int)(((i >> 3) / N) / kKc);
  const float x = w[(long long)n * K + slab * kSlabK + kc * 8 + e];
  const __half hi = __float2half_rn(x);
  const __half lo = __float2half_rn(x - __half2float(hi));
  const long long base = (long long)slab * 2 * kKc * N * 8 + ((long long)kc * N + n) * 8 + e;
  out[base] = *reinterpret_cast<const uint16_t*>(&hi);
  out[base + (long long)kKc * N * 8] = *reinterpret_cast<const uint16_t*>(&lo);
}

__device__ __forceinline__ void split8(const float4& u, const float4& v, uint4& hi, uint4& lo) {
  const float x[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x[2 * i] - back.x, x[2 * i + 1] - back.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void split4(const float4& u, uint2& hi, uint2& lo) {
  const float x[4] = {u.x, u.y, u.z, u.w};
  uint32_t h[2], l[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x[2 * i] - back.x, x[2 * i + 1] - back.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint2(h[0], h[1]);
  lo = make_uint2(l[0], l[1]);
}

template <int N>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 2) dense_split_kernel(const DenseParams P) {
  using C = Cfg<N>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + C::kOffBar);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + C::kOffBar + 64);
  const uint32_t bar0 = smem_u32(s_bar);
  // barriers: a_full[2] (4 producer warps), w_full[2] (tx), empty[2] (commit), acc_full
  auto bar_afull = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto bar_wfull = [&](int s) { return bar0 + 8u * (uint32_t)(2 + s); };
  auto bar_empty = [&](int s) { return bar0 + 8u * (uint32_t)(4 + s); };
  const uint32_t bar_acc = bar0 + 8u * 6u;
  const int n_slabs = (P.K1 + P.K2) / kSlabK;
  float* s_bias = reinterpret_cast<float*>(smem + C::kOffBias);
  for (int i = tid; i < N; i += kThreads) s_bias[i] = P.bias ? P.bias[i] : 0.f;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_afull(s), 4);
      mbar_init(bar_wfull(s), 1);
      mbar_init(bar_empty(s), kCluster);
    }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(s_tmem), C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // every CTA's barriers exist before a peer multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t crank = cluster_ctarank();

  if (warp < 4) {
    // =========================== A producers, then epilogue ===========================
    // Loads are COALESCED: one warp instruction covers 4 rows x 128 bytes (lane -> row 4i + lane/8, floats 4 (lane%8)
    // .. +3 of the slab) instead of 32 different lines (thread = row), which kept the L1 tag stage busy 8x longer for
    // the same bytes.  A lane then owns half of a 16-byte UMMA entry (4 of its 8 K values) and stores 8 bytes.
    const int lrow = lane >> 3, lq = lane & 7;
    const long long row0 = (long long)blockIdx.x * kBM + warp * 32 + lrow;      // + 4 i
    auto load_slab = [&](int slab, float4 (&r)[2 * kKc]) {
      const int k = slab * kSlabK;
      const float* src = (k < P.K1) ? P.a1 + row0 * P.lda1 + k : P.a2 + row0 * P.lda2 + (k - P.K1);
      const long long ld = (k < P.K1) ? P.lda1 : P.lda2;
#pragma unroll
      for (int i = 0; i < 2 * kKc; ++i)
        r[i] = (row0 + 4 * i < P.M) ? __ldg(reinterpret_cast<const float4*>(src + (long long)(4 * i) * ld) + lq)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    static_assert(kSlabK == 32 && 2 * kKc == 8, "a warp instruction covers 4 rows of one 32-column slab");
    // The operand rows are prefetched TWO slabs ahead in registers (three rotating sets; the loop is unrolled by three
    // so the rotation is a renaming): with one slab ahead a CTA had 16 KB in flight and every slab was a dependent
    // DRAM round trip -- the kernel ran at ~3 TB/s of operand traffic with the tensor pipe a third busy.
    const int st_off = (lq >> 1) * kALbo + (warp * 32 + lrow) * 16 + (lq & 1) * 8;    // + 64 i
    auto step = [&](int slab, const float4 (&r)[2 * kKc]) {
      const int st = slab % kStages;
      mbar_wait(bar_empty(st), ((slab / kStages) & 1u) ^ 1u, P.errflag, 21);
      uint8_t* a_hi = smem + st * C::kStageBytes + st_off;
      uint8_t* a_lo = a_hi + kAHalfBytes;
#pragma unroll
      for (int i = 0; i < 2 * kKc; ++i) {
        uint2 hi, lo;
        split4(r[i], hi, lo);
        *reinterpret_cast<uint2*>(a_hi + i * 64) = hi;
        *reinterpret_cast<uint2*>(a_lo + i * 64) = lo;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_afull(st));
    };
    float4 r0[2 * kKc], r1[2 * kKc], r2[2 * kKc];
    if (n_slabs > 0) load_slab(0, r0);
    if (n_slabs > 1) load_slab(1, r1);
    for (int slab = 0; slab < n_slabs; slab += 3) {
      if (slab + 2 < n_slabs) load_slab(slab + 2, r2);
      step(slab, r0);
      if (slab + 1 >= n_slabs) break;
      if (slab + 3 < n_slabs) load_slab(slab + 3, r0);
      step(slab + 1, r1);
      if (slab + 2 >= n_slabs) break;
      if (slab + 4 < n_slabs) load_slab(slab + 4, r1);
      step(slab + 2, r2);
    }
    // ---- epilogue: this thread owns row `row` (TMEM lane 32*warp + lane), 32 columns per trip ----
    // A thread holds 128 contiguous bytes of ITS row; storing them directly makes every store instruction of a warp
    // touch 32 rows x 16 bytes (half sectors, 1 KB apart) -- ncu: the epilogue was 45 % of a CTA's life, stalled at the
    // loop top on the previous trip's store addresses (the store queue).  The warp therefore transposes each 32 x 128 B
    // block through shared memory (the stage ring is free once the last MMA has completed; chunk index XOR row so that
    // both directions are conflict-free) and writes whole 128-byte lines: 4 rows per instruction.
    mbar_wait(bar_acc, 0u, P.errflag, 22);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint8_t* stg = smem + warp * 4096;
    const int orow = lane >> 3, och = lane & 7;
    const long long row_base = (long long)blockIdx.x * kBM + warp * 32;
#pragma unroll 1
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(taddr + (uint32_t)c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bj = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * j);
        float y[4] = {__uint_as_float(v[4 * j]) + bj.x, __uint_as_float(v[4 * j + 1]) + bj.y,
                      __uint_as_float(v[4 * j + 2]) + bj.z, __uint_as_float(v[4 * j + 3]) + bj.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (P.act == 1) y[i] = fmaxf(y[i], 0.f);
          else if (P.act == 2) y[i] = y[i] > 0.f ? y[i] : y[i] * P.slope;
        }
        *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(y[0], y[1], y[2], y[3]);
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = orow + 4 * k;
        const float4 t = *reinterpret_cast<const float4*>(stg + r * 128 + ((och ^ (r & 7)) << 4));
        if (row_base + r < P.M) *reinterpret_cast<float4*>(P.c + (row_base + r) * P.ldc + c0 + 4 * och) = t;
      }
      __syncwarp();
    }
    tc_fence_before();
  } else if (warp == 4) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc(0, kBM, N);
    for (int slab = 0; slab < n_slabs; ++slab) {
      const int st = slab % kStages;
      const uint32_t par = (slab / kStages) & 1u;
      mbar_wait_backoff(bar_afull(st), par, P.errflag, 23, 32);
      mbar_wait_backoff(bar_wfull(st), par, P.errflag, 24, 32);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + st * C::kStageBytes);
        const uint64_t a_hi = make_smem_desc(sa, kALbo, 128);
        const uint64_t a_lo = make_smem_desc(sa + kAHalfBytes, kALbo, 128);
        const uint64_t b_hi = make_smem_desc(sa + 2 * kAHalfBytes, N * 16, 128);
        const uint64_t b_lo = make_smem_desc(sa + 2 * kAHalfBytes + C::kWHalfBytes, N * 16, 128);
#pragma unroll
        for (int k = 0; k < kSlabK / 16; ++k) {
          const uint64_t da = (uint64_t)((k * 2 * kALbo) >> 4);
          const uint64_t db = (uint64_t)((k * 2 * N * 16) >> 4);
          // small terms first
          umma_f16(tmem_base, a_lo + da, b_hi + db, idesc, (slab | k) != 0 ? 1u : 0u);
          umma_f16(tmem_base, a_hi + da, b_lo + db, idesc, 1u);
          umma_f16(tmem_base, a_hi + da, b_hi + db, idesc, 1u);
        }
        umma_commit_mc(bar_empty(st), (uint16_t)((1u << kCluster) - 1u));   // frees the stage in every CTA
        if (slab == n_slabs - 1) umma_commit(bar_acc);
      }
      __syncwarp();
    }
  } else {
    // =========================== weight producer ===========================
    if (lane == 0) {
      for (int slab = 0; slab < n_slabs; ++slab) {
        const int st = slab % kStages;
        mbar_wait_backoff(bar_empty(st), ((slab / kStages) & 1u) ^ 1u, P.errflag, 25);
        mbar_expect_tx(bar_wfull(st), 2u * C::kWHalfBytes);                       // the whole slab lands here ...
        constexpr uint32_t part = 2u * C::kWHalfBytes / kCluster;                  // ... in kCluster multicast pieces
        bulk_g2s_mc(smem_u32(smem + st * C::kStageBytes + 2 * kAHalfBytes) + crank * part,
                    reinterpret_cast<const uint8_t*>(P.w) + (size_t)slab * 2 * C::kWHalfBytes + (size_t)crank * part, part,
                    bar_wfull(st), (uint16_t)((1u << kCluster) - 1u));
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // no CTA may exit while a peer can still multicast into its shared memory
  tc_fence_after();
  if (warp == 4) tmem_dealloc(tmem_base, C::kTmemCols);
}

// ---- a CHAIN of dense layers in one launch (the modulator: modulated_siren.py:325-343) ---------------------------------
// Layer l+1 of the modulator reads rows of layer l's output and of the latent -- only the SAME 128 rows the CTA has
// just produced -- so a CTA can walk all layers of its row tile without any grid-wide synchronisation: "all layers'
// modulations for a batch of latents in one launch" (north_star (2)).  The pipeline state runs on across layers (slab
// counter, stage parities); the producer warps are the epilogue warps, so a layer's accumulator has been read before the
// first operand slab of the next layer is produced, and the weight producer streams the next layer's slabs underneath
// the epilogue.  A layer's output goes to global memory (it is an output: mods[l]) and comes back as the next layer's
// operand through L2 (ld.global.cg after a block-level fence: each warp re-reads exactly the rows it stored).
// Because the stage ring stays live, the epilogue has its own 2 KB per warp of transpose space (16 columns per trip:
// 64-byte row segments, two full sectors).
constexpr int kMaxChain = 16;
struct ChainLayer {
  const float* a1; long long lda1; int K1;
  const float* a2; long long lda2; int K2;
  const uint16_t* w;      // packed by pack_split_kernel
  const float* bias;      // [N] or null
  float* c; long long ldc;
};
struct ChainParams {
  ChainLayer layer[kMaxChain];
  int n_layers;
  long long M;
  int act;                // 0 none, 1 relu, 2 leaky relu
  float slope;
  int32_t* errflag;
};

template <int N>
struct ChainCfg {
  static constexpr int kOffStg = kStages * Cfg<N>::kStageBytes;      // 4 warps x [32 rows][64 B]
  static constexpr int kOffBar = kOffStg + 4 * 2048;
  static constexpr int kSmemBytes = kOffBar + 64 + 16;
};

template <int N>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 2)
dense_chain_kernel(const __grid_constant__ ChainParams P) {
  using C = Cfg<N>;
  using CC = ChainCfg<N>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + CC::kOffBar);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + CC::kOffBar + 64);
  const uint32_t bar0 = smem_u32(s_bar);
  auto bar_afull = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto bar_wfull = [&](int s) { return bar0 + 8u * (uint32_t)(2 + s); };
  auto bar_empty = [&](int s) { return bar0 + 8u * (uint32_t)(4 + s); };
  const uint32_t bar_acc = bar0 + 8u * 6u;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_afull(s), 4);
      mbar_init(bar_wfull(s), 1);
      mbar_init(bar_empty(s), kCluster);
    }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(s_tmem), C::kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t crank = cluster_ctarank();

  if (warp < 4) {
    // =========================== operand producers and epilogue, layer by layer ===========================
    const int lrow = lane >> 3, lq = lane & 7;
    const long long row0 = (long long)blockIdx.x * kBM + warp * 32 + lrow;      // + 4 i
    const int st_off = (lq >> 1) * kALbo + (warp * 32 + lrow) * 16 + (lq & 1) * 8;    // + 64 i
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    uint8_t* stg = smem + CC::kOffStg + warp * 2048;
    const int orow = lane >> 2, och = lane & 3;
    const long long row_base = (long long)blockIdx.x * kBM + warp * 32;
    uint32_t sg = 0;                                         // slabs of the layers before this one
    for (int l = 0; l < P.n_layers; ++l) {
      const ChainLayer& Lr = P.layer[l];
      const int n_slabs = (Lr.K1 + Lr.K2) / kSlabK;
      auto load_slab = [&](int slab, float4 (&r)[2 * kKc]) {
        const int k = slab * kSlabK;
        const float* src = (k < Lr.K1) ? Lr.a1 + row0 * Lr.lda1 + k : Lr.a2 + row0 * Lr.lda2 + (k - Lr.K1);
        const long long ld = (k < Lr.K1) ? Lr.lda1 : Lr.lda2;
#pragma unroll
        for (int i = 0; i < 2 * kKc; ++i)      // .cg: the previous layer's rows were written by this CTA a moment ago
          r[i] = (row0 + 4 * i < P.M) ? __ldcg(reinterpret_cast<const float4*>(src + (long long)(4 * i) * ld) + lq)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      auto step = [&](int slab, const float4 (&r)[2 * kKc]) {
        const uint32_t s = sg + (uint32_t)slab;
        const int st = (int)(s % kStages);
        mbar_wait(bar_empty(st), ((s / kStages) & 1u) ^ 1u, P.errflag, 41);
        uint8_t* a_hi = smem + st * C::kStageBytes + st_off;
        uint8_t* a_lo = a_hi + kAHalfBytes;
#pragma unroll
        for (int i = 0; i < 2 * kKc; ++i) {
          uint2 hi, lo;
          split4(r[i], hi, lo);
          *reinterpret_cast<uint2*>(a_hi + i * 64) = hi;
          *reinterpret_cast<uint2*>(a_lo + i * 64) = lo;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_afull(st));
      };
      float4 r0[2 * kKc], r1[2 * kKc], r2[2 * kKc];
      if (n_slabs > 0) load_slab(0, r0);
      if (n_slabs > 1) load_slab(1, r1);
      for (int slab = 0; slab < n_slabs; slab += 3) {
        if (slab + 2 < n_slabs) load_slab(slab + 2, r2);
        step(slab, r0);
        if (slab + 1 >= n_slabs) break;
        if (slab + 3 < n_slabs) load_slab(slab + 3, r0);
        step(slab + 1, r1);
        if (slab + 2 >= n_slabs) break;
        if (slab + 4 < n_slabs) load_slab(slab + 4, r1);
        step(slab + 2, r2);
      }
      sg += (uint32_t)n_slabs;
      // ---- epilogue of layer l: TMEM lane = row; 32 x 64-byte blocks transposed through the warp's own shared memory
      mbar_wait(bar_acc, (uint32_t)(l & 1), P.errflag, 42);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 bj = make_float4(0.f, 0.f, 0.f, 0.f);
          if (Lr.bias) bj = __ldg(reinterpret_cast<const float4*>(Lr.bias + c0 + 4 * j));
          float y[4] = {__uint_as_float(v[4 * j]) + bj.x, __uint_as_float(v[4 * j + 1]) + bj.y,
                        __uint_as_float(v[4 * j + 2]) + bj.z, __uint_as_float(v[4 * j + 3]) + bj.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (P.act == 1) y[i] = fmaxf(y[i], 0.f);
            else if (P.act == 2) y[i] = y[i] > 0.f ? y[i] : y[i] * P.slope;
          }
          *reinterpret_cast<float4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = make_float4(y[0], y[1], y[2], y[3]);
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int r = orow + 8 * k;
          const float4 t = *reinterpret_cast<const float4*>(stg + r * 64 + ((och ^ ((r >> 1) & 3)) << 4));
          if (row_base + r < P.M) *reinterpret_cast<float4*>(Lr.c + (row_base + r) * Lr.ldc + c0 + 4 * och) = t;
        }
        __syncwarp();
      }
      tc_fence_before();
      __threadfence_block();       // this warp's rows of layer l are the rows it loads for layer l + 1
      __syncwarp();
    }
  } else if (warp == 4) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc(0, kBM, N);
    uint32_t sg = 0;
    for (int l = 0; l < P.n_layers; ++l) {
      const int n_slabs = (P.layer[l].K1 + P.layer[l].K2) / kSlabK;
      for (int slab = 0; slab < n_slabs; ++slab) {
        const uint32_t s = sg + (uint32_t)slab;
        const int st = (int)(s % kStages);
        const uint32_t par = (s / kStages) & 1u;
        mbar_wait_backoff(bar_afull(st), par, P.errflag, 43, 32);
        mbar_wait_backoff(bar_wfull(st), par, P.errflag, 44, 32);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + st * C::kStageBytes);
          const uint64_t a_hi = make_smem_desc(sa, kALbo, 128);
          const uint64_t a_lo = make_smem_desc(sa + kAHalfBytes, kALbo, 128);
          const uint64_t b_hi = make_smem_desc(sa + 2 * kAHalfBytes, N * 16, 128);
          const uint64_t b_lo = make_smem_desc(sa + 2 * kAHalfBytes + C::kWHalfBytes, N * 16, 128);
#pragma unroll
          for (int k = 0; k < kSlabK / 16; ++k) {
            const uint64_t da = (uint64_t)((k * 2 * kALbo) >> 4);
            const uint64_t db = (uint64_t)((k * 2 * N * 16) >> 4);
            umma_f16(tmem_base, a_lo + da, b_hi + db, idesc, (slab | k) != 0 ? 1u : 0u);      // small terms first
            umma_f16(tmem_base, a_hi + da, b_lo + db, idesc, 1u);
            umma_f16(tmem_base, a_hi + da, b_hi + db, idesc, 1u);
          }
          umma_commit_mc(bar_empty(st), (uint16_t)((1u << kCluster) - 1u));
          if (slab == n_slabs - 1) umma_commit(bar_acc);
        }
        __syncwarp();
      }
      sg += (uint32_t)n_slabs;
    }
  } else {
    // =========================== weight producer ===========================
    if (lane == 0) {
      uint32_t sg = 0;
      for (int l = 0; l < P.n_layers; ++l) {
        const int n_slabs = (P.layer[l].K1 + P.layer[l].K2) / kSlabK;
        for (int slab = 0; slab < n_slabs; ++slab) {
          const uint32_t s = sg + (uint32_t)slab;
          const int st = (int)(s % kStages);
          mbar_wait_backoff(bar_empty(st), ((s / kStages) & 1u) ^ 1u, P.errflag, 45);
          mbar_expect_tx(bar_wfull(st), 2u * C::kWHalfBytes);
          constexpr uint32_t part = 2u * C::kWHalfBytes / kCluster;
          bulk_g2s_mc(smem_u32(smem + st * C::kStageBytes + 2 * kAHalfBytes) + crank * part,
                      reinterpret_cast<const uint8_t*>(P.layer[l].w) + (size_t)slab * 2 * C::kWHalfBytes + (size_t)crank * part,
                      part, bar_wfull(st), (uint16_t)((1u << kCluster) - 1u));
        }
        sg += (uint32_t)n_slabs;
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // no CTA may exit while a peer can still multicast into its shared memory
  tc_fence_after();
  if (warp == 4) tmem_dealloc(tmem_base, C::kTmemCols);
}

template <int N>
static int launch_n(const DenseParams& P, cudaStream_t st) {
  MRINR_SMEM_OPT_IN((dense_split_kernel<N>), Cfg<N>::kSmemBytes);
  long long grid = (P.M + kBM - 1) / kBM;
  grid = (grid + kCluster - 1) / kCluster * kCluster;      // whole clusters; the extra CTAs only help with the weights
  dense_split_kernel<N><<<(unsigned)grid, kThreads, Cfg<N>::kSmemBytes, st>>>(P);
  count_launch();
  return check_launch("dense_split");
}

}  // namespace dense

bool dense_split_supported(int N, int K1, int K2) {
  return (N == 64 || N == 128 || N == 256) && K1 > 0 && K1 % dense::kSlabK == 0 && K2 >= 0 && K2 % dense::kSlabK == 0;
}

int run_pack_split(const float* w, int N, int K, uint16_t* out, cudaStream_t st) {
  const long long n = (long long)N * K;
  dense::pack_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, N, K, out);
  count_launch();
  return check_launch("pack_split");
}

// Chain of n layers with N = 256 outputs each in ONE launch: layer l reads a1[l] ([M,K1[l]], may be the output c[l-1] of
// the previous layer) and the optional a2[l].  Returns MRINR_E_UNSUPPORTED for shapes the chain kernel does not take.
int launch_dense_chain256(int n, const float* const* a1, const long long* lda1, const int* K1, const float* const* a2,
                          const long long* lda2, const int* K2, const uint16_t* const* w_packed,
                          const float* const* bias, float* const* c, const long long* ldc, int act, float slope,
                          long long M, int32_t* errflag, cudaStream_t st) {
  MRINR_REQUIRE(n >= 1 && n <= dense::kMaxChain, MRINR_E_UNSUPPORTED, "dense_chain: 1..%d layers (got %d)", dense::kMaxChain, n);
  if (M <= 0) return 0;
  dense::ChainParams P;
  for (int l = 0; l < n; ++l) {
    MRINR_REQUIRE(dense_split_supported(256, K1[l], K2[l]), MRINR_E_UNSUPPORTED, "dense_chain: unsupported K (%d + %d)", K1[l], K2[l]);
    MRINR_REQUIRE(aligned16(a1[l]) && (K2[l] == 0 || aligned16(a2[l])) && aligned16(c[l]) && lda1[l] % 4 == 0 &&
                      (K2[l] == 0 || lda2[l] % 4 == 0) && ldc[l] % 4 == 0 && (!bias[l] || aligned16(bias[l])),
                  MRINR_E_ALIGN, "dense_chain: operands must be 16-byte aligned with row strides that are multiples of 4");
    dense::ChainLayer& Lr = P.layer[l];
    Lr.a1 = a1[l]; Lr.lda1 = lda1[l]; Lr.K1 = K1[l];
    Lr.a2 = K2[l] ? a2[l] : nullptr; Lr.lda2 = K2[l] ? lda2[l] : 0; Lr.K2 = K2[l];
    Lr.w = w_packed[l]; Lr.bias = bias[l]; Lr.c = c[l]; Lr.ldc = ldc[l];
  }
  P.n_layers = n; P.M = M; P.act = act; P.slope = slope; P.errflag = errflag;
  MRINR_SMEM_OPT_IN((dense::dense_chain_kernel<256>), dense::ChainCfg<256>::kSmemBytes);
  long long grid = (M + dense::kBM - 1) / dense::kBM;
  grid = (grid + dense::kCluster - 1) / dense::kCluster * dense::kCluster;
  dense::dense_chain_kernel<256><<<(unsigned)grid, dense::kThreads, dense::ChainCfg<256>::kSmemBytes, st>>>(P);
  count_launch();
  return check_launch("dense_chain");
}

// act: 0 none, 1 relu, 2 leaky relu (slope)
int launch_dense_split(const float* a1, long long lda1, int K1, const float* a2, long long lda2, int K2,
                       const uint16_t* w_packed, const float* bias, int N, int act, float slope, float* c,
                       long long ldc, long long M, int32_t* errflag, cudaStream_t st) {
  MRINR_REQUIRE(dense_split_supported(N, K1, K2), MRINR_E_UNSUPPORTED,
                "dense_split: unsupported shape (N=%d K1=%d K2=%d): N in {64,128,256}, K multiples of 64", N, K1, K2);
  MRINR_REQUIRE(aligned16(a1) && (K2 == 0 || aligned16(a2)) && aligned16(c) && lda1 % 4 == 0 && lda2 % 4 == 0 &&
                    ldc % 4 == 0,
                MRINR_E_ALIGN, "dense_split: operands must be 16-byte aligned with row strides that are multiples of 4");
  if (M <= 0) return 0;
  dense::DenseParams P;
  P.a1 = a1; P.lda1 = lda1; P.K1 = K1; P.a2 = a2; P.lda2 = lda2; P.K2 = K2; P.w = w_packed; P.bias = bias;
  P.c = c; P.ldc = ldc; P.M = M; P.act = act; P.slope = slope; P.errflag = errflag;
  if (N == 64) return dense::launch_n<64>(P, st);
  if (N == 128) return dense::launch_n<128>(P, st);
  return dense::launch_n<256>(P, st);
}

}  // namespace mrinr
