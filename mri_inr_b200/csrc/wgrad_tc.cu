// Weight gradient of a 256 x 256 synthesis layer on the tensor cores (training backward, csrc/train.cu):
//
//   dW[n][k] += sum_m dz[m][n] * h[m][k]        dz, h: [M, 256] fp32 row-major, M = B * S * S rows (the reduction)
//
// i.e. D = A B^T with A = dz^T [256 x M] and B = h^T [256 x M]: the reduction index m is the MMA's K.  Both operands are
// stored with their NON-K index contiguous, so the K-major operand tiles are built while staging into shared memory:
// producer thread t owns column t of dz (an A row) and column t of h (a B row); for every 8 consecutive m it issues 8
// loads (consecutive lanes = consecutive columns: coalesced 128-byte rows), splits the 8 values into fp16 hi / lo halves
// and writes one 16-byte UMMA core-matrix entry each -- the [K/8 chunks][rows][8] K-major, no-swizzle layout of
// dense_tc.cu.  Precision as there: three MMAs per product (lo.hi + hi.lo + hi.hi), fp32 accumulation in TMEM.
// The gradients arrive pre-scaled by the power-of-two grad_scale (train.cu), so their fp16 halves are in range.
//
// One CTA per SM walks a contiguous range of 32-row slabs of M (split-K) with the full 256 x 256 result resident in
// TMEM: two M = 128 accumulators x 256 fp32 columns = all 512 columns.  Per slab and CTA: 64 KB of operands staged
// (2-stage ring), 12 MMAs (2 halves x 2 K steps x 3 split terms).  Each CTA writes its partial result to a scratch
// buffer; a second kernel adds the partials to dW in a fixed order (deterministic, no atomics).
//   warps 0-7  producers (thread = column), afterwards the TMEM -> scratch epilogue
//   warp  8    MMA issuer (converged warp + elect.sync), owns the TMEM allocation
#include "tc_ptx.cuh"

namespace mrinr {
namespace wgrad {

constexpr int kH = 256;
constexpr int kSlabM = 32;                       // reduction rows per stage
constexpr int kStages = 2;
constexpr int kThreads = 9 * 32;
constexpr int kAHalf = 128 * kSlabM * 2;         // one 128-row half of A, hi or lo: [4 kc][128][8] fp16 = 8 KB
constexpr int kBPart = 256 * kSlabM * 2;         // B hi or lo: [4 kc][256][8] fp16 = 16 KB
constexpr int kOffAhi = 0, kOffAlo = 2 * kAHalf, kOffBhi = 4 * kAHalf, kOffBlo = 4 * kAHalf + kBPart;
constexpr int kStageBytes = 4 * kAHalf + 2 * kBPart;     // 64 KB
constexpr int kOffBar = kStages * kStageBytes;
constexpr int kSmemBytes = kOffBar + 64 + 16;

__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
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

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const float* __restrict__ dz, const float* __restrict__ h, long long M, float* __restrict__ partial,
                int32_t* errflag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kOffBar + 64);
  const uint32_t bar0 = smem_u32(s_bar);
  auto bar_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (uint32_t)(2 + s); };
  const uint32_t bar_acc = bar0 + 8u * 4u;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 8);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // this CTA's slabs of the reduction: [s0, s1)
  const long long n_slabs = (M + kSlabM - 1) / kSlabM;
  const long long s0 = n_slabs * blockIdx.x / gridDim.x, s1 = n_slabs * (blockIdx.x + 1) / gridDim.x;
  const int my = (int)(s1 - s0);

  if (warp < 8) {
    // =========================== producers ===========================
    const int t = tid;                                 // column of dz (row of A) and of h (row of B)
    const int a_off = (t >> 7) * kAHalf + (t & 127) * 16;
    const int b_off = t * 16;
    for (int i = 0; i < my; ++i) {
      const int st = i % kStages;
      const long long m0 = (s0 + i) * kSlabM;
      float va[kSlabM], vb[kSlabM];
#pragma unroll
      for (int r = 0; r < kSlabM; ++r) {
        const long long m = m0 + r;
        const bool ok = m < M;
        va[r] = ok ? __ldg(dz + m * kH + t) : 0.f;
        vb[r] = ok ? __ldg(h + m * kH + t) : 0.f;
      }
      mbar_wait(bar_empty(st), ((i / kStages) & 1u) ^ 1u, errflag, 31);
      uint8_t* base = smem + st * kStageBytes;
#pragma unroll
      for (int kc = 0; kc < kSlabM / 8; ++kc) {
        float x[8];
        uint4 hi, lo;
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = va[kc * 8 + e];
        split8(x, hi, lo);
        *reinterpret_cast<uint4*>(base + kOffAhi + kc * (128 * 16) + a_off) = hi;
        *reinterpret_cast<uint4*>(base + kOffAlo + kc * (128 * 16) + a_off) = lo;
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = vb[kc * 8 + e];
        split8(x, hi, lo);
        *reinterpret_cast<uint4*>(base + kOffBhi + kc * (256 * 16) + b_off) = hi;
        *reinterpret_cast<uint4*>(base + kOffBlo + kc * (256 * 16) + b_off) = lo;
      }
      fence_proxy_async();
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(bar_full(st));
    }
    // =========================== epilogue: TMEM -> this CTA's partial result ===========================
    const int acc = warp >> 2;                         // accumulator (row half); TMEM lanes 32 (warp % 4) .. + 31
    const int row = acc * 128 + (warp & 3) * 32 + (tid & 31);
    float* dst = partial + ((size_t)blockIdx.x * kH + row) * kH;
    if (my > 0) {
      mbar_wait(bar_acc, 0u, errflag, 32);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)acc * 256u;
#pragma unroll 1
      for (int c0 = 0; c0 < kH; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(dst + c0 + 4 * q) =
              make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                          __uint_as_float(v[4 * q + 3]));
      }
      tc_fence_before();
    } else {
      for (int c0 = 0; c0 < kH; c0 += 4) *reinterpret_cast<float4*>(dst + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    // =========================== MMA issuer ===========================
    const uint32_t idesc = make_idesc(0, 128, kH);
    for (int i = 0; i < my; ++i) {
      const int st = i % kStages;
      mbar_wait_backoff(bar_full(st), (i / kStages) & 1u, errflag, 33, 32);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + st * kStageBytes);
        const uint64_t b_hi = make_smem_desc(sa + kOffBhi, 256 * 16, 128);
        const uint64_t b_lo = make_smem_desc(sa + kOffBlo, 256 * 16, 128);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint64_t a_hi = make_smem_desc(sa + kOffAhi + half * kAHalf, 128 * 16, 128);
          const uint64_t a_lo = make_smem_desc(sa + kOffAlo + half * kAHalf, 128 * 16, 128);
          const uint32_t d = tmem_base + (uint32_t)half * 256u;
#pragma unroll
          for (int k = 0; k < kSlabM / 16; ++k) {
            const uint64_t da = (uint64_t)((k * 2 * 128 * 16) >> 4);
            const uint64_t db = (uint64_t)((k * 2 * 256 * 16) >> 4);
            umma_f16(d, a_lo + da, b_hi + db, idesc, (i | k) != 0 ? 1u : 0u);     // small terms first
            umma_f16(d, a_hi + da, b_lo + db, idesc, 1u);
            umma_f16(d, a_hi + da, b_hi + db, idesc, 1u);
          }
        }
        umma_commit(bar_empty(st));
        if (i == my - 1) umma_commit(bar_acc);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// dW[i] += sum_c partial[c][i]   (fixed order: deterministic)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int n_parts, float* __restrict__ dW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kH * kH) return;
  float s = 0.f;
  for (int c = 0; c < n_parts; ++c) s += partial[(size_t)c * kH * kH + i];
  dW[i] += s;
}

}  // namespace wgrad

int64_t wgrad_tc_scratch_floats(int num_sms) { return (int64_t)num_sms * wgrad::kH * wgrad::kH; }

// dW [256,256] += dz^T h;  scratch: wgrad_tc_scratch_floats(num_sms) floats
int launch_wgrad_tc(const float* dz, const float* h, long long M, float* scratch, float* dW, int num_sms,
                    int32_t* errflag, cudaStream_t st) {
  if (M <= 0 || !dW) return 0;
  using namespace wgrad;
  MRINR_SMEM_OPT_IN(wgrad_tc_kernel, kSmemBytes);
  const long long n_slabs = (M + kSlabM - 1) / kSlabM;
  const int grid = (int)(n_slabs < num_sms ? n_slabs : num_sms);
  wgrad_tc_kernel<<<grid, kThreads, kSmemBytes, st>>>(dz, h, M, scratch, errflag);
  count_launch();
  int rc = check_launch("wgrad_tc");
  if (rc != 0) return rc;
  wgrad_reduce_kernel<<<(kH * kH + 255) / 256, 256, 0, st>>>(scratch, grid, dW);
  count_launch();
  return check_launch("wgrad_reduce");
}

}  // namespace mrinr
