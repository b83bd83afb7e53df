// Patch encoder, convolutional front, tensor-core version: FixedAutoencoder.encoder[0..3]
// (src/networks/encoding/siren_encoder.py:503-507)
//
//   x [B,1,32,32] -> Conv2d(1,16,3,stride 2,pad 1) -> LeakyReLU(0.2) -> Conv2d(16,32,3,stride 2,pad 1) -> LeakyReLU(0.2)
//     -> y [B,32,8,8]
//
// encoder_conv.cu evaluates both convolutions with FFMA and is bound by the second one (590 of the 664 kFLOP per
// patch) at ~64 % of the fp32 peak.  Here the second convolution runs on the tensor core as an implicit GEMM, WITHOUT
// an im2col copy:
//
//   D[(y, p, x), co] = sum over taps (ky, kx) and channels ci of  map_p[ci][2y+ky-1][2x+kx-1] . w2[co][ci][ky][kx]
//
// One MMA tile (M = 128) is a PAIR of patches p = 0, 1 (64 output positions each); one K step (K = 16) is one tap
// with all 16 input channels; N = 32 output channels.  The conv1 map is stored in shared memory in a layout in which
// the A operand of every tap is directly a canonical UMMA K-major (no swizzle) operand:
//   * channel-last, 8 channels = one 16-byte entry (the two 8-channel chunks are the two K chunks of a step: LBO),
//   * columns de-interleaved by parity: a row holds [9 odd-column entries, entry 0 = zero border | 8 even-column
//     entries], so that the 8 output columns x = 0..7 of a tap (source column 2x+kx-1) are 8 CONSECUTIVE entries = one
//     128-byte core matrix, starting at entry 0 (kx = 0), 9 (kx = 1) or 1 (kx = 2),
//   * rows de-interleaved by parity and the two patches interleaved row by row: plane row j of patch p sits at
//     (2j + p) x 272 bytes, so row group g = 2y + p of the tile is at g x 272 bytes (SBO) from the tap's base:
//     odd plane (row 0 = zero border) base row 0 for ky = 0 and row 1 for ky = 2, even plane for ky = 1.
// The zero padding of the convolution is the two border lines, written once.
//
// Precision: as in dense_tc.cu -- every fp32 operand is split into hi = rn16(x), lo = rn16(x - hi) and a product is
// three MMAs (lo.hi + hi.lo + hi.hi, fp32 accumulation in TMEM): ~1e-6 relative, tests hold 1e-5 (operand range as
// stated in dense_tc.cu: conv1 outputs of [0, 1]-normalised patches are O(1)).
//
// Roles per CTA (persistent, one per SM): kGroups warpgroups of 4 warps, each running its own pipeline over pairs
//   stage 2 patches -> conv1 (FFMA2, 2 warps per patch) -> [epilogue of the group's previous pair: tcgen05.ld, bias,
//   LeakyReLU, store] -> bias + LeakyReLU + split + write the map -> signal the MMA warp
// and one MMA warp that issues the 27 MMAs of a pair as soon as its map is complete.  A group's MMAs run while the
// group stages and convolves its next pair, and while the other group works.
#include "tc_ptx.cuh"

namespace mrinr {
namespace enctc {

#ifndef MRINR_ENC_GROUPS
#define MRINR_ENC_GROUPS 3
#endif
constexpr int kGroups = MRINR_ENC_GROUPS;
constexpr int kGroupThreads = 128;
constexpr int kWarpMma = kGroups * 4;
constexpr int kThreads = kGroups * kGroupThreads + 32;
// input tile: image row r-1 / column c-4 at [r * kInLd + skew(r) + c]; row 0 and column index 3 are the zero border.
// Column offset 4 keeps both the staging stores and conv1's loads (index 8q+3 | 8q+4..7 | 8q+8..11) 16-byte
// aligned; the 4-float skew of every other row pair puts the two rows of a quarter-warp on disjoint banks.
constexpr int kInLd = 40;
constexpr int kInSz = 33 * kInLd;
__device__ __forceinline__ int in_row(int r) { return r * kInLd + ((r >> 1) & 1) * 4; }
constexpr float kSlope = 0.2f;

constexpr int kRowBytes = 17 * 16;                     // 9 odd-column entries + 8 even-column entries
constexpr int kEvenRegion = 16 * kRowBytes;            // even plane: 8 rows x 2 patches
constexpr int kOddOff = kEvenRegion + 16;              // 16-byte skew: even / odd rows of a store hit different banks
constexpr int kOddRegion = 18 * kRowBytes;             // odd plane: (border + 8 rows) x 2 patches
constexpr int kChunkBytes = kOddOff + kOddRegion;      // one 8-channel chunk of one precision half
constexpr int kHalfBytes = 2 * kChunkBytes;
constexpr int kMapBytes = 2 * kHalfBytes;              // hi, lo
constexpr int kWTapBytes = 2 * 32 * 16;                // [2 chunks][32 co][8] fp16
constexpr int kWHalfBytes = 9 * kWTapBytes;
static_assert(kChunkBytes % 16 == 0, "descriptor strides are in 16-byte units");

constexpr int kOffMap = 0;
constexpr int kOffW = kOffMap + kGroups * kMapBytes;
constexpr int kOffIn = kOffW + 2 * kWHalfBytes;
constexpr int kOffW1 = kOffIn + kGroups * 2 * 2 * kInSz * 4; // input tiles: [group][buffer][patch]; then [ky][kx][co] f32
constexpr int kOffB1 = kOffW1 + 9 * 16 * 4;
constexpr int kOffB2 = kOffB1 + 16 * 4;
constexpr int kOffBar = kOffB2 + 32 * 4;
constexpr int kOffTmemPtr = kOffBar + 2 * kGroups * 8;
constexpr int kSmemBytes = kOffTmemPtr + 16;
constexpr int kTmemCols = kGroups <= 1 ? 32 : (kGroups == 2 ? 64 : 128);
static_assert(kSmemBytes <= 232448, "shared memory budget");

__device__ __forceinline__ float lrelu(float x) { return fmaxf(x, x * kSlope); }   // slope < 1

// 8 fp32 -> 8 x fp16 hi and 8 x fp16 lo (x ~= hi + lo to 22 bits)
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
encoder_conv_tc_kernel(const float* __restrict__ patches, long long B, const float* __restrict__ w1,
                       const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                       float* __restrict__ out, int32_t* errflag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  float* s_w1 = reinterpret_cast<float*>(smem + kOffW1);
  float* s_b1 = reinterpret_cast<float*>(smem + kOffB1);
  float* s_b2 = reinterpret_cast<float*>(smem + kOffB2);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kOffTmemPtr);
  const uint32_t bar0 = smem_u32(smem + kOffBar);
  auto bar_afull = [&](int g) { return bar0 + 8u * (uint32_t)g; };
  auto bar_acc = [&](int g) { return bar0 + 8u * (uint32_t)(kGroups + g); };

  // ---- once per CTA: zero the maps (their border lines stay zero) and the input tiles, split the conv2 weights ----
  for (int i = tid; i < (kOffW - kOffMap) / 16; i += kThreads) reinterpret_cast<uint4*>(smem + kOffMap)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < kGroups * 4 * kInSz; i += kThreads) reinterpret_cast<float*>(smem + kOffIn)[i] = 0.f;
  for (int i = tid; i < 9 * 16 * 32; i += kThreads) {
    const int co = i & 31, ci = (i >> 5) & 15, tap = i >> 9;
    const float x = w2[(co * 16 + ci) * 9 + tap];                    // Conv2d weight [co][ci][ky][kx]
    const __half hi = __float2half_rn(x);
    const __half lo = __float2half_rn(x - __half2float(hi));
    const int off = tap * kWTapBytes + (ci >> 3) * 512 + co * 16 + (ci & 7) * 2;
    *reinterpret_cast<__half*>(smem + kOffW + off) = hi;
    *reinterpret_cast<__half*>(smem + kOffW + kWHalfBytes + off) = lo;
  }
  for (int i = tid; i < 9 * 16; i += kThreads) s_w1[i] = w1[(i & 15) * 9 + (i >> 4)];
  if (tid < 16) s_b1[tid] = b1[tid];
  if (tid < 32) s_b2[tid] = b2[tid];
  if (tid == 0) {
    for (int g = 0; g < kGroups; ++g) { mbar_init(bar_afull(g), 4); mbar_init(bar_acc(g), 1); }
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc(smem_u32(s_tmem), kTmemCols);
  fence_proxy_async();                 // the zeroed borders and the weights are read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const long long n_pairs = (B + 1) / 2;

  if (warp < kWarpMma) {
    // =========================== worker warpgroups ===========================
    const int grp = warp >> 2, wg = warp & 3, gt = tid & (kGroupThreads - 1);
    float* s_in0 = reinterpret_cast<float*>(smem + kOffIn) + grp * 4 * kInSz;
    uint8_t* map = smem + kOffMap + grp * kMapBytes;
    // conv1 role: patch pp of the pair, rows y (0..15), columns x0..x0+3, all 16 channels
    const int pp = wg >> 1;
    const int cg = (wg & 1) * 32 + lane;
    const int y = cg >> 2, x0 = (cg & 3) * 4;
    // epilogue role: TMEM lane m = 32 wg + lane = 8 (2 ye + pe) + xe
    const int m = wg * 32 + lane;
    const int xe = m & 7, pe = (m >> 3) & 1, ye = m >> 4;
    const uint32_t taddr = tmem_base + (uint32_t)(grp * 32) + ((uint32_t)(wg * 32) << 16);
    uint32_t it = 0;
    long long prev_pair = -1;
    // the pair's 2 x 1024 inputs are fetched one iteration ahead (coalesced 16-byte loads, 4 per thread), so that the
    // global-memory latency hides behind the previous pair's arithmetic; a missing second patch is staged as zeros
    constexpr int kLd = 2 * 1024 / (kGroupThreads * 4);
    float4 nxt[kLd];
    auto fetch = [&](long long pair) {
#pragma unroll
      for (int i = 0; i < kLd; ++i) {
        const int e = (i * kGroupThreads + gt) * 4;
        const long long b = 2 * pair + (e >> 10);
        nxt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < B) nxt[i] = __ldg(reinterpret_cast<const float4*>(patches + b * 1024 + (e & 1023)));
      }
    };
    if (blockIdx.x + (long long)grp * gridDim.x < n_pairs) fetch(blockIdx.x + (long long)grp * gridDim.x);
    for (long long k = grp;; k += kGroups) {
      const long long pair = blockIdx.x + k * (long long)gridDim.x;
      const bool active = pair < n_pairs;
      float2 acc[4][8];
      // input tiles are double-buffered: the barrier after the staging of iteration i also orders every warp's conv1
      // reads of iteration i-1 before the staging of iteration i+1 (same buffer as i-1)
      float* s_in = s_in0 + (it & 1u) * 2 * kInSz;
      if (active) {
        // ---- stage the two patches ----
#pragma unroll
        for (int i = 0; i < kLd; ++i) {
          const int e = (i * kGroupThreads + gt) * 4;
          const int sp = e >> 10, r = (e >> 5) & 31, c = e & 31;
          *reinterpret_cast<float4*>(s_in + sp * kInSz + in_row(r + 1) + c + 4) = nxt[i];
        }
        const long long pair_next = pair + (long long)kGroups * gridDim.x;
        if (pair_next < n_pairs) fetch(pair_next);
        named_bar_sync(1 + grp, kGroupThreads);
        // ---- conv1: taps in (ky, kx) order, packed fma.rn.f32x2 over channel pairs ----
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float2 bb = *reinterpret_cast<const float2*>(&s_b1[2 * c]);     // accumulators start at the bias
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j][c] = bb;
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float* row = s_in + pp * kInSz + in_row(2 * y + ky) + 2 * x0 + 3;   // input row 2y+ky-1, col 2x0-1
          const float r0 = row[0];
          const float4 r1 = *reinterpret_cast<const float4*>(row + 1);
          const float4 r2 = *reinterpret_cast<const float4*>(row + 5);
          const float2 v[9] = {make_float2(r0, r0),     make_float2(r1.x, r1.x), make_float2(r1.y, r1.y),
                               make_float2(r1.z, r1.z), make_float2(r1.w, r1.w), make_float2(r2.x, r2.x),
                               make_float2(r2.y, r2.y), make_float2(r2.z, r2.z), make_float2(r2.w, r2.w)};
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            float2 w[8];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const float4 t = *reinterpret_cast<const float4*>(&s_w1[(ky * 3 + kx) * 16 + c4 * 4]);
              w[c4 * 2] = make_float2(t.x, t.y); w[c4 * 2 + 1] = make_float2(t.z, t.w);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int c = 0; c < 8; ++c) acc[j][c] = __ffma2_rn(v[2 * j + kx], w[c], acc[j][c]);
          }
        }
      }
      if (prev_pair >= 0) {
        // ---- epilogue of the group's previous pair: its MMAs ran underneath the staging and conv1 above ----
        mbar_wait(bar_acc(grp), (it - 1u) & 1u, errflag, 31);
        tc_fence_after();
        const long long b = 2 * prev_pair + pe;
        float* dst = out + b * 2048 + ye * 8 + xe;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t d[16];
          tmem_ld16(taddr + (uint32_t)(h * 16), d);
          tmem_ld_wait();
          if (b < B) {
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const float4 bb = *reinterpret_cast<const float4*>(&s_b2[h * 16 + c4 * 4]);
              float* o = dst + (h * 16 + c4 * 4) * 64;
              o[0] = lrelu(__uint_as_float(d[c4 * 4]) + bb.x);
              o[64] = lrelu(__uint_as_float(d[c4 * 4 + 1]) + bb.y);
              o[128] = lrelu(__uint_as_float(d[c4 * 4 + 2]) + bb.z);
              o[192] = lrelu(__uint_as_float(d[c4 * 4 + 3]) + bb.w);
            }
          }
        }
        tc_fence_before();
      }
      if (!active) break;
      // ---- bias + LeakyReLU + split, written in the tap-addressable layout (see the header) ----
      {
        const int j2 = y >> 1;
        const int rowoff = (y & 1) ? kOddOff + ((j2 + 1) * 2 + pp) * kRowBytes : (j2 * 2 + pp) * kRowBytes;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int xx = x0 + j;                                   // x0 is a multiple of 4: parity of xx = parity of j
          const int coloff = (j & 1) ? ((xx >> 1) + 1) * 16 : (9 + (xx >> 1)) * 16;
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            float v[8];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float2 a = acc[j][c2 * 4 + c];
              const float2 sl = __fmul2_rn(a, make_float2(kSlope, kSlope));
              v[2 * c] = fmaxf(a.x, sl.x);
              v[2 * c + 1] = fmaxf(a.y, sl.y);
            }
            uint4 hi, lo;
            split8(v, hi, lo);
            uint8_t* dst = map + c2 * kChunkBytes + rowoff + coloff;
            *reinterpret_cast<uint4*>(dst) = hi;
            *reinterpret_cast<uint4*>(dst + kHalfBytes) = lo;
          }
        }
      }
      fence_proxy_async();             // generic-proxy writes of the map -> visible to the tensor core's reads
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_afull(grp));
      prev_pair = pair;
      ++it;
    }
  } else {
    // =========================== MMA issuer (converged warp, one elected lane) ===========================
    const uint32_t idesc = make_idesc(0, 128, 32);
    const uint32_t map0 = smem_u32(smem + kOffMap), w0 = smem_u32(smem + kOffW);
    for (long long k = 0;; ++k) {
      const long long pair = blockIdx.x + k * (long long)gridDim.x;
      if (pair >= n_pairs) break;
      const int grp = (int)(k % kGroups);
      const uint32_t itg = (uint32_t)(k / kGroups);
      mbar_wait_backoff(bar_afull(grp), itg & 1u, errflag, 32, 32);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t mapg = map0 + (uint32_t)grp * kMapBytes;
        const uint32_t d_tmem = tmem_base + (uint32_t)(grp * 32);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap % 3;
          const uint32_t rowbase = ky == 1 ? 0u : (uint32_t)kOddOff + (ky == 2 ? 2u * kRowBytes : 0u);
          const uint32_t colbase = kx == 0 ? 0u : (kx == 1 ? 9u * 16u : 16u);
          const uint64_t a_hi = make_smem_desc(mapg + rowbase + colbase, kChunkBytes, kRowBytes);
          const uint64_t a_lo = make_smem_desc(mapg + kHalfBytes + rowbase + colbase, kChunkBytes, kRowBytes);
          const uint64_t b_hi = make_smem_desc(w0 + (uint32_t)tap * kWTapBytes, 512, 128);
          const uint64_t b_lo = make_smem_desc(w0 + kWHalfBytes + (uint32_t)tap * kWTapBytes, 512, 128);
          umma_f16(d_tmem, a_lo, b_hi, idesc, tap != 0 ? 1u : 0u);      // small terms first
          umma_f16(d_tmem, a_hi, b_lo, idesc, 1u);
          umma_f16(d_tmem, a_hi, b_hi, idesc, 1u);
        }
        umma_commit(bar_acc(grp));
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kWarpMma) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace enctc

int launch_encoder_conv_tc(const float* d_patches, long long B, const float* w1, const float* b1, const float* w2,
                           const float* b2, float* d_out, int num_sms, int32_t* errflag, cudaStream_t st) {
  MRINR_SMEM_OPT_IN((enctc::encoder_conv_tc_kernel), enctc::kSmemBytes);
  if (B <= 0) return 0;
  long long grid = (long long)num_sms;
  const long long pairs = (B + 1) / 2;
  if (grid > pairs) grid = pairs;
  enctc::encoder_conv_tc_kernel<<<(unsigned)grid, enctc::kThreads, enctc::kSmemBytes, st>>>(d_patches, B, w1, b1, w2, b2,
                                                                                         d_out, errflag);
  count_launch();
  return check_launch("encoder_conv_tc");
}

}  // namespace mrinr
