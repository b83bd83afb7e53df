// HBM-bound kernels of the mri-inr hot path: coordinate grid, overlapping-patch extraction with the
// black-patch classifier, overlap (weighted) reassembly, complex magnitude, min-max normalisation,
// and the black-patch compaction that feeds the synthesis kernels.
//
// All of these move each byte once: coalesced 16-byte accesses, grid sized to cover the data with
// 256-thread CTAs (several waves over 148 SMs at production batch sizes).  Roofline: HBM bandwidth;
// algorithmic bytes per slice are listed in DESIGN.md.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace mrinr {

// ------------------------------------------------------------------------------------------------
// grid: src/networks/modulated_siren.py:427-433.  torch.linspace(-1,1,S) on CPU evaluates
// step=(end-start)/(S-1) in fp32 and lin[i] = i < S/2 ? start + step*i : end - step*(S-1-i),
// each with a single rounding (vectorised fmadd).  __fmaf_rn reproduces it bit for bit.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float linspace_pm1(int i, int S, float step) {
  if (S == 1) return -1.0f;
  return (i < S / 2) ? __fmaf_rn(step, (float)i, -1.0f) : __fmaf_rn(-step, (float)(S - 1 - i), 1.0f);
}

__global__ void make_grid_kernel(int S, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= S * S) return;
  const float step = __fdiv_rn(2.0f, (float)(S - 1));
  float2 g;
  g.x = linspace_pm1(c / S, S, step);
  g.y = linspace_pm1(c % S, S, step);
  reinterpret_cast<float2*>(out)[c] = g;
}

// ------------------------------------------------------------------------------------------------
// image_to_patches (tiling.py:10-64) fused with classify_patches (tiling.py:184-198).
// One warp per output patch; the O*O/4 float4 of a patch are written fully coalesced, the source
// pixels are gathered through the reflect map (scalar read-only loads: each input pixel is read
// (O/I)^2 = 4 times, from L1/L2 after the first).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_idx(int t, int n) {
  t = t < 0 ? -t : t;
  return t >= n ? 2 * (n - 1) - t : t;
}

// Persistent form: a warp walks patches warp, warp + n_warps, ... (a few thousand CTAs' worth of launch and drain cost
// was the difference between 0.73 and the copy roofline for 52 800 short-lived CTAs); for the baseline geometry
// (O = 32: 8 float4 per lane) all 8 loads of a patch are issued before the first store.
template <int NQ>      // float4 per lane and patch (O*O/128) when known at compile time, 0 = generic
__global__ void __launch_bounds__(256)
image_to_patches_kernel(const float* __restrict__ img, long long n_patches, int H, int W, int O, int I,
                        int nV, int nH, float* __restrict__ patches, uint8_t* __restrict__ black) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int per_img = nV * nH;
  const int pad = (O - I) / 2;
  const int q_per_row = O / 4;
  const int total_q = O * q_per_row;
  // x0 + s is a multiple of 4 when I and pad are; rows are 16-byte aligned when W is a multiple of 4
  const bool vec_ok = ((I | pad | W) & 3) == 0 && aligned16(img);
  for (long long warp = warp0; warp < n_patches; warp += n_warps) {
    const long long n = warp / per_img;
    const int p = (int)(warp - n * per_img);
    const int py = p / nH, px = p - py * nH;
    const float* src = img + n * (long long)H * W;
    float4* dst = reinterpret_cast<float4*>(patches + warp * (long long)O * O);
    const int y0 = I * py - pad, x0 = I * px - pad;
    auto load_q = [&](int q) -> float4 {
      const int r = q / q_per_row;
      const int s = (q - r * q_per_row) * 4;
      const float* row = src + (long long)reflect_idx(y0 + r, H) * W;
      float4 v;
      const int xs = x0 + s;
      if (vec_ok && xs >= 0 && xs + 3 < W) {
        v = __ldg(reinterpret_cast<const float4*>(row + xs));      // interior: one aligned 16-byte load
      } else {
        v.x = __ldg(row + reflect_idx(xs + 0, W));
        v.y = __ldg(row + reflect_idx(xs + 1, W));
        v.z = __ldg(row + reflect_idx(xs + 2, W));
        v.w = __ldg(row + reflect_idx(xs + 3, W));
      }
      return v;
    };
    float sum = 0.f;
    if (NQ > 0) {
      float4 v[NQ > 0 ? NQ : 1];
#pragma unroll
      for (int k = 0; k < NQ; ++k) v[k] = load_q(lane + 32 * k);
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        __stcs(dst + lane + 32 * k, v[k]);            // written once, read once by the encoder: streaming store
        sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
      }
    } else {
#pragma unroll 4
      for (int q = lane; q < total_q; q += 32) {
        const float4 v = load_q(q);
        dst[q] = v;
        sum += (v.x + v.y) + (v.z + v.w);
      }
    }
    if (black != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      // tile.mean() < 1e-10 (tiling.py:194-195); the summation order differs from torch's, which can
      // only matter for a patch whose sum is within rounding of O*O*1e-10.
      if (lane == 0) black[warp] = (sum / (float)(O * O) < 1e-10f) ? 1 : 0;
    }
  }
}

// classify_patches (tiling.py:184-198) on already extracted patches: one warp per patch.
__global__ void __launch_bounds__(256)
classify_patches_kernel(const float* __restrict__ patches, long long n_patches, int elems,
                        uint8_t* __restrict__ black) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= n_patches) return;
  const float* src = patches + warp * (long long)elems;
  float sum = 0.f;
  if ((elems & 3) == 0 && aligned16(src)) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int q = lane; q < elems / 4; q += 32) {
      const float4 v = __ldg(s4 + q);
      sum += (v.x + v.y) + (v.z + v.w);
    }
  } else {
    for (int q = lane; q < elems; q += 32) sum += __ldg(src + q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) black[warp] = (sum / (float)elems < 1e-10f) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// patches_to_image_weighted_average (tiling.py:91-140) / patches_to_image (:143-181), gather form.
// Output pixel (y,x) of image n receives tile[p][y+padq-I*py][x+padq-I*px] * w[..] from every patch
// whose window covers it.  F.fold on CPU (col2im) adds contributions in increasing kernel-offset
// order (ky,kx), i.e. decreasing py then decreasing px; the same order is used here so that the sums
// -- and the final division -- are bit-identical to the fp32 CPU reference.
// One thread per output pixel; consecutive threads read consecutive tile columns.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
patches_to_image_kernel(const float* __restrict__ tiles, const float* __restrict__ weights,
                        const uint8_t* __restrict__ black, long long n_pix, int nV, int nH, int K, int I,
                        float* __restrict__ out) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_pix) return;
  const int OH = nV * I, OW = nH * I;
  const long long n = gid / ((long long)OH * OW);
  const int rem = (int)(gid - n * (long long)OH * OW);
  const int y = rem / OW, x = rem - y * OW;
  const int padq = (K - I) / 2;
  const int yp = y + padq, xp = x + padq;
  // patches py with 0 <= yp - I*py <= K-1
  const int py_hi = min(yp / I, nV - 1);
  const int py_lo = (yp - K + 1 <= 0) ? 0 : (yp - K + I) / I;   // ceil((yp-K+1)/I)
  const int px_hi = min(xp / I, nH - 1);
  const int px_lo = (xp - K + 1 <= 0) ? 0 : (xp - K + I) / I;
  float acc = 0.f, norm = 0.f;
  const long long pbase = n * (long long)nV * nH;
  // increasing ky == decreasing py
  for (int py = py_hi; py >= py_lo; --py) {
    const int ky = yp - I * py;
    for (int px = px_hi; px >= px_lo; --px) {
      const int kx = xp - I * px;
      const long long p = pbase + (long long)py * nH + px;
      const float w = weights ? __ldg(weights + ky * K + kx) : 1.0f;
      float t = 0.f;
      if (black == nullptr || black[p] == 0) t = __ldg(tiles + p * (long long)K * K + ky * K + kx);
      acc = __fadd_rn(acc, __fmul_rn(t, w));
      norm = __fadd_rn(norm, w);
    }
  }
  out[gid] = __fdiv_rn(acc, norm);
}

// Same gather, four adjacent output pixels per thread (one 16-byte store, 16-byte tile and weight loads).
// Valid when I, K and padq = (K-I)/2 are multiples of 4: the set of patches covering a pixel only changes at
// x + padq = 0 or K - I (mod I), both multiples of 4, so an aligned group of four pixels shares its patch set and
// its tile-row offsets are 16-byte aligned.  Per pixel the additions happen in the scalar kernel's order, so the
// result is bit-identical.
__global__ void __launch_bounds__(256)
patches_to_image_vec4_kernel(const float* __restrict__ tiles, const float* __restrict__ weights,
                             const uint8_t* __restrict__ black, long long n_quads, int nV, int nH, int K, int I,
                             float* __restrict__ out) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_quads) return;
  const int OH = nV * I, OW4 = (nH * I) >> 2;
  const int per_img = OH * OW4;
  const long long n = gid / per_img;
  const int rem = (int)(gid - n * per_img);
  const int y = rem / OW4, x = (rem - y * OW4) << 2;
  const int padq = (K - I) / 2;
  const int yp = y + padq, xp = x + padq;
  const int py_hi = min(yp / I, nV - 1);
  const int py_lo = (yp - K + 1 <= 0) ? 0 : (yp - K + I) / I;
  const int px_hi = min(xp / I, nH - 1);
  const int px_lo = (xp + 3 - K + 1 <= 0) ? 0 : (xp + 3 - K + I) / I;     // same for all four pixels of the group
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), norm = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long pbase = n * (long long)nV * nH;
  const int KK = K * K;
  for (int py = py_hi; py >= py_lo; --py) {
    const int ky = yp - I * py;
    for (int px = px_hi; px >= px_lo; --px) {
      const int kx = xp - I * px;
      const long long p = pbase + py * nH + px;
      float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
      if (weights) w = __ldg(reinterpret_cast<const float4*>(weights + ky * K + kx));
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (black == nullptr || black[p] == 0) t = __ldg(reinterpret_cast<const float4*>(tiles + p * KK + ky * K + kx));
      acc.x = __fadd_rn(acc.x, __fmul_rn(t.x, w.x)); norm.x = __fadd_rn(norm.x, w.x);
      acc.y = __fadd_rn(acc.y, __fmul_rn(t.y, w.y)); norm.y = __fadd_rn(norm.y, w.y);
      acc.z = __fadd_rn(acc.z, __fmul_rn(t.z, w.z)); norm.z = __fadd_rn(norm.z, w.z);
      acc.w = __fadd_rn(acc.w, __fmul_rn(t.w, w.w)); norm.w = __fadd_rn(norm.w, w.w);
    }
  }
  reinterpret_cast<float4*>(out)[gid] = make_float4(__fdiv_rn(acc.x, norm.x), __fdiv_rn(acc.y, norm.y),
                                                    __fdiv_rn(acc.z, norm.z), __fdiv_rn(acc.w, norm.w));
}

// The same gather for compile-time (I, K) with K <= 2 I (the reference's 16 / 24 and 16 / 32): at most 2 x 2 patches
// cover a pixel, so all (up to eight) loads of a thread are issued before the first addition instead of one
// dependent DRAM round trip per contribution, the window arithmetic has no run-time divisions, and the image index
// comes from blockIdx.y.  Additions in the scalar kernel's order (py descending, then px descending): bit-identical.
template <int I, int K>
__global__ void __launch_bounds__(256)
patches_to_image_fixed_kernel(const float* __restrict__ tiles, const float* __restrict__ weights,
                              const uint8_t* __restrict__ black, long long N, int nV, int nH, float* __restrict__ out) {
  static_assert(K <= 2 * I && I % 4 == 0 && K % 4 == 0 && ((K - I) / 2) % 4 == 0, "at most two patches per axis");
  constexpr int padq = (K - I) / 2, KK = K * K;
  // the K x K weight window is staged in shared memory once per CTA: four scattered 16-byte reads per thread and
  // image otherwise go through the L1 tag stage next to the tile reads
  __shared__ __align__(16) float s_w[KK];
  if (weights) {
    for (int i = threadIdx.x; i < KK; i += blockDim.x) s_w[i] = weights[i];
    __syncthreads();
  }
  const int OW4 = (nH * I) >> 2;
  const int per_img = nV * I * OW4;
  const int rem = blockIdx.x * blockDim.x + threadIdx.x;
  if (rem >= per_img) return;
  const int y = rem / OW4, x = (rem - y * OW4) << 2;
  const int yp = y + padq, xp = x + padq;
  const int py_hi = min(yp / I, nV - 1);
  const int py_lo = (yp - K + 1 <= 0) ? 0 : (yp - K + I) / I;
  const int px_hi = min(xp / I, nH - 1);
  const int px_lo = (xp + 3 - K + 1 <= 0) ? 0 : (xp + 3 - K + I) / I;
  const bool two_y = py_lo < py_hi, two_x = px_lo < px_hi;
  const int ky0 = yp - I * py_hi, ky1 = yp - I * py_lo, kx0 = xp - I * px_hi, kx1 = xp - I * px_lo;
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
  for (long long n = blockIdx.y; n < N; n += gridDim.y) {
    const long long pb = n * (long long)nV * nH;
    const long long p00 = pb + py_hi * nH + px_hi, p01 = pb + py_hi * nH + px_lo;
    const long long p10 = pb + py_lo * nH + px_hi, p11 = pb + py_lo * nH + px_lo;
    // every load first ...
    float4 w00 = one, w01 = one, w10 = one, w11 = one;
    if (weights) {
      w00 = *reinterpret_cast<const float4*>(s_w + ky0 * K + kx0);
      if (two_x) w01 = *reinterpret_cast<const float4*>(s_w + ky0 * K + kx1);
      if (two_y) w10 = *reinterpret_cast<const float4*>(s_w + ky1 * K + kx0);
      if (two_y && two_x) w11 = *reinterpret_cast<const float4*>(s_w + ky1 * K + kx1);
    }
    bool b00 = false, b01 = false, b10 = false, b11 = false;
    if (black) {
      b00 = black[p00] != 0;
      if (two_x) b01 = black[p01] != 0;
      if (two_y) b10 = black[p10] != 0;
      if (two_y && two_x) b11 = black[p11] != 0;
    }
    float4 t00 = zero, t01 = zero, t10 = zero, t11 = zero;
    if (!b00) t00 = __ldg(reinterpret_cast<const float4*>(tiles + p00 * KK + ky0 * K + kx0));
    if (two_x && !b01) t01 = __ldg(reinterpret_cast<const float4*>(tiles + p01 * KK + ky0 * K + kx1));
    if (two_y && !b10) t10 = __ldg(reinterpret_cast<const float4*>(tiles + p10 * KK + ky1 * K + kx0));
    if (two_y && two_x && !b11) t11 = __ldg(reinterpret_cast<const float4*>(tiles + p11 * KK + ky1 * K + kx1));
    // ... then the additions in fold order
    float4 acc = zero, norm = zero;
    auto add = [&](const float4& t, const float4& w) {
      acc.x = __fadd_rn(acc.x, __fmul_rn(t.x, w.x)); norm.x = __fadd_rn(norm.x, w.x);
      acc.y = __fadd_rn(acc.y, __fmul_rn(t.y, w.y)); norm.y = __fadd_rn(norm.y, w.y);
      acc.z = __fadd_rn(acc.z, __fmul_rn(t.z, w.z)); norm.z = __fadd_rn(norm.z, w.z);
      acc.w = __fadd_rn(acc.w, __fmul_rn(t.w, w.w)); norm.w = __fadd_rn(norm.w, w.w);
    };
    add(t00, w00);
    if (two_x) add(t01, w01);
    if (two_y) {
      add(t10, w10);
      if (two_x) add(t11, w11);
    }
    reinterpret_cast<float4*>(out)[n * per_img + rem] = make_float4(__fdiv_rn(acc.x, norm.x), __fdiv_rn(acc.y, norm.y),
                                                                   __fdiv_rn(acc.z, norm.z), __fdiv_rn(acc.w, norm.w));
  }
}

// Weighted reassembly as a BAND kernel (I = 16, K = 24: the reference's inner / siren patch sizes).  The gather kernels
// above keep their loads in registers, and with ~2.25 contributions per output quad a thread cannot hold enough bytes in
// flight: measured 0.69 of the HBM roofline however the work was cut (profiles/r02_hbm_kernels.txt).  Here the unit of
// work is a band -- the I output rows yp in [I py, I py + I) of patch row py of one image -- and a band needs patch row
// py (rows 0..I-1 of its tiles) and patch row py-1 (rows I..K-1: the overlap zone).  The nH tiles of a patch row are
// contiguous in memory, so a patch row arrives as ONE bulk-async copy (46 KB) in a shared-memory ring; one CTA per SM
// walks a contiguous range of (image, py) items, which makes row py-1 the previous item's stage: every tile is fetched
// exactly once (plus one halo row per CTA), with two rows in flight while a band is computed.  (Cutting the rows into
// per-tile pieces -- 40 copies of 768 / 1536 bytes per band, conflict-free strides -- ran into the copy engine's issue
// rate: ~50 ns per copy and SM, 0.53-0.84 of the roofline depending on how many CTAs shared an SM.)  The last band of
// an image also emits the padq rows below it (rows I..I+padq-1 of its own tiles).
// A thread keeps its positions -- quad column j, rows rr and rr+8 -- for all bands, so the normaliser F.fold(w) of the
// interior bands is computed ONCE (additions in fold order: py descending, then px descending) together with its
// correctly rounded reciprocal.  The quotient acc / norm is then  q0 = acc y;  q1 = q0 + (acc - q0 norm) y;
// q = q1 + (acc - q1 norm) y  with y = RN(1 / norm) and exact FMA residuals: q1 is a faithful quotient, and one more
// residual step from a faithful quotient with a correctly rounded reciprocal gives RN(acc / norm) (Markstein) -- the
// reference's value bit for bit, in 5 instructions instead of the ~12 of a full IEEE division.  The argument needs the
// residuals to be free of underflow / overflow: quads with an element outside [2^-100, 2^100] (other than an all-zero
// quad), the first band of an image (no overlap from above) and the extra rows of the last one take __fdiv_rn.
template <int I, int K>
struct BandGeom {
  static constexpr int D = K - I, padq = D / 2, RA = I + padq;
  static constexpr int kStages = 4;          // rows py-1 and py of the band being computed, two rows in flight
  static constexpr int kPhases = 8;          // row phases: thread (rr, j) owns rows rr, rr + 8, .. of quad column j
  static constexpr int kRowsPerThread = I / kPhases;
  static_assert(I % kPhases == 0 && padq <= kPhases, "row phases");
  static size_t smem_bytes(int nH) { return (size_t)kStages * nH * K * K * 4 + K * K * 4 + 8 * kStages; }
};

__device__ __forceinline__ float div_by_rcp(float a, float b, float y) {
  const float q0 = __fmul_rn(a, y);
  const float q1 = __fmaf_rn(__fmaf_rn(-q0, b, a), y, q0);
  return __fmaf_rn(__fmaf_rn(-q1, b, a), y, q1);
}

template <int I, int K, bool HASB>
__global__ void __launch_bounds__(640, 1)
patches_to_image_band_kernel(const float* __restrict__ tiles, const float* __restrict__ weights,
                             const uint8_t* __restrict__ black, long long N, int nV, int nH, float* __restrict__ out) {
  using G = BandGeom<I, K>;
  static_assert(K <= 2 * I && I % 4 == 0 && K % 4 == 0 && G::padq % 4 == 0, "at most two patches per axis");
  constexpr int D = G::D, padq = G::padq, KK = K * K, S = G::kStages, kRows = G::kRowsPerThread;
  extern __shared__ __align__(128) uint8_t band_smem[];
  float* s_data = reinterpret_cast<float*>(band_smem);
  const int row_words = nH * KK;                       // one patch row: nH tiles, contiguous in global memory too
  float* s_w = s_data + S * row_words;                 // the K x K weight window (with ~all of the SM's L1 carved out as
                                                       // shared memory a global read of it is an L2 round trip every time)
  const uint32_t bar0 = smem_u32(s_w + KK);
  // this CTA's items t = n nV + py: [t0, t1); its loads start one item earlier when the first band needs the row above
  const long long T = N * nV;
  const long long t0 = T * blockIdx.x / gridDim.x, t1 = T * (blockIdx.x + 1) / gridDim.x;
  if (t0 >= t1) return;
  const int i_c0 = (t0 % nV) > 0 ? 1 : 0;              // local index of the first computed item
  const long long tl0 = t0 - i_c0;
  const int n_items = (int)(t1 - tl0);

  const int OW4 = (nH * I) >> 2;
  const int rr = threadIdx.x / OW4, j = threadIdx.x - rr * OW4;      // row phase 0..7, quad column
  const int xp = 4 * j + padq;
  const int px_hi = min(xp / I, nH - 1);
  const int px_lo = (xp + 3 - K + 1 <= 0) ? 0 : (xp + 3 - K + I) / I;
  const bool two_x = px_lo < px_hi;
  const int kx0 = xp - I * px_hi, kx1 = xp - I * px_lo;
  const int o_hi = px_hi * KK + kx0, o_lo = px_lo * KK + kx1;        // + ky K, within a patch row
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bar0 + 8u * s, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < KK; i += blockDim.x) s_w[i] = weights ? __ldg(weights + i) : 1.f;
  __syncthreads();
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  auto weight = [&](int ky, int kx) -> float4 { return *reinterpret_cast<const float4*>(s_w + ky * K + kx); };
  auto fetch = [&](int i) {          // patch row of item tl0 + i -> stage i % S (one thread)
    const uint32_t bar = bar0 + 8u * (uint32_t)(i % S);
    mbar_expect_tx(bar, (uint32_t)row_words * 4u);
    bulk_g2s(smem_u32(s_data + (i % S) * row_words), tiles + (size_t)(tl0 + i) * row_words, (uint32_t)row_words * 4u, bar);
  };
  if (threadIdx.x == 0)
    for (int i = 0; i < i_c0 + 2 && i < n_items; ++i) fetch(i);

  // ---- per-position constants of an interior band (py > 0, rows < I), while the first rows are on their way ----
  float4 nrm[kRows], rcp[kRows];
  bool exact_ok = true;              // the reciprocal path's range argument holds for this thread's normalisers
#pragma unroll
  for (int k = 0; k < kRows; ++k) {
    const int r = rr + G::kPhases * k;
    float4 nm = zero;
    auto addw = [&](const float4& w) {
      nm.x = __fadd_rn(nm.x, w.x); nm.y = __fadd_rn(nm.y, w.y); nm.z = __fadd_rn(nm.z, w.z); nm.w = __fadd_rn(nm.w, w.w);
    };
    addw(weight(r, kx0));
    if (two_x) addw(weight(r, kx1));
    if (r < D) {
      addw(weight(I + r, kx0));
      if (two_x) addw(weight(I + r, kx1));
    }
    nrm[k] = nm;
    rcp[k] = make_float4(__frcp_rn(nm.x), __frcp_rn(nm.y), __frcp_rn(nm.z), __frcp_rn(nm.w));
    const float lo = fminf(fminf(nm.x, nm.y), fminf(nm.z, nm.w)), hi = fmaxf(fmaxf(nm.x, nm.y), fmaxf(nm.z, nm.w));
    exact_ok = exact_ok && lo >= 0x1p-20f && hi <= 0x1p20f;       // (false for NaN as well)
  }

  if (i_c0 == 1) mbar_wait(bar0, 0u, nullptr, 0);       // the halo row (it is nobody's "current" row)
  for (int i = i_c0; i < n_items; ++i) {
    if (threadIdx.x == 0 && i + 2 < n_items) fetch(i + 2);       // into the stage of item i-2: released by the barrier below
    const long long t = tl0 + i;
    const int py = (int)(t % nV);
    const bool first = py == 0, lastb = py == nV - 1;
    bool kA_hi = true, kA_lo = true, kB_hi = true, kB_lo = true;      // keep (= not black)
    if (HASB) {
      const uint8_t* bb = black + t * nH;
      kA_hi = __ldg(bb + px_hi) == 0;
      kA_lo = __ldg(bb + px_lo) == 0;
      if (!first) {
        kB_hi = __ldg(bb - nH + px_hi) == 0;
        kB_lo = __ldg(bb - nH + px_lo) == 0;
      }
    }
    mbar_wait(bar0 + 8u * (uint32_t)(i % S), (uint32_t)(i / S) & 1u, nullptr, 0);
    const float* sA = s_data + (i % S) * row_words;
    const float* sB = s_data + ((i + S - 1) % S) * row_words;
    float4* orow = reinterpret_cast<float4*>(out) + (t * I - padq) * (long long)OW4 + j;      // + r OW4 (r >= padq when t = 0)
    // one output quad: row r of the band; hoisted = index into nrm / rcp, or -1 for the on-the-fly path
    auto emit = [&](int r, bool two_y, int hoisted) {
      float4 acc = zero, nm = zero;
      auto add = [&](const float* sp, bool keep, int ky, int kx) {
        float4 tv = *reinterpret_cast<const float4*>(sp);
        if (HASB && !keep) tv = zero;
        const float4 w = weight(ky, kx);
        acc.x = __fadd_rn(acc.x, __fmul_rn(tv.x, w.x)); acc.y = __fadd_rn(acc.y, __fmul_rn(tv.y, w.y));
        acc.z = __fadd_rn(acc.z, __fmul_rn(tv.z, w.z)); acc.w = __fadd_rn(acc.w, __fmul_rn(tv.w, w.w));
        if (hoisted < 0) {
          nm.x = __fadd_rn(nm.x, w.x); nm.y = __fadd_rn(nm.y, w.y); nm.z = __fadd_rn(nm.z, w.z); nm.w = __fadd_rn(nm.w, w.w);
        }
      };
      add(sA + o_hi + r * K, kA_hi, r, kx0);
      if (two_x) add(sA + o_lo + r * K, kA_lo, r, kx1);
      if (two_y) {
        add(sB + o_hi + (I + r) * K, kB_hi, I + r, kx0);
        if (two_x) add(sB + o_lo + (I + r) * K, kB_lo, I + r, kx1);
      }
      float4 q;
      bool fast = false;
      if (hoisted >= 0) {
        const float ax = fabsf(acc.x), ay = fabsf(acc.y), az = fabsf(acc.z), aw = fabsf(acc.w);
        const float hi = fmaxf(fmaxf(ax, ay), fmaxf(az, aw)), lo = fminf(fminf(ax, ay), fminf(az, aw));
        fast = exact_ok && hi <= 0x1p100f && (lo >= 0x1p-100f || hi == 0.f);
        nm = nrm[hoisted];
      }
      if (fast) {
        const float4 y = rcp[hoisted];
        q = make_float4(div_by_rcp(acc.x, nm.x, y.x), div_by_rcp(acc.y, nm.y, y.y), div_by_rcp(acc.z, nm.z, y.z),
                        div_by_rcp(acc.w, nm.w, y.w));
      } else {
        q = make_float4(__fdiv_rn(acc.x, nm.x), __fdiv_rn(acc.y, nm.y), __fdiv_rn(acc.z, nm.z), __fdiv_rn(acc.w, nm.w));
      }
      orow[(long long)r * OW4] = q;
    };
    if (!first) {
#pragma unroll
      for (int k = 0; k < kRows; ++k) emit(rr + G::kPhases * k, rr + G::kPhases * k < D, k);
    } else {
#pragma unroll 1
      for (int k = 0; k < kRows; ++k)
        if (rr + G::kPhases * k >= padq) emit(rr + G::kPhases * k, false, -1);       // rows above the image do not exist
    }
    if (lastb && rr < padq) emit(I + rr, false, -1);
    __syncthreads();            // every thread is done with rows i-1 and i: the next fetch may overwrite row i-1's stage
  }
}

// ------------------------------------------------------------------------------------------------
// complex magnitude (fastmri.complex_abs at preprocessing.py:58): sqrt(re^2 + im^2), with the two
// squares rounded separately as torch's (data**2).sum(-1).sqrt() does.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cabs1(float re, float im) {
  return __fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
}
__global__ void __launch_bounds__(256)
complex_abs_kernel(const float2* __restrict__ in, long long n, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 v = __ldg(in + i);
  out[i] = cabs1(v.x, v.y);
}
// four complex values per thread and a grid-stride loop (two 16-byte loads, one 16-byte store, several in flight)
__global__ void __launch_bounds__(256)
complex_abs_vec4_kernel(const float4* __restrict__ in, long long n4, float4* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll 2
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = __ldg(in + 2 * i), b = __ldg(in + 2 * i + 1);
    out[i] = make_float4(cabs1(a.x, a.y), cabs1(a.z, a.w), cabs1(b.x, b.y), cabs1(b.z, b.w));
  }
}

// ------------------------------------------------------------------------------------------------
// min-max normalisation (visualization.py:113-126): two passes.  Pass 1 reduces min/max per group
// with order-independent atomics on the float bit patterns; pass 2 applies (x-min)/(max-min).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int f2ord(float f) {   // monotone float -> uint map
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void minmax_init_kernel(unsigned int* scratch, long long G) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G) { scratch[2 * g] = 0xffffffffu; scratch[2 * g + 1] = 0u; }
}

__global__ void __launch_bounds__(256)
minmax_reduce_kernel(const float* __restrict__ in, long long n, unsigned int* __restrict__ scratch) {
  const long long g = blockIdx.y;
  const float* src = in + g * n;
  float mn = INFINITY, mx = -INFINITY;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((n & 3) == 0 && aligned16(src)) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    const long long n4 = n / 4;
    for (; i + 3 * stride < n4; i += 4 * stride) {            // four independent 16-byte loads in flight
      const float4 a = __ldg(s4 + i), b = __ldg(s4 + i + stride), c = __ldg(s4 + i + 2 * stride),
                   d = __ldg(s4 + i + 3 * stride);
      mn = fminf(fminf(fminf(mn, fminf(a.x, a.y)), fminf(fminf(a.z, a.w), fminf(b.x, b.y))),
                 fminf(fminf(fminf(b.z, b.w), fminf(c.x, c.y)), fminf(fminf(c.z, c.w), fminf(fminf(d.x, d.y), fminf(d.z, d.w)))));
      mx = fmaxf(fmaxf(fmaxf(mx, fmaxf(a.x, a.y)), fmaxf(fmaxf(a.z, a.w), fmaxf(b.x, b.y))),
                 fmaxf(fmaxf(fmaxf(b.z, b.w), fmaxf(c.x, c.y)), fmaxf(fmaxf(c.z, c.w), fmaxf(fmaxf(d.x, d.y), fmaxf(d.z, d.w)))));
    }
    for (; i < n4; i += stride) {
      const float4 v = __ldg(s4 + i);
      mn = fminf(fminf(mn, v.x), fminf(fminf(v.y, v.z), v.w));
      mx = fmaxf(fmaxf(mx, v.x), fmaxf(fmaxf(v.y, v.z), v.w));
    }
  } else {
    for (; i < n; i += stride) { const float v = __ldg(src + i); mn = fminf(mn, v); mx = fmaxf(mx, v); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ float smn[8], smx[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { smn[w] = mn; smx[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { mn = fminf(mn, smn[k]); mx = fmaxf(mx, smx[k]); }
    atomicMin(scratch + 2 * g, f2ord(mn));
    atomicMax(scratch + 2 * g + 1, f2ord(mx));
  }
}

__global__ void __launch_bounds__(256)
minmax_apply_kernel(const float* __restrict__ in, long long n, const unsigned int* __restrict__ scratch,
                    float* __restrict__ out) {
  const long long g = blockIdx.y;
  const float mn = ord2f(scratch[2 * g]);
  const float range = __fsub_rn(ord2f(scratch[2 * g + 1]), mn);
  const float* src = in + g * n;
  float* dst = out + g * n;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((n & 3) == 0 && aligned16(src) && aligned16(dst)) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    const long long n4 = n / 4;
    auto norm4 = [&](float4 v) {
      v.x = __fdiv_rn(__fsub_rn(v.x, mn), range);
      v.y = __fdiv_rn(__fsub_rn(v.y, mn), range);
      v.z = __fdiv_rn(__fsub_rn(v.z, mn), range);
      v.w = __fdiv_rn(__fsub_rn(v.w, mn), range);
      return v;
    };
    for (; i + 3 * stride < n4; i += 4 * stride) {            // four independent 16-byte loads in flight
      const float4 a = __ldg(s4 + i), b = __ldg(s4 + i + stride), c = __ldg(s4 + i + 2 * stride),
                   d = __ldg(s4 + i + 3 * stride);
      d4[i] = norm4(a); d4[i + stride] = norm4(b); d4[i + 2 * stride] = norm4(c); d4[i + 3 * stride] = norm4(d);
    }
    for (; i < n4; i += stride) d4[i] = norm4(__ldg(s4 + i));
  } else {
    for (; i < n; i += stride) dst[i] = __fdiv_rn(__fsub_rn(__ldg(src + i), mn), range);
  }
}

// ------------------------------------------------------------------------------------------------
// Black-patch compaction (filter_and_remember_black_patches, tiling.py:244-271, as an index list):
// idx[0..n_active) = ascending indices of the non-black patches; black patches' output rows are
// zero-filled here (reintegrate_black_patches, tiling.py:274-303) so the synthesis kernel never
// touches them.  Three small kernels: per-block counts, single-block exclusive scan, scatter.
// ------------------------------------------------------------------------------------------------
constexpr int kCompactBlock = 1024;

__global__ void __launch_bounds__(kCompactBlock)
compact_count_kernel(const uint8_t* __restrict__ black, long long B, int32_t* __restrict__ blocksums) {
  const long long i = (long long)blockIdx.x * kCompactBlock + threadIdx.x;
  const int keep = (i < B && black[i] == 0) ? 1 : 0;
  const int c = __syncthreads_count(keep);
  if (threadIdx.x == 0) blocksums[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024)
compact_scan_kernel(int32_t* __restrict__ blocksums, int nblocks, int32_t* __restrict__ nactive) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblocks ? blocksums[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
      int t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u;
      }
      warp_tot[lane] = t;   // inclusive over warps
    }
    __syncthreads();
    const int before = carry + (w > 0 ? warp_tot[w - 1] : 0) + incl - v;
    if (i < nblocks) blocksums[i] = before;   // exclusive prefix
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *nactive = carry;
}

__global__ void __launch_bounds__(kCompactBlock)
compact_scatter_kernel(const uint8_t* __restrict__ black, long long B, const int32_t* __restrict__ blocksums,
                       int32_t* __restrict__ idx) {
  __shared__ int32_t warp_tot[32];
  const long long i = (long long)blockIdx.x * kCompactBlock + threadIdx.x;
  const int keep = (i < B && black[i] == 0) ? 1 : 0;
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) warp_tot[w] = __popc(m);
  __syncthreads();
  if (w == 0) {
    int t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += u;
    }
    warp_tot[lane] = t;
  }
  __syncthreads();
  if (keep) {
    const int pos = blocksums[blockIdx.x] + (w > 0 ? warp_tot[w - 1] : 0) + __popc(m & ((1u << lane) - 1u));
    idx[pos] = (int32_t)i;
  }
}

// zero the output rows of black patches: one warp per patch
__global__ void __launch_bounds__(256)
zero_black_rows_kernel(const uint8_t* __restrict__ black, long long B, int C, float* __restrict__ out) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B || black[warp] == 0) return;
  float* dst = out + warp * (long long)C;
  for (int c = lane; c < C; c += 32) dst[c] = 0.f;
}

int launch_compact_black(const uint8_t* d_black, int64_t B, int32_t C, int32_t* d_idx, int32_t* d_nactive,
                         int32_t* d_blocksums, float* d_out, cudaStream_t st) {
  const int nblocks = (int)((B + kCompactBlock - 1) / kCompactBlock);
  compact_count_kernel<<<nblocks, kCompactBlock, 0, st>>>(d_black, B, d_blocksums);
  compact_scan_kernel<<<1, 1024, 0, st>>>(d_blocksums, nblocks, d_nactive);
  compact_scatter_kernel<<<nblocks, kCompactBlock, 0, st>>>(d_black, B, d_blocksums, d_idx);
  const long long threads = B * 32;
  zero_black_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_black, B, C, d_out);
  count_launch(4);
  return check_launch("compact_black");
}

}  // namespace mrinr

using namespace mrinr;

extern "C" int mrinr_make_grid(int32_t S, float* d_out, void* stream) {
  MRINR_REQUIRE(d_out != nullptr, MRINR_E_ARG, "mrinr_make_grid: null output");
  MRINR_REQUIRE(S >= 1 && S <= 4096, MRINR_E_ARG, "mrinr_make_grid: S=%d out of range", S);
  const int n = S * S;
  make_grid_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(S, d_out);
  count_launch();
  return check_launch("make_grid");
}

extern "C" int mrinr_image_to_patches(const float* d_img, int64_t N, int32_t H, int32_t W, int32_t O, int32_t I,
                                      float* d_patches, uint8_t* d_black, void* stream) {
  if (N == 0) return 0;
  MRINR_REQUIRE(d_img && d_patches, MRINR_E_ARG, "mrinr_image_to_patches: null pointer");
  MRINR_REQUIRE(N >= 0 && H > 0 && W > 0 && O > 0 && I > 0, MRINR_E_ARG, "mrinr_image_to_patches: bad sizes");
  MRINR_REQUIRE(O >= I && ((O - I) % 2) == 0 && (O % 4) == 0, MRINR_E_UNSUPPORTED,
                "mrinr_image_to_patches: need O >= I, (O-I) even, O %% 4 == 0 (O=%d I=%d)", O, I);
  const int pad = (O - I) / 2;
  const int vpad = (I - H % I) % I, hpad = (I - W % I) % I;
  // torch reflect padding requires pad < dim on every side (tiling.py:40-44)
  MRINR_REQUIRE(pad + vpad < H && pad + hpad < W, MRINR_E_UNSUPPORTED,
                "mrinr_image_to_patches: reflect padding %d/%d must be smaller than the image %dx%d",
                pad + vpad, pad + hpad, H, W);
  MRINR_REQUIRE(aligned16(d_patches), MRINR_E_ALIGN, "mrinr_image_to_patches: patches not 16-byte aligned");
  if (N == 0) return 0;
  const int nV = (H + vpad) / I, nH = (W + hpad) / I;
  const long long n_patches = (long long)N * nV * nH;
  // persistent grid: 8 warps per CTA, up to 16 CTAs' worth of warps per SM; a warp walks patches with that stride
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) (void)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long ctas = (n_patches + 7) / 8;
  if (ctas > (long long)sms * 16) ctas = (long long)sms * 16;
  if (O == 32)
    image_to_patches_kernel<8><<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(d_img, n_patches, H, W, O, I, nV, nH,
                                                                               d_patches, d_black);
  else
    image_to_patches_kernel<0><<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(d_img, n_patches, H, W, O, I, nV, nH,
                                                                               d_patches, d_black);
  count_launch();
  return check_launch("image_to_patches");
}

extern "C" int mrinr_classify_patches(const float* d_patches, int64_t n_patches, int32_t elems, uint8_t* d_black,
                                      void* stream) {
  if (n_patches == 0) return 0;
  MRINR_REQUIRE(d_patches && d_black && n_patches >= 0 && elems > 0, MRINR_E_ARG, "mrinr_classify_patches: bad arguments");
  const long long threads = (long long)n_patches * 32;
  classify_patches_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_patches, n_patches,
                                                                                             elems, d_black);
  count_launch();
  return check_launch("classify_patches");
}

extern "C" int mrinr_patches_to_image(const float* d_tiles, const float* d_weights, const uint8_t* d_black,
                                      int64_t N, int32_t nV, int32_t nH, int32_t K, int32_t I, float* d_img,
                                      void* stream) {
  if (N == 0) return 0;
  MRINR_REQUIRE(d_tiles && d_img, MRINR_E_ARG, "mrinr_patches_to_image: null pointer");
  MRINR_REQUIRE(N >= 0 && nV > 0 && nH > 0 && K > 0 && I > 0 && K >= I && ((K - I) % 2) == 0, MRINR_E_ARG,
                "mrinr_patches_to_image: bad sizes (K=%d I=%d)", K, I);
  if (N == 0) return 0;
  const long long n_pix = (long long)N * nV * I * nH * I;
  if (I % 4 == 0 && K % 4 == 0 && ((K - I) / 2) % 4 == 0 && aligned16(d_tiles) && aligned16(d_img) &&
      (d_weights == nullptr || aligned16(d_weights))) {
    const long long n_quads = n_pix / 4;
    const long long per_img = n_quads / N;
    if (I == 16 && K == 24 && nH * I <= 320 && BandGeom<16, 24>::smem_bytes(nH) <= 227 * 1024) {
      using G = BandGeom<16, 24>;
      const int smem = (int)G::smem_bytes(nH);
      const cudaStream_t st = (cudaStream_t)stream;
      // persistent: one CTA per SM, each walks a contiguous range of (image, patch row) items
      int sms = 148;
      {
        int dev = 0;
        MRINR_CUDA(cudaGetDevice(&dev));
        MRINR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      }
      const long long items = (long long)N * nV;
      dim3 grid((unsigned)(items < sms ? items : sms));
      const unsigned threads = (unsigned)(G::kPhases * nH * I / 4);          // 8 row phases x (nH I / 4) quads
      if (d_black) {
        MRINR_SMEM_OPT_IN((patches_to_image_band_kernel<16, 24, true>), 227 * 1024);
        patches_to_image_band_kernel<16, 24, true><<<grid, threads, smem, st>>>(d_tiles, d_weights, d_black, N, nV, nH, d_img);
      } else {
        MRINR_SMEM_OPT_IN((patches_to_image_band_kernel<16, 24, false>), 227 * 1024);
        patches_to_image_band_kernel<16, 24, false><<<grid, threads, smem, st>>>(d_tiles, d_weights, d_black, N, nV, nH, d_img);
      }
      count_launch();
      return check_launch("patches_to_image_band");
    }
    if (I == 16 && (K == 24 || K == 32) && per_img < (1ll << 30)) {
      // a thread walks several images (same pixel position: the window arithmetic and the weights are reused)
      // four images per thread.  (Measured in round 2: fewer, longer-lived CTAs -- ~16 per SM, 44 images per thread,
      // which helped image_to_patches and minmax -- make this gather SLOWER: 0.62 instead of 0.69 weighted, 0.89
      // instead of 1.01 unit; it lives on many CTAs' worth of independent loads in flight.)
      long long gy = (N + 3) / 4;
      if (gy > 65535) gy = 65535;
      dim3 grid((unsigned)((per_img + 255) / 256), (unsigned)gy);
      if (K == 24)
        patches_to_image_fixed_kernel<16, 24><<<grid, 256, 0, (cudaStream_t)stream>>>(d_tiles, d_weights, d_black, N, nV,
                                                                                      nH, d_img);
      else
        patches_to_image_fixed_kernel<16, 32><<<grid, 256, 0, (cudaStream_t)stream>>>(d_tiles, d_weights, d_black, N, nV,
                                                                                      nH, d_img);
      count_launch();
      return check_launch("patches_to_image_fixed");
    }
    patches_to_image_vec4_kernel<<<(unsigned)((n_quads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_tiles, d_weights, d_black, n_quads, nV, nH, K, I, d_img);
    count_launch();
    return check_launch("patches_to_image_vec4");
  }
  patches_to_image_kernel<<<(unsigned)((n_pix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_tiles, d_weights, d_black, n_pix, nV, nH, K, I, d_img);
  count_launch();
  return check_launch("patches_to_image");
}

extern "C" int mrinr_complex_abs(const float* d_in, int64_t n, float* d_out, void* stream) {
  if (n == 0) return 0;
  MRINR_REQUIRE(d_in && d_out && n >= 0, MRINR_E_ARG, "mrinr_complex_abs: bad arguments");
  if ((n & 3) == 0 && aligned16(d_in) && aligned16(d_out)) {
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    complex_abs_vec4_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(d_in), n4,
                                                                             reinterpret_cast<float4*>(d_out));
  } else {
    complex_abs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(d_in), n, d_out);
  }
  count_launch();
  return check_launch("complex_abs");
}

extern "C" int mrinr_minmax_normalize(const float* d_in, int64_t G, int64_t n, float* d_out, float* d_scratch,
                                      void* stream) {
  MRINR_REQUIRE(d_in && d_out && d_scratch && G >= 0 && n > 0, MRINR_E_ARG, "mrinr_minmax_normalize: bad arguments");
  MRINR_REQUIRE(G <= 65535, MRINR_E_UNSUPPORTED, "mrinr_minmax_normalize: at most 65535 groups per call");
  if (G == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned int* scr = reinterpret_cast<unsigned int*>(d_scratch);
  minmax_init_kernel<<<(unsigned)((G + 255) / 256), 256, 0, st>>>(scr, G);
  // about 16 CTAs per SM over all groups, at least 4 float4 per thread: 96 volumes x 592 CTAs of one or two loads per
  // thread each (round 1) spent their time in launch / drain and one atomic per CTA (0.68 of the copy roofline)
  long long bx = (n / 4 + 1023) / 1024;
  const long long bx_cap = (148LL * 16 + G - 1) / G;
  if (bx > bx_cap) bx = bx_cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)G);
  minmax_reduce_kernel<<<grid, 256, 0, st>>>(d_in, n, scr);
  minmax_apply_kernel<<<grid, 256, 0, st>>>(d_in, n, scr, d_out);
  count_launch(3);
  return check_launch("minmax_normalize");
}
