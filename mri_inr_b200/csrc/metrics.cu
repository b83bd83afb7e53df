// Image-quality metrics of the evaluation loop on the device: src/util/error.py:23-84 as called by metrics_error
// (:256-269) -- PSNR, SSIM and NRMSE of (fully sampled, reconstructed) image pairs with the reference's data-range
// rule (max over both images - min over both images, :23-38).  The reference copies three images per slice to the
// host and calls scikit-image; here N pairs are reduced in three launches and 24 bytes per slice come back.
//
// Definitions (scikit-image, the defaults error.py uses):
//   PSNR  = 10 log10(R^2 / mean((o-p)^2))
//   NRMSE = sqrt(mean((o-p)^2)) / sqrt(mean(o^2))                       (euclidean normalisation)
//   SSIM  : 7x7 uniform window, K1 = 0.01, K2 = 0.03, sample covariance (x 49/48), mean over the image cropped by 3
//           pixels -- so only windows that lie completely inside the image contribute and no border rule is needed.
// Pass 1 reads both images once (8 B per pixel: range, squared error, energy, means), pass 2 reads them once more:
// a row-streaming SSIM kernel (running 7-row sums in registers, 7-column sums across lanes) for W % 4 == 0, a tiled
// shared-memory kernel (38x38 tiles, 1.41x halo) for other shapes; window sums in fp32, sums over windows in fp64.
#include "common.cuh"

namespace mrinr {
namespace metrics {

struct Acc {                 // per image pair, in the caller's scratch buffer
  unsigned int mn, mx;       // order-preserving encodings of the min / max over both images
  double sse, so2, ssim_sum;
  double so, sp;             // sums of both images (their means centre the SSIM window statistics)
};

__device__ __forceinline__ unsigned int f2ord(float f) {
  const unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void init_kernel(Acc* acc, long long N) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) { acc[i].mn = 0xffffffffu; acc[i].mx = 0u; acc[i].sse = 0.0; acc[i].so2 = 0.0; acc[i].ssim_sum = 0.0; acc[i].so = 0.0; acc[i].sp = 0.0; }
}

__global__ void __launch_bounds__(256)
reduce_kernel(const float* __restrict__ orig, const float* __restrict__ pred, long long n, Acc* __restrict__ acc) {
  const long long g = blockIdx.y;
  const float* o = orig + g * n;
  const float* p = pred + g * n;
  float mn = INFINITY, mx = -INFINITY;
  double sse = 0.0, so2 = 0.0, so = 0.0, sp = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if ((n & 3) == 0 && aligned16(o) && aligned16(p)) {
    // 16-byte loads, kU per image in flight per thread (all issued before the first use); fp32 partial sums over
    // 4 kU values, fp64 across them.  A thread makes several trips: the block reduction and the six atomics at the
    // end are paid once per ~100 loads, not once per 4.
    constexpr int kU = 4;
    const float4* o4 = reinterpret_cast<const float4*>(o);
    const float4* p4 = reinterpret_cast<const float4*>(p);
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += kU * stride) {
      float4 a[kU], b[kU];
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        const long long j = i + k * stride < n4 ? i + k * stride : i;        // the tail repeats element i: skipped below
        a[k] = __ldg(o4 + j);
        b[k] = __ldg(p4 + j);
      }
      float e = 0.f, q = 0.f, sa = 0.f, sb = 0.f;
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        if (k > 0 && i + k * stride >= n4) break;
        const float4 x = a[k], y = b[k];
        mn = fminf(mn, fminf(fminf(fminf(x.x, x.y), fminf(x.z, x.w)), fminf(fminf(y.x, y.y), fminf(y.z, y.w))));
        mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)), fmaxf(fmaxf(y.x, y.y), fmaxf(y.z, y.w))));
        const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        e = fmaf(d0, d0, e); e = fmaf(d1, d1, e); e = fmaf(d2, d2, e); e = fmaf(d3, d3, e);
        q = fmaf(x.x, x.x, q); q = fmaf(x.y, x.y, q); q = fmaf(x.z, x.z, q); q = fmaf(x.w, x.w, q);
        sa += (x.x + x.y) + (x.z + x.w);
        sb += (y.x + y.y) + (y.z + y.w);
      }
      sse += (double)e; so2 += (double)q; so += (double)sa; sp += (double)sb;
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float a = __ldg(o + i), b = __ldg(p + i);
      mn = fminf(mn, fminf(a, b));
      mx = fmaxf(mx, fmaxf(a, b));
      const float d = a - b;
      sse += (double)d * (double)d;
      so2 += (double)a * (double)a;
      so += (double)a;
      sp += (double)b;
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, s));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    sse += __shfl_xor_sync(0xffffffffu, sse, s);
    so2 += __shfl_xor_sync(0xffffffffu, so2, s);
    so += __shfl_xor_sync(0xffffffffu, so, s);
    sp += __shfl_xor_sync(0xffffffffu, sp, s);
  }
  __shared__ float smn[8], smx[8];
  __shared__ double sse_s[8], so2_s[8], so_s[8], sp_s[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { smn[w] = mn; smx[w] = mx; sse_s[w] = sse; so2_s[w] = so2; so_s[w] = so; sp_s[w] = sp; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      mn = fminf(mn, smn[k]); mx = fmaxf(mx, smx[k]); sse += sse_s[k]; so2 += so2_s[k]; so += so_s[k]; sp += sp_s[k];
    }
    atomicMin(&acc[g].mn, f2ord(mn));
    atomicMax(&acc[g].mx, f2ord(mx));
    atomicAdd(&acc[g].sse, sse);
    atomicAdd(&acc[g].so2, so2);
    atomicAdd(&acc[g].so, so);
    atomicAdd(&acc[g].sp, sp);
  }
}

constexpr int kT = 32;            // output tile
constexpr int kWin = 7;
constexpr int kIn = kT + kWin - 1; // 38

// one CTA: 32x32 window positions (top-left corners) of one image pair
__global__ void __launch_bounds__(256)
ssim_kernel(const float* __restrict__ orig, const float* __restrict__ pred, int H, int W, Acc* __restrict__ acc) {
  __shared__ float sx[kIn][kIn + 1], sy[kIn][kIn + 1];
  __shared__ float hs[5][kIn][kT + 1];
  __shared__ double red[8];
  const long long g = blockIdx.z;
  const float* o = orig + g * (long long)H * W;
  const float* p = pred + g * (long long)H * W;
  const int x0 = blockIdx.x * kT, y0 = blockIdx.y * kT;
  const int tid = threadIdx.x;
  for (int i = tid; i < kIn * kIn; i += 256) {
    const int r = i / kIn, c = i - r * kIn;
    const int y = y0 + r, x = x0 + c;
    const bool in = (y < H) && (x < W);
    sx[r][c] = in ? __ldg(o + (long long)y * W + x) : 0.f;
    sy[r][c] = in ? __ldg(p + (long long)y * W + x) : 0.f;
  }
  __syncthreads();
  // Shift both tiles by their top-left pixel before squaring: variances and the covariance do not change, but the
  // fp32 products are formed from numbers of the size of the local contrast instead of the absolute intensity, so
  // E[x^2] - E[x]^2 loses far fewer bits (scikit-image computes this in the input precision as well).
  const float x00 = sx[0][0], y00 = sy[0][0];
  __syncthreads();
  for (int i = tid; i < kIn * kIn; i += 256) {
    const int r = i / kIn, c = i - r * kIn;
    sx[r][c] -= x00;
    sy[r][c] -= y00;
  }
  __syncthreads();
  // horizontal 7-tap sums of x, y, xx, yy, xy
  for (int i = tid; i < kIn * kT; i += 256) {
    const int r = i / kT, c = i - r * kT;
    float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const float u = sx[r][c + k], v = sy[r][c + k];
      a += u; b += v; aa = fmaf(u, u, aa); bb = fmaf(v, v, bb); ab = fmaf(u, v, ab);
    }
    hs[0][r][c] = a; hs[1][r][c] = b; hs[2][r][c] = aa; hs[3][r][c] = bb; hs[4][r][c] = ab;
  }
  __syncthreads();
  const float R = ord2f(acc[g].mx) - ord2f(acc[g].mn);        // data range (pass 1 has completed: stream order)
  const float c1 = (0.01f * R) * (0.01f * R), c2 = (0.03f * R) * (0.03f * R);
  const float inv = 1.0f / 49.0f, cov_norm = 49.0f / 48.0f;
  double local = 0.0;
  const int nx = W - kWin + 1, ny = H - kWin + 1;            // number of complete windows
  for (int i = tid; i < kT * kT; i += 256) {
    const int r = i / kT, c = i - r * kT;
    if (y0 + r >= ny || x0 + c >= nx) continue;
    float s[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < kWin; ++k) t += hs[q][r + k][c];
      s[q] = t * inv;
    }
    const float vx = cov_norm * (s[2] - s[0] * s[0]), vy = cov_norm * (s[3] - s[1] * s[1]);
    const float vxy = cov_norm * (s[4] - s[0] * s[1]);
    const float ux = s[0] + x00, uy = s[1] + y00;
    const float num = (2.f * ux * uy + c1) * (2.f * vxy + c2);
    const float den = (ux * ux + uy * uy + c1) * (vx + vy + c2);
    local += (double)(num / den);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
  if ((tid & 31) == 0) red[tid >> 5] = local;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(&acc[g].ssim_sum, t);
  }
}

// SSIM as a row-streaming kernel (W a multiple of 4, 16-byte aligned images).  A CTA walks row segments of image
// pairs, a warp owns 120 window columns of them: lane l carries the four pixel columns 120 cb + 4 l .. + 3 (lanes 30 and
// 31 only supply the 8 halo columns) and walks down the rows keeping the VERTICAL 7-row sums of its columns in
// registers, then combines them HORIZONTALLY with the two neighbouring lanes' prefix sums (4 shuffles per quantity, 11
// additions without cancellation for the four 7-column windows).  No shared memory, no barriers; every loop bound is
// CTA-uniform (blockIdx only), so the shuffles sit in provably convergent code (the first version derived the row
// range from the warp index: ptxas wrapped all 24 shuffles of an iteration in WARPSYNC / collective bookkeeping).
//
// Arithmetic, per pixel column, with u = x - mean(x), v = y - mean(y) (image means from pass 1: E[x^2] - E[x]^2 is
// then formed from numbers of the size of the contrast) and "n" / "o" the rows entering / leaving the window:
//   d = n - o, s = u_n + u_o = (n + o) - 2 mean:   sum u += dx,   sum (u^2 + v^2) += dx sx + dy sy,
//                                                   sum 2uv += dx sy + sx dy         (u_n v_n - u_o v_o, twice)
// -- 12 operations per column PAIR as packed fp32x2 (a float4 load is two pairs), no separate "subtract the old row"
// pass: during the first 7 rows of a segment the leaving row is the constant (mean x, mean y), i.e. u = v = 0.
// Window formula with the divisions by 49 / 48 cancelled between numerator and denominator (S = 7x7 sums):
//   Ux = Sx + 49 mean x, C1 = 49^2 c1, C2 = 48 c2:
//   ssim = (2 Ux Uy + C1)(S2uv - (2/49) Sx Sy + C2) / ((Ux^2 + Uy^2 + C1)(Suu+vv - (Sx^2 + Sy^2)/49 + C2))
// also packed over window pairs.  ~55 instructions per window (first streaming version: 100; tiled kernel: 170).
constexpr int kColsPerWarp = 120;
constexpr int kStreamMaxWarps = 4;
constexpr int kRowsPerSeg = 32;

__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

__global__ void __launch_bounds__(32 * kStreamMaxWarps, 6)
ssim_stream_kernel(const float* __restrict__ orig, const float* __restrict__ pred, int H, int W, int n_colblocks,
                   int cb_groups, long long N, Acc* __restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int nx = W - kWin + 1, ny = H - kWin + 1;
  // Work units = (column-block group, image pair, segment of kRowsPerSeg window rows); gridDim.x persistent CTAs walk
  // equal contiguous ranges of them (CTA-uniform bounds).  The segmentation of an image is FIXED -- it depends neither
  // on the batch size nor on the grid -- so an image's result does not depend on the batch it is evaluated in; ten
  // segments per 320-row image keep the ranges balanced to 1 % (one (pair, half image) per CTA ran as two full waves
  // and a nearly empty third) at the price of six warm-up rows (vertical update only, a quarter of a full row) per segment.
  const int n_seg = (ny + kRowsPerSeg - 1) / kRowsPerSeg;
  const long long units = (long long)cb_groups * N * n_seg;
  const long long u_end = units * (blockIdx.x + 1) / gridDim.x;
  for (long long u = units * blockIdx.x / gridDim.x; u < u_end; ++u) {
  const int grp = (int)(u / (N * n_seg));
  const long long ug = u - (long long)grp * N * n_seg;
  const long long g = ug / n_seg;
  const int y_a = (int)(ug - g * n_seg) * kRowsPerSeg;
  const int y_b = min(ny, y_a + kRowsPerSeg);                               // window rows [y_a, y_b) of pair g
  const int cb = grp * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const float* o = orig + g * (long long)H * W;
  const float* p = pred + g * (long long)H * W;
  const int xc = cb * kColsPerWarp + 4 * lane;                               // this lane's first pixel column
  const bool in_cols = cb < n_colblocks && xc + 3 < W;
  const double n_pix = (double)H * (double)W;
  const float mo = (float)(acc[g].so / n_pix), mp = (float)(acc[g].sp / n_pix);
  const float R = ord2f(acc[g].mx) - ord2f(acc[g].mn);        // data range (pass 1 has completed: stream order)
  const float c1 = (0.01f * R) * (0.01f * R), c2 = (0.03f * R) * (0.03f * R);
  const uint64_t MO2 = pk2(2.f * mo, 2.f * mo), MP2 = pk2(2.f * mp, 2.f * mp);
  const uint64_t MO49 = pk2(49.f * mo, 49.f * mo), MP49 = pk2(49.f * mp, 49.f * mp);
  const uint64_t C1 = pk2(2401.f * c1, 2401.f * c1), C2 = pk2(48.f * c2, 48.f * c2);
  const uint64_t NEG2_49 = pk2(-2.f / 49.f, -2.f / 49.f), NEG1_49 = pk2(-1.f / 49.f, -1.f / 49.f);
  // 0 / 1 weights of this lane's four windows (halo lanes, columns past the last complete window)
  float wk[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) wk[k] = (cb < n_colblocks && lane < kColsPerWarp / 4 && xc + k < nx) ? 1.f : 0.f;
  const uint64_t WK[2] = {pk2(wk[0], wk[1]), pk2(wk[2], wk[3])};
  // vertical sums of u, v, u^2 + v^2, 2uv for the column pairs (0,1) and (2,3)
  uint64_t Vx[2] = {0ull, 0ull}, Vy[2] = {0ull, 0ull}, Vq[2] = {0ull, 0ull}, Vc[2] = {0ull, 0ull};
  // Raw rows are fetched ahead of their use (a warp's rows are otherwise one dependent DRAM round trip each): the
  // entering row two iterations ahead into a rotating set of three register slots (the loop is unrolled by three, so
  // the rotation is a renaming, not 16 moves), the leaving row (L1 / L2 resident) one ahead.  Lanes outside the image
  // never load: their slots stay at (mean x, mean y), i.e. centred value 0.
  struct Raw { float4 a, b; };
  const int y_end = y_b + kWin - 1;                  // one past the last input row of the segment
  const Raw outside = {make_float4(mo, mo, mo, mo), make_float4(mp, mp, mp, mp)};
  auto load = [&](Raw& r, int y) {
    if (in_cols && y < y_end) {
      r.a = __ldg(reinterpret_cast<const float4*>(o + (long long)y * W + xc));
      r.b = __ldg(reinterpret_cast<const float4*>(p + (long long)y * W + xc));
    }
  };
  double local = 0.0;
  Raw old = outside;
  auto body = [&](const Raw& cur, int y) {
    {
      const uint64_t A[2] = {pk2(cur.a.x, cur.a.y), pk2(cur.a.z, cur.a.w)}, B[2] = {pk2(cur.b.x, cur.b.y), pk2(cur.b.z, cur.b.w)};
      const uint64_t Ao[2] = {pk2(old.a.x, old.a.y), pk2(old.a.z, old.a.w)}, Bo[2] = {pk2(old.b.x, old.b.y), pk2(old.b.z, old.b.w)};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t dx = sub2(A[h], Ao[h]), sx = sub2(add2(A[h], Ao[h]), MO2);
        const uint64_t dy = sub2(B[h], Bo[h]), sy = sub2(add2(B[h], Bo[h]), MP2);
        Vx[h] = add2(Vx[h], dx);
        Vy[h] = add2(Vy[h], dy);
        Vq[h] = fma2(dy, sy, fma2(dx, sx, Vq[h]));
        Vc[h] = fma2(sx, dy, fma2(dx, sy, Vc[h]));
      }
    }
    if (y + 1 - y_a >= kWin) load(old, y + 1 - kWin);        // the row that leaves at the next iteration
    if (y - y_a < kWin - 1) return;                          // (CTA-uniform) the first complete window ends at row y_a + 6
    uint64_t S[4][2];                                        // window sums, windows (0,1) and (2,3)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint64_t* V = q == 0 ? Vx : q == 1 ? Vy : q == 2 ? Vq : Vc;
      float a0, a1, a2, a3;
      upk2(V[0], a0, a1);
      upk2(V[1], a2, a3);
      // prefix / suffix sums of the lane's own four columns; the neighbours contribute theirs: 4 shuffles, 11 additions
      const float l2 = a0 + a1, r2 = a2 + a3, l3 = l2 + a2, r3 = a1 + r2, t = l2 + r2;
      const float l3n = __shfl_down_sync(0xffffffffu, l3, 1), tn = __shfl_down_sync(0xffffffffu, t, 1);
      const float l1nn = __shfl_down_sync(0xffffffffu, a0, 2), l2nn = __shfl_down_sync(0xffffffffu, l2, 2);
      S[q][0] = pk2(t + l3n, r3 + tn);                       // columns 0..6, 1..7
      S[q][1] = pk2(r2 + (tn + l1nn), a3 + (tn + l2nn));     // columns 2..8, 3..9
    }
    uint64_t ROW = 0ull;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint64_t Sx = S[0][h], Sy = S[1][h];
      const uint64_t Ux = add2(Sx, MO49), Uy = add2(Sy, MP49);
      const uint64_t n1_ = fma2(add2(Ux, Ux), Uy, C1);
      const uint64_t d1_ = fma2(Uy, Uy, fma2(Ux, Ux, C1));
      const uint64_t n2_ = fma2(mul2(Sx, Sy), NEG2_49, add2(S[3][h], C2));
      const uint64_t d2_ = fma2(fma2(Sx, Sx, mul2(Sy, Sy)), NEG1_49, add2(S[2][h], C2));
      float den0, den1, r0, r1;
      upk2(mul2(d1_, d2_), den0, den1);
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(den0));
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(den1));
      ROW = fma2(mul2(mul2(n1_, n2_), pk2(r0, r1)), WK[h], ROW);
    }
    float row0, row1;
    upk2(ROW, row0, row1);
    local += (double)(row0 + row1);
  };
  Raw r0 = outside, r1 = outside, r2 = outside;
  load(r0, y_a);
  load(r1, y_a + 1);
  for (int y = y_a; y < y_end; y += 3) {
    load(r2, y + 2);
    body(r0, y);
    if (y + 1 >= y_end) break;
    load(r0, y + 3);
    body(r1, y + 1);
    if (y + 2 >= y_end) break;
    load(r1, y + 4);
    body(r2, y + 2);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
  if (lane == 0) atomicAdd(&acc[g].ssim_sum, local);
  }  // next unit of this CTA
}

__global__ void finalize_kernel(const Acc* __restrict__ acc, long long N, int H, int W, double* __restrict__ out) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  const double n = (double)H * (double)W;
  const double R = (double)ord2f(acc[g].mx) - (double)ord2f(acc[g].mn);
  const double mse = acc[g].sse / n;
  out[3 * g + 0] = 10.0 * log10(R * R / mse);
  const double windows = (double)(H - kWin + 1) * (double)(W - kWin + 1);
  out[3 * g + 1] = acc[g].ssim_sum / windows;
  out[3 * g + 2] = sqrt(mse) / sqrt(acc[g].so2 / n);
}

}  // namespace metrics
}  // namespace mrinr

using namespace mrinr;

extern "C" int64_t mrinr_image_metrics_scratch_bytes(int64_t N) {
  return N < 0 ? 0 : (int64_t)((size_t)N * sizeof(metrics::Acc));
}

extern "C" int mrinr_image_metrics(const float* d_original, const float* d_predicted, int64_t N, int32_t H, int32_t W,
                                   double* d_out, void* d_scratch, int64_t scratch_bytes, void* stream) {
  if (N == 0) return 0;
  MRINR_REQUIRE(d_original && d_predicted && d_out && d_scratch, MRINR_E_ARG, "mrinr_image_metrics: null pointer");
  MRINR_REQUIRE(N > 0 && H >= 7 && W >= 7, MRINR_E_ARG, "mrinr_image_metrics: needs N > 0 and images of at least 7x7 (got %lld x %d x %d)",
                (long long)N, H, W);
  MRINR_REQUIRE(N <= 65535, MRINR_E_UNSUPPORTED, "mrinr_image_metrics: at most 65535 image pairs per call");
  MRINR_REQUIRE(scratch_bytes >= mrinr_image_metrics_scratch_bytes(N), MRINR_E_ARG,
                "mrinr_image_metrics: scratch must hold mrinr_image_metrics_scratch_bytes(N) bytes");
  MRINR_REQUIRE((reinterpret_cast<uintptr_t>(d_scratch) & 7u) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 7u) == 0,
                MRINR_E_ALIGN, "mrinr_image_metrics: scratch and output must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  metrics::Acc* acc = static_cast<metrics::Acc*>(d_scratch);
  const long long n = (long long)H * W;
  metrics::init_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(acc, N);
  // reduction pass: ~100 pixels of each image per thread (the block reduction and the six atomics are per CTA).  The
  // split of an image over CTAs depends on the image size only, never on N: an image's partial sums are formed the
  // same way in every batch.
  long long bx = (n >> 2) / (256 * 24);
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  metrics::reduce_kernel<<<dim3((unsigned)bx, (unsigned)N), 256, 0, st>>>(d_original, d_predicted, n, acc);
  const int nx = W - metrics::kWin + 1, ny = H - metrics::kWin + 1;
  if (W % 4 == 0 && aligned16(d_original) && aligned16(d_predicted)) {
    // row-streaming kernel: one wave of persistent CTAs (<= 4 column blocks = warps each), every CTA an equal share of
    // the (pair, 32-row segment) units
    const int n_cb = (nx + metrics::kColsPerWarp - 1) / metrics::kColsPerWarp;
    const int wpb = n_cb < metrics::kStreamMaxWarps ? n_cb : metrics::kStreamMaxWarps;
    const int cb_groups = (n_cb + wpb - 1) / wpb;
    int sms = 148, per_sm = 4;
    {
      int dev = 0;
      MRINR_CUDA(cudaGetDevice(&dev));
      MRINR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      MRINR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, metrics::ssim_stream_kernel, 32 * wpb, 0));
      if (per_sm < 1) per_sm = 1;
    }
    const long long units = (long long)cb_groups * N * ((ny + metrics::kRowsPerSeg - 1) / metrics::kRowsPerSeg);
    long long grid = (long long)sms * per_sm;
    if (grid > units) grid = units;
    metrics::ssim_stream_kernel<<<(unsigned)grid, 32 * wpb, 0, st>>>(d_original, d_predicted, H, W, n_cb, cb_groups, N, acc);
  } else {
    dim3 grid((nx + metrics::kT - 1) / metrics::kT, (ny + metrics::kT - 1) / metrics::kT, (unsigned)N);
    metrics::ssim_kernel<<<grid, 256, 0, st>>>(d_original, d_predicted, H, W, acc);
  }
  metrics::finalize_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(acc, N, H, W, d_out);
  count_launch(4);
  return check_launch("image_metrics");
}
