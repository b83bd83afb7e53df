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
    // 16-byte loads, two in flight per image; fp32 partial sums over 8 values, fp64 across them
    const float4* o4 = reinterpret_cast<const float4*>(o);
    const float4* p4 = reinterpret_cast<const float4*>(p);
    const long long n4 = n >> 2;
    auto eat = [&](const float4& a, const float4& b, float& e, float& q, float& sa, float& sb) {
      mn = fminf(mn, fminf(fminf(fminf(a.x, a.y), fminf(a.z, a.w)), fminf(fminf(b.x, b.y), fminf(b.z, b.w))));
      mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w))));
      const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
      e = fmaf(d0, d0, e); e = fmaf(d1, d1, e); e = fmaf(d2, d2, e); e = fmaf(d3, d3, e);
      q = fmaf(a.x, a.x, q); q = fmaf(a.y, a.y, q); q = fmaf(a.z, a.z, q); q = fmaf(a.w, a.w, q);
      sa += (a.x + a.y) + (a.z + a.w);
      sb += (b.x + b.y) + (b.z + b.w);
    };
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
      const bool two = i + stride < n4;
      const float4 a0 = __ldg(o4 + i), b0 = __ldg(p4 + i);
      float4 a1 = a0, b1 = b0;
      if (two) { a1 = __ldg(o4 + i + stride); b1 = __ldg(p4 + i + stride); }
      float e = 0.f, q = 0.f, sa = 0.f, sb = 0.f;
      eat(a0, b0, e, q, sa, sb);
      if (two) eat(a1, b1, e, q, sa, sb);
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

// SSIM as a row-streaming kernel (W a multiple of 4, 16-byte aligned images).  A warp owns 120 window columns of a
// row segment of one image pair: lane l carries the four pixel columns 120 b + 4 l .. + 3 (lanes 30 and 31 only supply
// the 8 halo columns) and walks down the rows keeping the VERTICAL 7-row sums of x, y, xx, yy, xy of its columns in
// registers (x, y, xx + yy, xy: four quantities) -- the row entering the window is added, the row leaving it is read again (L1) and subtracted -- and for
// every row combines them HORIZONTALLY with the two neighbouring lanes' sums (6 shuffles per quantity, 13 additions
// without cancellation for the four 7-column windows).  ~60 instructions per window instead of ~170 for the tiled
// kernel below, no shared memory, no barriers.  Both images are centred by their means (pass 1) before squaring, so
// E[x^2] - E[x]^2 is formed from numbers of the size of the contrast; the running sums restart with every row segment.
constexpr int kColsPerWarp = 120;
__global__ void __launch_bounds__(128)
ssim_stream_kernel(const float* __restrict__ orig, const float* __restrict__ pred, int H, int W, int n_colblocks,
                   int n_seg, int rows_per_seg, Acc* __restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const int unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // (segment, column block) of this warp
  if (unit >= n_colblocks * n_seg) return;
  const int seg = unit / n_colblocks, cb = unit - seg * n_colblocks;
  const long long g = blockIdx.y;
  const float* o = orig + g * (long long)H * W;
  const float* p = pred + g * (long long)H * W;
  const int nx = W - kWin + 1, ny = H - kWin + 1;
  const int y_a = seg * rows_per_seg, y_b = min(ny, y_a + rows_per_seg);     // window rows [y_a, y_b)
  const int xc = cb * kColsPerWarp + 4 * lane;                               // this lane's first pixel column
  const bool in_cols = xc + 3 < W;
  const double n_pix = (double)H * (double)W;
  const float mo = (float)(acc[g].so / n_pix), mp = (float)(acc[g].sp / n_pix);
  const float R = ord2f(acc[g].mx) - ord2f(acc[g].mn);        // data range (pass 1 has completed: stream order)
  const float c1 = (0.01f * R) * (0.01f * R), c2 = (0.03f * R) * (0.03f * R);
  const float inv = 1.0f / 49.0f, cov_norm = 49.0f / 48.0f;
  float V[4][4];               // x, y, xx + yy, xy (the two variances only ever appear as their sum)
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int k = 0; k < 4; ++k) V[q][k] = 0.f;
  // raw rows are fetched ahead of their use (a warp's rows are otherwise one dependent DRAM round trip each: with
  // ~20 warps per SM that is ~20 KB in flight, a third of what the HBM needs): the entering row two iterations ahead,
  // the leaving row (L1 / L2 resident) one ahead
  struct Raw { float4 a, b; };
  const int y_end = y_b + kWin - 1;                  // one past the last input row of the segment
  auto fetch = [&](int y) -> Raw {
    Raw r;
    r.a = make_float4(mo, mo, mo, mo);               // outside the image: centred value 0
    r.b = make_float4(mp, mp, mp, mp);
    if (in_cols && y < y_end) {
      r.a = __ldg(reinterpret_cast<const float4*>(o + (long long)y * W + xc));
      r.b = __ldg(reinterpret_cast<const float4*>(p + (long long)y * W + xc));
    }
    return r;
  };
  auto centre = [&](const Raw& r, float (&u)[4], float (&v)[4]) {
    u[0] = r.a.x - mo; u[1] = r.a.y - mo; u[2] = r.a.z - mo; u[3] = r.a.w - mo;
    v[0] = r.b.x - mp; v[1] = r.b.y - mp; v[2] = r.b.z - mp; v[3] = r.b.w - mp;
  };
  double local = 0.0;
  Raw n0 = fetch(y_a), n1 = fetch(y_a + 1), old = fetch(y_a);      // old: first needed at y = y_a + 7
  for (int y = y_a; y < y_end; ++y) {
    const Raw n2 = fetch(y + 2);
    float u[4], v[4];
    centre(n0, u, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      V[0][k] += u[k]; V[1][k] += v[k];
      V[2][k] = fmaf(u[k], u[k], V[2][k]); V[2][k] = fmaf(v[k], v[k], V[2][k]); V[3][k] = fmaf(u[k], v[k], V[3][k]);
    }
    n0 = n1;
    n1 = n2;
    if (y - y_a >= kWin) {
      centre(old, u, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        V[0][k] -= u[k]; V[1][k] -= v[k];
        V[2][k] = fmaf(-u[k], u[k], V[2][k]); V[2][k] = fmaf(-v[k], v[k], V[2][k]); V[3][k] = fmaf(-u[k], v[k], V[3][k]);
      }
    }
    if (y - y_a >= kWin - 1) old = fetch(y + 1 - kWin);      // the row that leaves at the next iteration
    if (y - y_a < kWin - 1) continue;              // (warp-uniform) the first complete window ends at row y_a + 6
    float S[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float a0 = V[q][0], a1 = V[q][1], a2 = V[q][2], a3 = V[q][3];
      const float b0 = __shfl_down_sync(0xffffffffu, a0, 1), b1 = __shfl_down_sync(0xffffffffu, a1, 1);
      const float b2 = __shfl_down_sync(0xffffffffu, a2, 1), b3 = __shfl_down_sync(0xffffffffu, a3, 1);
      const float d0 = __shfl_down_sync(0xffffffffu, a0, 2), d1 = __shfl_down_sync(0xffffffffu, a1, 2);
      const float m = (a3 + b0) + (b1 + b2);       // common to the four windows
      const float a12 = a1 + a2, bd = b3 + d0;
      S[q][0] = (a0 + a12) + m;
      S[q][1] = a12 + (m + b3);
      S[q][2] = a2 + (m + bd);
      S[q][3] = m + (bd + d1);
    }
    float row = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float s0 = S[0][k] * inv, s1 = S[1][k] * inv;
      const float vsum = cov_norm * fmaf(-s1, s1, fmaf(-s0, s0, S[2][k] * inv));        // var x + var y
      const float vxy = cov_norm * fmaf(-s0, s1, S[3][k] * inv);
      const float ux = s0 + mo, uy = s1 + mp;
      const float num = fmaf(2.f * ux, uy, c1) * fmaf(2.f, vxy, c2);
      const float den = fmaf(uy, uy, fmaf(ux, ux, c1)) * (vsum + c2);
      if (lane < kColsPerWarp / 4 && xc + k < nx) row += __fdividef(num, den);
    }
    local += (double)row;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
  if (lane == 0) atomicAdd(&acc[g].ssim_sum, local);
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
  long long bx = (n + 256 * 8 - 1) / (256 * 8);
  if (bx < 1) bx = 1;
  if (bx > 64) bx = 64;
  metrics::reduce_kernel<<<dim3((unsigned)bx, (unsigned)N), 256, 0, st>>>(d_original, d_predicted, n, acc);
  const int nx = W - metrics::kWin + 1, ny = H - metrics::kWin + 1;
  if (W % 4 == 0 && aligned16(d_original) && aligned16(d_predicted)) {
    // row-streaming kernel: enough row segments for ~32 warps per SM, at least 16 window rows each
    const int n_cb = (nx + metrics::kColsPerWarp - 1) / metrics::kColsPerWarp;
    int sms = 148;
    {
      int dev = 0;
      MRINR_CUDA(cudaGetDevice(&dev));
      MRINR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    long long n_seg = (32ll * sms + N * n_cb - 1) / (N * n_cb);
    if (n_seg > (ny + 15) / 16) n_seg = (ny + 15) / 16;
    if (n_seg < 1) n_seg = 1;
    const int rows_per_seg = (int)((ny + n_seg - 1) / n_seg);
    n_seg = (ny + rows_per_seg - 1) / rows_per_seg;
    const int units = (int)(n_cb * n_seg);
    metrics::ssim_stream_kernel<<<dim3((unsigned)((units + 3) / 4), (unsigned)N), 128, 0, st>>>(
        d_original, d_predicted, H, W, n_cb, (int)n_seg, rows_per_seg, acc);
  } else {
    dim3 grid((nx + metrics::kT - 1) / metrics::kT, (ny + metrics::kT - 1) / metrics::kT, (unsigned)N);
    metrics::ssim_kernel<<<grid, 256, 0, st>>>(d_original, d_predicted, H, W, acc);
  }
  metrics::finalize_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(acc, N, H, W, d_out);
  count_launch(4);
  return check_launch("image_metrics");
}
