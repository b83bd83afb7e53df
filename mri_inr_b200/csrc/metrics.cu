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
// HBM-bound: pass 1 reads both images once (8 B per pixel), pass 2 reads them once more through 38x38 shared-memory
// tiles (1.41x halo), separable 7-tap sums in fp32, partial sums in fp64.
#include "common.cuh"

namespace mrinr {
namespace metrics {

struct Acc {                 // per image pair, in the caller's scratch buffer
  unsigned int mn, mx;       // order-preserving encodings of the min / max over both images
  double sse, so2, ssim_sum;
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
  if (i < N) { acc[i].mn = 0xffffffffu; acc[i].mx = 0u; acc[i].sse = 0.0; acc[i].so2 = 0.0; acc[i].ssim_sum = 0.0; }
}

__global__ void __launch_bounds__(256)
reduce_kernel(const float* __restrict__ orig, const float* __restrict__ pred, long long n, Acc* __restrict__ acc) {
  const long long g = blockIdx.y;
  const float* o = orig + g * n;
  const float* p = pred + g * n;
  float mn = INFINITY, mx = -INFINITY;
  double sse = 0.0, so2 = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float a = __ldg(o + i), b = __ldg(p + i);
    mn = fminf(mn, fminf(a, b));
    mx = fmaxf(mx, fmaxf(a, b));
    const float d = a - b;
    sse += (double)d * (double)d;
    so2 += (double)a * (double)a;
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, s));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    sse += __shfl_xor_sync(0xffffffffu, sse, s);
    so2 += __shfl_xor_sync(0xffffffffu, so2, s);
  }
  __shared__ float smn[8], smx[8];
  __shared__ double sse_s[8], so2_s[8];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { smn[w] = mn; smx[w] = mx; sse_s[w] = sse; so2_s[w] = so2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) { mn = fminf(mn, smn[k]); mx = fmaxf(mx, smx[k]); sse += sse_s[k]; so2 += so2_s[k]; }
    atomicMin(&acc[g].mn, f2ord(mn));
    atomicMax(&acc[g].mx, f2ord(mx));
    atomicAdd(&acc[g].sse, sse);
    atomicAdd(&acc[g].so2, so2);
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
  dim3 grid((nx + metrics::kT - 1) / metrics::kT, (ny + metrics::kT - 1) / metrics::kT, (unsigned)N);
  metrics::ssim_kernel<<<grid, 256, 0, st>>>(d_original, d_predicted, H, W, acc);
  metrics::finalize_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(acc, N, H, W, d_out);
  count_launch(4);
  return check_launch("image_metrics");
}
