// Training-mode forward and backward of the whole module: what `Trainer._train_iteration`
// (src/train/training.py:177-207) needs from `ModulatedSiren` -- `model(undersampled)` in train() mode with dropout
// after every hidden activation (src/networks/modulated_siren.py:124,154-156), and gradients with respect to every
// parameter (synthesis net, modulator, patch encoder) for a loss on the `[B,S,S]` output.
//
//   forward   z = encoder(x);  m_0..m_{L-1} = modulator(z)                                   (inference kernels, kept)
//             h_0 = drop(act_0(W_0 g + b_0)) * m_0;  h_l = drop(act(h_{l-1} W_l^T + b_l)) * m_l;  y = sin(w0 (h w_last + b))
//   backward  dy -> dw_last, db_last, dh_{L-1};  per layer: dm_l = sum_c dh * drop(act), dz = dh * m_l * keep * act',
//             db_l = sum dz, dW_l = dz^T h_{l-1}, dh_{l-1} = dz W_l;  then the modulator (ReLU chains with the latent
//             concatenated, modulated_siren.py:337-341) and the encoder (Linear, 8x8 conv as a GEMM, two strided 3x3
//             convolutions, LeakyReLU(0.2); siren_encoder.py:503-512).
//
// Arithmetic: everything is evaluated in fp32-class precision.  The [M,256] x [256,256] products of the forward and of
// dh = dz W run on the tensor cores through the split-fp16 kernel of dense_tc.cu (three MMAs per product, ~1e-6);
// the weight gradients of the synthesis layers, dz^T h (a reduction over M = B * S * S rows), too (wgrad_tc.cu: the
// reduction index is the MMA's K, the operand tiles are transposed while staging, split-K over the SMs).  The small
// modulator / encoder weight gradients (a reduction over B rows) run on a CUDA-core split-K SGEMM with fp32 atomics.
// The reference trains under fp16 autocast (training.py:29,197), so fp32-class results are inside its own noise.
// Gradient range: the split-fp16 operands hold 22 significant bits only above ~6e-5 (fp16 subnormals below), and the
// gradients of an MSE over B * S * S pixels are ~1e-6.  mrinr_train_backward therefore takes a power-of-two `grad_scale`
// (the caller picks it from max|dy|, like a loss scale but exact): dy is multiplied by it on the way in, every
// intermediate gradient carries it, and the parameter gradients are multiplied by 1/grad_scale at the end.
// Dropout: counter-based (a hash of seed, layer and element index), recomputed in the backward pass -- no mask is
// stored; an explicit keep-mask can be supplied instead (parity tests against the reference with a fixed mask).
#include "common.cuh"

#include <cstring>

using namespace mrinr;

extern "C" int mrinr_encoder_forward(const MrinrPacked* p, const float* d_patches, int64_t B, float* d_latent,
                                     void* d_workspace, int64_t workspace_bytes, void* stream);
extern "C" int64_t mrinr_encoder_workspace_bytes(int64_t B);
extern "C" int mrinr_modulator_forward(const MrinrPacked* p, const float* d_latent, int64_t B, float* d_mods, void* stream);

namespace mrinr {
namespace train {

constexpr int kH = 256;

// ---- dropout -------------------------------------------------------------------------------------------------
struct Drop {
  const uint8_t* mask;   // optional explicit keep-mask [L][M*H] (1 = keep); overrides the hash
  unsigned long long seed;
  unsigned int thresh;   // keep iff hash >= thresh  (thresh = p * 2^32)
  float inv_keep;        // 1 / (1 - p)
  int enabled;
};
__device__ __forceinline__ unsigned int mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (unsigned int)(x >> 32);
}
__device__ __forceinline__ float keep_scale(const Drop& d, int layer, long long idx, long long plane) {
  if (!d.enabled) return 1.f;
  bool keep;
  if (d.mask) keep = d.mask[(long long)layer * plane + idx] != 0;
  else keep = mix64(d.seed ^ ((unsigned long long)(layer + 1) << 48) ^ (unsigned long long)idx) >= d.thresh;
  return keep ? d.inv_keep : 0.f;
}

// the same for four consecutive elements idx .. idx + 3 (idx a multiple of 4)
__device__ __forceinline__ void keep_scale4(const Drop& d, int layer, long long idx, long long plane, float (&ks)[4]) {
  if (!d.enabled) {
    ks[0] = ks[1] = ks[2] = ks[3] = 1.f;
    return;
  }
  if (d.mask) {
    const uint8_t* mp = d.mask + (long long)layer * plane + idx;
    uint8_t k[4];
    if ((reinterpret_cast<uintptr_t>(mp) & 3u) == 0) {
      const uchar4 m = *reinterpret_cast<const uchar4*>(mp);
      k[0] = m.x; k[1] = m.y; k[2] = m.z; k[3] = m.w;
    } else {
      k[0] = mp[0]; k[1] = mp[1]; k[2] = mp[2]; k[3] = mp[3];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) ks[i] = k[i] ? d.inv_keep : 0.f;
    return;
  }
  const unsigned long long base = d.seed ^ ((unsigned long long)(layer + 1) << 48);
#pragma unroll
  for (int i = 0; i < 4; ++i) ks[i] = mix64(base ^ (unsigned long long)(idx + i)) >= d.thresh ? d.inv_keep : 0.f;
}

// ---- activations (modulated_siren.py:54, :80) and their derivatives --------------------------------------------
__device__ __forceinline__ float act_f(float z, float w0, int morlet) {
  const float s = sinf(w0 * z);
  return morlet ? s * expf(-0.5f * z * z) : s;
}
__device__ __forceinline__ void act_fd(float z, float w0, int morlet, float& a, float& da) {
  float s, c;
  sincosf(w0 * z, &s, &c);
  if (morlet) {
    const float e = expf(-0.5f * z * z);
    a = s * e;
    da = (w0 * c - z * s) * e;
  } else {
    a = s;
    da = w0 * c;
  }
}

// pre0[c][j] = W_0[j,:] . g_c + b_0[j]   (the same rounding sequence as the inference table, dense_fp32.cu)
__global__ void layer0_pre_kernel(const float* __restrict__ grid, const float* __restrict__ w, const float* __restrict__ b,
                                  int C, float* __restrict__ pre0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * kH) return;
  const int c = i / kH, j = i - c * kH;
  float pre = __fmaf_rn(grid[2 * c + 1], w[2 * j + 1], __fmul_rn(grid[2 * c], w[2 * j]));
  if (b) pre = __fadd_rn(pre, b[j]);
  pre0[i] = pre;
}

// h[m][j] = drop(act(pre)) * mod[b][j];  pre = pre0[c][j] for layer 0 (patch independent), else pre[m][j]
// grid (patch, coordinate chunk).  A thread owns FOUR consecutive features (16-byte loads and stores, four independent
// sine / hash chains in flight) of every fourth coordinate of the chunk: the kernels are issue-bound on the accurate
// sinf / sincosf and the dropout hash, and with one element per thread and trip each trip was one dependent chain
// (228 us per layer for 472 MB).  Rows are read and written coalesced, no index divisions.
constexpr int kChunks = 4;      // coordinate chunks per patch (CTAs per patch) of the element-wise kernels
constexpr int kRowsPerTrip = 4; // kH / 4 threads cover a row, so a 256-thread CTA covers 4 rows per trip
__global__ void __launch_bounds__(kH) act_fwd_kernel(const float* __restrict__ pre, int layer0, const float* __restrict__ mod,
                                                     long long M, int C, float w0, int morlet, Drop drop, int layer,
                                                     float* __restrict__ h) {
  const long long b = blockIdx.x;
  const int j = 4 * (threadIdx.x & (kH / 4 - 1)), r = threadIdx.x / (kH / 4);
  const int per = (C + kChunks - 1) / kChunks;
  const int c0 = blockIdx.y * per, c1 = min(C, c0 + per);
  const long long total = M * kH;
  const float4 mj = *reinterpret_cast<const float4*>(mod + b * kH + j);
  for (int c = c0 + r; c < c1; c += kRowsPerTrip) {
    const long long idx = (b * C + c) * kH + j;
    const float4 z = *reinterpret_cast<const float4*>(layer0 ? pre + (long long)c * kH + j : pre + idx);
    float ks[4];
    keep_scale4(drop, layer, idx, total, ks);
    float4 o;
    o.x = act_f(z.x, w0, morlet) * ks[0] * mj.x;
    o.y = act_f(z.y, w0, morlet) * ks[1] * mj.y;
    o.z = act_f(z.z, w0, morlet) * ks[2] * mj.z;
    o.w = act_f(z.w, w0, morlet) * ks[3] * mj.w;
    *reinterpret_cast<float4*>(h + idx) = o;
  }
}

// output layer (always sine, never modulated, no dropout; modulated_siren.py:211-213,233): one warp per row
__global__ void out_fwd_kernel(const float* __restrict__ h, const float* __restrict__ w_last, const float* __restrict__ b_last,
                               long long M, float w0, float* __restrict__ pre_last, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long m = warp; m < M; m += nwarps) {
    const float4* row = reinterpret_cast<const float4*>(h + m * kH);
    const float4* wv = reinterpret_cast<const float4*>(w_last);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float4 a = row[lane + 32 * i], w = wv[lane + 32 * i];
      s += a.x * w.x + a.y * w.y + a.z * w.z + a.w * w.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      const float p = s + (b_last ? *b_last : 0.f);
      pre_last[m] = p;
      y[m] = sinf(w0 * p);
    }
  }
}

// grid (patch, coordinate chunk), thread layout as in act_fwd_kernel (four features of every fourth coordinate):
// g = dy w0 cos(w0 pre_last);  dh = g w_last;  dw_last += g h;  db_last += g
__global__ void __launch_bounds__(kH) out_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ pre_last,
                                                     const float* __restrict__ h, const float* __restrict__ w_last, int C,
                                                     float w0, float grad_scale, float* __restrict__ dh,
                                                     float* __restrict__ dw_last, float* __restrict__ db_last) {
  __shared__ float red[kRowsPerTrip][kH];
  __shared__ float redb[kRowsPerTrip];
  const long long b = blockIdx.x;
  const int j = 4 * (threadIdx.x & (kH / 4 - 1)), r = threadIdx.x / (kH / 4);
  const int per = (C + kChunks - 1) / kChunks;
  const int c0 = blockIdx.y * per, c1 = min(C, c0 + per);
  const float4 wl = *reinterpret_cast<const float4*>(w_last + j);
  float4 accw = make_float4(0.f, 0.f, 0.f, 0.f);
  float accb = 0.f;
  for (int c = c0 + r; c < c1; c += kRowsPerTrip) {
    const long long m = b * C + c;
    const float g = dy[m] * grad_scale * w0 * cosf(w0 * pre_last[m]);
    const float4 hv = *reinterpret_cast<const float4*>(h + m * kH + j);
    *reinterpret_cast<float4*>(dh + m * kH + j) = make_float4(g * wl.x, g * wl.y, g * wl.z, g * wl.w);
    accw.x = fmaf(g, hv.x, accw.x);
    accw.y = fmaf(g, hv.y, accw.y);
    accw.z = fmaf(g, hv.z, accw.z);
    accw.w = fmaf(g, hv.w, accw.w);
    accb += g;
  }
  red[r][j] = accw.x; red[r][j + 1] = accw.y; red[r][j + 2] = accw.z; red[r][j + 3] = accw.w;
  if ((threadIdx.x & (kH / 4 - 1)) == 0) redb[r] = accb;
  __syncthreads();
  const int t = threadIdx.x;
  if (dw_last) atomicAdd(dw_last + t, (red[0][t] + red[1][t]) + (red[2][t] + red[3][t]));
  if (db_last && t == 0) atomicAdd(db_last, (redb[0] + redb[1]) + (redb[2] + redb[3]));
}

// grid (patch, coordinate chunk); thread layout as in act_fwd_kernel.  In: dh (gradient w.r.t. h_l).  Out (in place): dz
// (gradient w.r.t. the pre-activation);  dmod[b][j] = sum_c dh * drop(act);  db[j] += sum dz;  layer 0 also:
// dW_0[j][0..1] += sum dz * g_c.  The four row groups of a CTA are added in shared memory before the atomics.
__global__ void __launch_bounds__(kH) act_bwd_kernel(float* __restrict__ dh, const float* __restrict__ pre, int layer0,
                                                     const float* __restrict__ mod, const float* __restrict__ grid, int C,
                                                     long long M, float w0, int morlet, Drop drop, int layer,
                                                     float* __restrict__ dmod, float* __restrict__ db,
                                                     float* __restrict__ dw0) {
  __shared__ float red[4][kRowsPerTrip][kH];
  const long long b = blockIdx.x;
  const int j = 4 * (threadIdx.x & (kH / 4 - 1)), r = threadIdx.x / (kH / 4);
  const int per = (C + kChunks - 1) / kChunks;
  const int c0 = blockIdx.y * per, c1 = min(C, c0 + per);
  const float4 mj4 = *reinterpret_cast<const float4*>(mod + b * kH + j);
  const float mj[4] = {mj4.x, mj4.y, mj4.z, mj4.w};
  const long long total = M * kH;
  float accm[4] = {0.f, 0.f, 0.f, 0.f}, accb[4] = {0.f, 0.f, 0.f, 0.f}, accx[4] = {0.f, 0.f, 0.f, 0.f},
        accy[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c = c0 + r; c < c1; c += kRowsPerTrip) {
    const long long idx = (b * C + c) * kH + j;
    const float4 z4 = *reinterpret_cast<const float4*>(layer0 ? pre + (long long)c * kH + j : pre + idx);
    const float4 g4 = *reinterpret_cast<const float4*>(dh + idx);
    const float z[4] = {z4.x, z4.y, z4.z, z4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
    float ks[4], dz[4];
    keep_scale4(drop, layer, idx, total, ks);
    float gx = 0.f, gy = 0.f;
    if (layer0) { gx = grid[2 * c]; gy = grid[2 * c + 1]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a, da;
      act_fd(z[i], w0, morlet, a, da);
      accm[i] = fmaf(g[i], a * ks[i], accm[i]);
      dz[i] = g[i] * mj[i] * ks[i] * da;
      accb[i] += dz[i];
      if (layer0) {
        accx[i] = fmaf(dz[i], gx, accx[i]);
        accy[i] = fmaf(dz[i], gy, accy[i]);
      }
    }
    *reinterpret_cast<float4*>(dh + idx) = make_float4(dz[0], dz[1], dz[2], dz[3]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[0][r][j + i] = accm[i];
    red[1][r][j + i] = accb[i];
    red[2][r][j + i] = accx[i];
    red[3][r][j + i] = accy[i];
  }
  __syncthreads();
  const int t = threadIdx.x;
  auto total_of = [&](int q) { return (red[q][0][t] + red[q][1][t]) + (red[q][2][t] + red[q][3][t]); };
  atomicAdd(dmod + b * kH + t, total_of(0));     // kChunks partial sums per element (dmod is zero-initialised)
  if (db) atomicAdd(db + t, total_of(1));
  if (layer0 && dw0) {
    atomicAdd(dw0 + 2 * t, total_of(2));
    atomicAdd(dw0 + 2 * t + 1, total_of(3));
  }
}

// ---- C[Na,Nb] += A[M,Na]^T B[M,Nb]  (reduction over the M rows; split over row ranges, fp32 atomics) ---------------
constexpr int kGT = 128, kGK = 8;
__global__ void __launch_bounds__(256) gemm_tn_atomic_kernel(const float* __restrict__ A, long long lda, int Na,
                                                             const float* __restrict__ Bm, long long ldb, int Nb,
                                                             long long M, int rows_per_cta, float* __restrict__ Cm,
                                                             long long ldc) {
  __shared__ __align__(16) float As[kGK][kGT];
  __shared__ __align__(16) float Bs[kGK][kGT];
  const int a0 = blockIdx.x * kGT, b0 = blockIdx.y * kGT;
  const long long m0 = (long long)blockIdx.z * rows_per_cta;
  const long long m1 = m0 + rows_per_cta < M ? m0 + rows_per_cta : M;
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;          // 16 x 16 threads, 8 x 8 outputs each: rows ty*4+{0..3}, 64+ty*4+{0..3}
  const int lr = t >> 5, lc = (t & 31) * 4;    // loader: row lr of the K slab, 4 consecutive columns
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;
  for (long long m = m0; m < m1; m += kGK) {
    {
      const long long r = m + lr;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (r < m1) {
        const float* pa = A + r * lda + a0 + lc;
        const float* pb = Bm + r * ldb + b0 + lc;
        if (a0 + lc + 3 < Na) va = *reinterpret_cast<const float4*>(pa);
        else {
          if (a0 + lc + 0 < Na) va.x = pa[0];
          if (a0 + lc + 1 < Na) va.y = pa[1];
          if (a0 + lc + 2 < Na) va.z = pa[2];
        }
        if (b0 + lc + 3 < Nb) vb = *reinterpret_cast<const float4*>(pb);
        else {
          if (b0 + lc + 0 < Nb) vb.x = pb[0];
          if (b0 + lc + 1 < Nb) vb.y = pb[1];
          if (b0 + lc + 2 < Nb) vb.z = pb[2];
        }
      }
      *reinterpret_cast<float4*>(&As[lr][lc]) = va;
      *reinterpret_cast<float4*>(&Bs[lr][lc]) = vb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kGK; ++k) {
      const float4 a_lo = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a_hi = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b_lo = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b_hi = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float av[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
      const float bv[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[i][q] = fmaf(av[i], bv[q], acc[i][q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int a = a0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (a >= Na) continue;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int bcol = b0 + (q < 4 ? tx * 4 + q : 64 + tx * 4 + q - 4);
      if (bcol < Nb) atomicAdd(Cm + (long long)a * ldc + bcol, acc[i][q]);
    }
  }
}

static int gemm_tn_atomic(const float* A, long long lda, int Na, const float* Bm, long long ldb, int Nb, long long M,
                          float* Cm, long long ldc, cudaStream_t st) {
  if (M <= 0 || !Cm) return 0;
  MRINR_REQUIRE(aligned16(A) && aligned16(Bm) && lda % 4 == 0 && ldb % 4 == 0, MRINR_E_ALIGN,
                "gemm_tn: operands must be 16-byte aligned with row strides that are multiples of 4");
  // slices of the reduction per output tile: enough CTAs to fill the GPU twice, at least 32 rows each
  const int tiles = ((Na + kGT - 1) / kGT) * ((Nb + kGT - 1) / kGT);
  long long slices = (2 * 148 + tiles - 1) / tiles;
  if (slices > (M + 31) / 32) slices = (M + 31) / 32;
  if (slices < 1) slices = 1;
  const int rows = (int)(((M + slices - 1) / slices + kGK - 1) / kGK * kGK);
  dim3 grid((Na + kGT - 1) / kGT, (Nb + kGT - 1) / kGT, (unsigned)((M + rows - 1) / rows));
  gemm_tn_atomic_kernel<<<grid, 256, 0, st>>>(A, lda, Na, Bm, ldb, Nb, M, rows, Cm, ldc);
  count_launch();
  return check_launch("gemm_tn_atomic");
}

// ---- small element-wise / reduction kernels of the modulator and encoder backward -----------------------------
// out[j] += sum_r X[r][j]
__global__ void colsum_atomic_kernel(const float* __restrict__ X, long long ld, long long R, int N, float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const long long r0 = (long long)blockIdx.y * 256, r1 = r0 + 256 < R ? r0 + 256 : R;
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += X[r * ld + j];
  atomicAdd(out + j, s);
}
static int colsum_atomic(const float* X, long long ld, long long R, int N, float* out, cudaStream_t st) {
  if (!out || R <= 0) return 0;
  dim3 grid((N + 127) / 128, (unsigned)((R + 255) / 256));
  colsum_atomic_kernel<<<grid, 128, 0, st>>>(X, ld, R, N, out);
  count_launch();
  return check_launch("colsum");
}
// G = (dm + carry) * (h > 0)      (carry may be null)
__global__ void relu_bwd_kernel(const float* __restrict__ dm, const float* __restrict__ carry, const float* __restrict__ h,
                                long long n, float* __restrict__ G) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = dm[i] + (carry ? carry[i] : 0.f);
  G[i] = h[i] > 0.f ? g : 0.f;
}
// D *= (out > 0 ? 1 : slope)   -- LeakyReLU backward from the layer's OUTPUT (slope > 0 keeps the sign)
__global__ void lrelu_bwd_kernel(float* __restrict__ D, const float* __restrict__ out, long long n, float slope) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!(out[i] > 0.f)) D[i] *= slope;
}
__global__ void scale_kernel(float* __restrict__ Y, long long n, float s) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) Y[i] *= s;
}
__global__ void add_kernel(float* __restrict__ Y, const float* __restrict__ X, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) Y[i] += X[i];
}
// dst[k][n] = src[n][k]
__global__ void transpose_kernel(const float* __restrict__ src, int N, int K, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const int n = i / K, k = i - n * K;
  dst[(long long)k * N + n] = src[i];
}
#define EW_LAUNCH(kernel, n, ...)                                                       \
  do {                                                                                  \
    const long long n_ = (n);                                                           \
    if (n_ > 0) {                                                                       \
      kernel<<<(unsigned)((n_ + 255) / 256), 256, 0, st>>>(__VA_ARGS__);                \
      count_launch();                                                                   \
      const int rc_ = check_launch(#kernel);                                            \
      if (rc_ != 0) return rc_;                                                         \
    }                                                                                   \
  } while (0)

// ---- the two strided 3x3 convolutions of the encoder, backward (siren_encoder.py:503-507) ---------------------------
// One CTA walks patches b = blockIdx.x, + gridDim.x, ...: recompute c1 = lrelu(conv1(x)) (the fused forward kernel never
// writes it), back-propagate g2 = d(loss)/d(conv2 pre-activation) [32][8][8] to conv1's pre-activation, and accumulate the
// four parameter gradients in registers; one atomicAdd per output and CTA at the end.
constexpr int kConvSmemFloats = 34 * 34 + 16 * 18 * 18 + 16 * 16 * 16 + 32 * 64 + 32 * 16 * 9 + 16 * 16 * 16;
__global__ void __launch_bounds__(256) encoder_conv_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g2all,
                                                               long long B, const float* __restrict__ w1,
                                                               const float* __restrict__ b1, const float* __restrict__ w2,
                                                               float* __restrict__ dw1, float* __restrict__ db1,
                                                               float* __restrict__ dw2, float* __restrict__ db2) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                         // [34][34]   input with a zero border (index = coordinate + 1)
  float* c1s = xs + 34 * 34;              // [16][18][18] conv1 output with a zero border
  float* f1 = c1s + 16 * 18 * 18;         // [16][16][16] LeakyReLU slope factor of conv1's pre-activation
  float* g2 = f1 + 16 * 16 * 16;          // [32][8][8]
  float* w2s = g2 + 32 * 64;              // [32][16][9]
  float* d1s = w2s + 32 * 16 * 9;         // [16][16][16] gradient w.r.t. conv1's pre-activation
  const int t = threadIdx.x;
  for (int i = t; i < 32 * 16 * 9; i += 256) w2s[i] = w2[i];
  float acc2[18];
#pragma unroll
  for (int i = 0; i < 18; ++i) acc2[i] = 0.f;
  float acc1 = 0.f, accb = 0.f;           // threads 0..143: dw1; 144..159: db1; 160..191: db2
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = t; i < 34 * 34; i += 256) {
      const int r = i / 34, c = i - r * 34;
      xs[i] = (r >= 1 && r <= 32 && c >= 1 && c <= 32) ? x[b * 1024 + (r - 1) * 32 + (c - 1)] : 0.f;
    }
    for (int i = t; i < 16 * 18 * 18; i += 256) c1s[i] = 0.f;
    for (int i = t; i < 32 * 64; i += 256) g2[i] = g2all[b * 2048 + i];
    __syncthreads();
    // conv1 forward: Conv2d(1,16,3,stride 2,pad 1) + LeakyReLU(0.2)
    for (int i = t; i < 16 * 256; i += 256) {
      const int o = i >> 8, y = (i >> 4) & 15, xx = i & 15;
      float s = b1[o];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) s = fmaf(w1[o * 9 + ky * 3 + kx], xs[(2 * y + ky) * 34 + 2 * xx + kx], s);
      const bool pos = s > 0.f;
      c1s[o * 324 + (y + 1) * 18 + xx + 1] = pos ? s : 0.2f * s;
      f1[i] = pos ? 1.f : 0.2f;
    }
    __syncthreads();
    // gradient w.r.t. conv1's output: dc1[i][u][v] = sum_o sum_{ky,kx} g2[o][y][x] w2[o][i][ky][kx], u = 2y + ky - 1
    for (int i = t; i < 16 * 256; i += 256) {
      const int ci = i >> 8, u = (i >> 4) & 15, v = i & 15;
      float s = 0.f;
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = u + 1 - ky;
        if (yy & 1) continue;
        const int y = yy >> 1;
        if (y < 0 || y > 7) continue;
        for (int kx = 0; kx < 3; ++kx) {
          const int xx2 = v + 1 - kx;
          if (xx2 & 1) continue;
          const int xq = xx2 >> 1;
          if (xq < 0 || xq > 7) continue;
          for (int o = 0; o < 32; ++o) s = fmaf(g2[o * 64 + y * 8 + xq], w2s[(o * 16 + ci) * 9 + ky * 3 + kx], s);
        }
      }
      d1s[i] = s * f1[i];
    }
    __syncthreads();
    // dw2[o][i][ky][kx] += sum_{y,x} g2[o][y][x] c1[i][2y+ky-1][2x+kx-1]
#pragma unroll
    for (int r = 0; r < 18; ++r) {
      const int idx = t + 256 * r;
      const int o = idx / 144, rem = idx - o * 144, ci = rem / 9, k = rem - ci * 9, ky = k / 3, kx = k - ky * 3;
      float s = 0.f;
      for (int y = 0; y < 8; ++y)
#pragma unroll
        for (int xq = 0; xq < 8; ++xq) s = fmaf(g2[o * 64 + y * 8 + xq], c1s[ci * 324 + (2 * y + ky) * 18 + 2 * xq + kx], s);
      acc2[r] += s;
    }
    if (t < 144) {
      const int o = t / 9, k = t - o * 9, ky = k / 3, kx = k - ky * 3;
      float s = 0.f;
      for (int y = 0; y < 16; ++y)
        for (int xx = 0; xx < 16; ++xx) s = fmaf(d1s[o * 256 + y * 16 + xx], xs[(2 * y + ky) * 34 + 2 * xx + kx], s);
      acc1 += s;
    } else if (t < 160) {
      const int o = t - 144;
      float s = 0.f;
      for (int i = 0; i < 256; ++i) s += d1s[o * 256 + i];
      accb += s;
    } else if (t < 192) {
      const int o = t - 160;
      float s = 0.f;
      for (int i = 0; i < 64; ++i) s += g2[o * 64 + i];
      accb += s;
    }
  }
  if (dw2) {
#pragma unroll
    for (int r = 0; r < 18; ++r) atomicAdd(dw2 + t + 256 * r, acc2[r]);
  }
  if (t < 144) { if (dw1) atomicAdd(dw1 + t, acc1); }
  else if (t < 160) { if (db1) atomicAdd(db1 + (t - 144), accb); }
  else if (t < 192) { if (db2) atomicAdd(db2 + (t - 160), accb); }
}

// ---- split-fp16 operand packing with arbitrary source strides (layout of dense_tc.cu: pack_split_kernel) --------------
// packed W'[n][k] = src[n * sn + k * sk],  n < N, k < K  ->  [K/32 slabs][hi, lo][4 kc][N][8] fp16
__global__ void pack_split_strided_kernel(const float* __restrict__ src, int N, int K, long long sn, long long sk,
                                          uint16_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * K) return;
  const int e = (int)(i & 7);
  const int n = (int)((i >> 3) % N);
  const int kc = (int)(((i >> 3) / N) % 4);
  const int slab = (int)(((i >> 3) / N) / 4);
  const float v = src[(long long)n * sn + (long long)(slab * 32 + kc * 8 + e) * sk];
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  const long long base = (long long)slab * 2 * 4 * N * 8 + ((long long)kc * N + n) * 8 + e;
  out[base] = *reinterpret_cast<const uint16_t*>(&hi);
  out[base + (long long)4 * N * 8] = *reinterpret_cast<const uint16_t*>(&lo);
}
static int pack_split_strided(const float* src, int N, int K, long long sn, long long sk, uint16_t* out, cudaStream_t st) {
  const long long n = (long long)N * K;
  pack_split_strided_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, N, K, sn, sk, out);
  count_launch();
  return check_launch("pack_split_strided");
}

// ---- workspace ---------------------------------------------------------------------------------------------------
struct Ws {
  // element offsets (floats) into the workspace; every section starts at a multiple of 64 floats (256 bytes)
  size_t enc, z, mods, pre0, h, pre, pre_last, pk, dA, dB, dmods, G, carry, dzlat, tmpz, dc3, dc2, wg, total;
  size_t plane;   // M * H
};
static inline size_t al(size_t x) { return (x + 63) & ~(size_t)63; }
static Ws layout(const MrinrPacked* p, int64_t B) {
  Ws w;
  const size_t L = p->L, H = p->H, Z = p->Z, C = p->C, M = (size_t)B * C;
  w.plane = M * H;
  size_t o = 0;
  w.enc = o; o += al((size_t)B * (2048 + 64));
  w.z = o; o += al((size_t)B * Z);
  w.mods = o; o += al(L * B * H);
  w.pre0 = o; o += al(C * H);
  w.h = o; o += L * al(w.plane);
  w.pre = o; o += (L - 1) * al(w.plane);
  w.pre_last = o; o += al(M);
  // packed split-fp16 operands (two uint16 per weight = one float slot each): the largest user is one [256,256] matrix
  // at a time for the synthesis net, the modulator halves, and the 2048 x 64 / Z x 64 encoder matrices
  w.pk = o; o += al((size_t)2048 * 64 > H * (H + Z) ? (size_t)2048 * 64 : H * (H + Z));
  w.dA = o; o += al(w.plane);
  w.dB = o; o += al(w.plane);
  w.dmods = o; o += al(L * B * H);
  w.G = o; o += al((size_t)B * H);
  w.carry = o; o += al((size_t)B * H);
  w.dzlat = o; o += al((size_t)B * Z);
  w.tmpz = o; o += al((size_t)B * (Z > 64 ? Z : 64));
  w.dc3 = o; o += al((size_t)B * 64);
  w.dc2 = o; o += al((size_t)B * 2048);
  w.wg = o; o += al((size_t)wgrad_tc_scratch_floats(p->num_sms));      // per-CTA partial results of the weight gradient
  w.total = o;
  return w;
}

static int check_cfg(const MrinrPacked* p, const char* who) {
  MRINR_REQUIRE(p->H == kH, MRINR_E_UNSUPPORTED, "%s: the training path needs dim_hidden == 256 (got %d)", who, p->H);
  MRINR_REQUIRE(p->has_encoder, MRINR_E_ARG, "%s: no encoder weights were packed", who);
  MRINR_REQUIRE(p->Z == 64 || p->Z == 128 || p->Z == 256, MRINR_E_UNSUPPORTED, "%s: latent_dim must be 64, 128 or 256", who);
  return 0;
}

static Drop make_drop(float p, unsigned long long seed, const uint8_t* mask) {
  Drop d;
  d.mask = mask;
  d.seed = seed;
  d.enabled = p > 0.f ? 1 : 0;
  const double t = (double)p * 4294967296.0;
  d.thresh = t >= 4294967295.0 ? 4294967295u : (unsigned int)t;
  d.inv_keep = p < 1.f ? 1.f / (1.f - p) : 0.f;
  return d;
}

}  // namespace train
}  // namespace mrinr

using namespace mrinr::train;

extern "C" int64_t mrinr_train_workspace_bytes(const MrinrPacked* p, int64_t B) {
  if (!p || B <= 0) return 0;
  return (int64_t)(layout(p, B).total * sizeof(float));
}

extern "C" int mrinr_train_forward(const MrinrPacked* p, const MrinrWeightsView* v, const float* d_tiles, int64_t B,
                                   float dropout_p, uint64_t seed, const uint8_t* d_keep_mask, float* d_out,
                                   void* d_workspace, int64_t workspace_bytes, void* stream) {
  if (B == 0) return 0;
  MRINR_REQUIRE(p && v && d_tiles && d_out && d_workspace, MRINR_E_ARG, "mrinr_train_forward: null pointer");
  MRINR_REQUIRE(B > 0 && dropout_p >= 0.f && dropout_p < 1.f, MRINR_E_ARG, "mrinr_train_forward: bad batch / dropout");
  int rc = check_cfg(p, "mrinr_train_forward");
  if (rc != 0) return rc;
  MRINR_REQUIRE(workspace_bytes >= mrinr_train_workspace_bytes(p, B), MRINR_E_ARG,
                "mrinr_train_forward: needs a workspace of mrinr_train_workspace_bytes(B) bytes");
  MRINR_REQUIRE(aligned16(d_tiles) && aligned16(d_out) && (reinterpret_cast<uintptr_t>(d_workspace) & 255u) == 0,
                MRINR_E_ALIGN, "mrinr_train_forward: buffers must be 16-byte aligned (workspace: 256)");
  cudaStream_t st = (cudaStream_t)stream;
  const Ws w = layout(p, B);
  float* ws = static_cast<float*>(d_workspace);
  const int L = p->L, C = p->C;
  const long long M = (long long)B * C;
  const int morlet = p->activation == MRINR_ACT_MORLET;
  const Drop drop = make_drop(dropout_p, seed, d_keep_mask);
  const size_t hs = al(w.plane);

  // encoder + modulator: the inference kernels; their intermediates (conv2 map, conv3 output, latent, modulations)
  // stay in the workspace for the backward pass
  rc = mrinr_encoder_forward(p, d_tiles, B, ws + w.z, ws + w.enc, mrinr_encoder_workspace_bytes(B), stream);
  if (rc != 0) return rc;
  rc = mrinr_modulator_forward(p, ws + w.z, B, ws + w.mods, stream);
  if (rc != 0) return rc;
  // layer 0: patch-independent pre-activation table, then activation, dropout, modulation
  layer0_pre_kernel<<<(C * kH + 255) / 256, 256, 0, st>>>(v->d_grid, v->d_net_weight[0],
                                                          v->d_net_bias ? v->d_net_bias[0] : nullptr, C, ws + w.pre0);
  count_launch();
  if ((rc = check_launch("layer0_pre")) != 0) return rc;
  const dim3 ew_grid((unsigned)B, kChunks);
  act_fwd_kernel<<<ew_grid, kH, 0, st>>>(ws + w.pre0, 1, ws + w.mods, M, C, p->w0_initial, morlet, drop, 0, ws + w.h);
  count_launch();
  if ((rc = check_launch("act_fwd")) != 0) return rc;
  for (int l = 1; l < L; ++l) {
    uint16_t* pk = reinterpret_cast<uint16_t*>(ws + w.pk);
    if ((rc = pack_split_strided(v->d_net_weight[l], kH, kH, kH, 1, pk, st)) != 0) return rc;
    float* pre = ws + w.pre + (size_t)(l - 1) * hs;
    rc = launch_dense_split(ws + w.h + (size_t)(l - 1) * hs, kH, kH, nullptr, 0, 0, pk,
                            v->d_net_bias ? v->d_net_bias[l] : nullptr, kH, 0, 0.f, pre, kH, M, p->d_errflag, st);
    if (rc != 0) return rc;
    act_fwd_kernel<<<ew_grid, kH, 0, st>>>(pre, 0, ws + w.mods + (size_t)l * B * kH, M, C, p->w0, morlet, drop, l,
                                          ws + w.h + (size_t)l * hs);
    count_launch();
    if ((rc = check_launch("act_fwd")) != 0) return rc;
  }
  const unsigned og = (unsigned)((M + 7) / 8 < 148 * 8 ? (M + 7) / 8 : 148 * 8);
  out_fwd_kernel<<<og, 256, 0, st>>>(ws + w.h + (size_t)(L - 1) * hs, v->d_last_weight, v->d_last_bias, M, p->w0,
                                     ws + w.pre_last, d_out);
  count_launch();
  return check_launch("out_fwd");
}

// grads: the same struct as the weights view, every pointer a ZERO-INITIALISED fp32 buffer of the parameter's shape
// (null = not wanted).  d_grid is ignored (the grid is a buffer, not a parameter).
extern "C" int mrinr_train_backward(const MrinrPacked* p, const MrinrWeightsView* v, const float* d_tiles,
                                    const float* d_dout, int64_t B, float dropout_p, uint64_t seed,
                                    const uint8_t* d_keep_mask, float grad_scale, const MrinrWeightsView* grads,
                                    void* d_workspace, int64_t workspace_bytes, void* stream) {
  if (B == 0) return 0;
  MRINR_REQUIRE(p && v && d_tiles && d_dout && grads && d_workspace, MRINR_E_ARG, "mrinr_train_backward: null pointer");
  MRINR_REQUIRE(grad_scale > 0.f && grad_scale < 3e38f, MRINR_E_ARG, "mrinr_train_backward: grad_scale must be positive");
  int rc = check_cfg(p, "mrinr_train_backward");
  if (rc != 0) return rc;
  MRINR_REQUIRE(B > 0 && workspace_bytes >= mrinr_train_workspace_bytes(p, B), MRINR_E_ARG,
                "mrinr_train_backward: needs the workspace that mrinr_train_forward filled");
  MRINR_REQUIRE(aligned16(d_tiles) && aligned16(d_dout) && (reinterpret_cast<uintptr_t>(d_workspace) & 255u) == 0,
                MRINR_E_ALIGN, "mrinr_train_backward: buffers must be 16-byte aligned (workspace: 256)");
  cudaStream_t st = (cudaStream_t)stream;
  const Ws w = layout(p, B);
  float* ws = static_cast<float*>(d_workspace);
  const int L = p->L, C = p->C, Z = p->Z;
  const long long M = (long long)B * C;
  const int morlet = p->activation == MRINR_ACT_MORLET;
  const Drop drop = make_drop(dropout_p, seed, d_keep_mask);
  const size_t hs = al(w.plane);
  uint16_t* pk = reinterpret_cast<uint16_t*>(ws + w.pk);
  // the gradient view reuses the weights-view struct: its pointers are writable buffers
  float** g_net_w = (float**)grads->d_net_weight;
  float** g_net_b = (float**)grads->d_net_bias;
  float** g_mod_w = (float**)grads->d_mod_weight;
  float** g_mod_b = (float**)grads->d_mod_bias;

  // ---- output layer
  float* cur = ws + w.dA;      // gradient w.r.t. h_l, then (in place) w.r.t. the pre-activation of layer l
  float* nxt = ws + w.dB;
  MRINR_CUDA(cudaMemsetAsync(ws + w.dmods, 0, (size_t)L * B * kH * sizeof(float), st));
  const dim3 ew_grid((unsigned)B, kChunks);
  out_bwd_kernel<<<ew_grid, kH, 0, st>>>(d_dout, ws + w.pre_last, ws + w.h + (size_t)(L - 1) * hs, v->d_last_weight,
                                            C, p->w0, grad_scale, cur, (float*)grads->d_last_weight,
                                            (float*)grads->d_last_bias);
  count_launch();
  if ((rc = check_launch("out_bwd")) != 0) return rc;
  // ---- synthesis layers L-1 .. 0
  for (int l = L - 1; l >= 0; --l) {
    const float* pre = l == 0 ? ws + w.pre0 : ws + w.pre + (size_t)(l - 1) * hs;
    act_bwd_kernel<<<ew_grid, kH, 0, st>>>(cur, pre, l == 0, ws + w.mods + (size_t)l * B * kH, v->d_grid, C, M,
                                              l == 0 ? p->w0_initial : p->w0, morlet, drop, l,
                                              ws + w.dmods + (size_t)l * B * kH, g_net_b ? g_net_b[l] : nullptr,
                                              l == 0 && g_net_w ? g_net_w[0] : nullptr);
    count_launch();
    if ((rc = check_launch("act_bwd")) != 0) return rc;
    if (l == 0) break;
    // dW_l[n][k] += sum_m dz[m][n] h_{l-1}[m][k]: the big reduction (M rows) on the tensor cores (wgrad_tc.cu)
#ifdef MRINR_WGRAD_SGEMM   // diagnostic build only (make wgsgemm): the CUDA-core split-K SGEMM instead
    if (g_net_w && g_net_w[l])
      if ((rc = gemm_tn_atomic(cur, kH, kH, ws + w.h + (size_t)(l - 1) * hs, kH, kH, M, g_net_w[l], kH, st)) != 0) return rc;
#else
    if (g_net_w && g_net_w[l])
      if ((rc = launch_wgrad_tc(cur, ws + w.h + (size_t)(l - 1) * hs, M, ws + w.wg, g_net_w[l], p->num_sms, p->d_errflag,
                                st)) != 0) return rc;
#endif
    // dh_{l-1} = dz W_l : the packed operand is W_l^T, rows = k (output column), K index = n
    if ((rc = pack_split_strided(v->d_net_weight[l], kH, kH, 1, kH, pk, st)) != 0) return rc;
    rc = launch_dense_split(cur, kH, kH, nullptr, 0, 0, pk, nullptr, kH, 0, 0.f, nxt, kH, M, p->d_errflag, st);
    if (rc != 0) return rc;
    float* tswap = cur; cur = nxt; nxt = tswap;
  }

  // ---- modulator (modulated_siren.py:325-343): h_0 = relu(A_0 z + c_0), h_i = relu(A_i [h_{i-1}, z] + c_i)
  float* G = ws + w.G;
  float* carry = ws + w.carry;
  float* dzlat = ws + w.dzlat;
  float* tmpz = ws + w.tmpz;
  const float* z = ws + w.z;
  MRINR_CUDA(cudaMemsetAsync(dzlat, 0, (size_t)B * Z * sizeof(float), st));
  for (int i = L - 1; i >= 0; --i) {
    const float* hi = ws + w.mods + (size_t)i * B * kH;
    EW_LAUNCH(relu_bwd_kernel, (long long)B * kH, ws + w.dmods + (size_t)i * B * kH, i == L - 1 ? nullptr : carry, hi,
              (long long)B * kH, G);
    if ((rc = colsum_atomic(G, kH, B, kH, g_mod_b ? g_mod_b[i] : nullptr, st)) != 0) return rc;
    const float* Ai = v->d_mod_weight[i];
    const int ldA = i == 0 ? Z : kH + Z;
    float* dAi = g_mod_w ? g_mod_w[i] : nullptr;
    if (i > 0) {
      const float* hprev = ws + w.mods + (size_t)(i - 1) * B * kH;
      if (dAi && (rc = gemm_tn_atomic(G, kH, kH, hprev, kH, kH, B, dAi, ldA, st)) != 0) return rc;
      if (dAi && (rc = gemm_tn_atomic(G, kH, kH, z, Z, Z, B, dAi + kH, ldA, st)) != 0) return rc;
      // carry = G A_i[:, :H]   (packed operand rows = k < H, K index = n)
      if ((rc = pack_split_strided(Ai, kH, kH, 1, ldA, pk, st)) != 0) return rc;
      if ((rc = launch_dense_split(G, kH, kH, nullptr, 0, 0, pk, nullptr, kH, 0, 0.f, carry, kH, B, p->d_errflag, st)) != 0) return rc;
      // dz += G A_i[:, H:]
      if ((rc = pack_split_strided(Ai + kH, Z, kH, 1, ldA, pk, st)) != 0) return rc;
      if ((rc = launch_dense_split(G, kH, kH, nullptr, 0, 0, pk, nullptr, Z, 0, 0.f, tmpz, Z, B, p->d_errflag, st)) != 0) return rc;
    } else {
      if (dAi && (rc = gemm_tn_atomic(G, kH, kH, z, Z, Z, B, dAi, ldA, st)) != 0) return rc;
      if ((rc = pack_split_strided(Ai, Z, kH, 1, ldA, pk, st)) != 0) return rc;
      if ((rc = launch_dense_split(G, kH, kH, nullptr, 0, 0, pk, nullptr, Z, 0, 0.f, tmpz, Z, B, p->d_errflag, st)) != 0) return rc;
    }
    EW_LAUNCH(add_kernel, (long long)B * Z, dzlat, tmpz, (long long)B * Z);
  }

  // ---- encoder (siren_encoder.py:503-512): Linear(64, Z) <- LeakyReLU <- Conv2d(32,64,8) <- LeakyReLU <- conv2 <- conv1
  const float* c2 = ws + w.enc;                        // [B][2048]  conv2 output (after LeakyReLU)
  const float* c3 = ws + w.enc + (size_t)B * 2048;     // [B][64]    conv3 output (after LeakyReLU)
  float* dc3 = ws + w.dc3;
  float* dc2 = ws + w.dc2;
  if (grads->d_enc_fc_weight)
    if ((rc = gemm_tn_atomic(dzlat, Z, Z, c3, 64, 64, B, (float*)grads->d_enc_fc_weight, 64, st)) != 0) return rc;
  if ((rc = colsum_atomic(dzlat, Z, B, Z, (float*)grads->d_enc_fc_bias, st)) != 0) return rc;
  // dc3 = dz W_fc  (W_fc [Z][64]: packed operand rows = 64 outputs, K index = z)
  if ((rc = pack_split_strided(v->d_enc_fc_weight, 64, Z, 1, 64, pk, st)) != 0) return rc;
  if ((rc = launch_dense_split(dzlat, Z, Z, nullptr, 0, 0, pk, nullptr, 64, 0, 0.f, dc3, 64, B, p->d_errflag, st)) != 0) return rc;
  EW_LAUNCH(lrelu_bwd_kernel, (long long)B * 64, dc3, c3, (long long)B * 64, 0.2f);
  if (grads->d_enc_conv3_weight)
    if ((rc = gemm_tn_atomic(dc3, 64, 64, c2, 2048, 2048, B, (float*)grads->d_enc_conv3_weight, 2048, st)) != 0) return rc;
  if ((rc = colsum_atomic(dc3, 64, B, 64, (float*)grads->d_enc_conv3_bias, st)) != 0) return rc;
  // dc2 = dc3 W_3  (W_3 [64][2048]): 8 column blocks of 256
  for (int blk = 0; blk < 8; ++blk) {
    if ((rc = pack_split_strided(v->d_enc_conv3_weight + blk * 256, 256, 64, 1, 2048, pk, st)) != 0) return rc;
    if ((rc = launch_dense_split(dc3, 64, 64, nullptr, 0, 0, pk, nullptr, 256, 0, 0.f, dc2 + blk * 256, 2048, B,
                                 p->d_errflag, st)) != 0) return rc;
  }
  EW_LAUNCH(lrelu_bwd_kernel, (long long)B * 2048, dc2, c2, (long long)B * 2048, 0.2f);
  {
    const size_t smem = (size_t)kConvSmemFloats * sizeof(float);
    MRINR_SMEM_OPT_IN(encoder_conv_bwd_kernel, smem);
    const unsigned grid = (unsigned)(B < 148 * 2 ? B : 148 * 2);
    encoder_conv_bwd_kernel<<<grid, 256, smem, st>>>(d_tiles, dc2, B, v->d_enc_conv1_weight, v->d_enc_conv1_bias,
                                                     v->d_enc_conv2_weight, (float*)grads->d_enc_conv1_weight,
                                                     (float*)grads->d_enc_conv1_bias, (float*)grads->d_enc_conv2_weight,
                                                     (float*)grads->d_enc_conv2_bias);
    count_launch();
    if ((rc = check_launch("encoder_conv_bwd")) != 0) return rc;
  }
  // ---- undo the gradient scale (exact: a power of two)
  if (grad_scale != 1.f) {
    const float inv = 1.f / grad_scale;
    const int H = kH;
    for (int l = 0; l < L; ++l) {
      if (g_net_w && g_net_w[l]) EW_LAUNCH(scale_kernel, (long long)H * (l == 0 ? 2 : H), g_net_w[l], (long long)H * (l == 0 ? 2 : H), inv);
      if (g_net_b && g_net_b[l]) EW_LAUNCH(scale_kernel, (long long)H, g_net_b[l], (long long)H, inv);
      if (g_mod_w && g_mod_w[l]) EW_LAUNCH(scale_kernel, (long long)H * (l == 0 ? Z : H + Z), g_mod_w[l], (long long)H * (l == 0 ? Z : H + Z), inv);
      if (g_mod_b && g_mod_b[l]) EW_LAUNCH(scale_kernel, (long long)H, g_mod_b[l], (long long)H, inv);
    }
    if (grads->d_last_weight) EW_LAUNCH(scale_kernel, (long long)H, (float*)grads->d_last_weight, (long long)H, inv);
    if (grads->d_last_bias) EW_LAUNCH(scale_kernel, 1LL, (float*)grads->d_last_bias, 1LL, inv);
    float* enc[8] = {(float*)grads->d_enc_conv1_weight, (float*)grads->d_enc_conv1_bias, (float*)grads->d_enc_conv2_weight,
                     (float*)grads->d_enc_conv2_bias,   (float*)grads->d_enc_conv3_weight, (float*)grads->d_enc_conv3_bias,
                     (float*)grads->d_enc_fc_weight,    (float*)grads->d_enc_fc_bias};
    const long long encn[8] = {144, 16, 4608, 32, 64LL * 2048, 64, (long long)Z * 64, Z};
    for (int i = 0; i < 8; ++i)
      if (enc[i]) EW_LAUNCH(scale_kernel, encn[i], enc[i], encn[i], inv);
  }
  return 0;
}
