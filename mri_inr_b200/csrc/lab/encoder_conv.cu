// Patch encoder, convolutional front: FixedAutoencoder.encoder[0..3] (src/networks/encoding/siren_encoder.py:503-507)
//
//   x [B,1,32,32] -> Conv2d(1,16,3,stride 2,pad 1) -> LeakyReLU(0.2) -> Conv2d(16,32,3,stride 2,pad 1) -> LeakyReLU(0.2)
//     -> y [B,32,8,8]   (= the [B,2048] operand of Conv2d(32,64,8), which dense_tc.cu evaluates as a plain product)
//
// fp32 FFMA on the CUDA cores (664 kFLOP per patch, K = 9 and K = 144: too thin for the tensor core without an
// im2col expansion).  Persistent CTAs of 256 threads, 8 patches per pass, ONE WARP PER PATCH (4 patches per CTA and two CTAs per SM
// measured slower):
//   conv1: a thread owns 4 consecutive output columns x all 16 channels (64 accumulators, 27 input values), twice,
//   conv2: a thread owns 4 consecutive output columns x 16 channels (64 accumulators); per (ci, ky) it reads 9 input
//          values and 3 x 16 weights (four broadcast LDS.128 per tap) for 192 FFMA -- 0.16 shared-memory
//          instructions per FFMA (the first version, 4 x 8 outputs per thread, had 0.26 and was LSU-bound at 53 % of
//          the FFMA peak).
// Both convolutions keep their zero padding as a border in shared memory (only index -1 is ever touched: stride 2,
// pad 1, even sizes).  Summation order: taps in (ci, ky, kx) order, bias added last -- the result agrees with
// cuDNN / the CPU reference to ~1e-6.
#include "common.cuh"

namespace mrinr {
namespace enc {

constexpr int kP = 8;             // patches per pass (one warp each)
constexpr int kThreads = 32 * kP;
constexpr int kInLd = 36;         // input tile [33][36]: row / col index + 1 (border at 0)
constexpr int kInSz = 33 * kInLd;
constexpr int kC1Ld = 18;         // conv1 map [17][18] per channel
constexpr int kC1Sz = 17 * kC1Ld;
constexpr float kSlope = 0.2f;

struct Smem {
  float w2[16 * 9 * 32];          // [ci][ky][kx][co]
  float w1[9 * 16];               // [ky][kx][co]
  float b1[16];
  float b2[32];
  float in[kP][kInSz];
  float c1[kP][16][kC1Sz];
};

__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : x * kSlope; }

__global__ void __launch_bounds__(kThreads, 1)
encoder_conv_kernel(const float* __restrict__ patches, long long B, const float* __restrict__ w1, const float* __restrict__ b1,
                    const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem& S = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x;
  // ---- once per CTA: weights (re-ordered so that the output channel is the fastest index) and the zero borders ----
  for (int i = tid; i < 16 * 9 * 32; i += kThreads) {
    const int co = i & 31, tap = (i >> 5) % 9, ci = (i >> 5) / 9;
    S.w2[i] = w2[(co * 16 + ci) * 9 + tap];                 // Conv2d weight [co][ci][ky][kx]
  }
  for (int i = tid; i < 9 * 16; i += kThreads) S.w1[i] = w1[(i & 15) * 9 + (i >> 4)];
  if (tid < 16) S.b1[tid] = b1[tid];
  if (tid < 32) S.b2[tid] = b2[tid];
  for (int i = tid; i < kP * kInSz; i += kThreads) (&S.in[0][0])[i] = 0.f;
  for (int i = tid; i < kP * 16 * kC1Sz; i += kThreads) (&S.c1[0][0][0])[i] = 0.f;
  __syncthreads();

  const int p = tid >> 5;          // patch slot of this warp
  const int lane = tid & 31;
  const long long n_groups = (B + kP - 1) / kP;
  for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const long long b0 = grp * kP;
    // ---- stage kP input patches (coalesced 16-byte loads) ----
#pragma unroll
    for (int i = 0; i < kP * 1024 / (kThreads * 4); ++i) {
      const int e = (i * kThreads + tid) * 4;                // element index within the kP x 1024 block
      const int pp = e >> 10, r = (e >> 5) & 31, c = e & 31;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b0 + pp < B) v = __ldg(reinterpret_cast<const float4*>(patches + (b0 + pp) * 1024 + r * 32 + c));
      float* dst = &S.in[pp][(r + 1) * kInLd + c + 1];
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __syncthreads();
    // ---- conv1 + LeakyReLU: thread -> output row y, columns x0..x0+3, all 16 channels; two position groups ----
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int g = half * 32 + lane;
      const int y = g >> 2, x0 = (g & 3) * 4;
      float2 acc[4][8];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[j][c] = make_float2(0.f, 0.f);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float* row = &S.in[p][(2 * y + ky) * kInLd + 2 * x0];       // input row 2y+ky-1, col 2x0-1 (border + 1)
        float2 v[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) v[j] = make_float2(row[j], row[j]);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          float2 w[8];
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float4 t = *reinterpret_cast<const float4*>(&S.w1[(ky * 3 + kx) * 16 + c4 * 4]);
            w[c4 * 2] = make_float2(t.x, t.y); w[c4 * 2 + 1] = make_float2(t.z, t.w);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[j][c] = __ffma2_rn(v[2 * j + kx], w[c], acc[j][c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          S.c1[p][2 * c][(y + 1) * kC1Ld + x0 + j + 1] = lrelu(acc[j][c].x + S.b1[2 * c]);
          S.c1[p][2 * c + 1][(y + 1) * kC1Ld + x0 + j + 1] = lrelu(acc[j][c].y + S.b1[2 * c + 1]);
        }
    }
    __syncwarp();        // a patch's conv1 map is produced and consumed by the same warp
    // ---- conv2 + LeakyReLU: thread -> output row y, columns x0..x0+3, channels co0..co0+15 ----
    {
      const int co0 = (lane >> 4) * 16;
      const int y = lane & 7, x0 = ((lane >> 3) & 1) * 4;
      float2 acc[4][8];              // [output column][channel pair]: one fma.rn.f32x2 per pair (same rounding as fmaf)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[j][c] = make_float2(0.f, 0.f);
#pragma unroll 1
      for (int ci = 0; ci < 16; ++ci) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float* row = &S.c1[p][ci][(2 * y + ky) * kC1Ld + 2 * x0];
          float2 v[9];
#pragma unroll
          for (int j = 0; j < 9; ++j) v[j] = make_float2(row[j], row[j]);
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float* wp = &S.w2[((ci * 3 + ky) * 3 + kx) * 32 + co0];
            float2 w[8];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const float4 t = *reinterpret_cast<const float4*>(wp + c4 * 4);
              w[c4 * 2] = make_float2(t.x, t.y); w[c4 * 2 + 1] = make_float2(t.z, t.w);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int c = 0; c < 8; ++c) acc[j][c] = __ffma2_rn(v[2 * j + kx], w[c], acc[j][c]);
          }
        }
      }
      if (b0 + p < B) {
        float* dst = out + (b0 + p) * 2048 + y * 8 + x0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float ba = S.b2[co0 + 2 * c], bb = S.b2[co0 + 2 * c + 1];
          *reinterpret_cast<float4*>(dst + (co0 + 2 * c) * 64) = make_float4(
              lrelu(acc[0][c].x + ba), lrelu(acc[1][c].x + ba), lrelu(acc[2][c].x + ba), lrelu(acc[3][c].x + ba));
          *reinterpret_cast<float4*>(dst + (co0 + 2 * c + 1) * 64) = make_float4(
              lrelu(acc[0][c].y + bb), lrelu(acc[1][c].y + bb), lrelu(acc[2][c].y + bb), lrelu(acc[3][c].y + bb));
        }
      }
    }
    __syncthreads();     // the next pass overwrites S.in / S.c1
  }
}

}  // namespace enc

int launch_encoder_conv(const float* d_patches, long long B, const float* w1, const float* b1, const float* w2,
                        const float* b2, float* d_out, int num_sms, cudaStream_t st) {
  const size_t smem = sizeof(enc::Smem);
  MRINR_SMEM_OPT_IN((enc::encoder_conv_kernel), (int)smem);
  if (B <= 0) return 0;
  long long grid = (long long)num_sms;
  const long long groups = (B + enc::kP - 1) / enc::kP;
  if (grid > groups) grid = groups;
  enc::encoder_conv_kernel<<<(unsigned)grid, enc::kThreads, smem, st>>>(d_patches, B, w1, b1, w2, b2, d_out);
  count_launch();
  return check_launch("encoder_conv");
}

}  // namespace mrinr
