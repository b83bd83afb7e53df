// Micro-benchmark of the synthesis epilogue's arithmetic (registers only, no TMEM / shared memory):
// per 16 activations: NP of them take the FMA-pipe polynomial sine, 16 - NP take MUFU.SIN; then multiply by the
// modulation and pack to f16x2.   How many cycles per 16-activation chunk per SM sub-partition?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epi_bench epi_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float vsin(float x) { float y; asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float vmul(float a, float b) { float y; asm volatile("mul.f32 %0, %1, %2;" : "=f"(y) : "f"(a), "f"(b)); return y; }
__device__ __forceinline__ uint32_t vpack2(float lo, float hi) { uint32_t y; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }

constexpr float kInv2Pi = 0.15915494309189535f;
constexpr float kMagic = 12582912.0f;
constexpr float c1 = 6.283055782318115f, c3 = -41.331214904785156f, c5 = 81.36681365966797f, c7 = -74.4780044555664f,
                c9 = 32.781394958496094f;

// scalar polynomial sine times modulation
__device__ __forceinline__ float psin_mul(float x, float m) {
  const float t = fmaf(x, kInv2Pi, kMagic);
  const float n = t - kMagic;
  const float f = fmaf(x, kInv2Pi, -n);
  const float f2 = f * f;
  float p = fmaf(c9, f2, c7);
  p = fmaf(p, f2, c5);
  p = fmaf(p, f2, c3);
  p = fmaf(p, f2, c1);
  return p * (f * m);
}

struct f2 { float x, y; };
__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
// packed polynomial sine times modulation for two activations
__device__ __forceinline__ void psin_mul2(float x0, float x1, float m0, float m1, float& h0, float& h1) {
  const unsigned long long X = pk(x0, x1), K = pk(kInv2Pi, kInv2Pi), MG = pk(kMagic, kMagic), NMG = pk(-kMagic, -kMagic);
  const unsigned long long T = fma2(X, K, MG);
  const unsigned long long N = add2(T, NMG);
  float n0, n1; upk(N, n0, n1);
  const unsigned long long F = fma2(X, K, pk(-n0, -n1));
  const unsigned long long F2 = mul2(F, F);
  unsigned long long P = fma2(pk(c9, c9), F2, pk(c7, c7));
  P = fma2(P, F2, pk(c5, c5));
  P = fma2(P, F2, pk(c3, c3));
  P = fma2(P, F2, pk(c1, c1));
  const unsigned long long H = mul2(P, mul2(F, pk(m0, m1)));
  upk(H, h0, h1);
}

// MODE 0: scalar polynomial, 1: packed f32x2 polynomial, 2: mul by modulation with mul.f32x2 (MUFU for all sines)
template <int NP, int MODE>
__global__ void bench(float* out, int iters, long long* cycles, const float* src) {
  float x[16], m[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = src[(threadIdx.x + i) & 255]; m[i] = src[(threadIdx.x * 3 + i) & 255]; }
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float h[16];
    if (MODE == 2) {
      float s[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) s[i] = vsin(x[i]);
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const unsigned long long H = mul2(pk(s[i], s[i + 1]), pk(m[i], m[i + 1]));
        upk(H, h[i], h[i + 1]);
      }
    } else {
      // elements i with (i % 16) < NP... interleave: every (16/NP)-th pair goes to the polynomial
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        constexpr int NPP = NP / 2;                               // pairs taking the polynomial
        const bool poly = (NPP > 0) && ((i / 2) % (8 / (NPP > 0 ? NPP : 1)) == 0);
        if (poly) {
          if (MODE == 1) psin_mul2(x[i], x[i + 1], m[i], m[i + 1], h[i], h[i + 1]);
          else { h[i] = psin_mul(x[i], m[i]); h[i + 1] = psin_mul(x[i + 1], m[i + 1]); }
        } else {
          h[i] = vmul(vsin(x[i]), m[i]);
          h[i + 1] = vmul(vsin(x[i + 1]), m[i + 1]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const uint32_t p = vpack2(h[i], h[i + 1]);
      sink ^= p;
      x[i] = __uint_as_float((p & 0x007fffffu) | 0x3f800000u);      // feed back so nothing is loop-invariant
      x[i + 1] = x[i] + 0.25f;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(sink);
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int NP, int MODE>
void run(int warps_per_smsp, float* d_out, long long* d_cyc, const float* d_src) {
  const int iters = 4000;
  bench<NP, MODE><<<148, warps_per_smsp * 4 * 32>>>(d_out, iters, d_cyc, d_src);
  bench<NP, MODE><<<148, warps_per_smsp * 4 * 32>>>(d_out, iters, d_cyc, d_src);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
  const char* names[] = {"scalar poly", "f32x2 poly", "f32x2 modulation mul"};
  printf("warps/SMSP %d  %2d/16 %-22s: %7.2f cycles per 16-activation chunk per SMSP  (%.2f per activation-warp)\n",
         warps_per_smsp, NP, names[MODE], (double)c / (iters * (double)warps_per_smsp), (double)c / (iters * 16.0 * warps_per_smsp));
}

__global__ void check_poly(float* err) {
  float worst = 0.f;
  for (int i = threadIdx.x; i < 2000000; i += blockDim.x) {
    const float x = -40.f + 80.f * (float)i / 2000000.f;
    float h0, h1;
    psin_mul2(x, x + 0.5f, 1.f, 1.f, h0, h1);
    worst = fmaxf(worst, fabsf(h0 - (float)sin((double)x)));
    worst = fmaxf(worst, fabsf(psin_mul(x, 1.f) - (float)sin((double)x)));
  }
  atomicMax(reinterpret_cast<int*>(err), __float_as_int(worst));
}

int main() {
  float *d_out, *d_src; long long* d_cyc;
  cudaMalloc(&d_out, 148 * 1024 * 4); cudaMalloc(&d_cyc, 8); cudaMalloc(&d_src, 1024);
  float h[256];
  for (int i = 0; i < 256; ++i) h[i] = 0.01f * i - 1.0f;
  cudaMemcpy(d_src, h, 1024, cudaMemcpyHostToDevice);
  float* d_err; cudaMalloc(&d_err, 4); cudaMemset(d_err, 0, 4);
  check_poly<<<1, 256>>>(d_err);
  float e; cudaMemcpy(&e, d_err, 4, cudaMemcpyDeviceToHost);
  printf("polynomial sine, |x| <= 40: max abs error %.3e\n", e);
  for (int w = 2; w <= 4; w *= 2) {
    run<0, 0>(w, d_out, d_cyc, d_src);
    run<2, 0>(w, d_out, d_cyc, d_src); run<4, 0>(w, d_out, d_cyc, d_src); run<8, 0>(w, d_out, d_cyc, d_src); run<16, 0>(w, d_out, d_cyc, d_src);
    run<2, 1>(w, d_out, d_cyc, d_src); run<4, 1>(w, d_out, d_cyc, d_src); run<8, 1>(w, d_out, d_cyc, d_src); run<16, 1>(w, d_out, d_cyc, d_src);
    run<0, 2>(w, d_out, d_cyc, d_src);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
