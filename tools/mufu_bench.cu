// Micro-benchmark: does MUFU.SIN (quarter rate) overlap with FMA-pipe instructions from the same / other warps of an
// SM sub-partition?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../mri_inr_b200/csrc/tc_ptx.cuh"
namespace mrinr { void set_error(const char*, ...) {} void count_launch(int) {} int check_launch(const char*) { return 0; } }
using namespace mrinr;

// MODE 0: K x FFMA   1: K x F2FP (cvt.rn.f16x2.f32)   2: K x STS.128   3: K x LDS.128   4: K x FMUL.RZ-like (mul.rz)
template <int K, int MODE>
__global__ void bench(float* out, int iters, long long* cycles, const float4* gsrc) {
  __shared__ uint4 sbuf[1024];
  __shared__ uint32_t s_tmem;
  uint32_t tm = 0;
  if (MODE == 6) {
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&s_tmem), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    tm = s_tmem + ((uint32_t)(((threadIdx.x >> 5) & 3) * 32) << 16);
  }
  float x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 0.001f + i; y[i] = 1.0f + i * 1e-3f; }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s;
      asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(x[i]));
      x[i] = s;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(y[(i + k) & 7]) : "f"(1.0001f));
        if (MODE == 1) { unsigned r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y[(i + k) & 7]), "f"(y[(i + k + 1) & 7])); y[(i + k) & 7] = __uint_as_float(r | 0x3f000000u); }
        if (MODE == 2) asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" :: "r"((unsigned)__cvta_generic_to_shared(&sbuf[threadIdx.x & 1023])), "f"(y[(i + k) & 7]) : "memory");
        if (MODE == 3) { float4 v; asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(&sbuf[(threadIdx.x + k) & 1023]))); y[(i + k) & 7] += v.x; }
        if (MODE == 5) { float4 v = __ldg(gsrc + ((i + k) & 63)); y[(i + k) & 7] += v.x; }
        if (MODE == 6) { uint32_t v[16]; tmem_ld16(tm + ((i * 16 + k * 64) & 255), v); tmem_ld_wait(); y[(i + k) & 7] += __uint_as_float(v[0]) * 0.f; }
        if (MODE == 4) asm volatile("mul.rz.f32 %0, %0, %1;" : "+f"(y[(i + k) & 7]) : "f"(1.0001f));
      }
      // MODE 7: the synthesis epilogue's mix per 8 sines: 8 FMUL + 4 F2FP, and K STS.128 per 16 sines (K = 0, 1, 2, 4)
      if (MODE == 7) {
        asm volatile("mul.f32 %0, %0, %1;" : "+f"(y[i]) : "f"(1.0001f));
        if (i & 1) { unsigned r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y[i]), "f"(y[i - 1])); y[i] = __uint_as_float(r | 0x3f000000u); }
      }
      if (MODE == 7 && K >= 1 && i == 7 && ((it & 1) || K >= 2)) {
        const int n = K >= 2 ? K / 2 : 1;
        for (int q = 0; q < n; ++q)
          asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" :: "r"((unsigned)__cvta_generic_to_shared(&sbuf[(threadIdx.x + q * 256) & 1023])), "f"(y[q & 7]) : "memory");
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += x[i] + y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  if (MODE == 6) { tc_fence_before(); __syncthreads(); tc_fence_after(); if (threadIdx.x < 32) tmem_dealloc(s_tmem, 512); }
}

template <int K, int MODE>
void run(int warps_per_smsp, float* d_out, long long* d_cyc) {
  const int iters = 2000;
  bench<K, MODE><<<148, warps_per_smsp * 4 * 32>>>(d_out, iters, d_cyc, (const float4*)d_out);
  bench<K, MODE><<<148, warps_per_smsp * 4 * 32>>>(d_out, iters, d_cyc, (const float4*)d_out);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
  // per SMSP: warps_per_smsp warps x iters x 8 groups of (1 MUFU(+FMUL.RZ) + K FFMA)
  const char* names[] = {"FFMA", "F2FP", "STS.128", "LDS.128", "FMUL.RZ", "LDG.128u", "LDTM.x16", "STS.128 per 16 sines (epilogue mix)"};
  printf("warps/SMSP %d, %2d %-8s per MUFU.SIN: %.2f cycles per group per SMSP\n", warps_per_smsp, K, names[MODE],
         (double)c / (iters * 8.0 * warps_per_smsp));
}

int main() {
  float* d_out; long long* d_cyc;
  cudaMalloc(&d_out, 148 * 1024 * 4); cudaMalloc(&d_cyc, 8);
  for (int w = 2; w <= 4; w *= 2) {
    run<0, 0>(w, d_out, d_cyc); run<2, 0>(w, d_out, d_cyc); run<4, 0>(w, d_out, d_cyc); run<8, 0>(w, d_out, d_cyc);
    run<1, 1>(w, d_out, d_cyc); run<2, 1>(w, d_out, d_cyc); run<4, 1>(w, d_out, d_cyc);
    run<1, 2>(w, d_out, d_cyc); run<2, 2>(w, d_out, d_cyc);
    run<1, 3>(w, d_out, d_cyc); run<2, 3>(w, d_out, d_cyc);
    run<2, 4>(w, d_out, d_cyc); run<4, 4>(w, d_out, d_cyc);
    run<1, 5>(w, d_out, d_cyc); run<2, 5>(w, d_out, d_cyc);
    run<1, 6>(w, d_out, d_cyc);
    run<0, 7>(w, d_out, d_cyc); run<1, 7>(w, d_out, d_cyc); run<2, 7>(w, d_out, d_cyc); run<4, 7>(w, d_out, d_cyc);
  }
  return 0;
}
