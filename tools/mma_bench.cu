// Micro-benchmark: sustained tcgen05.mma issue rate (cycles per M x 256 x 16 f16 MMA) for the operand layouts
// and CTA-group modes considered in DESIGN.md.  No epilogue, operands are zeros; one persistent CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu && ./mma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "../mri_inr_b200/csrc/tc_ptx.cuh"

namespace mrinr {
void set_error(const char*, ...) {}
void count_launch(int) {}
int check_launch(const char*) { return 0; }
}  // namespace mrinr
using namespace mrinr;

__device__ __forceinline__ uint64_t make_desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// mode: 0 = cta_group::1 no swizzle, 1 = cta_group::1 swizzle128, 2 = cta_group::2 no swizzle, 3 = cta_group::2 swizzle128
template <int PAIR>
__global__ void __launch_bounds__(288, 1) bench_kernel(int swz, int iters, int busy_kind, int commit_every, int commit_kind, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t s_bar;
  __shared__ uint64_t s_bar2;
  __shared__ uint64_t s_bar3;
  __shared__ float s_sink[288];
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (65536 + 131072) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&s_bar), 1); mbar_init(smem_u32(&s_bar2), 1); mbar_init(smem_u32(&s_bar3), 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 0) { if (PAIR) tmem_alloc_pair(smem_u32(&s_tmem), 512); else tmem_alloc(smem_u32(&s_tmem), 512); }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 65536);
  const bool leader = !PAIR || cluster_ctarank() == 0;
  __shared__ uint64_t s_ring[8];
  if (threadIdx.x < 8) mbar_init(smem_u32(&s_ring[threadIdx.x]), 1);
  __syncthreads();
  const bool elect_style = (commit_kind == 11);
  bool issuer = (threadIdx.x == 256);
  if (elect_style && warp == 8) {
    uint32_t is_elected = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(is_elected));
    issuer = is_elected != 0;
  }
  if (issuer && leader) {
    const uint32_t idesc = make_idesc(0, PAIR ? 256 : 128, 256);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        uint64_t ad, bd;
        if (!swz) {
          ad = make_desc_sw(sA + k * 4096, 2048, 128, 0);
          bd = PAIR ? make_desc_sw(sB + k * 4096, 2048, 128, 0) : make_desc_sw(sB + k * 8192, 4096, 128, 0);
        } else {
          // swizzle-128B, K-major: 64-wide K blocks; A block = 16 KB, B block = 32 KB (16 KB per CTA of a pair)
          ad = make_desc_sw(sA + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, 2);
          bd = make_desc_sw(sB + (k >> 2) * (PAIR ? 16384 : 32768) + (k & 3) * 32, 16, 1024, 2);
        }
        const uint32_t dcol = (commit_kind == 12) ? 0u : (uint32_t)(it & 1) * 256u;
        const uint32_t accf = (commit_kind == 12) ? 1u : (uint32_t)(k != 0);
        if (PAIR) umma_f16_pair(tmem + dcol, ad, bd, idesc, accf);
        else umma_f16(tmem + dcol, ad, bd, idesc, accf);
        if (commit_every && (k % commit_every) == commit_every - 1) {
          if (!PAIR) umma_commit(smem_u32(&s_bar2));
          else if (commit_kind == 0) umma_commit_pair(smem_u32(&s_bar2), 1);
          else if (commit_kind == 1) umma_commit_pair(smem_u32(&s_bar2), 3);
          else if (commit_kind == 10) umma_commit_pair(smem_u32(&s_ring[(it * 16 + k) / commit_every % 8]), 3);
          else if (commit_kind == 11 || commit_kind == 12) umma_commit_pair(smem_u32(&s_bar2), 3);
          else if (commit_kind == 2)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar2)) : "memory");
          else { umma_commit_pair(smem_u32(&s_bar2), 3); umma_commit_pair(smem_u32(&s_bar3), 3); }
        }
      }
    }
    if (PAIR) umma_commit_pair(smem_u32(&s_bar), 1); else umma_commit(smem_u32(&s_bar));
    mbar_wait(smem_u32(&s_bar), 0, nullptr, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (threadIdx.x < 256 && busy_kind == 1) {
    // shared-memory store traffic from 8 warps while the MMAs run (the epilogue's operand stores)
    uint4* p = reinterpret_cast<uint4*>(smem + 65536 + 131072 - 65536) + threadIdx.x;
    for (int i = 0; i < iters * 48; ++i) { p[(i & 15) * 256] = make_uint4(i, i, i, i); }
  } else if (threadIdx.x < 256 && busy_kind == 2) {
    // TMEM loads from 8 warps (the epilogue's accumulator reads), from the accumulator NOT being written
    float acc = 0.f;
    for (int i = 0; i < iters * 4; ++i) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (((i >> 2) & 1) ^ 1) * 256 + (warp >> 2) * 128 + (i & 3) * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += __uint_as_float(v[j]);
    }
    s_sink[threadIdx.x] = acc;
  } else if (threadIdx.x < 256 && busy_kind == 3) {
    float x = threadIdx.x * 0.001f, acc = 0.f;
    for (int i = 0; i < iters * 256; ++i) { acc += __sinf(x); x += 0.01f; }
    s_sink[threadIdx.x] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  if (warp == 0) { if (PAIR) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512); }
}

template <int PAIR>
static void run(int swz, int busy, int commit_every, int commit_kind, long long* d_out) {
  const int iters = 400;
  const size_t smem = 65536 + 131072;
  cudaFuncSetAttribute(bench_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(288); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = PAIR ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaLaunchKernelEx(&cfg, bench_kernel<PAIR>, swz, iters, busy, commit_every, commit_kind, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
  }
  long long cyc = 0;
  cudaMemcpy(&cyc, d_out, sizeof(cyc), cudaMemcpyDeviceToHost);
  const char* kinds[] = {"idle", "sts", "tmem-ld", "mufu"};
  const char* ck[] = {"mcast mask1", "mcast mask3", "non-mcast", "2x mcast mask3", "", "", "", "", "", "", "rotating barriers", "elect.sync issuer", "same D, accumulate"};
  printf("cta_group::%d %-10s other warps: %-8s commit every %2d (%s) : %.1f cycles per MMA (M=%d N=256 K=16)\n", PAIR ? 2 : 1,
         swz ? "swizzle128" : "no-swizzle", kinds[busy], commit_every, PAIR ? ck[commit_kind] : "cta_group::1", (double)cyc / (iters * 16), PAIR ? 256 : 128);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  run<1>(0, 0, 16, 10, d_out);
  run<1>(0, 0, 4, 10, d_out);
  run<1>(0, 0, 16, 11, d_out);
  run<1>(0, 0, 4, 11, d_out);
  run<1>(0, 0, 16, 12, d_out);
  run<1>(0, 0, 4, 12, d_out);
  run<1>(0, 0, 0, 12, d_out);
  run<0>(0, 0, 0, 0, d_out);
  run<0>(0, 0, 4, 0, d_out);
  run<0>(0, 0, 16, 0, d_out);
  run<0>(0, 0, 1, 0, d_out);
  for (int ck = 0; ck < 4; ++ck) {
    run<1>(0, 0, 4, ck, d_out);
    run<1>(0, 0, 16, ck, d_out);
  }
  run<1>(0, 0, 1, 1, d_out);
  run<1>(0, 2, 16, 1, d_out);
  run<1>(0, 3, 16, 1, d_out);
  return 0;
}
