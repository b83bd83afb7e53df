// fp32 CUDA-core kernels: weight re-tiling, the patch-independent layer-0 table, the modulator
// (all layers, one launch) and the exact-mode (MRINR_PREC_FP32) synthesis kernel.
//
// Both dense kernels use the same scheme: a CTA owns TM=32 rows (patches for the modulator,
// coordinates for the synthesis net), one thread per output column, activations ping-pong between
// two shared-memory buffers across layers so nothing but the final result goes back to HBM; weights
// are read through L1/L2 from a transposed [K][H] copy so that a warp's loads are one 128-byte line.
#include "common.cuh"

namespace mrinr {

constexpr int TM = 32;

// ---- pack-time kernels ----------------------------------------------------------------------------
// wT[k][n] = W[n][k]
__global__ void transpose_kernel(const float* __restrict__ w, int N, int K, float* __restrict__ wT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const int k = i / N, n = i - k * N;
  wT[i] = w[(long long)n * K + k];
}

// UMMA K-major / no-swizzle operand layout: out[kc][n][e] = cvt(W[n][8*kc + e]), kc = k/8.
// A K=16 MMA step reads two consecutive kc slabs; slab stride (LBO) = H*16 bytes, 8-row group
// stride (SBO) = 128 bytes.  See siren_tc.cu.
__global__ void pack_w16_kernel(const float* __restrict__ w, int H, int use_bf16, uint16_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * H) return;
  const int e = i & 7;
  const int n = (i >> 3) % H;
  const int kc = (i >> 3) / H;
  const float v = w[(long long)n * H + kc * 8 + e];
  uint16_t bits;
  if (use_bf16) {
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    bits = *reinterpret_cast<const uint16_t*>(&b);
  } else {
    const __half h = __float2half_rn(v);
    bits = *reinterpret_cast<const uint16_t*>(&h);
  }
  out[i] = bits;
}

// cta_group::2 variant: CTA rank r of a pair holds output rows n in [128r, 128r+128) of every layer:
// out[r][kc][n'][e] = cvt(W[128r + n'][8*kc + e]).  A K=64 slab of one rank is 16 KB contiguous.
__global__ void pack_w16_pair_kernel(const float* __restrict__ w, int H, int use_bf16, uint16_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * H) return;
  const int half_rows = H / 2;
  const int e = i & 7;
  const int np = (i >> 3) % half_rows;
  const int kc = ((i >> 3) / half_rows) % (H / 8);
  const int r = (i >> 3) / half_rows / (H / 8);
  const float v = w[(long long)(r * half_rows + np) * H + kc * 8 + e];
  uint16_t bits;
  if (use_bf16) {
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    bits = *reinterpret_cast<const uint16_t*>(&b);
  } else {
    const __half h = __float2half_rn(v);
    bits = *reinterpret_cast<const uint16_t*>(&h);
  }
  out[i] = bits;
}

// cta_group::2 + bias-in-MMA variant: per (rank r): [kc 0..31][n' 0..127][8] weights as above, followed by one
// extra K=16 step [kc 32..33][n'][8] whose first two K slots hold the bias split into hi + lo 16-bit parts
// (b = hi + lo to ~2^-22 relative); the A operand has ones in those two slots, so D = A W^T + b.
// out is [(2)][34][128][8] = 2 x 69632 bytes per layer.
__global__ void pack_w16_pair_bias_kernel(const float* __restrict__ w, const float* __restrict__ bias, int H,
                                          int use_bf16, uint16_t* __restrict__ out, int part) {
  const int per_rank = (H / 8 + 2) * (H / 2) * 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * per_rank) return;
  const int r = i / per_rank;
  const int j = i - r * per_rank;
  const int e = j & 7;
  const int np = (j >> 3) % (H / 2);
  const int kc = (j >> 3) / (H / 2);
  const int n = r * (H / 2) + np;
  float v = 0.f;
  if (kc < H / 8) {
    v = w[(long long)n * H + kc * 8 + e];
    if (part == 1) v -= __half2float(__float2half_rn(v));      // fp16x3 mode: the residual half of W = hi + lo
  } else if (part == 0 && kc == H / 8 && e < 2 && bias != nullptr) {
    const float b = bias[n];
    float hi;
    if (use_bf16) hi = __bfloat162float(__float2bfloat16_rn(b)); else hi = __half2float(__float2half_rn(b));
    v = (e == 0) ? hi : (b - hi);
  }
  uint16_t bits;
  if (use_bf16) {
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    bits = *reinterpret_cast<const uint16_t*>(&b);
  } else {
    const __half h = __float2half_rn(v);
    bits = *reinterpret_cast<const uint16_t*>(&h);
  }
  out[i] = bits;
}

__device__ __forceinline__ float act_exact(float pre, float w0, int activation) {
  // Sine: modulated_siren.py:54 ; Morlet: :80 (Gaussian on the un-scaled pre-activation)
  const float s = sinf(w0 * pre);
  return activation == MRINR_ACT_MORLET ? s * expf(-0.5f * (pre * pre)) : s;
}

// table0[c][j] = act_0(W_0[j,:] . g_c + b_0[j])  -- every patch shares the coordinates
// (modulated_siren.py:448), so layer 0 before modulation is patch independent.
__global__ void layer0_table_kernel(const float* __restrict__ grid, const float* __restrict__ w0w,
                                    const float* __restrict__ b0, int C, int H, float w0_initial,
                                    int activation, float* __restrict__ table) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * H) return;
  const int c = i / H, j = i - c * H;
  const float gx = grid[2 * c], gy = grid[2 * c + 1];
  float pre = __fmaf_rn(gy, w0w[2 * j + 1], __fmul_rn(gx, w0w[2 * j]));
  if (b0) pre = __fadd_rn(pre, b0[j]);
  table[i] = act_exact(pre, w0_initial, activation);
}

// ---- shared dense step ------------------------------------------------------------------------------
// acc[r] += sum_k s_in[r*ld + k] * wT[k*H + j]   (k ascending, fp32 FFMA)
__device__ __forceinline__ void dense_accumulate(const float* __restrict__ s_in, int ld, int K,
                                                 const float* __restrict__ wT, int H, int j, float (&acc)[TM]) {
  for (int k = 0; k < K; k += 4) {
    const float w0 = __ldg(wT + (long long)(k + 0) * H + j);
    const float w1 = __ldg(wT + (long long)(k + 1) * H + j);
    const float w2 = __ldg(wT + (long long)(k + 2) * H + j);
    const float w3 = __ldg(wT + (long long)(k + 3) * H + j);
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(s_in + r * ld + k);
      acc[r] = fmaf(x.x, w0, acc[r]);
      acc[r] = fmaf(x.y, w1, acc[r]);
      acc[r] = fmaf(x.z, w2, acc[r]);
      acc[r] = fmaf(x.w, w3, acc[r]);
    }
  }
}

// ---- modulator: modulated_siren.py:325-343 ------------------------------------------------------------
// smem: z [TM][Z] | h [2][TM][H]
__global__ void __launch_bounds__(512)
modulator_kernel(const float* __restrict__ latent, long long B, int Z, int H, int L,
                 const float* __restrict__ wT, const float* __restrict__ bias, float* __restrict__ mods) {
  extern __shared__ __align__(16) float smem[];
  float* s_z = smem;
  float* s_h = smem + TM * Z;
  const int j = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * TM;
  for (int i = threadIdx.x; i < TM * Z; i += blockDim.x) {
    const int r = i / Z;
    s_z[i] = (row0 + r < B) ? latent[(row0 + r) * Z + (i - r * Z)] : 0.f;
  }
  __syncthreads();
  const float* w = wT;
  for (int l = 0; l < L; ++l) {
    float acc[TM];
#pragma unroll
    for (int r = 0; r < TM; ++r) acc[r] = 0.f;
    float* h_out = s_h + (l & 1) * TM * H;
    if (l > 0) {
      // cat((x, z), dim=1): previous hidden first, then the latent (modulated_siren.py:341)
      dense_accumulate(s_h + ((l - 1) & 1) * TM * H, H, H, w, H, j, acc);
      w += (long long)H * H;
    }
    dense_accumulate(s_z, Z, Z, w, H, j, acc);
    w += (long long)Z * H;
    const float b = bias[l * H + j];
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      const float v = fmaxf(acc[r] + b, 0.f);
      h_out[r * H + j] = v;
      if (row0 + r < B) mods[((long long)l * B + row0 + r) * H + j] = v;
    }
    __syncthreads();
  }
}

// ---- modulator, register-tiled version (H == 256, Z <= 256): 64 patches per CTA, 8 x 8 outputs per thread -------
// Activations live in shared memory as [k][patch] (row stride 68 floats) so that the 8 patch values a thread needs
// for one k are two 16-byte broadcast loads; a thread's 8 output columns are tx + 32 j so that its weight loads
// (wT[k][col]) and its global stores are 128-byte coalesced per warp.  64 FFMA per 2 LDS.128 + 8 LDG.
constexpr int MT_P = 64;        // patches per CTA
constexpr int MT_LD = 68;       // smem row stride (floats)

__global__ void __launch_bounds__(256, 1)
modulator_tiled_kernel(const float* __restrict__ latent, long long B, int Z, int L,
                       const float* __restrict__ wT, const float* __restrict__ bias, float* __restrict__ mods) {
  constexpr int H = 256;
  extern __shared__ __align__(16) float smem[];
  float* s_z = smem;                       // [Z][68]
  float* s_h0 = smem + 256 * MT_LD;        // [256][68]
  float* s_h1 = s_h0 + 256 * MT_LD;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long p0 = (long long)blockIdx.x * MT_P;
  // latent [patch][Z] -> s_z[k][patch]
  for (int i = threadIdx.x; i < MT_P * Z; i += 256) {
    const int p = i / Z, k = i - p * Z;
    s_z[k * MT_LD + p] = (p0 + p < B) ? latent[(p0 + p) * Z + k] : 0.f;
  }
  __syncthreads();
  const float* w = wT;
  for (int l = 0; l < L; ++l) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const float* s_prev = (l & 1) ? s_h0 : s_h1;     // layer l-1 wrote buffer (l-1)&1: 0 -> s_h0, 1 -> s_h1
    float* s_next = (l & 1) ? s_h1 : s_h0;
    const int n_seg = (l == 0) ? 1 : 2;
    for (int seg = 0; seg < n_seg; ++seg) {
      const bool zseg = (l == 0) || (seg == 1);        // cat((h, z)): hidden first, then latent (:341)
      const float* s_in = zseg ? s_z : s_prev;
      const int K = zseg ? Z : H;
#pragma unroll 2
      for (int k = 0; k < K; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(s_in + k * MT_LD + ty * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(s_in + k * MT_LD + ty * 8 + 4);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float b[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = __ldg(w + (long long)k * H + tx + 32 * j);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      w += (long long)K * H;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = tx + 32 * j;
      const float bj = bias[l * H + col];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float v = fmaxf(acc[i][j] + bj, 0.f);
        const int p = ty * 8 + i;
        s_next[col * MT_LD + p] = v;
        if (p0 + p < B) mods[((long long)l * B + p0 + p) * H + col] = v;
      }
    }
    __syncthreads();
  }
}

// ---- exact-mode synthesis kernel: SirenNet.forward, modulated_siren.py:215-233 -------------------------
// smem: h [2][TM][H] | row bookkeeping
__global__ void __launch_bounds__(512)
siren_fp32_kernel(const float* __restrict__ table0, const float* __restrict__ wT, const float* __restrict__ bias,
                  const float* __restrict__ last_w, const float* __restrict__ last_b,
                  const float* __restrict__ mods, const int32_t* __restrict__ idx,
                  const int32_t* __restrict__ nactive, long long B, int C, int H, int L, float w0,
                  int activation, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* s_h = smem;
  int* s_patch = reinterpret_cast<int*>(smem + 2 * TM * H);   // original patch index per row, -1 = padding
  int* s_c = s_patch + TM;
  const long long n_act = nactive ? (long long)*nactive : B;
  const long long total_rows = n_act * C;
  const long long n_tiles = (total_rows + TM - 1) / TM;
  const int j = threadIdx.x;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * TM;
    __syncthreads();
    if (threadIdx.x < TM) {
      const long long R = row0 + threadIdx.x;
      int patch = -1, c = 0;
      if (R < total_rows) {
        const long long pc = R / C;
        c = (int)(R - pc * C);
        patch = idx ? idx[pc] : (int)pc;
      }
      s_patch[threadIdx.x] = patch;
      s_c[threadIdx.x] = c;
    }
    __syncthreads();
    // layer 0 from the table, then modulation (modulated_siren.py:231)
    for (int r = 0; r < TM; ++r) {
      const int patch = s_patch[r];
      float v = 0.f;
      if (patch >= 0) v = table0[(long long)s_c[r] * H + j] * mods[((long long)0 * B + patch) * H + j];
      s_h[r * H + j] = v;
    }
    __syncthreads();
    for (int l = 1; l < L; ++l) {
      float acc[TM];
#pragma unroll
      for (int r = 0; r < TM; ++r) acc[r] = 0.f;
      const float* h_in = s_h + ((l - 1) & 1) * TM * H;
      float* h_out = s_h + (l & 1) * TM * H;
      dense_accumulate(h_in, H, H, wT + (long long)(l - 1) * H * H, H, j, acc);
      const float b = bias[l * H + j];
#pragma unroll
      for (int r = 0; r < TM; ++r) {
        const int patch = s_patch[r];
        float v = 0.f;
        if (patch >= 0) v = act_exact(acc[r] + b, w0, activation) * mods[((long long)l * B + patch) * H + j];
        h_out[r * H + j] = v;
      }
      __syncthreads();
    }
    // output layer: always sine, not modulated (modulated_siren.py:211-213, :233)
    const float* h_fin = s_h + ((L - 1) & 1) * TM * H;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int r = w; r < TM; r += nw) {
      float s = 0.f;
      for (int k = lane; k < H; k += 32) s = fmaf(h_fin[r * H + k], __ldg(last_w + k), s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const int patch = s_patch[r];
      if (lane == 0 && patch >= 0) out[(long long)patch * C + s_c[r]] = sinf(w0 * (s + (last_b ? *last_b : 0.f)));
    }
  }
}

int run_transpose(const float* w, int N, int K, float* wT, cudaStream_t st) {
  transpose_kernel<<<(N * K + 255) / 256, 256, 0, st>>>(w, N, K, wT);
  count_launch();
  return check_launch("transpose");
}

int run_pack_w16(const float* w, int H, int use_bf16, uint16_t* out, cudaStream_t st) {
  pack_w16_kernel<<<(H * H + 255) / 256, 256, 0, st>>>(w, H, use_bf16, out);
  count_launch();
  return check_launch("pack_w16");
}

int run_pack_w16_pair(const float* w, int H, int use_bf16, uint16_t* out, cudaStream_t st) {
  pack_w16_pair_kernel<<<(H * H + 255) / 256, 256, 0, st>>>(w, H, use_bf16, out);
  count_launch();
  return check_launch("pack_w16_pair");
}

int run_pack_w16_pair_bias(const float* w, const float* bias, int H, int use_bf16, uint16_t* out, cudaStream_t st,
                           int part) {
  const int n = 2 * (H / 8 + 2) * (H / 2) * 8;
  pack_w16_pair_bias_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, bias, H, use_bf16, out, part);
  count_launch();
  return check_launch("pack_w16_pair_bias");
}

__global__ void table16_kernel(const float* __restrict__ table, long long n, int use_bf16, uint16_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (use_bf16) {
    const __nv_bfloat16 b = __float2bfloat16_rn(table[i]);
    out[i] = *reinterpret_cast<const uint16_t*>(&b);
  } else {
    const __half h = __float2half_rn(table[i]);
    out[i] = *reinterpret_cast<const uint16_t*>(&h);
  }
}

int run_table16(const float* table, long long n, int use_bf16, uint16_t* out, cudaStream_t st) {
  table16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(table, n, use_bf16, out);
  count_launch();
  return check_launch("table16");
}

int run_layer0_table(const float* grid, const float* w, const float* b, int C, int H, float w0_initial,
                     int activation, float* table, cudaStream_t st) {
  layer0_table_kernel<<<(C * H + 255) / 256, 256, 0, st>>>(grid, w, b, C, H, w0_initial, activation, table);
  count_launch();
  return check_launch("layer0_table");
}

int launch_siren_fp32(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx,
                      const int32_t* d_nactive, int64_t B, float* d_out, cudaStream_t st) {
  const size_t smem = (size_t)2 * TM * p->H * sizeof(float) + 2 * TM * sizeof(int);
  // once per device, sized for the largest supported configuration (H = 512)
  MRINR_SMEM_OPT_IN(siren_fp32_kernel, (size_t)2 * TM * 512 * sizeof(float) + 2 * TM * sizeof(int));
  const long long n_tiles = (B * p->C + TM - 1) / TM;
  long long grid = (long long)p->num_sms * 2;
  if (grid > n_tiles) grid = n_tiles;
  if (grid < 1) grid = 1;
  siren_fp32_kernel<<<(unsigned)grid, p->H, smem, st>>>(p->d_table0, p->d_net_wT, p->d_net_bias, p->d_last_w,
                                                      p->d_last_b, d_mods, d_idx, d_nactive, B, p->C, p->H, p->L,
                                                      p->w0, p->activation, d_out);
  count_launch();
  return check_launch("siren_fp32");
}

}  // namespace mrinr

using namespace mrinr;

extern "C" int mrinr_modulator_forward(const MrinrPacked* p, const float* d_latent, int64_t B, float* d_mods,
                                       void* stream) {
  if (B == 0) return 0;
  MRINR_REQUIRE(p && d_latent && d_mods, MRINR_E_ARG, "mrinr_modulator_forward: null pointer");
  MRINR_REQUIRE(B >= 0, MRINR_E_ARG, "mrinr_modulator_forward: negative batch");
  if (B == 0) return 0;
  if (p->mod_tc) {
    // tensor cores: split-fp16 products; layer l reads h_{l-1} = d_mods[l-1] and the latent
    MRINR_REQUIRE(aligned16(d_latent) && aligned16(d_mods), MRINR_E_ALIGN, "mrinr_modulator_forward: buffers must be 16-byte aligned");
    const size_t plane = (size_t)B * p->H;
    if (p->H == 256 && p->L <= 16) {
      // every layer in ONE launch: a CTA walks all layers of its 128 patches (dense_tc.cu: dense_chain_kernel)
      const float* a1[16]; const float* a2[16]; const uint16_t* w[16]; const float* bias[16]; float* c[16];
      long long lda1[16], lda2[16], ldc[16];
      int K1[16], K2[16];
      for (int l = 0; l < p->L; ++l) {
        a1[l] = l == 0 ? d_latent : d_mods + (size_t)(l - 1) * plane;
        lda1[l] = l == 0 ? p->Z : p->H;
        K1[l] = l == 0 ? p->Z : p->H;
        a2[l] = l == 0 ? nullptr : d_latent;
        lda2[l] = l == 0 ? 0 : p->Z;
        K2[l] = l == 0 ? 0 : p->Z;
        w[l] = p->d_mod_ws + p->mod_ws_off[l];
        bias[l] = p->d_mod_bias + (size_t)l * p->H;
        c[l] = d_mods + (size_t)l * plane;
        ldc[l] = p->H;
      }
      return launch_dense_chain256(p->L, a1, lda1, K1, a2, lda2, K2, w, bias, c, ldc, /*relu*/ 1, 0.f, B, p->d_errflag,
                                   (cudaStream_t)stream);
    }
    int rc = launch_dense_split(d_latent, p->Z, p->Z, nullptr, 0, 0, p->d_mod_ws + p->mod_ws_off[0], p->d_mod_bias, p->H,
                                /*relu*/ 1, 0.f, d_mods, p->H, B, p->d_errflag, (cudaStream_t)stream);
    for (int l = 1; l < p->L && rc == 0; ++l)
      rc = launch_dense_split(d_mods + (size_t)(l - 1) * plane, p->H, p->H, d_latent, p->Z, p->Z,
                              p->d_mod_ws + p->mod_ws_off[l], p->d_mod_bias + (size_t)l * p->H, p->H, 1, 0.f,
                              d_mods + (size_t)l * plane, p->H, B, p->d_errflag, (cudaStream_t)stream);
    return rc;
  }
  if (p->H == 256 && p->Z <= 256) {
    const size_t smem_t = (size_t)3 * 256 * MT_LD * sizeof(float);
    MRINR_SMEM_OPT_IN((modulator_tiled_kernel), (int)smem_t);
    const long long grid_t = (B + MT_P - 1) / MT_P;
    modulator_tiled_kernel<<<(unsigned)grid_t, 256, smem_t, (cudaStream_t)stream>>>(d_latent, B, p->Z, p->L, p->d_mod_wT,
                                                                                  p->d_mod_bias, d_mods);
    count_launch();
    return check_launch("modulator_tiled");
  }
  const size_t smem = ((size_t)TM * p->Z + (size_t)2 * TM * p->H) * sizeof(float);
  MRINR_REQUIRE(smem <= 227 * 1024, MRINR_E_UNSUPPORTED, "modulator: H=%d Z=%d needs %zu bytes of shared memory", p->H,
                p->Z, smem);
  MRINR_SMEM_OPT_IN(modulator_kernel, 227 * 1024);
  const long long grid = (B + TM - 1) / TM;
  modulator_kernel<<<(unsigned)grid, p->H, smem, (cudaStream_t)stream>>>(d_latent, B, p->Z, p->H, p->L, p->d_mod_wT,
                                                                        p->d_mod_bias, d_mods);
  count_launch();
  return check_launch("modulator");
}
