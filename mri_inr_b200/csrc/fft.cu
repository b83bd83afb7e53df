// k-space front end: src/data/preprocessing.py:49-58 (load_mri_scan) --
//   apply_mask (random column mask) -> fastmri.ifft2c -> fastmri.complex_abs
// fastmri.ifft2c(x) = fftshift(ifft2(ifftshift(x), norm="ortho")) over the two image axes (fastmri/fftc.py; the
// package is an un-vendored dependency of the reference, requirements.txt:1 -- parity against it is UNPINNED and is
// anchored on torch.fft instead, tests/test_gpu_parity.py::test_kspace_front_end).
//
// Hand-written FFT, two HBM passes per image; length 320 (the reference's slices) runs register-resident as 16 x 20
// (fft320.cuh, below), any other product of 2, 3 and 5 as a mixed-radix (2, 3, 4, 5) Stockham FFT in shared memory:
//   pass 1 (rows):    k-space [N,H,W] complex -> mask x ifftshift folded into the load index -> W-point transform of
//                     8 rows per CTA -> fftshift folded into the store index -> workspace [N,H,W] complex
//   pass 2 (columns): 8 adjacent columns per CTA (64-byte row segments), H-point transform, shift on store, and either
//                     the complex result or its magnitude sqrt(re^2 + im^2) (one 4-byte store per pixel)
// Algorithmic HBM bytes per pixel: 8 in + 8 out (pass 1) + 8 in + 4 out (pass 2) = 28 B (+1 B/pixel-row of mask).
// The twiddle table exp(+-2 pi i j / n) is built per CTA with sincospif (exact argument reduction), so the
// transform is accurate to a few fp32 ulps of the largest element.
#include "common.cuh"
#include "fft320.cuh"

namespace mrinr {
namespace fft {

constexpr int kMaxN = 1024;
constexpr int kG = 8;              // sequences per CTA
constexpr int kThreads = 256;
constexpr int kMaxStages = 12;

struct Plan {
  int n;
  int n_stages;
  int radix[kMaxStages];
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// R-point DFT with roots exp(sgn 2 pi i q t / R), in place
template <int R>
__device__ __forceinline__ void small_dft(float2 (&v)[R], float sgn) {
  if (R == 2) {
    const float2 a = v[0], b = v[1];
    v[0] = make_float2(a.x + b.x, a.y + b.y);
    v[1] = make_float2(a.x - b.x, a.y - b.y);
  } else if (R == 4) {
    const float2 a = v[0], b = v[1], c = v[2], d = v[3];
    const float2 s0 = make_float2(a.x + c.x, a.y + c.y), s1 = make_float2(a.x - c.x, a.y - c.y);
    const float2 s2 = make_float2(b.x + d.x, b.y + d.y), s3 = make_float2(b.x - d.x, b.y - d.y);
    const float2 is3 = make_float2(-sgn * s3.y, sgn * s3.x);          // (sgn i) * (b - d)
    v[0] = make_float2(s0.x + s2.x, s0.y + s2.y);
    v[1] = make_float2(s1.x + is3.x, s1.y + is3.y);
    v[2] = make_float2(s0.x - s2.x, s0.y - s2.y);
    v[3] = make_float2(s1.x - is3.x, s1.y - is3.y);
  } else {
    // small prime (3, 5): u_q = sum_t v_t w^(q t), w = exp(sgn 2 pi i / R)
    float2 w[R];
    w[0] = make_float2(1.f, 0.f);
    if (R == 3) {
      w[1] = make_float2(-0.5f, sgn * 0.86602540378443865f);
      w[2] = make_float2(-0.5f, -sgn * 0.86602540378443865f);
    } else {
      w[1] = make_float2(0.30901699437494742f, sgn * 0.95105651629515357f);
      w[2] = make_float2(-0.80901699437494742f, sgn * 0.58778525229247313f);
      w[R - 2] = make_float2(-0.80901699437494742f, -sgn * 0.58778525229247313f);
      w[R - 1] = make_float2(0.30901699437494742f, -sgn * 0.95105651629515357f);
    }
    float2 u[R];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      float2 acc = v[0];
#pragma unroll
      for (int t = 1; t < R; ++t) {
        const float2 p = cmul(v[t], w[(q * t) % R]);
        acc.x += p.x;
        acc.y += p.y;
      }
      u[q] = acc;
    }
#pragma unroll
    for (int q = 0; q < R; ++q) v[q] = u[q];
  }
}

// One Stockham pass of radix R over kG sequences of length n held in shared memory.
// Element i of sequence g lives at x[g * ss + i * es].  Thread mapping without run-time divisions: for row
// transforms a warp owns one sequence and its lanes stride over the butterflies; for column transforms the
// sequence index is the fastest (kG is a power of two).  k = j mod Ns uses a mask while Ns is a power of two
// (radices 4 and 2 come first in the plan), a modulo only in the trailing radix-3/5 stages.
// NC / NSC: compile-time n and Ns (0 = run time).  With both known every index computation folds to shifts, masks and
// immediate offsets; the generic instantiation spends more than half of its instructions on them (ncu: ALU pipe
// 39-51 %, FMA 27-33 %).  The arithmetic and its order are the same, so the results are bit-identical.
template <int R, bool COLS, int NC = 0, int NSC = 0>
__device__ __forceinline__ void stage(const float2* __restrict__ x, float2* __restrict__ y, int n_rt, int Ns_rt, int ss, int es,
                                      const float2* __restrict__ tw, float sgn) {
  const int n = NC ? NC : n_rt;
  const int Ns = NSC ? NSC : Ns_rt;
  const int nb = n / R;
  const int tstep = n / (Ns * R);
  const bool pow2 = (Ns & (Ns - 1)) == 0;
  auto butterfly = [&](int g, int j) {
    const int k = pow2 ? (j & (Ns - 1)) : (j % Ns);
    float2 v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) {
      v[t] = x[g * ss + (j + t * nb) * es];
      if (t > 0) v[t] = cmul(v[t], tw[t * k * tstep]);
    }
    small_dft<R>(v, sgn);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int q = 0; q < R; ++q) y[g * ss + (j0 + q * Ns) * es] = v[q];
  };
  if (COLS) {
    for (int w = threadIdx.x; w < nb * kG; w += kThreads) butterfly(w & (kG - 1), w / kG);
  } else {
    static_assert(kThreads / 32 == kG, "one warp per sequence");
    const int g = threadIdx.x >> 5;
    for (int j = threadIdx.x & 31; j < nb; j += 32) butterfly(g, j);
  }
}

template <bool COLS>
__device__ __forceinline__ const float2* run_stages(float2* a, float2* b, const Plan& plan, int ss, int es, const float2* tw,
                                                    float sgn) {
  int Ns = 1;
  float2* x = a;
  float2* y = b;
  for (int s = 0; s < plan.n_stages; ++s) {
    const int r = plan.radix[s];
    if (r == 4) stage<4, COLS>(x, y, plan.n, Ns, ss, es, tw, sgn);
    else if (r == 2) stage<2, COLS>(x, y, plan.n, Ns, ss, es, tw, sgn);
    else if (r == 3) stage<3, COLS>(x, y, plan.n, Ns, ss, es, tw, sgn);
    else stage<5, COLS>(x, y, plan.n, Ns, ss, es, tw, sgn);
    Ns *= r;
    __syncthreads();
    float2* t = x; x = y; y = t;
  }
  return x;
}

// The same plan as make_plan() (radices 4, then 2, then 3, then 5) unrolled at compile time for a fixed length.
template <bool COLS, int NC, int NS, int REM>
struct FixedStages {
  static __device__ __forceinline__ const float2* run(float2* x, float2* y, int ss, int es, const float2* tw, float sgn) {
    constexpr int R = (REM % 4 == 0) ? 4 : (REM % 2 == 0) ? 2 : (REM % 3 == 0) ? 3 : 5;
    static_assert(REM % R == 0, "length must be a product of 2, 3 and 5");
    stage<R, COLS, NC, NS>(x, y, NC, NS, ss, es, tw, sgn);
    __syncthreads();
    return FixedStages<COLS, NC, NS * R, REM / R>::run(y, x, ss, es, tw, sgn);
  }
};
template <bool COLS, int NC, int NS>
struct FixedStages<COLS, NC, NS, 1> {
  static __device__ __forceinline__ const float2* run(float2* x, float2*, int, int, const float2*, float) { return x; }
};
template <bool COLS, int NC>
__device__ __forceinline__ const float2* run_plan(float2* a, float2* b, const Plan& plan, int ss, int es, const float2* tw,
                                                  float sgn) {
  if (NC != 0) return FixedStages<COLS, NC ? NC : 1, 1, NC ? NC : 1>::run(a, b, ss, es, tw, sgn);
  return run_stages<COLS>(a, b, plan, ss, es, tw, sgn);
}

__device__ __forceinline__ void build_twiddles(float2* tw, int n, float sgn) {
  for (int i = threadIdx.x; i < n; i += kThreads) {
    float s, c;
    sincospif(2.0f * (float)i / (float)n, &s, &c);
    tw[i] = make_float2(c, sgn * s);
  }
}

// rows: in [n_rows, n] complex -> out [n_rows, n] complex; centred (shift by n/2 on both sides), orthonormal
template <int NC>
__global__ void __launch_bounds__(kThreads)
rows_kernel(const float2* __restrict__ in, const uint8_t* __restrict__ colmask, float2* __restrict__ out, long long n_rows,
            Plan plan, float sgn, float scale) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int n = NC ? NC : plan.n;
  const int ss = n + 1;
  float2* tw = reinterpret_cast<float2*>(smem_raw);
  float2* a = tw + n;
  float2* b = a + kG * ss;
  build_twiddles(tw, n, sgn);
  const long long row0 = (long long)blockIdx.x * kG;
  const int half = n / 2;
  const int g = threadIdx.x >> 5;                  // one warp per row
  for (int i = threadIdx.x & 31; i < n; i += 32) {   // i = index in the source row (coalesced)
    float2 v = make_float2(0.f, 0.f);
    if (row0 + g < n_rows) {
      v = __ldg(in + (row0 + g) * n + i);
      if (colmask && !colmask[i]) v = make_float2(0.f, 0.f);
    }
    // ifftshift: a[m] = src[(m + n/2) % n]  <=>  source index i goes to m = (i - n/2) mod n
    int m = i - half;
    if (m < 0) m += n;
    a[g * ss + m] = v;
  }
  __syncthreads();
  const float2* r = run_plan<false, NC>(a, b, plan, ss, 1, tw, sgn);
  if (row0 + g < n_rows) {
    for (int o = threadIdx.x & 31; o < n; o += 32) {   // o = index in the destination row (coalesced)
      // fftshift: out[o] = r[(o - n/2) mod n]
      int m = o - half;
      if (m < 0) m += n;
      const float2 v = r[g * ss + m];
      out[(row0 + g) * n + o] = make_float2(v.x * scale, v.y * scale);
    }
  }
}

// columns: in [N, H, W] complex, transform along H for kG adjacent columns; complex or magnitude output
template <bool ABS, int NC>
__global__ void __launch_bounds__(kThreads)
cols_kernel(const float2* __restrict__ in, float2* __restrict__ out_c, float* __restrict__ out_abs, int H, int W, Plan plan,
            float sgn, float scale) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int n = NC ? NC : plan.n;                 // == H
  float2* tw = reinterpret_cast<float2*>(smem_raw);
  float2* a = tw + n;
  float2* b = a + kG * n;
  build_twiddles(tw, n, sgn);
  const long long img = blockIdx.y;
  const int x0 = blockIdx.x * kG;
  const float2* src = in + img * (long long)H * W;
  const int half = n / 2;
  for (int w = threadIdx.x; w < kG * n; w += kThreads) {
    const int g = w & (kG - 1), i = w / kG;        // g fastest: 64-byte row segments
    float2 v = make_float2(0.f, 0.f);
    if (x0 + g < W) v = __ldg(src + (long long)i * W + x0 + g);
    int m = i - half;
    if (m < 0) m += n;
    a[m * kG + g] = v;
  }
  __syncthreads();
  const float2* r = run_plan<true, NC>(a, b, plan, 1, kG, tw, sgn);
  for (int w = threadIdx.x; w < kG * n; w += kThreads) {
    const int g = w & (kG - 1), o = w / kG;
    if (x0 + g >= W) continue;
    int m = o - half;
    if (m < 0) m += n;
    float2 v = r[m * kG + g];
    v.x *= scale;
    v.y *= scale;
    const long long dst = img * (long long)H * W + (long long)o * W + x0 + g;
    if (ABS) out_abs[dst] = sqrtf(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)));   // fastmri.complex_abs
    else out_c[dst] = v;
  }
}

// ---- length 320 (the reference's knee slices): register-resident transform, fft320.cuh ----
// 320 = 16 x 20: a thread takes the 16 inputs x'[20 n1 + n2] of one residue n2 straight from global memory (the
// ifftshift is a permutation of n1), does a 16-point DFT in registers, multiplies by w320^(n2 k1) (table in shared
// memory, orthonormal scale folded in) and hands the values over through shared memory ONCE; a thread then takes the 20
// values of one k1, does a 20-point prime-factor DFT in registers and stores X[k1 + 16 k2] (the fftshift is a
// permutation of k2).  ~40 instructions per point instead of ~129 for the four shared-memory Stockham passes above
// (profiles/r02_metrics_fft_ncu.txt): the passes become HBM-bound.
constexpr int kSeq320 = 16;          // sequences (rows / adjacent columns) per CTA
constexpr int kThreads320 = 20 * kSeq320;

// tw[k1 * 20 + n2] = scale * w320^(n2 k1): indexed by the consumer's (k1, n2), so that the 20 threads of a row read 20
// consecutive entries (a table indexed by the exponent n2 k1 is read with stride k1: up to 20-way bank conflicts)
template <bool INV>
__device__ __forceinline__ void build_twiddles320(float2* tw, float scale) {
  for (int i = threadIdx.x; i < 320; i += kThreads320) {
    const int k1 = i / 20, n2 = i - 20 * k1;
    float sn, cs;
    sincospif(2.0f * (float)fft320::twiddle_index(n2, k1) / 320.0f, &sn, &cs);
    tw[i] = make_float2(cs * scale, (INV ? sn : -sn) * scale);
  }
}

template <bool INV>
__global__ void __launch_bounds__(kThreads320)
rows320_kernel(const float2* __restrict__ in, const uint8_t* __restrict__ colmask, float2* __restrict__ out, long long n_rows,
               float scale) {
  using namespace fft320;
  __shared__ float2 tw[320];
  __shared__ float2 ys[kSeq320][20 * 17];          // [row][n2 * 17 + k1]: odd stride, conflict-free both ways
  build_twiddles320<INV>(tw, scale);
  const long long row0 = (long long)blockIdx.x * kSeq320;
  {
    const int r = threadIdx.x / 20, t = threadIdx.x - 20 * r;
    const long long row = row0 + r;
    // all 16 loads first, then the mask: written as one loop the compiler chained load -> mask -> select per value,
    // i.e. 16 dependent DRAM round trips per thread (ncu: 63 % long-scoreboard stalls, 23 % issue, 0.64 ms)
    float2 raw[16];
    uint8_t keep[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      raw[n1] = make_float2(0.f, 0.f);
      if (row < n_rows) raw[n1] = __ldg(in + row * 320 + src_index(n1, t));
    }
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) keep[n1] = colmask ? __ldg(colmask + src_index(n1, t)) : (uint8_t)1;
    C a[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) a[n1] = keep[n1] ? mk(raw[n1].x, raw[n1].y) : mk(0.f, 0.f);
    dft16<INV>(a);
    __syncthreads();                                // twiddle table complete
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      const float2 w = tw[k1 * 20 + t];
      const C y = mulc(a[k1], w.x, w.y);            // (k1 = 0: w = scale)
      ys[r][t * 17 + k1] = make_float2(y.x, y.y);
    }
  }
  __syncthreads();
  if (threadIdx.x < 16 * kSeq320) {
    const int r = threadIdx.x >> 4, k1 = threadIdx.x & 15;
    const long long row = row0 + r;
    C b[20];
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) {
      const float2 v = ys[r][n2 * 17 + k1];
      b[n2] = mk(v.x, v.y);
    }
    dft20<INV>(b);
    if (row < n_rows) {
#pragma unroll
      for (int k2 = 0; k2 < 20; ++k2) out[row * 320 + dst_index(k1, k2)] = make_float2(b[k2].x, b[k2].y);
    }
  }
}

// columns of [N, 320, W]: 16 adjacent columns per CTA (128-byte row segments), complex or magnitude output
template <bool INV, bool ABS>
__global__ void __launch_bounds__(kThreads320)
cols320_kernel(const float2* __restrict__ in, float2* __restrict__ out_c, float* __restrict__ out_abs, int W, float scale) {
  using namespace fft320;
  __shared__ float2 tw[320];
  __shared__ float2 ys[320 * kSeq320];             // [(n2 * 16 + k1) * 16 + g]
  build_twiddles320<INV>(tw, scale);
  const long long img = blockIdx.y;
  const int x0 = blockIdx.x * kSeq320;
  const float2* src = in + img * 320ll * W;
  {
    const int g = threadIdx.x & (kSeq320 - 1), n2 = threadIdx.x / kSeq320;
    const bool live = x0 + g < W;
    C a[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      float2 v = make_float2(0.f, 0.f);
      if (live) v = __ldg(src + (long long)src_index(n1, n2) * W + x0 + g);
      a[n1] = mk(v.x, v.y);
    }
    dft16<INV>(a);
    __syncthreads();
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      const float2 w = tw[k1 * 20 + n2];
      const C y = mulc(a[k1], w.x, w.y);
      ys[(n2 * 16 + k1) * kSeq320 + g] = make_float2(y.x, y.y);
    }
  }
  __syncthreads();
  if (threadIdx.x < 16 * kSeq320) {
    const int g = threadIdx.x & (kSeq320 - 1), k1 = threadIdx.x / kSeq320;
    C b[20];
#pragma unroll
    for (int n2 = 0; n2 < 20; ++n2) {
      const float2 v = ys[(n2 * 16 + k1) * kSeq320 + g];
      b[n2] = mk(v.x, v.y);
    }
    dft20<INV>(b);
    if (x0 + g < W) {
#pragma unroll
      for (int k2 = 0; k2 < 20; ++k2) {
        const long long dst = img * 320ll * W + (long long)dst_index(k1, k2) * W + x0 + g;
        if (ABS) out_abs[dst] = sqrtf(__fadd_rn(__fmul_rn(b[k2].x, b[k2].x), __fmul_rn(b[k2].y, b[k2].y)));   // fastmri.complex_abs
        else out_c[dst] = make_float2(b[k2].x, b[k2].y);
      }
    }
  }
}

static bool make_plan(int n, Plan* p) {
  if (n < 2 || n > kMaxN) return false;
  p->n = n;
  p->n_stages = 0;
  int m = n;
  const int radices[4] = {4, 2, 3, 5};
  for (int ri = 0; ri < 4; ++ri) {
    const int r = radices[ri];
    while (m % r == 0) {
      if (p->n_stages == kMaxStages) return false;
      p->radix[p->n_stages++] = r;
      m /= r;
    }
  }
  return m == 1;
}

static int run(const float* d_in, const uint8_t* d_colmask, long long N, int H, int W, int inverse, float* d_out_c,
               float* d_out_abs, void* d_ws, cudaStream_t st) {
  Plan pw, ph;
  MRINR_REQUIRE(make_plan(W, &pw) && make_plan(H, &ph), MRINR_E_UNSUPPORTED,
                "fft2c: sizes must be products of 2, 3 and 5 in [2, %d] (got %d x %d)", kMaxN, H, W);
  const float sgn = inverse ? 1.f : -1.f;
  const size_t smem_r = (size_t)(W + 2 * kG * (W + 1)) * sizeof(float2);
  const size_t smem_c = (size_t)(H + 2 * kG * H) * sizeof(float2);
  const long long n_rows = N * H;
  float2* tmp = static_cast<float2*>(d_ws);
  const float2* in2 = reinterpret_cast<const float2*>(d_in);
  const unsigned grid_r = (unsigned)((n_rows + kG - 1) / kG);
  const dim3 grid_c((W + kG - 1) / kG, (unsigned)N);
  const float sw = 1.0f / sqrtf((float)W), sh = 1.0f / sqrtf((float)H);
  // (the shared-memory size depends on the length, so the opt-in is renewed on every call)
  // lengths with a compile-time plan (the reference's knee slices are 320 x 320); everything else: run-time plan
#define MRINR_FFT_ROWS(NC)                                                                                  \
  do {                                                                                                      \
    MRINR_CUDA(cudaFuncSetAttribute(rows_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r)); \
    rows_kernel<NC><<<grid_r, kThreads, smem_r, st>>>(in2, d_colmask, tmp, n_rows, pw, sgn, sw);            \
  } while (0)
#define MRINR_FFT_COLS(NC)                                                                                  \
  do {                                                                                                      \
    if (d_out_abs) {                                                                                        \
      MRINR_CUDA(cudaFuncSetAttribute(cols_kernel<true, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c)); \
      cols_kernel<true, NC><<<grid_c, kThreads, smem_c, st>>>(tmp, nullptr, d_out_abs, H, W, ph, sgn, sh);  \
    } else {                                                                                                \
      MRINR_CUDA(cudaFuncSetAttribute(cols_kernel<false, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c)); \
      cols_kernel<false, NC><<<grid_c, kThreads, smem_c, st>>>(tmp, reinterpret_cast<float2*>(d_out_c), nullptr, H, W, ph, \
                                                               sgn, sh);                                    \
    }                                                                                                       \
  } while (0)
  if (W == 320) {
    const unsigned g320 = (unsigned)((n_rows + kSeq320 - 1) / kSeq320);
    if (inverse) rows320_kernel<true><<<g320, kThreads320, 0, st>>>(in2, d_colmask, tmp, n_rows, sw);
    else rows320_kernel<false><<<g320, kThreads320, 0, st>>>(in2, d_colmask, tmp, n_rows, sw);
  } else if (W == 256) MRINR_FFT_ROWS(256); else MRINR_FFT_ROWS(0);
  if (H == 320) {
    const dim3 g320((W + kSeq320 - 1) / kSeq320, (unsigned)N);
    float2* oc = reinterpret_cast<float2*>(d_out_c);
    if (d_out_abs) {
      if (inverse) cols320_kernel<true, true><<<g320, kThreads320, 0, st>>>(tmp, nullptr, d_out_abs, W, sh);
      else cols320_kernel<false, true><<<g320, kThreads320, 0, st>>>(tmp, nullptr, d_out_abs, W, sh);
    } else {
      if (inverse) cols320_kernel<true, false><<<g320, kThreads320, 0, st>>>(tmp, oc, nullptr, W, sh);
      else cols320_kernel<false, false><<<g320, kThreads320, 0, st>>>(tmp, oc, nullptr, W, sh);
    }
  } else if (H == 256) MRINR_FFT_COLS(256); else MRINR_FFT_COLS(0);
#undef MRINR_FFT_ROWS
#undef MRINR_FFT_COLS
  count_launch(2);
  return check_launch("fft2c");
}

}  // namespace fft
}  // namespace mrinr

using namespace mrinr;

extern "C" int64_t mrinr_fft2c_workspace_bytes(int64_t N, int32_t H, int32_t W) {
  if (N < 0 || H < 0 || W < 0) return 0;
  return (int64_t)((size_t)N * H * W * sizeof(float2));
}

static int check_fft_args(const char* who, const void* in, const void* out, const void* ws, int64_t N, int32_t H, int32_t W,
                          int64_t ws_bytes) {
  MRINR_REQUIRE(in && out && ws, MRINR_E_ARG, "%s: null pointer", who);
  MRINR_REQUIRE(N > 0 && N <= 65535, MRINR_E_ARG, "%s: between 1 and 65535 images per call (got %lld)", who, (long long)N);
  MRINR_REQUIRE(ws_bytes >= mrinr_fft2c_workspace_bytes(N, H, W), MRINR_E_ARG,
                "%s: workspace must hold mrinr_fft2c_workspace_bytes(N,H,W) bytes", who);
  MRINR_REQUIRE((reinterpret_cast<uintptr_t>(in) & 7u) == 0 && (reinterpret_cast<uintptr_t>(out) & 7u) == 0 &&
                    (reinterpret_cast<uintptr_t>(ws) & 7u) == 0,
                MRINR_E_ALIGN, "%s: buffers must be 8-byte aligned", who);
  return 0;
}

extern "C" int mrinr_fft2c(const float* d_in, int64_t N, int32_t H, int32_t W, int32_t inverse, float* d_out,
                           void* d_workspace, int64_t workspace_bytes, void* stream) {
  if (N == 0) return 0;
  const int rc = check_fft_args("mrinr_fft2c", d_in, d_out, d_workspace, N, H, W, workspace_bytes);
  if (rc != 0) return rc;
  return fft::run(d_in, nullptr, N, H, W, inverse, d_out, nullptr, d_workspace, (cudaStream_t)stream);
}

extern "C" int mrinr_kspace_to_image(const float* d_kspace, const uint8_t* d_colmask, int64_t N, int32_t H, int32_t W,
                                     float* d_mag, void* d_workspace, int64_t workspace_bytes, void* stream) {
  if (N == 0) return 0;
  const int rc = check_fft_args("mrinr_kspace_to_image", d_kspace, d_mag, d_workspace, N, H, W, workspace_bytes);
  if (rc != 0) return rc;
  return fft::run(d_kspace, d_colmask, N, H, W, /*inverse*/ 1, nullptr, d_mag, d_workspace, (cudaStream_t)stream);
}
