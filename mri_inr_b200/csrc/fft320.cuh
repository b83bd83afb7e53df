// In-register building blocks of the 320-point transform (fft.cu): 320 = 16 x 20 with
//   X[k1 + 16 k2] = sum_{n2 < 20} w20^(n2 k2) [ w320^(n2 k1) sum_{n1 < 16} x[20 n1 + n2] w16^(n1 k1) ]
// i.e. a 16-point DFT per thread (n2 fixed), one twiddle per value, an exchange through shared memory, and a 20-point
// DFT per thread (k1 fixed).  The 16-point DFT is 4 x 4 (nine constant twiddles), the 20-point DFT is the prime-factor
// form 4 x 5 (4 and 5 are coprime: n = 5 a + 4 b, k = 5 c + 16 d mod 20, no twiddles at all).  Every index map is a
// compile-time permutation of registers.  w = exp(sgn 2 pi i / n), sgn = +1 for the inverse transform (INV).
// Plain C++ apart from the MRINR_HD qualifier and the packed device versions of add / sub / axpy / scale:
// tests/test_fft320_host.py compiles it with g++ and checks the three routines against naive fp64 DFTs (the GPU
// tests check the device versions against numpy fp64).
#pragma once
#ifdef __CUDACC__
#define MRINR_HD __device__ __forceinline__
#else
#define MRINR_HD inline
#endif

namespace mrinr {
namespace fft320 {

struct C {
  float x, y;
};
MRINR_HD C mk(float x, float y) { C c; c.x = x; c.y = y; return c; }
#ifdef __CUDA_ARCH__
// On the device a complex value is one packed fp32 pair: complex add / subtract / real scaling are ONE instruction
// (add.f32x2 / sub.f32x2 / fma.rn.f32x2) instead of two -- the columns pass of the 320-point transform is issue-bound.
// Each packed operation rounds its two halves exactly like the two scalar operations it replaces.
__device__ __forceinline__ unsigned long long pk(C a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ C upk(unsigned long long v) {
  C c;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(v));
  return c;
}
__device__ __forceinline__ C add(C a, C b) {
  unsigned long long r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
  return upk(r);
}
__device__ __forceinline__ C sub(C a, C b) {
  unsigned long long r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b)));
  return upk(r);
}
// acc + c * t (c real)
__device__ __forceinline__ C axpy(C acc, float c, C t) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(t)), "l"(pk(mk(c, c))), "l"(pk(acc)));
  return upk(r);
}
// c * t (c real)
__device__ __forceinline__ C scale(float c, C t) {
  unsigned long long r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(t)), "l"(pk(mk(c, c))));
  return upk(r);
}
#else
MRINR_HD C add(C a, C b) { return mk(a.x + b.x, a.y + b.y); }
MRINR_HD C sub(C a, C b) { return mk(a.x - b.x, a.y - b.y); }
MRINR_HD C axpy(C acc, float c, C t) { return mk(acc.x + c * t.x, acc.y + c * t.y); }
MRINR_HD C scale(float c, C t) { return mk(c * t.x, c * t.y); }
#endif
// a * (c + i s)
MRINR_HD C mulc(C a, float c, float s) { return mk(a.x * c - a.y * s, a.x * s + a.y * c); }
// a * (sgn i)
template <bool INV>
MRINR_HD C muli(C a) { return INV ? mk(-a.y, a.x) : mk(a.y, -a.x); }

// 4-point DFT: v[q] <- sum_t v[t] w4^(q t)
template <bool INV>
MRINR_HD void dft4(C& v0, C& v1, C& v2, C& v3) {
  const C s0 = add(v0, v2), s1 = sub(v0, v2), s2 = add(v1, v3), s3 = muli<INV>(sub(v1, v3));
  v0 = add(s0, s2);
  v1 = add(s1, s3);
  v2 = sub(s0, s2);
  v3 = sub(s1, s3);
}

// 5-point DFT: v[q] <- sum_t v[t] w5^(q t)
template <bool INV>
MRINR_HD void dft5(C& v0, C& v1, C& v2, C& v3, C& v4) {
  constexpr float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;     // cos(2 pi / 5), cos(4 pi / 5)
  constexpr float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;      // sin(2 pi / 5), sin(4 pi / 5)
  const C t1 = add(v1, v4), t2 = add(v2, v3), t3 = sub(v1, v4), t4 = sub(v2, v3);
  const C m1 = axpy(axpy(v0, c1, t1), c2, t2);
  const C m2 = axpy(axpy(v0, c2, t1), c1, t2);
  const C n1 = muli<INV>(axpy(scale(s1, t3), s2, t4));
  const C n2 = muli<INV>(axpy(scale(s2, t3), -s1, t4));
  v0 = add(v0, add(t1, t2));
  v1 = add(m1, n1);
  v4 = sub(m1, n1);
  v2 = add(m2, n2);
  v3 = sub(m2, n2);
}

// 16-point DFT in place: a[k] <- sum_n a[n] w16^(n k).   n = 4 p + q, k = r + 4 s.
template <bool INV>
MRINR_HD void dft16(C (&a)[16]) {
  constexpr float sg = INV ? 1.f : -1.f;
  constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f;      // cos, sin of pi / 8
  constexpr float h = 0.70710678118654752f;
  // over p for every q: B[q][r] stored at a[4 r + q]
#pragma unroll
  for (int q = 0; q < 4; ++q) dft4<INV>(a[q], a[4 + q], a[8 + q], a[12 + q]);
  // twiddles w16^(q r)
  a[4 * 1 + 1] = mulc(a[4 * 1 + 1], c1, sg * s1);      // q r = 1
  a[4 * 1 + 2] = mulc(a[4 * 1 + 2], h, sg * h);        // 2
  a[4 * 1 + 3] = mulc(a[4 * 1 + 3], s1, sg * c1);      // 3
  a[4 * 2 + 1] = mulc(a[4 * 2 + 1], h, sg * h);        // 2
  a[4 * 2 + 2] = muli<INV>(a[4 * 2 + 2]);              // 4
  a[4 * 2 + 3] = mulc(a[4 * 2 + 3], -h, sg * h);       // 6
  a[4 * 3 + 1] = mulc(a[4 * 3 + 1], s1, sg * c1);      // 3
  a[4 * 3 + 2] = mulc(a[4 * 3 + 2], -h, sg * h);       // 6
  a[4 * 3 + 3] = mulc(a[4 * 3 + 3], -c1, -sg * s1);    // 9
  // over q for every r: A[r + 4 s] lands at a[4 r + s]; un-permute
#pragma unroll
  for (int r = 0; r < 4; ++r) dft4<INV>(a[4 * r], a[4 * r + 1], a[4 * r + 2], a[4 * r + 3]);
  C t[16];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int s = 0; s < 4; ++s) t[r + 4 * s] = a[4 * r + s];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = t[i];
}

// 20-point DFT in place (prime-factor 4 x 5): a[k] <- sum_n a[n] w20^(n k).
template <bool INV>
MRINR_HD void dft20(C (&a)[20]) {
  C b[4][5];      // b[n1][n2] = a[(5 n1 + 4 n2) mod 20]
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1)
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) b[n1][n2] = a[(5 * n1 + 4 * n2) % 20];
#pragma unroll
  for (int n2 = 0; n2 < 5; ++n2) dft4<INV>(b[0][n2], b[1][n2], b[2][n2], b[3][n2]);      // -> b[k1][n2]
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft5<INV>(b[k1][0], b[k1][1], b[k1][2], b[k1][3], b[k1][4]);      // -> b[k1][k2]
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = 0; k2 < 5; ++k2) a[(5 * k1 + 16 * k2) % 20] = b[k1][k2];
}

// Centred transform (fastmri's fftshift(fft(ifftshift(x)))) folded into the index maps, for n = 320:
// the thread with n2 = t reads its n1-th input from source index src_index(n1, t) -- x'[20 n1 + t] with
// x'[m] = src[(m + 160) mod 320] -- and the thread with k1 writes X[k1 + 16 k2] to destination index dst_index(k1, k2).
MRINR_HD constexpr int src_index(int n1, int t) { return 20 * ((n1 + 8) & 15) + t; }
MRINR_HD constexpr int dst_index(int k1, int k2) { return k1 + 16 * ((k2 + 10) % 20); }
// exponent of the w320 twiddle between the two stages
MRINR_HD constexpr int twiddle_index(int n2, int k1) { return n2 * k1; }

}  // namespace fft320
}  // namespace mrinr
