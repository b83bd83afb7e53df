// Host check of csrc/fft320.cuh: the in-register 16- and 20-point DFTs against naive fp64 DFTs (both signs), and the
// composed, centred 320-point transform (16-point DFTs, w320 twiddles, 20-point DFTs, shift folded into the index maps)
// against fftshift(dft(ifftshift(x))).  Prints OK and the largest errors.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../mri_inr_b200/csrc/fft320.cuh"

using namespace mrinr::fft320;
typedef std::complex<double> cd;

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

template <int N, bool INV, typename F>
static double check_small(F f) {
  double worst = 0;
  for (int trial = 0; trial < 50; ++trial) {
    C a[N];
    cd x[N];
    for (int i = 0; i < N; ++i) {
      a[i] = mk((float)frand(), (float)frand());
      x[i] = cd(a[i].x, a[i].y);
    }
    f(a);
    for (int k = 0; k < N; ++k) {
      cd s = 0;
      for (int n = 0; n < N; ++n) s += x[n] * std::polar(1.0, (INV ? 2.0 : -2.0) * M_PI * (double)((n * k) % N) / N);
      worst = std::fmax(worst, std::abs(s - cd(a[k].x, a[k].y)));
    }
  }
  return worst;
}

template <bool INV>
static double check_320() {
  const int n = 320;
  double worst = 0;
  std::vector<cd> tw(n);
  for (int i = 0; i < n; ++i) tw[i] = std::polar(1.0, (INV ? 2.0 : -2.0) * M_PI * i / n);
  for (int trial = 0; trial < 5; ++trial) {
    std::vector<cd> src(n), want(n);
    for (auto& v : src) v = cd((float)frand(), (float)frand());
    // reference: out[o] = X[(o - 160) mod n], X = DFT(x'), x'[m] = src[(m + 160) mod n]
    for (int o = 0; o < n; ++o) {
      const int k = (o + n - n / 2) % n;
      cd s = 0;
      for (int m = 0; m < n; ++m) s += src[(m + n / 2) % n] * tw[(int)(((long long)m * k) % n)];
      want[o] = s;
    }
    // the kernel's decomposition
    std::vector<C> y(20 * 16);
    for (int t = 0; t < 20; ++t) {
      C a[16];
      for (int n1 = 0; n1 < 16; ++n1) a[n1] = mk((float)src[src_index(n1, t)].real(), (float)src[src_index(n1, t)].imag());
      dft16<INV>(a);
      for (int k1 = 0; k1 < 16; ++k1) {
        const cd w = tw[twiddle_index(t, k1)];
        y[t * 16 + k1] = k1 ? mulc(a[k1], (float)w.real(), (float)w.imag()) : a[k1];
      }
    }
    for (int k1 = 0; k1 < 16; ++k1) {
      C b[20];
      for (int n2 = 0; n2 < 20; ++n2) b[n2] = y[n2 * 16 + k1];
      dft20<INV>(b);
      for (int k2 = 0; k2 < 20; ++k2) worst = std::fmax(worst, std::abs(want[dst_index(k1, k2)] - cd(b[k2].x, b[k2].y)));
    }
  }
  return worst;
}

int main() {
  srand(1234);
  const double e16f = check_small<16, false>([](C(&a)[16]) { dft16<false>(a); });
  const double e16i = check_small<16, true>([](C(&a)[16]) { dft16<true>(a); });
  const double e20f = check_small<20, false>([](C(&a)[20]) { dft20<false>(a); });
  const double e20i = check_small<20, true>([](C(&a)[20]) { dft20<true>(a); });
  const double e320f = check_320<false>(), e320i = check_320<true>();
  const bool ok = e16f < 2e-5 && e16i < 2e-5 && e20f < 2e-5 && e20i < 2e-5 && e320f < 3e-4 && e320i < 3e-4;
  printf("%s dft16 %.2e %.2e  dft20 %.2e %.2e  fft320 %.2e %.2e\n", ok ? "OK" : "FAIL", e16f, e16i, e20f, e20i, e320f, e320i);
  return ok ? 0 : 1;
}
