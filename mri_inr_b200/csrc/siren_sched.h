// Static tile schedule of the fused synthesis kernel (siren_tc5.cu), shared by every warp role of a cluster.
// Plain C++ (no CUDA constructs beyond the MRINR_HD qualifier) so that tests/test_sched_host.py can compile it with g++
// and enumerate the walk on the CPU: every (patch, coordinate block) of every cluster must be visited exactly once.
#pragma once
#ifdef __CUDACC__
#define MRINR_HD __device__ __forceinline__
#else
#define MRINR_HD inline
#endif

namespace mrinr {
namespace v5 {

constexpr int kTileM = 128;
constexpr int kMaxSub = 2;                     // patches sharing a remainder tile

// The schedule of one cluster; every role derives it from the same inputs, so nothing is communicated.
// The cluster's patches are walked in sub-blocks of kSubBlock patches, block-major inside a sub-block: the modulation
// vectors of a sub-block (5 KB per patch, read once per coordinate block) then stay in L2 between their
// C/128 + 1 uses (74 clusters x 128 patches x 5 KB = 47 MB) instead of streaming from HBM every time.
constexpr int kSubBlock = 128;
struct Sched {
  long long pa;          // this cluster's patches: compacted indices [pa, pa + np)
  int np;
  int n_full, rem, ksub; // C = 128 n_full + rem; ksub = patches per remainder tile (0 if rem == 0)
  int n_types;           // coordinate blocks per patch: n_full (+ 1 if rem)
  int tpi;               // tiles per cluster iteration: 2 CTAs x tile slots per CTA (4, or 2 in the fp16x3 mode)
  long long total;       // cluster iterations (tpi tiles each)
};
MRINR_HD int iters_rem(const Sched& s, int n) {   // remainder-block iterations for n patches
  return s.rem ? ((n + s.ksub - 1) / s.ksub + s.tpi - 1) / s.tpi : 0;
}
MRINR_HD Sched make_sched(long long n_act, int C, long long cluster_id, long long n_clusters, int tpi) {
  Sched s;
  s.tpi = tpi;
  s.pa = n_act * cluster_id / n_clusters;
  s.np = (int)(n_act * (cluster_id + 1) / n_clusters - s.pa);
  s.n_full = C / kTileM;
  s.rem = C - s.n_full * kTileM;
  s.ksub = s.rem ? (kTileM / s.rem < kMaxSub ? kTileM / s.rem : kMaxSub) : 0;
  s.n_types = s.n_full + (s.rem ? 1 : 0);
  const int blocks = s.np / kSubBlock, tail = s.np - blocks * kSubBlock;
  s.total = (long long)blocks * (s.n_full * (kSubBlock / tpi) + iters_rem(s, kSubBlock));
  if (tail) s.total += s.n_full * ((tail + tpi - 1) / tpi) + iters_rem(s, tail);
  return s;
}
struct Walk {          // (sub-block, coordinate block, iteration within the block), advanced without divisions
  int type = 0;
  int j = 0;
  int base = 0;        // first patch of the sub-block, relative to Sched::pa
  int nps = 0;         // patches in the sub-block
  int itf = 0, itr = 0;
  MRINR_HD void set_block(const Sched& s) {
    nps = s.np - base < kSubBlock ? s.np - base : kSubBlock;
    itf = (nps + s.tpi - 1) / s.tpi;
    itr = iters_rem(s, nps);
  }
  MRINR_HD void next(const Sched& s) {
    if (++j == (type < s.n_full ? itf : itr)) {
      j = 0;
      if (++type == s.n_types) { type = 0; base += kSubBlock; set_block(s); }
    }
  }
};

}  // namespace v5
}  // namespace mrinr
