// Host enumeration of the synthesis kernel's static schedule (mri_inr_b200/csrc/siren_sched.h): mirrors how the epilogue
// warps and the producer derive, per cluster iteration, the tile of every (slot, CTA rank) -- see siren_tc5.cu -- and
// checks that every (patch, coordinate block) is produced exactly once, for tpi = 4 (two tile slots) and tpi = 2 (fp16x3).
// usage: sched_check <n_act> <C> <n_clusters> <tpi>   -> prints "OK <tiles> <iterations>" or "FAIL ..."
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../mri_inr_b200/csrc/siren_sched.h"

using namespace mrinr::v5;

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  const long long n_act = atoll(argv[1]);
  const int C = atoi(argv[2]);
  const long long n_clusters = atoll(argv[3]);
  const int tpi = atoi(argv[4]);
  const int slots = tpi / 2;
  const int n_full = C / kTileM, rem = C % kTileM;
  const int types = n_full + (rem ? 1 : 0);
  std::vector<int> seen((size_t)n_act * types, 0);
  long long tiles = 0, iters = 0, phantom = 0;
  for (long long cl = 0; cl < n_clusters; ++cl) {
    const Sched S = make_sched(n_act, C, cl, n_clusters, tpi);
    Walk w;
    w.set_block(S);
    for (long long it = 0; it < S.total; ++it, ++iters) {
      for (int slot = 0; slot < slots; ++slot) {
        for (int rank = 0; rank < 2; ++rank) {
          const int ti = w.j * tpi + slot * 2 + rank;
          ++tiles;
          if (w.type < S.n_full) {
            const int pl = ti;
            if (pl < w.nps) ++seen[(size_t)(S.pa + w.base + pl) * types + w.type]; else ++phantom;
          } else {
            bool any = false;
            for (int sub = 0; sub < S.ksub; ++sub) {
              const int pl = ti * S.ksub + sub;
              if (pl < w.nps) { ++seen[(size_t)(S.pa + w.base + pl) * types + w.type]; any = true; }
            }
            if (!any) ++phantom;
          }
          if (w.base >= S.np) { printf("FAIL walk ran past the cluster's patches (cluster %lld)\n", cl); return 1; }
        }
      }
      w.next(S);
    }
    if (S.total > 0 && !(w.base >= S.np || (w.type == 0 && w.j == 0))) {
      printf("FAIL cluster %lld: walk did not end on a sub-block boundary\n", cl);
      return 1;
    }
  }
  for (size_t i = 0; i < seen.size(); ++i)
    if (seen[i] != 1) {
      printf("FAIL patch %zu block %zu visited %d times\n", i / types, i % types, seen[i]);
      return 1;
    }
  printf("OK %lld %lld %lld\n", tiles, iters, phantom);
  return 0;
}
