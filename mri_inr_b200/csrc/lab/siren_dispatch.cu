// Variant selection for the tensor-core synthesis kernel.  Default: version 5 (CTA pairs, two tile slots,
// coordinate-block tiles with the layer-0 table in registers).  MRINR_TC_VARIANT=1|2|3|4 selects the earlier kernels,
// 6 the experiment with layer 0 on dedicated warps (correct, but slower: see the header of siren_tc6.cu) (kept for A/B measurements, see profiles/).
#include "common.cuh"

#include <cstdlib>

namespace mrinr {
namespace v1 {
int launch_siren_tc_v1(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                       int64_t B, float* d_out, cudaStream_t st);
}
namespace v2 {
int launch_siren_tc_v2(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                       int64_t B, float* d_out, cudaStream_t st);
}
namespace v3 {
int launch_siren_tc_v3(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                       int64_t B, float* d_out, cudaStream_t st);
}
namespace v4 {
int launch_siren_tc_v4(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                       int64_t B, float* d_out, cudaStream_t st);
}
namespace v7 {
int launch_siren_tc_v7(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                       int64_t B, float* d_out, cudaStream_t st);
}
namespace v6 {
int launch_siren_tc_v6(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                       int64_t B, float* d_out, cudaStream_t st);
}
namespace v5 {
int launch_siren_tc_v5(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                       int64_t B, float* d_out, cudaStream_t st);
}

int launch_siren_tc(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx, const int32_t* d_nactive,
                    int64_t B, float* d_out, cudaStream_t st) {
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("MRINR_TC_VARIANT");
    variant = (e && e[0] >= '1' && e[0] <= '7') ? (e[0] - '0') : 5;
  }
  if (variant == 1 || p->L < 3) return v1::launch_siren_tc_v1(p, d_mods, d_idx, d_nactive, B, d_out, st);
  if (variant == 2) return v2::launch_siren_tc_v2(p, d_mods, d_idx, d_nactive, B, d_out, st);
  if (variant == 3) return v3::launch_siren_tc_v3(p, d_mods, d_idx, d_nactive, B, d_out, st);
  if (variant == 4) return v4::launch_siren_tc_v4(p, d_mods, d_idx, d_nactive, B, d_out, st);
  if (variant == 7) return v7::launch_siren_tc_v7(p, d_mods, d_idx, d_nactive, B, d_out, st);
  if (variant == 5) return v5::launch_siren_tc_v5(p, d_mods, d_idx, d_nactive, B, d_out, st);
  return v6::launch_siren_tc_v6(p, d_mods, d_idx, d_nactive, B, d_out, st);
}

}  // namespace mrinr
