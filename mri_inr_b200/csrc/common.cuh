// Shared host/device helpers for libmrinr (sm_100a only).
#pragma once
#include <atomic>

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mrinr.h"

namespace mrinr {

// ---- error plumbing (api.cu) -------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int  check_launch(const char* what);   // cudaPeekAtLastError -> code (+message)

#define MRINR_REQUIRE(cond, code, ...)            \
  do {                                            \
    if (!(cond)) {                                \
      ::mrinr::set_error(__VA_ARGS__);            \
      return (code);                              \
    }                                             \
  } while (0)

#define MRINR_CUDA(expr)                                                          \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      ::mrinr::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));         \
      return (int)_e;                                                             \
    }                                                                             \
  } while (0)

// "Do this once per device": the opt-in to more than 48 KB of dynamic shared memory (cudaFuncSetAttribute) is a
// per-device property of a kernel, so a process-wide `static bool` is wrong as soon as one process drives two GPUs
// (and racy between host threads).  One bit per device ordinal; setting the attribute twice is harmless.
struct DeviceOnce {
  std::atomic<unsigned long long> mask[4];          // 256 device ordinals
  bool need(int dev) const { return !(mask[(dev >> 6) & 3].load(std::memory_order_acquire) & (1ull << (dev & 63))); }
  void done(int dev) { mask[(dev >> 6) & 3].fetch_or(1ull << (dev & 63), std::memory_order_release); }
};
#define MRINR_SMEM_OPT_IN(kernel_expr, bytes)                                                                     \
  do {                                                                                                            \
    static DeviceOnce once_;                                                                                      \
    int dev_ = 0;                                                                                                 \
    MRINR_CUDA(cudaGetDevice(&dev_));                                                                             \
    if (once_.need(dev_)) {                                                                                       \
      MRINR_CUDA(cudaFuncSetAttribute((kernel_expr), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      once_.done(dev_);                                                                                           \
    }                                                                                                             \
  } while (0)


__host__ __device__ static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- packed weights (opaque to callers) ----------------------------------------------------------
}  // namespace mrinr

struct MrinrPacked {
  int32_t H, L, Z, S, C;
  float w0, w0_initial;
  int32_t activation, precision;
  int32_t device;
  int32_t num_sms;
  int32_t synth_clusters;   // CTA pairs the synthesis kernel may use (0 = one per SM pair); mrinr_set_synthesis_clusters
  float*    d_table0;    // [C,H]  act_0(W_0 g_c + b_0), fp32 (before modulation)
  uint16_t* d_table16;   // [C,H]  the same table in the tensor-core operand format (fp16 / bf16)
  float*    d_net_wT;    // [(L-1)][H(k)][H(n)]  fp32 transposed hidden weights (fp32 mode)
  uint16_t* d_net_w16;   // [(L-1)][H/8 (kc)][H (n)][8]  fp16/bf16 UMMA K-major no-swizzle operand layout
  uint16_t* d_net_w16p;  // [(L-1)][2 (rank)][H/8 (kc)][H/2 (n)][8]  same, split by output-row half for cta_group::2
  uint16_t* d_net_w16q;  // [(L-1)][2][H/8 + 2][H/2][8]  pair layout + one K=16 step carrying the bias (hi, lo)
  uint16_t* d_net_w16x3; // [(L-1)][2 (hi, lo)][2][H/8 + 2][H/2][8]  MRINR_PREC_FP16X3: W = hi + lo, bias step in the hi part
  float*    d_net_bias;  // [L][H]   (zeros when use_bias=False)
  float*    d_layer0;    // [3][H]   W_0[:,0], W_0[:,1], b_0 (layer 0 evaluated in-kernel by the cta_group::2 path)
  float*    d_grid;      // [C][2]   copy of the grid buffer
  float*    d_last_w;    // [H]
  float*    d_last_b;    // [1]
  float*    d_mod_wT;    // layer 0: [Z][H]; layer i>=1: [(H+Z)][H] -- concatenated, transposed
  float*    d_mod_bias;  // [L][H]
  int32_t*  d_errflag;   // device-side watchdog flag (mbarrier timeouts)
  // split-fp16 tensor-core operands (dense_tc.cu): [K/64 slabs][hi, lo][8][N][8] per layer
  uint16_t* d_mod_ws;    // modulator layers, concatenated (layer l at mod_ws_off[l], in uint16 elements)
  size_t    mod_ws_off[17];
  int32_t   mod_tc;      // 1: the modulator runs on the tensor cores
  // patch encoder (optional)
  int32_t   has_encoder;
  float*    d_enc_c1w;   // [16,1,3,3]
  float*    d_enc_c1b;   // [16]
  float*    d_enc_c2w;   // [32,16,3,3]
  float*    d_enc_c2b;   // [32]
  uint16_t* d_enc_w3s;   // Conv2d(32,64,8) as [64, 2048], split-packed
  float*    d_enc_b3;    // [64]
  uint16_t* d_enc_wfs;   // Linear(64, Z), split-packed
  float*    d_enc_bf;    // [Z]
};

namespace mrinr {

int launch_siren_fp32(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx,
                      const int32_t* d_nactive, int64_t B, float* d_out, cudaStream_t st);
int launch_siren_tc(const MrinrPacked* p, const float* d_mods, const int32_t* d_idx,
                    const int32_t* d_nactive, int64_t B, float* d_out, cudaStream_t st);
int run_transpose(const float* w, int N, int K, float* wT, cudaStream_t st);
int run_pack_w16(const float* w, int H, int use_bf16, uint16_t* out, cudaStream_t st);
int run_pack_w16_pair(const float* w, int H, int use_bf16, uint16_t* out, cudaStream_t st);
int run_pack_w16_pair_bias(const float* w, const float* bias, int H, int use_bf16, uint16_t* out, cudaStream_t st,
                           int part = 0);   // part 1: the fp16 residual w - rn16(w) (zero bias step)
int run_table16(const float* table, long long n, int use_bf16, uint16_t* out, cudaStream_t st);
int run_layer0_table(const float* grid, const float* w, const float* b, int C, int H, float w0_initial,
                     int activation, float* table, cudaStream_t st);
bool dense_split_supported(int N, int K1, int K2);
int run_pack_split(const float* w, int N, int K, uint16_t* out, cudaStream_t st);
int launch_dense_split(const float* a1, long long lda1, int K1, const float* a2, long long lda2, int K2,
                       const uint16_t* w_packed, const float* bias, int N, int act, float slope, float* c,
                       long long ldc, long long M, int32_t* errflag, cudaStream_t st);
int launch_dense_chain256(int n, const float* const* a1, const long long* lda1, const int* K1, const float* const* a2,
                          const long long* lda2, const int* K2, const uint16_t* const* w_packed,
                          const float* const* bias, float* const* c, const long long* ldc, int act, float slope,
                          long long M, int32_t* errflag, cudaStream_t st);
int launch_encoder_conv(const float* d_patches, long long B, const float* w1, const float* b1, const float* w2,
                        const float* b2, float* d_out, int num_sms, cudaStream_t st);
int launch_encoder_conv_tc(const float* d_patches, long long B, const float* w1, const float* b1, const float* w2,
                           const float* b2, float* d_out, int num_sms, int32_t* errflag, cudaStream_t st);
int64_t wgrad_tc_scratch_floats(int num_sms);
int launch_wgrad_tc(const float* dz, const float* h, long long M, float* scratch, float* dW, int num_sms,
                    int32_t* errflag, cudaStream_t st);
int launch_compact_black(const uint8_t* d_black, int64_t B, int32_t C, int32_t* d_idx, int32_t* d_nactive,
                         int32_t* d_blocksums, float* d_out, cudaStream_t st);

}  // namespace mrinr
