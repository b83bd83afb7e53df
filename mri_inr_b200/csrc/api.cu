// C-ABI entry points that are not pure kernel wrappers: error plumbing, weight packing, dispatch of
// the synthesis forward by precision mode.
#include "common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace mrinr {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Error codes that a kernel launch can never produce but that other components of the process leave behind in the
// runtime's (non-sticky) last-error slot: peer access enabled twice (this library's lazy CUDA-IPC mapping and a
// collective library's P2P transport may both enable it towards the same GPU), host memory registered twice.  They are
// not ours to report: a stale one would turn the next launch into a spurious failure (that mechanism -- a handled
// cudaHostRegister failure, then a launch check tripping over its stale code -- was the N=4 rc=1 of round 1's sweep).
static bool benign_stale_error(cudaError_t e) {
  return e == cudaErrorPeerAccessAlreadyEnabled || e == cudaErrorPeerAccessNotEnabled ||
         e == cudaErrorHostMemoryAlreadyRegistered || e == cudaErrorHostMemoryNotRegistered;
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (benign_stale_error(e)) {
    (void)cudaGetLastError();          // clear it, then look again
    e = cudaPeekAtLastError();
  }
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return (int)e;
  }
  return 0;
}

}  // namespace mrinr

using namespace mrinr;

extern "C" int mrinr_abi_version(void) { return MRINR_ABI_VERSION; }
extern "C" const char* mrinr_last_error(void) { return g_err; }
extern "C" int64_t mrinr_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" void mrinr_free_packed(MrinrPacked* p) {
  if (!p) return;
  cudaFree(p->d_table0);
  cudaFree(p->d_table16);    // lab builds only (null otherwise)
  cudaFree(p->d_net_wT);
  cudaFree(p->d_net_w16);    // lab
  cudaFree(p->d_net_w16p);   // lab
  cudaFree(p->d_net_w16q);
  cudaFree(p->d_net_w16x3);
  cudaFree(p->d_layer0);
  cudaFree(p->d_grid);
  cudaFree(p->d_net_bias);
  cudaFree(p->d_last_w);
  cudaFree(p->d_last_b);
  cudaFree(p->d_mod_wT);
  cudaFree(p->d_mod_bias);
  cudaFree(p->d_errflag);
  cudaFree(p->d_mod_ws);
  cudaFree(p->d_enc_c1w);
  cudaFree(p->d_enc_c1b);
  cudaFree(p->d_enc_c2w);
  cudaFree(p->d_enc_c2b);
  cudaFree(p->d_enc_w3s);
  cudaFree(p->d_enc_b3);
  cudaFree(p->d_enc_wfs);
  cudaFree(p->d_enc_bf);
  delete p;
}

static bool view_has_encoder(const MrinrWeightsView* v) {
  return v->d_enc_conv1_weight && v->d_enc_conv1_bias && v->d_enc_conv2_weight && v->d_enc_conv2_bias &&
         v->d_enc_conv3_weight && v->d_enc_conv3_bias && v->d_enc_fc_weight && v->d_enc_fc_bias;
}
static bool view_has_any_encoder(const MrinrWeightsView* v) {
  return v->d_enc_conv1_weight || v->d_enc_conv1_bias || v->d_enc_conv2_weight || v->d_enc_conv2_bias ||
         v->d_enc_conv3_weight || v->d_enc_conv3_bias || v->d_enc_fc_weight || v->d_enc_fc_bias;
}

// Everything derived from the parameter VALUES: re-tiled / split operand copies, the layer-0 table, bias copies.
// Writes into the buffers of an allocated MrinrPacked, all on `st`, no allocation and no synchronisation -- so that a
// training loop can refresh the handle after every optimizer step (mrinr_refresh_weights) instead of rebuilding it.
static int fill_packed(MrinrPacked* p, const MrinrWeightsView* v, cudaStream_t st) {
  const int H = p->H, L = p->L, Z = p->Z, C = p->C, precision = p->precision;
  const size_t w16q_layer = (size_t)2 * (H / 8 + 2) * (H / 2) * 8;     // elements per layer: [2 ranks][H/8+2][H/2][8]
  int rc = 0;
#define PK_RC(expr)             \
  do {                          \
    rc = (expr);                \
    if (rc != 0) return rc;     \
  } while (0)
  MRINR_CUDA(cudaMemsetAsync(p->d_net_bias, 0, (size_t)L * H * sizeof(float), st));
  MRINR_CUDA(cudaMemsetAsync(p->d_last_b, 0, sizeof(float), st));
  const float* b0 = (v->d_net_bias && v->d_net_bias[0]) ? v->d_net_bias[0] : nullptr;
  // layer-0 parameters as three [H] rows: W_0[:,0], W_0[:,1], b_0
  MRINR_CUDA(cudaMemsetAsync(p->d_layer0, 0, (size_t)3 * H * sizeof(float), st));
  MRINR_CUDA(cudaMemcpy2DAsync(p->d_layer0, sizeof(float), v->d_net_weight[0], 2 * sizeof(float), sizeof(float), H,
                               cudaMemcpyDeviceToDevice, st));
  MRINR_CUDA(cudaMemcpy2DAsync(p->d_layer0 + H, sizeof(float), v->d_net_weight[0] + 1, 2 * sizeof(float), sizeof(float), H,
                               cudaMemcpyDeviceToDevice, st));
  if (b0) MRINR_CUDA(cudaMemcpyAsync(p->d_layer0 + 2 * H, b0, H * sizeof(float), cudaMemcpyDeviceToDevice, st));
  MRINR_CUDA(cudaMemcpyAsync(p->d_grid, v->d_grid, (size_t)C * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  PK_RC(run_layer0_table(v->d_grid, v->d_net_weight[0], b0, C, H, p->w0_initial, p->activation, p->d_table0, st));
#ifdef MRINR_LAB
  PK_RC(run_table16(p->d_table0, (long long)C * H, precision == MRINR_PREC_BF16, p->d_table16, st));
#endif
  for (int l = 0; l < L; ++l) {
    if (v->d_net_bias && v->d_net_bias[l])
      MRINR_CUDA(cudaMemcpyAsync(p->d_net_bias + (size_t)l * H, v->d_net_bias[l], H * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st));
    MRINR_CUDA(cudaMemcpyAsync(p->d_mod_bias + (size_t)l * H, v->d_mod_bias[l], H * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
  }
  for (int l = 1; l < L; ++l) {
    PK_RC(run_transpose(v->d_net_weight[l], H, H, p->d_net_wT + (size_t)(l - 1) * H * H, st));
    if (precision != MRINR_PREC_FP32) {
#ifdef MRINR_LAB
      PK_RC(run_pack_w16(v->d_net_weight[l], H, precision == MRINR_PREC_BF16, p->d_net_w16 + (size_t)(l - 1) * H * H, st));
      PK_RC(run_pack_w16_pair(v->d_net_weight[l], H, precision == MRINR_PREC_BF16,
                              p->d_net_w16p + (size_t)(l - 1) * H * H, st));
#endif
      const float* bl = v->d_net_bias ? v->d_net_bias[l] : nullptr;
      if (precision == MRINR_PREC_FP16X3) {
        for (int part = 0; part < 2; ++part)
          PK_RC(run_pack_w16_pair_bias(v->d_net_weight[l], bl, H, 0,
                                       p->d_net_w16x3 + ((size_t)(l - 1) * 2 + part) * w16q_layer, st, part));
      } else {
        PK_RC(run_pack_w16_pair_bias(v->d_net_weight[l], bl, H, precision == MRINR_PREC_BF16,
                                     p->d_net_w16q + (size_t)(l - 1) * w16q_layer, st));
      }
    }
  }
  MRINR_CUDA(cudaMemcpyAsync(p->d_last_w, v->d_last_weight, H * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (v->d_last_bias)
    MRINR_CUDA(cudaMemcpyAsync(p->d_last_b, v->d_last_bias, sizeof(float), cudaMemcpyDeviceToDevice, st));
  // modulator: layer 0 is [H,Z]; layers >= 1 are [H, H+Z] with the hidden columns first
  float* dst = p->d_mod_wT;
  PK_RC(run_transpose(v->d_mod_weight[0], H, Z, dst, st));
  dst += (size_t)Z * H;
  for (int l = 1; l < L; ++l) {
    PK_RC(run_transpose(v->d_mod_weight[l], H, H + Z, dst, st));
    dst += (size_t)(H + Z) * H;
  }
  if (p->mod_tc) {      // modulator on the tensor cores: split-fp16 operands per layer (dense_tc.cu)
    for (int l = 0; l < L; ++l)
      PK_RC(run_pack_split(v->d_mod_weight[l], H, l == 0 ? Z : H + Z, p->d_mod_ws + p->mod_ws_off[l], st));
  }
  if (p->has_encoder) {
    MRINR_CUDA(cudaMemcpyAsync(p->d_enc_c1w, v->d_enc_conv1_weight, 16 * 9 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    MRINR_CUDA(cudaMemcpyAsync(p->d_enc_c1b, v->d_enc_conv1_bias, 16 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    MRINR_CUDA(cudaMemcpyAsync(p->d_enc_c2w, v->d_enc_conv2_weight, 32 * 16 * 9 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    MRINR_CUDA(cudaMemcpyAsync(p->d_enc_c2b, v->d_enc_conv2_bias, 32 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    MRINR_CUDA(cudaMemcpyAsync(p->d_enc_b3, v->d_enc_conv3_bias, 64 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    MRINR_CUDA(cudaMemcpyAsync(p->d_enc_bf, v->d_enc_fc_bias, (size_t)Z * sizeof(float), cudaMemcpyDeviceToDevice, st));
    PK_RC(run_pack_split(v->d_enc_conv3_weight, 64, 2048, p->d_enc_w3s, st));
    PK_RC(run_pack_split(v->d_enc_fc_weight, Z, 64, p->d_enc_wfs, st));
  }
#undef PK_RC
  return 0;
}

static int check_view_pointers(const MrinrWeightsView* v, const char* who) {
  MRINR_REQUIRE(v->d_grid && v->d_net_weight && v->d_last_weight && v->d_mod_weight && v->d_mod_bias, MRINR_E_ARG,
                "%s: null weight pointer", who);
  for (int l = 0; l < v->num_layers; ++l)
    MRINR_REQUIRE(v->d_net_weight[l] && v->d_mod_weight[l] && v->d_mod_bias[l], MRINR_E_ARG,
                  "%s: null pointer for layer %d", who, l);
  if (view_has_any_encoder(v))
    MRINR_REQUIRE(view_has_encoder(v), MRINR_E_ARG, "%s: the encoder needs all eight tensors (or none)", who);
  return 0;
}

extern "C" int mrinr_pack_weights(const MrinrWeightsView* v, int precision, void* stream, MrinrPacked** out) {
  MRINR_REQUIRE(v && out, MRINR_E_ARG, "mrinr_pack_weights: null argument");
  *out = nullptr;
  MRINR_REQUIRE(v->dim_in == 2, MRINR_E_UNSUPPORTED, "mrinr_pack_weights: dim_in must be 2 (got %d)", v->dim_in);
  MRINR_REQUIRE(v->dim_out == 1, MRINR_E_UNSUPPORTED,
                "mrinr_pack_weights: dim_out must be 1 (got %d); the reference squeezes it (modulated_siren.py:451)",
                v->dim_out);
  MRINR_REQUIRE(v->num_layers >= 2 && v->num_layers <= 16, MRINR_E_UNSUPPORTED,
                "mrinr_pack_weights: num_layers must be in [2,16] (got %d)", v->num_layers);
  MRINR_REQUIRE(v->dim_hidden >= 32 && v->dim_hidden <= 512 && v->dim_hidden % 32 == 0, MRINR_E_UNSUPPORTED,
                "mrinr_pack_weights: dim_hidden must be a multiple of 32 in [32,512] (got %d)", v->dim_hidden);
  MRINR_REQUIRE(v->latent_dim >= 4 && v->latent_dim <= 1024 && v->latent_dim % 4 == 0, MRINR_E_UNSUPPORTED,
                "mrinr_pack_weights: latent_dim must be a multiple of 4 in [4,1024] (got %d)", v->latent_dim);
  MRINR_REQUIRE(v->siren_patch_size >= 1 && v->siren_patch_size <= 1024, MRINR_E_UNSUPPORTED,
                "mrinr_pack_weights: siren_patch_size out of range (%d)", v->siren_patch_size);
  MRINR_REQUIRE(v->activation == MRINR_ACT_SINE || v->activation == MRINR_ACT_MORLET, MRINR_E_ARG,
                "mrinr_pack_weights: unknown activation %d", v->activation);
  MRINR_REQUIRE(precision == MRINR_PREC_FP16 || precision == MRINR_PREC_BF16 || precision == MRINR_PREC_FP32 ||
                    precision == MRINR_PREC_FP16X3,
                MRINR_E_ARG, "mrinr_pack_weights: unknown precision %d", precision);
  if (precision != MRINR_PREC_FP32) {
    MRINR_REQUIRE(v->dim_hidden == 256, MRINR_E_UNSUPPORTED,
                  "mrinr_pack_weights: the tensor-core path requires dim_hidden == 256 (got %d); use MRINR_PREC_FP32",
                  v->dim_hidden);
    MRINR_REQUIRE(v->siren_patch_size * v->siren_patch_size >= 128, MRINR_E_UNSUPPORTED,
                  "mrinr_pack_weights: the tensor-core path requires siren_patch_size^2 >= 128 (got %d)",
                  v->siren_patch_size);
  }
  int rc = check_view_pointers(v, "mrinr_pack_weights");
  if (rc != 0) return rc;
  if (view_has_any_encoder(v))
    MRINR_REQUIRE(v->outer_patch_size == 32 && dense_split_supported(v->latent_dim, 64, 0), MRINR_E_UNSUPPORTED,
                  "mrinr_pack_weights: the encoder is hard-wired to 32x32 patches (siren_encoder.py:498-512) and needs "
                  "latent_dim in {64,128,256} (got O=%d Z=%d)", v->outer_patch_size, v->latent_dim);

  int dev = 0;
  MRINR_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  MRINR_CUDA(cudaGetDeviceProperties(&prop, dev));
  MRINR_REQUIRE(prop.major == 10, MRINR_E_ARCH, "mrinr_pack_weights: device %d is sm_%d%d; this library is sm_100a only",
                dev, prop.major, prop.minor);

  cudaStream_t st = (cudaStream_t)stream;
  MrinrPacked* p = new (std::nothrow) MrinrPacked();
  MRINR_REQUIRE(p, MRINR_E_ARG, "mrinr_pack_weights: out of host memory");
  std::memset(p, 0, sizeof(*p));
  p->H = v->dim_hidden; p->L = v->num_layers; p->Z = v->latent_dim; p->S = v->siren_patch_size;
  p->C = p->S * p->S;
  p->w0 = v->w0; p->w0_initial = v->w0_initial; p->activation = v->activation; p->precision = precision;
  p->device = dev; p->num_sms = prop.multiProcessorCount;
  const int H = p->H, L = p->L, Z = p->Z, C = p->C;

#define PK_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      set_error("mrinr_pack_weights: %s failed: %s", #expr, cudaGetErrorString(_e));          \
      rc = (int)_e;                                                                           \
      goto fail;                                                                              \
    }                                                                                         \
  } while (0)

  {
    // ---- allocation (sizes depend on the configuration only) ...
    const size_t mod_w_elems = (size_t)Z * H + (size_t)(L - 1) * (H + Z) * H;
    const size_t w16q_layer = (size_t)2 * (H / 8 + 2) * (H / 2) * 8;
    PK_CUDA(cudaMalloc(&p->d_table0, (size_t)C * H * sizeof(float)));
    PK_CUDA(cudaMalloc(&p->d_net_wT, (size_t)(L - 1) * H * H * sizeof(float)));
#ifdef MRINR_LAB   // operand copies of the retired kernel variants (lab/): not part of the product library
    PK_CUDA(cudaMalloc(&p->d_table16, (size_t)C * H * sizeof(uint16_t)));
    PK_CUDA(cudaMalloc(&p->d_net_w16, (size_t)(L - 1) * H * H * sizeof(uint16_t)));
    PK_CUDA(cudaMalloc(&p->d_net_w16p, (size_t)(L - 1) * H * H * sizeof(uint16_t)));
#endif
    if (precision == MRINR_PREC_FP16 || precision == MRINR_PREC_BF16)
      PK_CUDA(cudaMalloc(&p->d_net_w16q, (size_t)(L - 1) * w16q_layer * sizeof(uint16_t)));
    if (precision == MRINR_PREC_FP16X3)
      PK_CUDA(cudaMalloc(&p->d_net_w16x3, (size_t)(L - 1) * 2 * w16q_layer * sizeof(uint16_t)));
    PK_CUDA(cudaMalloc(&p->d_layer0, (size_t)3 * H * sizeof(float)));
    PK_CUDA(cudaMalloc(&p->d_grid, (size_t)C * 2 * sizeof(float)));
    PK_CUDA(cudaMalloc(&p->d_net_bias, (size_t)L * H * sizeof(float)));
    PK_CUDA(cudaMalloc(&p->d_last_w, (size_t)H * sizeof(float)));
    PK_CUDA(cudaMalloc(&p->d_last_b, sizeof(float)));
    PK_CUDA(cudaMalloc(&p->d_mod_wT, mod_w_elems * sizeof(float)));
    PK_CUDA(cudaMalloc(&p->d_mod_bias, (size_t)L * H * sizeof(float)));
    PK_CUDA(cudaMalloc(&p->d_errflag, sizeof(int32_t)));
    PK_CUDA(cudaMemsetAsync(p->d_errflag, 0, sizeof(int32_t), st));
    // modulator on the tensor cores: split-fp16 operands per layer (dense_tc.cu)
    p->mod_tc = (precision != MRINR_PREC_FP32) && dense_split_supported(H, Z, 0) && dense_split_supported(H, H, Z);
    if (p->mod_tc) {
      PK_CUDA(cudaMalloc(&p->d_mod_ws, mod_w_elems * 2 * sizeof(uint16_t)));
      size_t off = 0;
      for (int l = 0; l < L; ++l) {
        p->mod_ws_off[l] = off;
        off += (size_t)2 * H * (l == 0 ? Z : H + Z);
      }
    }
    if (view_has_encoder(v)) {      // patch encoder (optional)
      PK_CUDA(cudaMalloc(&p->d_enc_c1w, 16 * 9 * sizeof(float)));
      PK_CUDA(cudaMalloc(&p->d_enc_c1b, 16 * sizeof(float)));
      PK_CUDA(cudaMalloc(&p->d_enc_c2w, 32 * 16 * 9 * sizeof(float)));
      PK_CUDA(cudaMalloc(&p->d_enc_c2b, 32 * sizeof(float)));
      PK_CUDA(cudaMalloc(&p->d_enc_w3s, (size_t)2 * 64 * 2048 * sizeof(uint16_t)));
      PK_CUDA(cudaMalloc(&p->d_enc_b3, 64 * sizeof(float)));
      PK_CUDA(cudaMalloc(&p->d_enc_wfs, (size_t)2 * Z * 64 * sizeof(uint16_t)));
      PK_CUDA(cudaMalloc(&p->d_enc_bf, (size_t)Z * sizeof(float)));
      p->has_encoder = 1;
    }
    // ---- ... then everything derived from the values
    rc = fill_packed(p, v, st);
    if (rc != 0) goto fail;
    PK_CUDA(cudaStreamSynchronize(st));
  }
  *out = p;
  return 0;
fail:
  mrinr_free_packed(p);
  return rc;
#undef PK_CUDA
}

extern "C" int mrinr_refresh_weights(MrinrPacked* p, const MrinrWeightsView* v, void* stream) {
  MRINR_REQUIRE(p && v, MRINR_E_ARG, "mrinr_refresh_weights: null argument");
  MRINR_REQUIRE(v->dim_in == 2 && v->dim_out == 1 && v->dim_hidden == p->H && v->num_layers == p->L &&
                    v->latent_dim == p->Z && v->siren_patch_size == p->S && v->activation == p->activation &&
                    v->w0 == p->w0 && v->w0_initial == p->w0_initial && (view_has_encoder(v) ? 1 : 0) == p->has_encoder,
                MRINR_E_ARG, "mrinr_refresh_weights: the view does not have the configuration the handle was packed for");
  const int rc = check_view_pointers(v, "mrinr_refresh_weights");
  if (rc != 0) return rc;
  return fill_packed(p, v, (cudaStream_t)stream);
}

extern "C" int mrinr_set_synthesis_clusters(MrinrPacked* p, int32_t clusters) {
  MRINR_REQUIRE(p && clusters >= 0, MRINR_E_ARG, "mrinr_set_synthesis_clusters: bad argument");
  p->synth_clusters = clusters;
  return 0;
}

extern "C" int mrinr_packed_layer0_table(const MrinrPacked* p, float* d_out, void* stream) {
  MRINR_REQUIRE(p && d_out, MRINR_E_ARG, "mrinr_packed_layer0_table: null argument");
  MRINR_CUDA(cudaMemcpyAsync(d_out, p->d_table0, (size_t)p->C * p->H * sizeof(float), cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
  return 0;
}

// workspace layout: [0,16) n_active (int32) | idx int32[B] | blocksums int32[ceil(B/1024)]
static inline size_t ws_idx_off() { return 16; }
static inline size_t ws_bsum_off(int64_t B) { return 16 + (((size_t)B * 4 + 15) & ~(size_t)15); }

extern "C" int64_t mrinr_siren_workspace_bytes(int64_t B) {
  if (B < 0) return 0;
  return (int64_t)(ws_bsum_off(B) + (((size_t)((B + 1023) / 1024) * 4 + 15) & ~(size_t)15) + 16);
}

extern "C" int mrinr_siren_forward(const MrinrPacked* p, const float* d_mods, const uint8_t* d_black, int64_t B,
                                   float* d_out, void* d_workspace, int64_t workspace_bytes, void* stream) {
  if (B == 0) return 0;
  MRINR_REQUIRE(p && d_mods && d_out, MRINR_E_ARG, "mrinr_siren_forward: null pointer");
  MRINR_REQUIRE(B >= 0 && B * (int64_t)p->C < ((int64_t)1 << 40), MRINR_E_ARG, "mrinr_siren_forward: bad batch %lld",
                (long long)B);
  MRINR_REQUIRE(B <= 0x7fffffff, MRINR_E_UNSUPPORTED, "mrinr_siren_forward: at most 2^31-1 patches per call");
  MRINR_REQUIRE(aligned16(d_mods) && aligned16(d_out), MRINR_E_ALIGN, "mrinr_siren_forward: buffers must be 16-byte aligned");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t* d_idx = nullptr;
  const int32_t* d_nactive = nullptr;
  if (d_black != nullptr) {
    MRINR_REQUIRE(d_workspace != nullptr && workspace_bytes >= mrinr_siren_workspace_bytes(B), MRINR_E_ARG,
                  "mrinr_siren_forward: a black mask needs a workspace of mrinr_siren_workspace_bytes(B) bytes");
    MRINR_REQUIRE(aligned16(d_workspace), MRINR_E_ALIGN, "mrinr_siren_forward: workspace must be 16-byte aligned");
    char* ws = static_cast<char*>(d_workspace);
    int32_t* nact = reinterpret_cast<int32_t*>(ws);
    int32_t* idx = reinterpret_cast<int32_t*>(ws + ws_idx_off());
    int32_t* bsum = reinterpret_cast<int32_t*>(ws + ws_bsum_off(B));
    const int rc = launch_compact_black(d_black, B, p->C, idx, nact, bsum, d_out, st);
    if (rc != 0) return rc;
    d_idx = idx;
    d_nactive = nact;
  }
  if (p->precision == MRINR_PREC_FP32) return launch_siren_fp32(p, d_mods, d_idx, d_nactive, B, d_out, st);
  return launch_siren_tc(p, d_mods, d_idx, d_nactive, B, d_out, st);
}

// ---- patch encoder ---------------------------------------------------------------------------------
// workspace: conv2 output [B,2048] fp32 | conv3 output [B,64] fp32
extern "C" int64_t mrinr_encoder_workspace_bytes(int64_t B) {
  if (B < 0) return 0;
  return (int64_t)((size_t)B * (2048 + 64) * sizeof(float));
}

extern "C" int mrinr_encoder_forward(const MrinrPacked* p, const float* d_patches, int64_t B, float* d_latent,
                                     void* d_workspace, int64_t workspace_bytes, void* stream) {
  if (B == 0) return 0;
  MRINR_REQUIRE(p && d_patches && d_latent, MRINR_E_ARG, "mrinr_encoder_forward: null pointer");
  MRINR_REQUIRE(B > 0, MRINR_E_ARG, "mrinr_encoder_forward: negative batch");
  MRINR_REQUIRE(p->has_encoder, MRINR_E_ARG, "mrinr_encoder_forward: no encoder weights were packed");
  MRINR_REQUIRE(d_workspace && workspace_bytes >= mrinr_encoder_workspace_bytes(B), MRINR_E_ARG,
                "mrinr_encoder_forward: needs a workspace of mrinr_encoder_workspace_bytes(B) bytes");
  MRINR_REQUIRE(aligned16(d_patches) && aligned16(d_latent) && aligned16(d_workspace), MRINR_E_ALIGN,
                "mrinr_encoder_forward: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  float* c2 = static_cast<float*>(d_workspace);
  float* c3 = c2 + (size_t)B * 2048;
  // conv1 + conv2: tensor-core implicit GEMM for conv2 (encoder_conv_tc.cu)
  int rc;
#ifdef MRINR_LAB   // lab library only: MRINR_ENC_VARIANT=ffma selects the retired all-FFMA kernel (lab/encoder_conv.cu)
  static int enc_variant = -1;
  if (enc_variant < 0) {
    const char* e = getenv("MRINR_ENC_VARIANT");
    enc_variant = (e && e[0] == 'f') ? 0 : 1;
  }
  if (!enc_variant)
    rc = launch_encoder_conv(d_patches, B, p->d_enc_c1w, p->d_enc_c1b, p->d_enc_c2w, p->d_enc_c2b, c2, p->num_sms, st);
  else
#endif
  rc = launch_encoder_conv_tc(d_patches, B, p->d_enc_c1w, p->d_enc_c1b, p->d_enc_c2w, p->d_enc_c2b, c2, p->num_sms,
                              p->d_errflag, st);
  if (rc != 0) return rc;
  // Conv2d(32,64,8) on the 8x8 map == [B,2048] x [2048,64] (weight [64,32,8,8] flattens in the same (c,y,x) order)
  rc = launch_dense_split(c2, 2048, 2048, nullptr, 0, 0, p->d_enc_w3s, p->d_enc_b3, 64, /*leaky*/ 2, 0.2f, c3, 64, B,
                          p->d_errflag, st);
  if (rc != 0) return rc;
  return launch_dense_split(c3, 64, 64, nullptr, 0, 0, p->d_enc_wfs, p->d_enc_bf, p->Z, /*none*/ 0, 0.f, d_latent, p->Z,
                            B, p->d_errflag, st);
}

// ---- peer memory (include/mrinr.h: the exchange step fused into the reassembly kernel) ----------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == MRINR_PEER_HANDLE_BYTES, "IPC handle size");

extern "C" int mrinr_peer_alloc(int64_t bytes, void** d_ptr, uint8_t* handle) {
  MRINR_REQUIRE(bytes > 0 && d_ptr && handle, MRINR_E_ARG, "mrinr_peer_alloc: bad arguments");
  void* p = nullptr;
  MRINR_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  memcpy(handle, &h, sizeof(h));
  *d_ptr = p;
  return 0;
}

extern "C" int mrinr_peer_open(const uint8_t* handle, void** d_ptr) {
  MRINR_REQUIRE(handle && d_ptr, MRINR_E_ARG, "mrinr_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  MRINR_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  // the lazy peer-access enable can leave cudaErrorPeerAccessAlreadyEnabled behind when NCCL got there first
  if (benign_stale_error(cudaPeekAtLastError())) (void)cudaGetLastError();
  *d_ptr = p;
  return 0;
}

extern "C" int mrinr_peer_close(void* d_ptr) {
  if (!d_ptr) return 0;
  MRINR_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return 0;
}

extern "C" int mrinr_peer_free(void* d_ptr) {
  if (!d_ptr) return 0;
  MRINR_CUDA(cudaFree(d_ptr));
  return 0;
}
