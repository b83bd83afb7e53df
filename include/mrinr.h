/*
 * mrinr.h -- C ABI of libmrinr.so: the B200 (sm_100a) implementation of the mri-inr hot path
 * (modulated-SIREN dense forward + the tiling either side of it).
 *
 * Conventions
 *   - Every pointer argument named `d_*` (or documented "device") is a CUDA device pointer owned by
 *     the caller; the library never allocates or frees caller-visible buffers.  The only library-owned
 *     object is the opaque `MrinrPacked` handle (re-tiled weights + layer-0 table).
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises the
 *     device, so every entry point is CUDA-graph capturable except pack/free.
 *   - Return value: 0 = success; <0 = argument / shape / unsupported-configuration error (MRINR_E_*);
 *     >0 = a cudaError_t.  `mrinr_last_error()` returns a thread-local message for the last failure.
 *   - Row-major, fp32 at the boundary.  Integer indices (patch order, reflect indices, black mask)
 *     are bit-exact with the reference; floating-point tolerances are stated per function.
 *
 * Reference = MatteoWohlrapp/mri-inr (file:line cited per entry point; the reference is pure
 * Python/PyTorch and has no FFI of its own -- these are the functions a binding of this path
 * would replace; INTEGRATION.md shows the ctypes stub).
 */
#ifndef MRINR_H_
#define MRINR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRINR_ABI_VERSION 4

#if defined(__GNUC__)
#define MRINR_API __attribute__((visibility("default")))
#else
#define MRINR_API
#endif

/* error codes (negative) */
#define MRINR_E_ARG          (-1)  /* null pointer / non-positive size */
#define MRINR_E_UNSUPPORTED  (-2)  /* configuration outside what the kernels implement */
#define MRINR_E_ALIGN        (-3)  /* pointer not 16-byte aligned */
#define MRINR_E_ARCH         (-4)  /* device is not sm_100 */

/* activation of the hidden synthesis layers: src/networks/modulated_siren.py:31-80 */
#define MRINR_ACT_SINE   0
#define MRINR_ACT_MORLET 1

/* operand precision of the hidden-layer contractions in mrinr_siren_forward */
#define MRINR_PREC_FP16  0   /* tcgen05 kind::f16, fp16 operands (11-bit significand), fp32 accumulate  -- default */
#define MRINR_PREC_BF16  1   /* tcgen05 kind::f16, bf16 operands, fp32 accumulate                       */
#define MRINR_PREC_FP32  2   /* CUDA-core FFMA, fp32 throughout (exact-mode reference kernel)            */
#define MRINR_PREC_FP16X3 3  /* tcgen05 kind::f16 with split operands (hi + lo fp16 halves of activations and
                              * weights, three MMAs per product): fp32-class accuracy on the tensor cores for
                              * weights / modulations too large for an 11-bit significand        */

/* A view of the reference module's parameters (state_dict tensors, fp32, contiguous, on the device).
 * Layout = src/networks/modulated_siren.py: SirenNet (:160-213), Modulator (:304-323), grid (:427-433). */
typedef struct MrinrWeightsView {
  int32_t dim_in;            /* must be 2                                                      */
  int32_t dim_hidden;        /* H; tensor-core path requires 256, fp32 path H % 32 == 0, <=512 */
  int32_t dim_out;           /* must be 1 (modulated_siren.py:451-455 squeezes it)             */
  int32_t num_layers;        /* L hidden layers, 2..16                                         */
  int32_t latent_dim;        /* Z, multiple of 4, <= 1024 (patch encoder: 64, 128 or 256)      */
  int32_t siren_patch_size;  /* S; coordinates per patch C = S*S                               */
  float   w0;                /* hidden layers 1..L-1 and the output layer                      */
  float   w0_initial;        /* layer 0                                                        */
  int32_t activation;        /* MRINR_ACT_*; the output layer is always sine (:211-213)        */
  int32_t reserved;
  const float* d_grid;             /* [C,2]   "grid" buffer (the coordinates actually used, :448)   */
  const float* const* d_net_weight;/* host array of L device pointers: [H,2], then [H,H]            */
  const float* const* d_net_bias;  /* host array of L device pointers [H], or NULL (use_bias=False) */
  const float* d_last_weight;      /* [1,H]                                                         */
  const float* d_last_bias;        /* [1] or NULL                                                   */
  const float* const* d_mod_weight;/* host array of L device pointers: [H,Z], then [H,H+Z]          */
  const float* const* d_mod_bias;  /* host array of L device pointers [H]                           */
  /* Patch encoder (FixedAutoencoder.encoder, src/networks/encoding/siren_encoder.py:503-512), optional: all eight
   * pointers NULL = no encoder packed (mrinr_encoder_forward then fails with MRINR_E_ARG). */
  int32_t outer_patch_size;        /* O; the encoder is hard-wired to 32x32 inputs (siren_encoder.py:498-512) */
  int32_t reserved2;
  const float* d_enc_conv1_weight; /* [16,1,3,3]  encoder.encoder.encoder.0.weight */
  const float* d_enc_conv1_bias;   /* [16]                                         */
  const float* d_enc_conv2_weight; /* [32,16,3,3] encoder.encoder.encoder.2.weight */
  const float* d_enc_conv2_bias;   /* [32]                                         */
  const float* d_enc_conv3_weight; /* [64,32,8,8] encoder.encoder.encoder.4.weight */
  const float* d_enc_conv3_bias;   /* [64]                                         */
  const float* d_enc_fc_weight;    /* [Z,64]      encoder.encoder.encoder.7.weight */
  const float* d_enc_fc_bias;      /* [Z]                                          */
} MrinrWeightsView;

typedef struct MrinrPacked MrinrPacked;   /* opaque */

/* ---- library ------------------------------------------------------------------------------------ */
MRINR_API int         mrinr_abi_version(void);
MRINR_API const char* mrinr_last_error(void);
/* Number of kernel launches this library has enqueued since load (all threads). */
MRINR_API int64_t     mrinr_launch_count(void);

/* ---- weights ------------------------------------------------------------------------------------ */
/* Re-tiles the parameters for the kernels (tensor-core operand layout for W_1..W_{L-1}, transposed fp32
 * copies, the patch-independent layer-0 activation table sin(w0_initial*(W_0 g_c + b_0)) [C,H]).
 * Replaces what nn.Module construction + load_state_dict + .to(device) prepare in the reference
 * (test_mod_siren.py:96-120).  Synchronises `stream` before returning. */
MRINR_API int  mrinr_pack_weights(const MrinrWeightsView* view, int precision, void* stream, MrinrPacked** out);
MRINR_API void mrinr_free_packed(MrinrPacked* p);
/* Copies the layer-0 table [C,H] fp32 to d_out (test/introspection hook). */
MRINR_API int  mrinr_packed_layer0_table(const MrinrPacked* p, float* d_out, void* stream);

/* Re-derive every operand copy of an existing handle from NEW parameter values of the SAME configuration (same
 * dimensions, activation, w0, precision, encoder present or not): no allocation, no synchronisation, all work on
 * `stream`.  What a training loop calls after every optimizer step instead of free + pack. */
MRINR_API int mrinr_refresh_weights(MrinrPacked* p, const MrinrWeightsView* weights, void* stream);

/* The persistent synthesis kernel launches one CTA pair per SM pair (74 on a B200).  `clusters` > 0 caps that number,
 * leaving SMs free for kernels of another stream (the front end of the next chunk); 0 restores the default. */
MRINR_API int mrinr_set_synthesis_clusters(MrinrPacked* p, int32_t clusters);

/* ---- coordinate grid: modulated_siren.py:427-433 -------------------------------------------------- */
/* d_out [S*S,2]: out[c] = (lin[c / S], lin[c % S]), lin = torch.linspace(-1,1,S).  Bit-exact. */
MRINR_API int mrinr_make_grid(int32_t S, float* d_out, void* stream);

/* ---- patch encoder: modulated_siren.py:282-301 (Encoder.forward) -> siren_encoder.py:565-577,503-512 ---------- */
/* d_patches [B,32,32] -> d_latent [B,Z]: Conv(1,16,3,s2,p1)+LeakyReLU(.2), Conv(16,32,3,s2,p1)+LeakyReLU,
 * Conv(32,64,8)+LeakyReLU, Flatten, Linear(64,Z).  The two strided convolutions run as fp32 FFMA, the 8x8
 * convolution and the linear layer as split-fp16 tcgen05 products (three MMAs per product, ~1e-6 relative).
 * d_workspace: mrinr_encoder_workspace_bytes(B) bytes, 16-byte aligned.  Tolerance vs the reference: 1e-5. */
MRINR_API int64_t mrinr_encoder_workspace_bytes(int64_t B);
MRINR_API int mrinr_encoder_forward(const MrinrPacked* p, const float* d_patches, int64_t B, float* d_latent,
                          void* d_workspace, int64_t workspace_bytes, void* stream);

/* ---- modulator: modulated_siren.py:325-343 (Modulator.forward) ------------------------------------ */
/* d_latent [B,Z] -> d_mods [L,B,H]: h_0 = relu(A_0 z + c_0), h_i = relu(A_i [h_{i-1}; z] + c_i).
 * MRINR_PREC_FP16/BF16 with H in {64,128,256} and Z % 64 == 0: one split-fp16 tcgen05 launch per layer (three
 * MMAs per product, fp32 accumulate, ~1e-6 relative); otherwise one fp32 FFMA launch for all layers.
 * Tolerance vs the reference 1e-5 relative. */
MRINR_API int mrinr_modulator_forward(const MrinrPacked* p, const float* d_latent, int64_t B, float* d_mods,
                            void* stream);

/* ---- synthesis network: modulated_siren.py:215-233 (SirenNet.forward) over the grid of every patch -- */
/* d_mods [L,B,H] (>= 0), d_black [B] (nullable; 1 = patch is black -> its output rows are zero and its
 * coordinates are never evaluated: filter_and_remember_black_patches / reintegrate_black_patches,
 * src/util/tiling.py:244-303), d_out [B, S*S].  When d_black is given, d_workspace must hold
 * mrinr_siren_workspace_bytes(B) bytes (compacted index list); otherwise it may be NULL.
 * Precision as chosen at pack time.  Tolerance: max-abs 1e-3 on the output (north_star); the
 * fp32 mode agrees with the reference to ~1e-5. */
MRINR_API int64_t mrinr_siren_workspace_bytes(int64_t B);
MRINR_API int mrinr_siren_forward(const MrinrPacked* p, const float* d_mods, const uint8_t* d_black, int64_t B,
                        float* d_out, void* d_workspace, int64_t workspace_bytes, void* stream);

/* ---- tiling: src/util/tiling.py ------------------------------------------------------------------- */
/* image_to_patches (tiling.py:10-64) for N same-sized images + classify_patches (:184-198).
 * d_img [N,H,W] -> d_patches [N*nV*nH, O, O], d_black [N*nV*nH] (nullable; 1 iff fp32 mean < 1e-10).
 * nV = ceil(H/I), nH = ceil(W/I); reflect padding (O-I)/2 plus bottom/right padding to a multiple
 * of I.  Indices bit-exact.  Requires O >= I, (O-I) even, O % 4 == 0, pad < H, W. */
MRINR_API int mrinr_image_to_patches(const float* d_img, int64_t N, int32_t H, int32_t W, int32_t O, int32_t I,
                           float* d_patches, uint8_t* d_black, void* stream);

/* classify_patches (tiling.py:184-198) over n_patches already-extracted patches of `elems` floats each:
 * d_black[i] = 1 iff the fp32 mean of patch i is < 1e-10. */
MRINR_API int mrinr_classify_patches(const float* d_patches, int64_t n_patches, int32_t elems, uint8_t* d_black,
                           void* stream);

/* patches_to_image_weighted_average (tiling.py:91-140) / patches_to_image (:143-181) for N images.
 * d_tiles [N*nV*nH, K, K]; d_weights [K,K] (generate_weight_matrix, :67-88) or NULL for unit
 * weights; d_black nullable (black tiles contribute 0 with their weight, :287-301);
 * d_img [N, nV*I, nH*I].  Accumulation order equals F.fold's on CPU, so the result is bit-exact
 * against the fp32 CPU reference. */
MRINR_API int mrinr_patches_to_image(const float* d_tiles, const float* d_weights, const uint8_t* d_black,
                           int64_t N, int32_t nV, int32_t nH, int32_t K, int32_t I, float* d_img,
                           void* stream);

/* ---- magnitude / normalisation epilogue ------------------------------------------------------------ */
/* fastmri.complex_abs as called at src/data/preprocessing.py:58: d_in [n,2] -> d_out [n],
 * sqrt(re*re + im*im) with separately rounded products (bit-exact against torch fp32). */
MRINR_API int mrinr_complex_abs(const float* d_in, int64_t n, float* d_out, void* stream);
/* normalize_scan (src/util/visualization.py:113-126; per volume at preprocessing.py:127-137):
 * d_in [G, n] -> d_out [G, n], per group (x - min) / (max - min).  d_scratch: 2*G floats. Bit-exact. */
MRINR_API int mrinr_minmax_normalize(const float* d_in, int64_t G, int64_t n, float* d_out, float* d_scratch,
                           void* stream);

/* ---- k-space front end: src/data/preprocessing.py:49-58 (apply_mask -> fastmri.ifft2c -> fastmri.complex_abs) ---- */
/* Centred orthonormal 2-D FFT, fftshift(fft2(ifftshift(x), norm="ortho")) (inverse != 0: ifft2), of N complex images
 * d_in [N,H,W,2] -> d_out [N,H,W,2].  Hand-written mixed-radix Stockham kernels: H, W products of 2, 3, 5, <= 1024.
 * d_workspace: mrinr_fft2c_workspace_bytes(N,H,W) bytes.  Tolerance vs torch.fft: 2e-6 of the largest |element|. */
MRINR_API int64_t mrinr_fft2c_workspace_bytes(int64_t N, int32_t H, int32_t W);
MRINR_API int mrinr_fft2c(const float* d_in, int64_t N, int32_t H, int32_t W, int32_t inverse, float* d_out,
                void* d_workspace, int64_t workspace_bytes, void* stream);
/* load_mri_scan (preprocessing.py:33-60) after the file read: d_kspace [N,H,W,2] x column mask d_colmask [W]
 * (nullable = fully sampled) -> ifft2c -> magnitude d_mag [N,H,W], in two HBM passes (the magnitude is fused into
 * the column pass).  Follow with mrinr_minmax_normalize per volume (normalize_scan, preprocessing.py:127-137). */
MRINR_API int mrinr_kspace_to_image(const float* d_kspace, const uint8_t* d_colmask, int64_t N, int32_t H, int32_t W,
                          float* d_mag, void* d_workspace, int64_t workspace_bytes, void* stream);

/* ---- image-quality metrics of the evaluation loop: src/util/error.py:23-84 as called by metrics_error (:256-269) -- */
/* d_original, d_predicted [N,H,W] -> d_out [N,3] (fp64): PSNR, SSIM, NRMSE per pair with the reference's data range
 * (max over both images - min over both images).  scikit-image definitions with the defaults error.py uses: SSIM
 * with a 7x7 uniform window, K1 = .01, K2 = .03, sample covariance, 3-pixel border cropped; NRMSE euclidean.
 * d_scratch: mrinr_image_metrics_scratch_bytes(N) bytes, 8-byte aligned.  Tolerance vs fp64 scikit-image
 * arithmetic: 1e-3 dB / 1e-4 / 1e-6 relative (window sums are fp32, as scikit-image's are for fp32 images). */
MRINR_API int64_t mrinr_image_metrics_scratch_bytes(int64_t N);
MRINR_API int mrinr_image_metrics(const float* d_original, const float* d_predicted, int64_t N, int32_t H, int32_t W,
                        double* d_out, void* d_scratch, int64_t scratch_bytes, void* stream);

/* ---- training: forward in train() mode (dropout after every hidden activation, modulated_siren.py:124,154-156) and
 * the backward pass of the whole module, as Trainer._train_iteration needs them (src/train/training.py:177-207).
 * mrinr_train_forward: d_tiles [B,32,32] -> d_out [B,C]; the intermediates stay in d_workspace
 *   (mrinr_train_workspace_bytes(p, B) bytes, 256-byte aligned) for mrinr_train_backward, which must be given the same
 *   weights view, tiles, dropout_p, seed and mask.  `p` are the packed weights of the CURRENT parameter values;
 *   `weights` the fp32 parameters themselves.  Dropout: element kept iff hash(seed, layer, index) >= dropout_p * 2^32,
 *   kept values scaled by 1/(1-p); d_keep_mask (optional, uint8 [L][B*C][H], 1 = keep) replaces the hash.
 * mrinr_train_backward: d_dout [B,C] = d(loss)/d(out).  `grads` reuses the MrinrWeightsView layout: every non-null
 *   pointer is a ZERO-INITIALISED, writable fp32 buffer of the parameter's shape that receives d(loss)/d(parameter)
 *   (accumulated with atomics); the scalar fields and d_grid are ignored.  `grad_scale`: a power of two that brings
 *   max|d_dout| * grad_scale to ~2^10 (the tensor-core products split their operands into fp16 halves, which resolve
 *   22 bits only above 6e-5); applied to d_dout on the way in and divided out of the results.  1.0 = no scaling.
 * Requires dim_hidden 256, latent_dim in {64,128,256}, an encoder in `p`.  fp32-class arithmetic throughout. */
MRINR_API int64_t mrinr_train_workspace_bytes(const MrinrPacked* p, int64_t B);
MRINR_API int mrinr_train_forward(const MrinrPacked* p, const MrinrWeightsView* weights, const float* d_tiles, int64_t B,
                                  float dropout_p, uint64_t seed, const uint8_t* d_keep_mask, float* d_out,
                                  void* d_workspace, int64_t workspace_bytes, void* stream);
MRINR_API int mrinr_train_backward(const MrinrPacked* p, const MrinrWeightsView* weights, const float* d_tiles,
                                   const float* d_dout, int64_t B, float dropout_p, uint64_t seed,
                                   const uint8_t* d_keep_mask, float grad_scale, const MrinrWeightsView* grads,
                                   void* d_workspace, int64_t workspace_bytes, void* stream);

/* ---- peer memory for the one exchange step (gather of reconstructed slices to one rank) ---- */
/* The reference has no distributed code; the multi-GPU sweep gathers every rank's reconstructed slices on one rank.
 * Instead of a collective after the reassembly, the gathering rank exports its [n_total,H,W] buffer and every other
 * process maps it (CUDA IPC over NVLink / NVSwitch peer access), so that mrinr_patches_to_image on rank r stores
 * its slices straight into the final buffer: compute and exchange are one kernel.
 *   mrinr_peer_alloc : cudaMalloc `bytes` on the current device and export it -> d_ptr, handle (MRINR_PEER_HANDLE_BYTES)
 *   mrinr_peer_open  : map an exported buffer into this process, peer access from the current device enabled lazily
 *   mrinr_peer_close : unmap (opener), mrinr_peer_free: release (owner)
 * All four are host-synchronous set-up calls (not on the data path). */
#define MRINR_PEER_HANDLE_BYTES 64
MRINR_API int mrinr_peer_alloc(int64_t bytes, void** d_ptr, uint8_t* handle);
MRINR_API int mrinr_peer_open(const uint8_t* handle, void** d_ptr);
MRINR_API int mrinr_peer_close(void* d_ptr);
MRINR_API int mrinr_peer_free(void* d_ptr);

#ifdef __cplusplus
}
#endif
#endif /* MRINR_H_ */
