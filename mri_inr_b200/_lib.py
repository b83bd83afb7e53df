"""ctypes binding of ``libmrinr.so`` (the C ABI declared in ``include/mrinr.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C mri_inr_b200/csrc`` for
sm_100a only.  There is no CPU or PyTorch fallback: if the shared library is missing or a call
fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# MRINR_LIB: development override (an instrumented build of the same library, e.g. build/libmrinr_tl.so)
LIB_PATH = os.environ.get("MRINR_LIB") or os.path.join(_HERE, "lib", "libmrinr.so")

# include/mrinr.h
ABI_VERSION = 4
ACT_SINE, ACT_MORLET = 0, 1
PREC_FP16, PREC_BF16, PREC_FP32, PREC_FP16X3 = 0, 1, 2, 3
PRECISIONS = {"fp16": PREC_FP16, "bf16": PREC_BF16, "fp32": PREC_FP32, "fp16x3": PREC_FP16X3}
ACTIVATIONS = {"sine": ACT_SINE, "morlet": ACT_MORLET}

# every symbol include/mrinr.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = (
    "mrinr_abi_version", "mrinr_last_error", "mrinr_launch_count",
    "mrinr_pack_weights", "mrinr_free_packed", "mrinr_packed_layer0_table",
    "mrinr_make_grid", "mrinr_encoder_workspace_bytes", "mrinr_encoder_forward", "mrinr_modulator_forward",
    "mrinr_siren_workspace_bytes", "mrinr_siren_forward",
    "mrinr_image_to_patches", "mrinr_classify_patches", "mrinr_patches_to_image",
    "mrinr_complex_abs", "mrinr_minmax_normalize",
    "mrinr_image_metrics_scratch_bytes", "mrinr_image_metrics",
    "mrinr_fft2c_workspace_bytes", "mrinr_fft2c", "mrinr_kspace_to_image",
    "mrinr_peer_alloc", "mrinr_peer_open", "mrinr_peer_close", "mrinr_peer_free",
    "mrinr_train_workspace_bytes", "mrinr_train_forward", "mrinr_train_backward",
    "mrinr_set_synthesis_clusters", "mrinr_refresh_weights",
)


class WeightsView(ctypes.Structure):
    """``MrinrWeightsView`` (include/mrinr.h)."""

    _fields_ = [
        ("dim_in", c_int32), ("dim_hidden", c_int32), ("dim_out", c_int32), ("num_layers", c_int32),
        ("latent_dim", c_int32), ("siren_patch_size", c_int32),
        ("w0", c_float), ("w0_initial", c_float),
        ("activation", c_int32), ("reserved", c_int32),
        ("d_grid", c_void_p),
        ("d_net_weight", POINTER(c_void_p)), ("d_net_bias", POINTER(c_void_p)),
        ("d_last_weight", c_void_p), ("d_last_bias", c_void_p),
        ("d_mod_weight", POINTER(c_void_p)), ("d_mod_bias", POINTER(c_void_p)),
        ("outer_patch_size", c_int32), ("reserved2", c_int32),
        ("d_enc_conv1_weight", c_void_p), ("d_enc_conv1_bias", c_void_p),
        ("d_enc_conv2_weight", c_void_p), ("d_enc_conv2_bias", c_void_p),
        ("d_enc_conv3_weight", c_void_p), ("d_enc_conv3_bias", c_void_p),
        ("d_enc_fc_weight", c_void_p), ("d_enc_fc_bias", c_void_p),
    ]


_lib: Optional[ctypes.CDLL] = None


def _declare(lib: ctypes.CDLL) -> None:
    lib.mrinr_abi_version.restype = c_int
    lib.mrinr_abi_version.argtypes = []
    lib.mrinr_last_error.restype = c_char_p
    lib.mrinr_last_error.argtypes = []
    lib.mrinr_launch_count.restype = c_int64
    lib.mrinr_launch_count.argtypes = []
    lib.mrinr_pack_weights.restype = c_int
    lib.mrinr_pack_weights.argtypes = [POINTER(WeightsView), c_int, c_void_p, POINTER(c_void_p)]
    lib.mrinr_free_packed.restype = None
    lib.mrinr_free_packed.argtypes = [c_void_p]
    lib.mrinr_packed_layer0_table.restype = c_int
    lib.mrinr_packed_layer0_table.argtypes = [c_void_p, c_void_p, c_void_p]
    lib.mrinr_make_grid.restype = c_int
    lib.mrinr_make_grid.argtypes = [c_int32, c_void_p, c_void_p]
    lib.mrinr_encoder_workspace_bytes.restype = c_int64
    lib.mrinr_encoder_workspace_bytes.argtypes = [c_int64]
    lib.mrinr_encoder_forward.restype = c_int
    lib.mrinr_encoder_forward.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p]
    lib.mrinr_modulator_forward.restype = c_int
    lib.mrinr_modulator_forward.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
    lib.mrinr_siren_workspace_bytes.restype = c_int64
    lib.mrinr_siren_workspace_bytes.argtypes = [c_int64]
    lib.mrinr_siren_forward.restype = c_int
    lib.mrinr_siren_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p]
    lib.mrinr_image_to_patches.restype = c_int
    lib.mrinr_image_to_patches.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                           c_void_p]
    lib.mrinr_classify_patches.restype = c_int
    lib.mrinr_classify_patches.argtypes = [c_void_p, c_int64, c_int32, c_void_p, c_void_p]
    lib.mrinr_patches_to_image.restype = c_int
    lib.mrinr_patches_to_image.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32,
                                           c_void_p, c_void_p]
    lib.mrinr_image_metrics_scratch_bytes.restype = c_int64
    lib.mrinr_image_metrics_scratch_bytes.argtypes = [c_int64]
    lib.mrinr_image_metrics.restype = c_int
    lib.mrinr_image_metrics.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_void_p]
    lib.mrinr_fft2c_workspace_bytes.restype = c_int64
    lib.mrinr_fft2c_workspace_bytes.argtypes = [c_int64, c_int32, c_int32]
    lib.mrinr_fft2c.restype = c_int
    lib.mrinr_fft2c.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_void_p]
    lib.mrinr_kspace_to_image.restype = c_int
    lib.mrinr_kspace_to_image.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_int64, c_void_p]
    lib.mrinr_complex_abs.restype = c_int
    lib.mrinr_complex_abs.argtypes = [c_void_p, c_int64, c_void_p, c_void_p]
    lib.mrinr_minmax_normalize.restype = c_int
    lib.mrinr_minmax_normalize.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]
    lib.mrinr_peer_alloc.restype = c_int
    lib.mrinr_peer_alloc.argtypes = [c_int64, POINTER(c_void_p), c_void_p]
    lib.mrinr_peer_open.restype = c_int
    lib.mrinr_peer_open.argtypes = [c_void_p, POINTER(c_void_p)]
    lib.mrinr_peer_close.restype = c_int
    lib.mrinr_peer_close.argtypes = [c_void_p]
    lib.mrinr_peer_free.restype = c_int
    lib.mrinr_peer_free.argtypes = [c_void_p]
    lib.mrinr_refresh_weights.restype = c_int
    lib.mrinr_refresh_weights.argtypes = [c_void_p, POINTER(WeightsView), c_void_p]
    lib.mrinr_set_synthesis_clusters.restype = c_int
    lib.mrinr_set_synthesis_clusters.argtypes = [c_void_p, c_int32]
    lib.mrinr_train_workspace_bytes.restype = c_int64
    lib.mrinr_train_workspace_bytes.argtypes = [c_void_p, c_int64]
    lib.mrinr_train_forward.restype = c_int
    lib.mrinr_train_forward.argtypes = [c_void_p, POINTER(WeightsView), c_void_p, c_int64, c_float, ctypes.c_uint64,
                                        c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
    lib.mrinr_train_backward.restype = c_int
    lib.mrinr_train_backward.argtypes = [c_void_p, POINTER(WeightsView), c_void_p, c_void_p, c_int64, c_float,
                                         ctypes.c_uint64, c_void_p, c_float, POINTER(WeightsView), c_void_p, c_int64,
                                         c_void_p]


def load() -> ctypes.CDLL:
    """Load ``libmrinr.so`` (once).  Raises ``RuntimeError`` when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"mri_inr_b200: {LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C mri_inr_b200/csrc` (needs nvcc with sm_100a support). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    _declare(lib)
    got = lib.mrinr_abi_version()
    if got != ABI_VERSION:
        raise RuntimeError(f"mri_inr_b200: libmrinr.so has ABI version {got}, expected {ABI_VERSION}; rebuild it")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    """Raise ``RuntimeError`` for a non-zero return code of the C ABI."""
    if rc != 0:
        msg = load().mrinr_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libmrinr {what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().mrinr_launch_count())


def stream_ptr(device) -> int:
    import torch

    return int(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name: str, dtype=None):
    """Boundary checks shared by the wrappers: CUDA, contiguous, expected dtype."""
    import torch

    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: mri_inr_b200 has no CPU path")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    return t
