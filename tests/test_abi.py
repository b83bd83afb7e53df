"""CPU: the C-ABI shared library loads and exports every symbol include/mrinr.h declares.  No compute calls."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mrinr.h")


@pytest.fixture(scope="module")
def lib():
    from mri_inr_b200 import _lib

    if not os.path.isfile(_lib.LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "mri_inr_b200", "csrc"), "-j4"], check=True)
    return _lib.load()


def header_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"MRINR_API[^;(]*?\b(mrinr_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_symbols():
    from mri_inr_b200 import _lib

    assert header_symbols() == sorted(_lib.EXPORTED_SYMBOLS)
    assert len(header_symbols()) == 31


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), f"libmrinr.so does not export {name}"
        assert isinstance(getattr(lib, name), ctypes._CFuncPtr)


def test_abi_version_and_host_only_entry_points(lib):
    text = open(HEADER).read()
    want = int(re.search(r"#define MRINR_ABI_VERSION (\d+)", text).group(1))
    assert lib.mrinr_abi_version() == want
    assert isinstance(lib.mrinr_last_error(), bytes)
    assert lib.mrinr_launch_count() >= 0
    # workspace sizing is pure host arithmetic: n_active + index list + block sums, 16-byte aligned pieces
    assert lib.mrinr_siren_workspace_bytes(0) >= 16
    assert lib.mrinr_siren_workspace_bytes(400) >= 16 + 400 * 4
    assert lib.mrinr_siren_workspace_bytes(4_136_000) >= 4_136_000 * 4


def test_error_codes_match_header(lib):
    from mri_inr_b200 import _lib

    text = open(HEADER).read()
    for name, val in (("MRINR_PREC_FP16", _lib.PREC_FP16), ("MRINR_PREC_BF16", _lib.PREC_BF16),
                      ("MRINR_PREC_FP32", _lib.PREC_FP32), ("MRINR_ACT_SINE", _lib.ACT_SINE),
                      ("MRINR_ACT_MORLET", _lib.ACT_MORLET)):
        assert int(re.search(rf"#define {name}\s+(-?\d+)", text).group(1)) == val


def test_argument_errors_do_not_need_a_gpu(lib):
    """Null / bad arguments are rejected before any CUDA call: negative code + message."""
    rc = lib.mrinr_make_grid(24, None, None)
    assert rc == -1 and b"null" in lib.mrinr_last_error()
    rc = lib.mrinr_image_to_patches(None, 1, 320, 320, 32, 16, None, None, None)
    assert rc == -1
    rc = lib.mrinr_image_to_patches(ctypes.c_void_p(16), 1, 320, 320, 30, 16, ctypes.c_void_p(16), None, None)
    assert rc == -2 and b"O % 4" in lib.mrinr_last_error()
    # reflect padding must be smaller than the image (torch raises for the same input, tiling.py:40-44)
    rc = lib.mrinr_image_to_patches(ctypes.c_void_p(16), 1, 8, 8, 32, 16, ctypes.c_void_p(16), None, None)
    assert rc == -2
    assert lib.mrinr_image_to_patches(None, 0, 320, 320, 32, 16, None, None, None) == 0      # empty batch
    assert lib.mrinr_pack_weights(None, 0, None, None) == -1


def test_new_entry_points_validate_arguments_without_a_gpu(lib):
    """Encoder, k-space front end and metrics: sizing is host arithmetic; null / unsupported / undersized arguments are
    rejected before any CUDA call; empty batches succeed."""
    p = ctypes.c_void_p
    # workspace sizing
    assert lib.mrinr_encoder_workspace_bytes(0) == 0
    assert lib.mrinr_encoder_workspace_bytes(94000) == 94000 * (2048 + 64) * 4
    assert lib.mrinr_fft2c_workspace_bytes(3, 320, 320) == 3 * 320 * 320 * 8
    assert lib.mrinr_image_metrics_scratch_bytes(10) >= 10 * 32
    # empty batches are no-ops
    assert lib.mrinr_encoder_forward(None, None, 0, None, None, 0, None) == 0
    assert lib.mrinr_fft2c(None, 0, 320, 320, 1, None, None, 0, None) == 0
    assert lib.mrinr_kspace_to_image(None, None, 0, 320, 320, None, None, 0, None) == 0
    assert lib.mrinr_image_metrics(None, None, 0, 320, 320, None, None, 0, None) == 0
    # null pointers
    assert lib.mrinr_encoder_forward(None, None, 4, None, None, 0, None) == -1
    assert lib.mrinr_fft2c(None, 1, 320, 320, 1, None, None, 0, None) == -1 and b"null" in lib.mrinr_last_error()
    assert lib.mrinr_image_metrics(None, None, 2, 320, 320, None, None, 0, None) == -1
    # undersized workspace / scratch
    rc = lib.mrinr_fft2c(p(64), 1, 320, 320, 1, p(64), p(64), 16, None)
    assert rc == -1 and b"workspace" in lib.mrinr_last_error()
    rc = lib.mrinr_image_metrics(p(64), p(64), 2, 320, 320, p(64), p(64), 8, None)
    assert rc == -1 and b"scratch" in lib.mrinr_last_error()
    # images smaller than the 7x7 SSIM window
    assert lib.mrinr_image_metrics(p(64), p(64), 1, 5, 320, p(64), p(64), 1 << 20, None) == -1
    # misaligned buffers
    rc = lib.mrinr_kspace_to_image(p(68), None, 1, 320, 320, p(64), p(64), 1 << 30, None)
    assert rc == -3
    # sizes with a prime factor other than 2, 3, 5 (or above 1024) are refused, not silently mis-transformed
    rc = lib.mrinr_fft2c(p(64), 1, 14, 320, 1, p(64), p(64), 1 << 30, None)
    assert rc == -2 and b"products of 2, 3 and 5" in lib.mrinr_last_error()
    assert lib.mrinr_fft2c(p(64), 1, 2048, 320, 1, p(64), p(64), 1 << 30, None) == -2
