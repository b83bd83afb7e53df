"""Drop-in for the evaluation helpers of the reference's ``src/util/error.py`` that sit on the hot path:
``metrics_error`` (:200-271) and ``calculate_data_range / psnr / ssim / nrmse`` (:23-84), same names, argument
meaning and return values -- but nothing leaves the device until the three floats are read.

The reference's ``metrics_error`` filters the black patches with a Python loop (400 device syncs per slice), runs the
model, re-inserts the black patches with another loop, reassembles three images, copies them to the host and calls
scikit-image three times.  Here: one classifier launch, the fused forward with the black mask (black patches are
skipped and zero-filled by the synthesis kernel: the same result as filter + reintegrate), two reassembly launches and
``mrinr_image_metrics``.  ``visual_error`` / ``calculate_difference`` (PNG output through matplotlib) are out of scope.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import ops
from .metrics import calculate_data_range, calculate_nrmse, calculate_psnr, calculate_ssim  # noqa: F401  (re-export)
from .tiling import _weights_on

__all__ = ["metrics_error", "calculate_data_range", "calculate_psnr", "calculate_ssim", "calculate_nrmse"]


@torch.no_grad()
def metrics_error(model, fully_sampled: torch.Tensor, undersampled: torch.Tensor,
                  img_information: Sequence[Tuple[int, int]], device, outer_patch_size: int, inner_patch_size: int,
                  siren_patch_size: int) -> Tuple[float, float, float]:
    """error.py:200-271.  ``fully_sampled`` / ``undersampled``: the ``[nV*nH, outer, outer]`` patches of ONE image as
    returned by ``image_to_patches`` (test_mod_siren.py:206-229); returns ``(psnr, ssim, nrmse)`` of the fully sampled
    image against the reconstruction."""
    dev = torch.device(device)
    nv, nh = img_information[0]
    under = undersampled.to(dev, torch.float32).contiguous()
    full = fully_sampled.to(dev, torch.float32).contiguous()
    if under.shape[0] != nv * nh:
        raise RuntimeError(f"{under.shape[0]} patches do not match an image of {nv}x{nh} patches")
    model._check_inference()
    # filter_and_remember_black_patches + model + reintegrate_black_patches (error.py:230-243)
    black = ops.classify_patches(under)
    mods = model.modulations(under)
    tiles = model.synthesize(mods, black=black)
    # patches_to_image_weighted_average / patches_to_image (error.py:245-254)
    recon = ops.patches_to_image(tiles, 1, (nv, nh), inner_patch_size, weights=_weights_on(siren_patch_size, dev),
                                 black=black)
    full_img = ops.patches_to_image(full, 1, (nv, nh), inner_patch_size)
    m = ops.image_metrics(full_img, recon)[0].tolist()      # the only device -> host transfer: 24 bytes
    return m[0], m[1], m[2]
