"""Mirror of the metric helpers of ``src/util/error.py`` (:23-84) on the device.

The reference converts every image to numpy on the host and calls scikit-image (``peak_signal_noise_ratio``,
``structural_similarity``, ``normalized_root_mse``, error.py:10-12) with ``data_range = max(both) - min(both)``
(:23-38).  Here the same numbers come from ``mrinr_image_metrics`` (csrc/metrics.cu) for CUDA tensors; a batch of
slices costs four launches and 24 bytes per slice of device-to-host traffic.  CUDA only, like the rest of the package.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import ops

__all__ = ["calculate_data_range", "calculate_psnr", "calculate_ssim", "calculate_nrmse", "image_metrics"]

image_metrics = ops.image_metrics


def _pair(original: torch.Tensor, predicted: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    if not (torch.is_tensor(original) and torch.is_tensor(predicted) and original.is_cuda and predicted.is_cuda):
        raise RuntimeError("mri_inr_b200.metrics works on CUDA tensors (there is no CPU path); the reference's "
                           "numpy/scikit-image helpers live in src/util/error.py")
    return original.squeeze().to(torch.float32), predicted.squeeze().to(torch.float32)


def calculate_data_range(original: torch.Tensor, predicted: torch.Tensor) -> float:
    """error.py:23-38."""
    o, p = _pair(original, predicted)
    return float(torch.maximum(o.max(), p.max()) - torch.minimum(o.min(), p.min()))


def calculate_psnr(original: torch.Tensor, predicted: torch.Tensor) -> float:
    """error.py:41-54."""
    o, p = _pair(original, predicted)
    return float(ops.image_metrics(o, p)[0, 0])


def calculate_ssim(original: torch.Tensor, predicted: torch.Tensor) -> float:
    """error.py:57-70."""
    o, p = _pair(original, predicted)
    return float(ops.image_metrics(o, p)[0, 1])


def calculate_nrmse(original: torch.Tensor, predicted: torch.Tensor) -> float:
    """error.py:73-84."""
    o, p = _pair(original, predicted)
    return float(ops.image_metrics(o, p)[0, 2])
