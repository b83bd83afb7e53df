"""Synthetic single-coil k-space slices for benchmarks and examples (there is no network for fastMRI data).

Follows the reference's preprocessing chain (src/data/preprocessing.py:33-60,127-137): k-space -> random column
mask (acceleration / center fraction) -> centred orthonormal inverse FFT -> magnitude -> per-volume min-max
normalisation.  The k-space front end -- mask, centred inverse FFT, magnitude (``ops.kspace_to_image``) and the
normalisation -- runs on this package's kernels; ``torch.fft`` (cuFFT) is only used to *make* the phantom's k-space.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _fft2c(x: torch.Tensor) -> torch.Tensor:
    return torch.fft.fftshift(torch.fft.fft2(torch.fft.ifftshift(x, dim=(-2, -1)), norm="ortho"), dim=(-2, -1))


def _ifft2c(x: torch.Tensor) -> torch.Tensor:
    return torch.fft.fftshift(torch.fft.ifft2(torch.fft.ifftshift(x, dim=(-2, -1)), norm="ortho"), dim=(-2, -1))


def column_mask(width: int, acceleration: int, center_fraction: float, seed: int) -> np.ndarray:
    """fastMRI-style random column mask: ``round(W*cf)`` centre columns plus Bernoulli columns so that the
    expected sampling rate is ``1/acceleration``."""
    rs = np.random.RandomState(seed)
    num_low = int(round(width * center_fraction))
    prob = (width / acceleration - num_low) / (width - num_low)
    mask = rs.uniform(size=width) < prob
    pad = (width - num_low + 1) // 2
    mask[pad:pad + num_low] = True
    return mask


@torch.no_grad()
def synthetic_slices(n_slices: int, height: int = 320, width: int = 320, *, device="cuda", seed: int = 1234,
                     acceleration: int = 6, center_fraction: float = 0.05, slices_per_volume: int = 11,
                     undersampled: bool = True, chunk: int = 440) -> torch.Tensor:
    """``[n_slices, H, W]`` fp32 normalised magnitude images of undersampled synthetic k-space."""
    device = torch.device(device)
    out = torch.empty(n_slices, height, width, dtype=torch.float32, device=device)
    yy = torch.linspace(-1, 1, height, device=device)[:, None]
    xx = torch.linspace(-1, 1, width, device=device)[None, :]
    ky = torch.fft.fftfreq(height, device=device)[:, None]
    kx = torch.fft.fftfreq(width, device=device)[None, :]
    lowpass = torch.exp(-(ky ** 2 + kx ** 2) / (2 * 0.03 ** 2))
    mask = torch.from_numpy(column_mask(width, acceleration, center_fraction, seed)).to(device)
    gen = torch.Generator(device=device)
    chunk = max(slices_per_volume, (chunk // slices_per_volume) * slices_per_volume)
    for s0 in range(0, n_slices, chunk):
        n = min(chunk, n_slices - s0)
        gen.manual_seed(seed + s0)
        field = torch.rand(n, height, width, device=device, generator=gen)
        field = torch.fft.ifft2(torch.fft.fft2(field) * lowpass).real
        ax = 0.55 + 0.3 * torch.rand(n, 1, 1, device=device, generator=gen)
        ay = 0.65 + 0.3 * torch.rand(n, 1, 1, device=device, generator=gen)
        support = ((xx / ax) ** 2 + (yy / ay) ** 2 <= 1.0).to(torch.float32)
        phantom = (field - field.amin(dim=(1, 2), keepdim=True)) * support
        k = torch.view_as_real(_fft2c(phantom.to(torch.complex64))).contiguous()
        # load_mri_scan (preprocessing.py:49-58): apply_mask -> ifft2c -> complex_abs
        mag = ops.kspace_to_image(k, mask if undersampled else None)
        # per-volume normalisation (preprocessing.py:127-137); a trailing partial volume is its own group
        full = (n // slices_per_volume) * slices_per_volume
        if full:
            out[s0:s0 + full] = ops.minmax_normalize(mag[:full].contiguous(), groups=full // slices_per_volume)
        if n > full:
            out[s0 + full:s0 + n] = ops.minmax_normalize(mag[full:n].contiguous(), groups=1)
    return out
