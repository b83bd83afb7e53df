"""CPU restatement of the reference's per-slice evaluation flow in the reference's own op vocabulary
(``F.pad``/``F.unfold``/``F.fold``, ``F.linear``, ``torch.sin``, in-place broadcast multiply), used as the timed
CPU arm of ``bench.py`` (``cpu_baseline`` / ``--impl reference``, kind "port") and cross-checked against the
index-arithmetic oracle in ``tests/``.  Test / measurement infrastructure only (see ``oracle/__init__.py``).

It materialises what the reference materialises (``grid.repeat(B,1,1)``, one ``[B,576,256]`` tensor per op), so
its run time is representative of ``test_mod_siren.py:196-234`` + ``src/util/error.py:230-248`` without the
skimage metrics and file I/O.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import siren
from .tiling import weight_matrix


def extract_patches(img: torch.Tensor, outer: int, inner: int) -> Tuple[torch.Tensor, Tuple[int, int]]:
    """src/util/tiling.py:10-64 for one ``[H,W]`` image."""
    h, w = img.shape
    pad = (outer - inner) // 2
    vpad, hpad = (inner - h % inner) % inner, (inner - w % inner) % inner
    x = F.pad(img[None, None], (pad, pad + hpad, pad, pad + vpad), mode="reflect")
    cols = F.unfold(x, kernel_size=(outer, outer), stride=inner)              # [1, outer*outer, P]
    patches = cols.transpose(1, 2).reshape(-1, outer, outer).contiguous()
    return patches, ((h + vpad) // inner, (w + hpad) // inner)


def fold_average(tiles: torch.Tensor, grid_shape: Tuple[int, int], k: int, inner: int, weighted: bool) -> torch.Tensor:
    """src/util/tiling.py:91-181 -> ``[nV*inner, nH*inner]``."""
    nv, nh = grid_shape
    size = (nv * inner, nh * inner)
    pad = (k - inner) // 2
    w = torch.from_numpy(weight_matrix(k)) if weighted else torch.ones(k, k)
    num = (tiles * w).reshape(-1, k * k).t()
    den = (torch.ones_like(tiles) * w).reshape(-1, k * k).t()
    img = F.fold(num, size, kernel_size=(k, k), stride=inner, padding=pad)
    img = img / F.fold(den, size, kernel_size=(k, k), stride=inner, padding=pad)
    return img.reshape(size)


def synthesis_reference_ops(sd: Dict[str, torch.Tensor], mods, num_layers: int, w0: float, w0_initial: float,
                            activation: str) -> torch.Tensor:
    """SirenNet.forward with the reference's op sequence (modulated_siren.py:448, :154-156, :227-233)."""
    b = mods[0].shape[0]
    x = sd["grid"].clone().repeat(b, 1, 1)
    for l in range(num_layers):
        pre = F.linear(x, sd[f"net.layers.{l}.weight"], sd.get(f"net.layers.{l}.bias"))
        lw0 = w0_initial if l == 0 else w0
        x = torch.sin(lw0 * pre)
        if activation == "morlet":
            x = x * torch.exp(-0.5 * pre ** 2)
        x *= mods[l].unsqueeze(1)
    pre = F.linear(x, sd["net.last_layer.weight"], sd.get("net.last_layer.bias"))
    return torch.sin(w0 * pre).squeeze(2)


@torch.no_grad()
def reconstruct_slice(sd: Dict[str, torch.Tensor], under: torch.Tensor, *, num_layers: int = 5, w0: float = 1.0,
                      w0_initial: float = 30.0, activation: str = "sine", outer: int = 32, inner: int = 16,
                      s: int = 24) -> torch.Tensor:
    """One undersampled ``[H,W]`` slice -> reconstructed ``[nV*inner, nH*inner]`` (error.py:230-248)."""
    patches, grid_shape = extract_patches(under, outer, inner)
    keep = [i for i in range(patches.shape[0]) if not (patches[i].mean() < 1e-10)]      # tiling.py:244-271
    kept = patches[keep]
    full = torch.zeros(patches.shape[0], s, s)
    if len(keep):
        z = siren.encoder_forward(sd, kept)
        mods = siren.modulator_forward(sd, z, num_layers)
        out = synthesis_reference_ops(sd, mods, num_layers, w0, w0_initial, activation)
        full[keep] = out.reshape(-1, s, s)
    return fold_average(full, grid_shape, s, inner, weighted=True)
