"""CPU restatement of one training iteration of the reference (src/train/training.py:177-207): the module's forward in
``train()`` mode -- dropout after every hidden activation (src/networks/modulated_siren.py:124,154-156), applied
BEFORE the modulation (:227-231) -- MSE against the centre crop of the fully sampled patches
(``extract_center_batch``, src/util/tiling.py:306-322) and autograd with respect to every parameter.
Test infrastructure only (see ``oracle/__init__.py``); pinned against goldens produced by the unmodified reference's own
autograd (``tests/golden/training.npz``, generator ``oracle/make_golden.py``).

The reference draws its dropout masks from torch's RNG; parity is therefore stated for dropout off and for a FIXED
keep-mask applied on both sides (kept values scaled by 1/(1-p), as ``nn.Dropout`` does).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from . import siren


def extract_center_batch(batch: torch.Tensor, outer: int, inner: int) -> torch.Tensor:
    """tiling.py:306-322."""
    pad = (outer - inner) // 2
    return batch[:, pad:pad + inner, pad:pad + inner]


def forward_train(sd: Dict[str, torch.Tensor], tiles: torch.Tensor, keep: Optional[torch.Tensor], p: float, *,
                  num_layers: int = 5, w0: float = 1.0, w0_initial: float = 30.0, activation: str = "sine",
                  siren_patch_size: int = 24) -> torch.Tensor:
    """Differentiable forward: ``keep [L, B*C, H]`` (1 = keep) or ``None`` for dropout off."""
    z = siren.encoder_forward(sd, tiles)
    mods = siren.modulator_forward(sd, z, num_layers)
    b = tiles.shape[0]
    x = sd["grid"].unsqueeze(0).expand(b, -1, -1)
    for l in range(num_layers):
        pre = F.linear(x, sd[f"net.layers.{l}.weight"], sd.get(f"net.layers.{l}.bias"))
        x = siren._activation(pre, w0_initial if l == 0 else w0, activation)
        if keep is not None and p > 0:
            x = x * keep[l].view_as(x).to(x.dtype) / (1.0 - p)
        x = x * mods[l].unsqueeze(1)
    pre = F.linear(x, sd["net.last_layer.weight"], sd.get("net.last_layer.bias"))
    return torch.sin(w0 * pre).squeeze(2).reshape(b, siren_patch_size, siren_patch_size)


def train_iteration(sd: Dict[str, torch.Tensor], under: torch.Tensor, full: torch.Tensor, keep: Optional[torch.Tensor],
                    p: float, **kw) -> Tuple[torch.Tensor, float, Dict[str, torch.Tensor]]:
    """``(outputs, loss, grads)`` of one iteration with the MSE criterion (training.py:108-114,199-201)."""
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "grid"}
    live = dict(params, grid=sd["grid"])
    out = forward_train(live, under, keep, p, **kw)
    s = out.shape[-1]
    loss = F.mse_loss(out, extract_center_batch(full, under.shape[-1], s).float())
    grads = torch.autograd.grad(loss, list(params.values()))
    return out.detach(), float(loss.item()), dict(zip(params.keys(), grads))
