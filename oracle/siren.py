"""CPU restatement (torch fp32, functional, state_dict-driven) of the reference model path.

Test infrastructure only (see ``oracle/__init__.py``).  Every function cites the
reference lines it follows; the reference is ``/root/reference``.

The reference is module-structured (``nn.Module`` tree); this restatement is a
set of pure functions over the reference's ``state_dict`` keys, so the same
tensors can be fed to the reference (when present), to this oracle and to the
CUDA path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# coordinate grid -- src/networks/modulated_siren.py:427-433
# ----------------------------------------------------------------------------------------------
def linspace_pm1(steps: int) -> np.ndarray:
    """``torch.linspace(-1, 1, steps)`` in fp32, bit for bit (modulated_siren.py:428-429).

    torch's CPU kernel evaluates ``step = (end-start)/(steps-1)`` in fp32 and then, per index,
    the *symmetric* form ``i < steps/2 ? start + step*i : end - step*(steps-1-i)`` with one
    rounding (a fused multiply-add).  float64 arithmetic on fp32 inputs followed by one rounding
    to fp32 reproduces a single-rounding FMA exactly for these magnitudes.
    """
    if steps == 1:
        return np.array([-1.0], dtype=np.float32)
    start, end = np.float32(-1.0), np.float32(1.0)
    step = np.float32((end - start) / np.float32(steps - 1))
    out = np.empty(steps, dtype=np.float32)
    half = steps // 2
    for i in range(steps):
        if i < half:
            out[i] = np.float32(np.float64(step) * i + np.float64(start))
        else:
            out[i] = np.float32(np.float64(end) - np.float64(step) * (steps - 1 - i))
    return out


def make_grid(siren_patch_size: int) -> np.ndarray:
    """``grid`` buffer ``[(h w), 2]``: ``grid[c] = (lin[c // S], lin[c % S])``
    (meshgrid ``indexing="ij"`` + ``rearrange("h w b -> (h w) b")``, modulated_siren.py:431-433)."""
    s = siren_patch_size
    lin = linspace_pm1(s)
    g = np.empty((s * s, 2), dtype=np.float32)
    c = np.arange(s * s)
    g[:, 0] = lin[c // s]
    g[:, 1] = lin[c % s]
    return g


# ----------------------------------------------------------------------------------------------
# encoder -- src/networks/encoding/siren_encoder.py:503-512 (layers), :565-577 (forward)
# ----------------------------------------------------------------------------------------------
def encoder_forward(sd: Dict[str, torch.Tensor], tiles: torch.Tensor,
                    prefix: str = "encoder.encoder.encoder.") -> torch.Tensor:
    """``FixedEncoder``: unsqueeze(1) -> Conv(1,16,3,s2,p1) LReLU(.2) -> Conv(16,32,3,s2,p1) LReLU
    -> Conv(32,64,8) LReLU -> Flatten -> Linear(64, latent).  ``Encoder.fc`` is Identity
    (modulated_siren.py:255, :296-301)."""
    x = tiles.unsqueeze(1)
    x = F.leaky_relu(F.conv2d(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"], stride=2, padding=1), 0.2)
    x = F.leaky_relu(F.conv2d(x, sd[prefix + "2.weight"], sd[prefix + "2.bias"], stride=2, padding=1), 0.2)
    x = F.leaky_relu(F.conv2d(x, sd[prefix + "4.weight"], sd[prefix + "4.bias"]), 0.2)
    x = x.flatten(1)
    return F.linear(x, sd[prefix + "7.weight"], sd[prefix + "7.bias"])


# ----------------------------------------------------------------------------------------------
# modulator -- src/networks/modulated_siren.py:304-343
# ----------------------------------------------------------------------------------------------
def modulator_forward(sd: Dict[str, torch.Tensor], z: torch.Tensor, num_layers: int) -> List[torch.Tensor]:
    """``h0 = relu(A0 z + c0)``, ``h_i = relu(A_i cat(h_{i-1}, z) + c_i)`` -- hidden first, then z
    (modulated_siren.py:337-341).  Returns all ``num_layers`` hiddens."""
    mods = []
    x = z
    for i in range(num_layers):
        w = sd[f"modulator.layers.{i}.0.weight"]
        b = sd[f"modulator.layers.{i}.0.bias"]
        h = torch.relu(F.linear(x, w, b))
        mods.append(h)
        x = torch.cat((h, z), dim=1)
    return mods


# ----------------------------------------------------------------------------------------------
# synthesis network -- src/networks/modulated_siren.py:144-157 (layer), :215-233 (net)
# ----------------------------------------------------------------------------------------------
def _activation(pre: torch.Tensor, w0: float, kind: str) -> torch.Tensor:
    # Sine: modulated_siren.py:54 ; Morlet: modulated_siren.py:80 (Gaussian on the *pre-activation*).
    if kind == "morlet":
        return torch.sin(w0 * pre) * torch.exp(-0.5 * pre ** 2)
    return torch.sin(w0 * pre)


def siren_forward(sd: Dict[str, torch.Tensor], coords: torch.Tensor, mods: Sequence[torch.Tensor],
                  num_layers: int, w0: float, w0_initial: float, activation: str) -> torch.Tensor:
    """``coords [C,2]`` shared by every patch, ``mods[l] [B,H]`` -> ``[B, C]``.

    Hidden layer l: ``x = act_l(x W_l^T + b_l)`` then ``x *= mods[l][:, None, :]``
    (modulated_siren.py:154-156, :227-231); output layer ``sin(w0 (x w_last^T + b_last))`` --
    always Sine, never modulated (modulated_siren.py:211-213, :233).  Dropout is the identity in
    eval mode (test_mod_siren.py:131-132)."""
    b = mods[0].shape[0]
    x = coords.unsqueeze(0).expand(b, -1, -1)
    for l in range(num_layers):
        w = sd[f"net.layers.{l}.weight"]
        bias = sd.get(f"net.layers.{l}.bias")
        pre = F.linear(x, w, bias)
        x = _activation(pre, w0_initial if l == 0 else w0, activation)
        x = x * mods[l].unsqueeze(1)
    pre = F.linear(x, sd["net.last_layer.weight"], sd.get("net.last_layer.bias"))
    return torch.sin(w0 * pre).squeeze(2)


def model_forward(sd: Dict[str, torch.Tensor], tiles: torch.Tensor, *, num_layers: int = 5,
                  w0: float = 1.0, w0_initial: float = 30.0, activation: str = "sine",
                  siren_patch_size: int = 24, return_intermediates: bool = False):
    """``ModulatedSiren.forward`` (modulated_siren.py:435-457): tiles ``[B,32,32]`` -> ``[B,S,S]``.
    The coordinates are the ``grid`` entry of the state_dict (the buffer the reference reads at
    :448), not regenerated ones."""
    with torch.no_grad():
        z = encoder_forward(sd, tiles)
        mods = modulator_forward(sd, z, num_layers)
        out = siren_forward(sd, sd["grid"], mods, num_layers, w0, w0_initial, activation)
        out = out.reshape(tiles.shape[0], siren_patch_size, siren_patch_size)
    if return_intermediates:
        return out, z, mods
    return out


# ----------------------------------------------------------------------------------------------
# deterministic synthetic weights (numpy legacy RandomState: stream frozen across versions)
# ----------------------------------------------------------------------------------------------
def synth_state_dict(seed: int = 0, *, num_layers: int = 5, dim_hidden: int = 256, latent_dim: int = 256,
                     siren_patch_size: int = 24, use_bias: bool = True, w0: float = 1.0,
                     mod_bias_shift: float = 0.0, hidden_weight_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """A full 31-key ``state_dict`` with the reference's init *ranges*
    (``Siren.init_`` modulated_siren.py:126-142: first layer U(+-1/dim_in), others
    U(+-sqrt(6/dim_in)/w0); ``nn.Linear``/``nn.Conv2d`` default U(+-1/sqrt(fan_in))), drawn from
    ``numpy.random.RandomState(seed)`` so that fixtures need to store no weights.

    ``mod_bias_shift`` / ``hidden_weight_scale`` give the "trained-like" regime of SURVEY.md H2
    (modulations O(0.5) instead of O(0.02)), where operand rounding actually shows."""
    rs = np.random.RandomState(seed)

    def uni(shape, bound):
        return torch.from_numpy(rs.uniform(-bound, bound, size=shape).astype(np.float32))

    sd: Dict[str, torch.Tensor] = {}
    sd["grid"] = torch.from_numpy(make_grid(siren_patch_size))
    for l in range(num_layers):
        din = 2 if l == 0 else dim_hidden
        bound = (1.0 / din) if l == 0 else math.sqrt(6.0 / din) / w0
        scale = 1.0 if l == 0 else hidden_weight_scale
        sd[f"net.layers.{l}.weight"] = uni((dim_hidden, din), bound) * scale
        if use_bias:
            sd[f"net.layers.{l}.bias"] = uni((dim_hidden,), bound)
    bound = math.sqrt(6.0 / dim_hidden) / w0
    sd["net.last_layer.weight"] = uni((1, dim_hidden), bound)
    if use_bias:
        sd["net.last_layer.bias"] = uni((1,), bound)
    for l in range(num_layers):
        din = latent_dim if l == 0 else dim_hidden + latent_dim
        bound = 1.0 / math.sqrt(din)
        sd[f"modulator.layers.{l}.0.weight"] = uni((dim_hidden, din), bound)
        sd[f"modulator.layers.{l}.0.bias"] = uni((dim_hidden,), bound) + mod_bias_shift
    p = "encoder.encoder.encoder."
    for idx, shape in (("0", (16, 1, 3, 3)), ("2", (32, 16, 3, 3)), ("4", (64, 32, 8, 8))):
        fan_in = shape[1] * shape[2] * shape[3]
        sd[p + idx + ".weight"] = uni(shape, 1.0 / math.sqrt(fan_in))
        sd[p + idx + ".bias"] = uni((shape[0],), 1.0 / math.sqrt(fan_in))
    sd[p + "7.weight"] = uni((latent_dim, 64), 1.0 / 8.0)
    sd[p + "7.bias"] = uni((latent_dim,), 1.0 / 8.0)
    return sd


def state_dict_key_order(num_layers: int = 5, use_bias: bool = True) -> List[str]:
    """Key order of the reference ``state_dict`` (verified by instantiating the reference)."""
    keys = ["grid"]
    for l in range(num_layers):
        keys.append(f"net.layers.{l}.weight")
        if use_bias:
            keys.append(f"net.layers.{l}.bias")
    keys.append("net.last_layer.weight")
    if use_bias:
        keys.append("net.last_layer.bias")
    for l in range(num_layers):
        keys += [f"modulator.layers.{l}.0.weight", f"modulator.layers.{l}.0.bias"]
    for idx in ("0", "2", "4", "7"):
        keys += [f"encoder.encoder.encoder.{idx}.weight", f"encoder.encoder.encoder.{idx}.bias"]
    return keys
