"""Drop-in ``ModulatedSiren`` for the reference's ``src/networks/modulated_siren.py``.

Same 17 constructor keyword arguments (modulated_siren.py:349-367, call site test_mod_siren.py:96-114),
same ``state_dict`` layout (31 tensors: ``grid``, ``net.layers.{i}.{weight,bias}``, ``net.last_layer.*``,
``modulator.layers.{i}.0.*``, ``encoder.encoder.encoder.{0,2,4,7}.*``) and the same
``forward(tiles [B,32,32]) -> [B,S,S]`` (modulated_siren.py:435-457), running on the hand-written sm_100a
kernels of ``libmrinr.so``.  Under ``eval()`` + ``torch.no_grad()`` (test_mod_siren.py:131-132,193-194):

* patch encoder: two fused strided convolutions (fp32 FFMA) + the 8x8 convolution and the linear layer as
  split-fp16 tcgen05 products (``mrinr_encoder_forward``),
* modulator: one split-fp16 tcgen05 launch per layer (``mrinr_modulator_forward``; fp32 FFMA in "fp32" mode),
* synthesis net over the coordinate grid of every patch: one persistent tcgen05 kernel
  (``mrinr_siren_forward``); the coordinates are the module's ``grid`` buffer, as in the reference (:448).

In ``train()`` mode or with gradients enabled (``Trainer._train_iteration``, src/train/training.py:177-207) the
forward applies dropout after every hidden activation (modulated_siren.py:124,154-156) and is differentiable with
respect to every parameter -- ``mrinr_train_forward`` / ``mrinr_train_backward`` (csrc/train.cu) behind a
``torch.autograd.Function``.

There is no CPU path and no library (cuDNN / cuBLAS) route: CPU tensors and unsupported shapes raise.
"""
from __future__ import annotations

import math
import os
import pathlib
from typing import Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import ops

__all__ = ["ModulatedSiren", "SirenNet", "Siren", "Modulator", "Encoder", "FixedEncoder", "make_grid_host"]


def make_grid_host(siren_patch_size: int) -> torch.Tensor:
    """The ``grid`` buffer ``[(h w), 2]`` exactly as ``torch.linspace``/``meshgrid(indexing="ij")`` build it
    (modulated_siren.py:427-433), computed with the single-rounding symmetric formula."""
    s = siren_patch_size
    lin = np.empty(s, dtype=np.float32)
    if s == 1:
        lin[0] = -1.0
    else:
        step = np.float32(np.float32(2.0) / np.float32(s - 1))
        for i in range(s):
            if i < s // 2:
                lin[i] = np.float32(np.float64(step) * i - 1.0)
            else:
                lin[i] = np.float32(1.0 - np.float64(step) * (s - 1 - i))
    c = np.arange(s * s)
    return torch.from_numpy(np.stack([lin[c // s], lin[c % s]], axis=1).astype(np.float32))


class _TrainFunction(torch.autograd.Function):
    """``ModulatedSiren.forward`` in training mode as one differentiable op: the forward and the backward are each a
    sequence of launches of csrc/train.cu through the C ABI.  ``params`` is ``model._train_params()`` (passed so that
    autograd tracks them); their gradients come back in the same order."""

    @staticmethod
    def forward(ctx, model, tiles, seed, keep_mask, *params):
        packed = model._packed()
        view, keep = model._weights_view()
        B = tiles.shape[0]
        ws = torch.empty(ops.train_workspace_bytes(packed, B), dtype=torch.uint8, device=tiles.device)
        out = ops.train_forward(packed, view, tiles, model.dropout if model.training else 0.0, seed, keep_mask, ws)
        ctx.model, ctx.tiles, ctx.ws, ctx.seed, ctx.keep_mask = model, tiles, ws, seed, keep_mask
        ctx.p = model.dropout if model.training else 0.0
        ctx.view, ctx.keep, ctx.packed = view, keep, packed
        ctx.needs = [p is not None and p.requires_grad for p in params]
        s = model.siren_patch_size
        return out.view(B, s, s)

    @staticmethod
    def backward(ctx, dout):
        model = ctx.model
        params = model._train_params()
        # every existing parameter gets a zero-initialised fp32 buffer (the kernels accumulate with atomics)
        grads = [None if p is None else torch.zeros(p.shape, dtype=torch.float32, device=p.device) for p in params]
        gview, gkeep = model._weights_view(tensors=grads)
        ops.train_backward(ctx.packed, ctx.view, gview, ctx.tiles,
                           dout.contiguous().view(ctx.tiles.shape[0], -1).to(torch.float32), ctx.p, ctx.seed,
                           ctx.keep_mask, ctx.ws)
        del gkeep
        ctx.ws = None
        return (None, None, None, None) + tuple(g if need else None for g, need in zip(grads, ctx.needs))


class Siren(nn.Module):
    """Parameter holder of one synthesis layer (``weight [out,in]``, ``bias [out]``); initialised with the
    reference's ranges (``Siren.init_``, modulated_siren.py:126-142).  It has no forward of its own: the
    whole network is evaluated by one fused kernel."""

    def __init__(self, dim_in, dim_out, w0=1.0, c=6.0, is_first=False, use_bias=True, activation=None, dropout=0.0):
        super().__init__()
        self.dim_in, self.dim_out, self.w0, self.is_first = dim_in, dim_out, w0, is_first
        self.activation_name = "morlet" if activation == "morlet" else "sine"
        self.dropout_p = float(dropout)
        bound = (1.0 / dim_in) if is_first else math.sqrt(c / dim_in) / w0
        self.weight = nn.Parameter(torch.empty(dim_out, dim_in).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.empty(dim_out).uniform_(-bound, bound)) if use_bias else None

    def forward(self, x):  # pragma: no cover - documented refusal
        raise RuntimeError("mri_inr_b200.Siren holds parameters only; call ModulatedSiren / SirenNet instead")


class SirenNet(nn.Module):
    """``SirenNet`` (modulated_siren.py:160-233): ``num_layers`` hidden layers + a sine output layer."""

    def __init__(self, dim_in, dim_hidden, dim_out, num_layers, w0, w0_initial, use_bias, dropout, activation):
        super().__init__()
        self.num_layers, self.dim_hidden = num_layers, dim_hidden
        self.w0, self.w0_initial = float(w0), float(w0_initial)
        self.layers = nn.ModuleList([
            Siren(dim_in if i == 0 else dim_hidden, dim_hidden, w0=w0_initial if i == 0 else w0, is_first=(i == 0),
                  use_bias=use_bias, activation=activation, dropout=dropout)
            for i in range(num_layers)
        ])
        self.last_layer = Siren(dim_hidden, dim_out, w0=w0, use_bias=use_bias)

    def forward(self, x, mods=None):  # pragma: no cover - documented refusal
        raise RuntimeError("mri_inr_b200.SirenNet is evaluated through ModulatedSiren.forward / "
                           "ModulatedSiren.synthesize(mods); per-call coordinates are not supported")


class Modulator(nn.Module):
    """``Modulator`` (modulated_siren.py:304-343).  ``forward`` runs the one-launch CUDA kernel."""

    def __init__(self, dim_in, dim_hidden, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.Sequential(nn.Linear(dim_in if i == 0 else dim_hidden + dim_in, dim_hidden), nn.ReLU())
            for i in range(num_layers)
        ])
        self._owner = None  # set by ModulatedSiren (not a submodule reference: avoids a module cycle)

    def forward(self, z):
        if self._owner is None:
            raise RuntimeError("Modulator must belong to a ModulatedSiren")
        mods = ops.modulator_forward(self._owner()._packed(), z.contiguous())
        return tuple(mods[i] for i in range(mods.shape[0]))


def _encoder_stack(latent_dim: int) -> nn.Sequential:
    # FixedAutoencoder.encoder (src/networks/encoding/siren_encoder.py:503-512): indices 0,2,4,7 carry parameters
    return nn.Sequential(
        nn.Conv2d(1, 16, 3, stride=2, padding=1), nn.LeakyReLU(0.2),
        nn.Conv2d(16, 32, 3, stride=2, padding=1), nn.LeakyReLU(0.2),
        nn.Conv2d(32, 64, 8), nn.LeakyReLU(0.2),
        nn.Flatten(), nn.Linear(64, latent_dim),
    )


class FixedEncoder(nn.Module):
    """``FixedEncoder`` (siren_encoder.py:551-577): the encoder half of the custom autoencoder, loaded from a
    checkpoint with key ``"state_dict"`` (:544-549).  ``model_path=None`` leaves it randomly initialised."""

    def __init__(self, model_path, device, latent_dim: int = 256):
        super().__init__()
        self.encoder = _encoder_stack(latent_dim)
        if model_path is not None:
            ckpt = torch.load(pathlib.Path(model_path), map_location=device)
            sd = ckpt["state_dict"]
            enc = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
            self.encoder.load_state_dict(enc, strict=True)
        self.encoder.to(device)

    def forward(self, x):
        return self.encoder(x.unsqueeze(1))


class Encoder(nn.Module):
    """``Encoder`` (modulated_siren.py:236-301).  Only ``encoder_type="custom"`` is in scope."""

    def __init__(self, latent_dim, encoder_path, device, encoder_type="custom"):
        super().__init__()
        self.latent_dim, self.encoder_type = latent_dim, encoder_type
        if encoder_type != "custom":
            raise NotImplementedError(
                f"encoder_type={encoder_type!r}: only the 'custom' encoder is part of the B200 hot path "
                "(the VGG ablation encoder is out of scope, SURVEY.md section 2)")
        self.encoder = FixedEncoder(encoder_path, device, latent_dim)
        self.fc = nn.Identity()
        self._owner = None  # set by ModulatedSiren (weak reference)

    def params(self):
        """conv1 w,b, conv2 w,b, conv3 w,b, fc w,b (siren_encoder.py:503-512: Sequential indices 0, 2, 4, 7)."""
        seq = self.encoder.encoder
        return [seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias, seq[4].weight, seq[4].bias,
                seq[7].weight, seq[7].bias]

    def forward(self, x, workspace=None):
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise RuntimeError("Encoder must belong to a ModulatedSiren")
        if x.dim() != 3 or tuple(x.shape[1:]) != (32, 32):
            # FixedAutoencoder is hard-wired to 32x32 inputs (siren_encoder.py:498-512: two stride-2 convolutions, then
            # an 8x8 kernel on the 8x8 map); other sizes fail in the reference as well (Linear(64, .) shape mismatch)
            raise RuntimeError(f"the custom encoder takes [B,32,32] patches, got {tuple(x.shape)}")
        return ops.encoder_forward(owner._packed(), x.to(torch.float32).contiguous(), workspace=workspace)


class ModulatedSiren(nn.Module):
    """Drop-in for ``src.networks.modulated_siren.ModulatedSiren``.

    Extra (non-reference) attribute: ``precision`` -- arithmetic of the hidden-layer contractions:

    * ``"fp16"`` (default): tcgen05, fp16 operands (11-bit significand), fp32 accumulation -- the fast path;
    * ``"bf16"``: the same with bf16 operands (meets 1e-3 only for random-init-scale weights);
    * ``"fp16x3"``: tcgen05 with split operands (hi + lo fp16 halves of activations and weights, three MMAs per
      product): fp32-class accuracy for weight / modulation scales where an 11-bit significand no longer meets the
      1e-3 bound;
    * ``"fp32"``: the exact CUDA-core kernel;
    * ``"auto"``: the cheapest of fp16 -> fp16x3 -> fp32 whose output agrees with the fp32 kernel to
      ``AUTO_TOLERANCE`` on a sample of the first batch (a one-off self-check per set of weights; the choice is
      kept in ``precision_selected`` until the parameters change).

    The default can be set with the environment variable ``MRINR_PRECISION``."""

    AUTO_TOLERANCE = 5e-4        # max-abs against the fp32 kernel on the sample: a factor 2 under north_star's 1e-3
                                 # (trained-like baseline-scale weights sit at 2e-4 .. 4e-4 in fp16 and keep the fast path)
    AUTO_SAMPLE = 256            # patches of the first batch used by the self-check

    def __init__(self, dim_in, dim_hidden, dim_out, num_layers, latent_dim, w0, w0_initial, use_bias, dropout,
                 modulate, encoder_type, encoder_path, outer_patch_size, inner_patch_size, siren_patch_size, device,
                 activation):
        super().__init__()
        self.dim_hidden, self.dim_out, self.num_layers, self.latent_dim = dim_hidden, dim_out, num_layers, latent_dim
        self.modulate = modulate          # stored, never read: modulation is always applied (modulated_siren.py:397,446)
        self.encoder_type = encoder_type
        self.outer_patch_size, self.inner_patch_size = outer_patch_size, inner_patch_size
        self.siren_patch_size = siren_patch_size
        self.activation = activation
        self.dropout = float(dropout)
        self.precision = os.environ.get("MRINR_PRECISION", "fp16")
        if outer_patch_size != 32 or latent_dim not in (64, 128, 256):
            raise ValueError("mri_inr_b200: the custom patch encoder (siren_encoder.py:498-512) is built for "
                             f"outer_patch_size=32 and latent_dim in (64, 128, 256); got {outer_patch_size}, {latent_dim}")

        self.net = SirenNet(dim_in, dim_hidden, dim_out, num_layers, w0, w0_initial, use_bias, dropout, activation)
        self.modulator = Modulator(latent_dim, dim_hidden, num_layers)
        self.encoder = Encoder(latent_dim, encoder_path, device, encoder_type)
        self.register_buffer("grid", make_grid_host(siren_patch_size))
        import weakref

        self.modulator._owner = weakref.ref(self)
        self.encoder._owner = weakref.ref(self)
        self._pack_key = None
        self._pack = None
        self.precision_selected = None      # the mode "auto" resolved to (else the mode in use)
        self._auto_key = None

    # ---- packed weights: derived state, rebuilt when parameters change (load_state_dict, .to(), optimizer step)
    def _pack_tensors(self) -> Sequence[Optional[torch.Tensor]]:
        ts = [self.grid]
        for layer in self.net.layers:
            ts += [layer.weight, layer.bias]
        ts += [self.net.last_layer.weight, self.net.last_layer.bias]
        for seq in self.modulator.layers:
            ts += [seq[0].weight, seq[0].bias]
        ts += self.encoder.params()
        return ts

    def _weights_key(self):
        return (self.activation,) + tuple(
            (None if t is None else (t.data_ptr(), t._version, str(t.device))) for t in self._pack_tensors())

    def _effective_precision(self) -> str:
        if self.precision != "auto":
            self.precision_selected = self.precision
            return self.precision
        if self._auto_key != self._weights_key():
            # not resolved yet for these weights: the fast mode, until synthesize() has seen data to check it on
            self.precision_selected = None
            return "fp16"
        return self.precision_selected

    def _packed(self) -> ops.PackedWeights:
        return self._packed_as(self._effective_precision())

    def _pack_kwargs(self):
        return dict(
            grid=self.grid,
            net_weights=[l.weight for l in self.net.layers],
            net_biases=[l.bias for l in self.net.layers],
            last_weight=self.net.last_layer.weight, last_bias=self.net.last_layer.bias,
            mod_weights=[s[0].weight for s in self.modulator.layers],
            mod_biases=[s[0].bias for s in self.modulator.layers],
            w0=self.net.w0, w0_initial=self.net.w0_initial, activation=self.activation,
            siren_patch_size=self.siren_patch_size, encoder_params=self.encoder.params(),
            outer_patch_size=self.outer_patch_size)

    def _build_pack(self, precision: str) -> ops.PackedWeights:
        return ops.PackedWeights(precision=precision, **self._pack_kwargs())

    def _resolve_auto(self, mods: torch.Tensor) -> None:
        """One-off self-check of ``precision="auto"``: evaluate a strided sample of the batch's modulations in every
        tensor-core mode and keep the cheapest one that agrees with the fp32 kernel to ``AUTO_TOLERANCE``."""
        B = mods.shape[1]
        n = min(B, self.AUTO_SAMPLE)
        sel = torch.linspace(0, B - 1, n, device=mods.device).round().long()
        sample = mods[:, sel].contiguous()
        exact_pack = self._build_pack("fp32")
        try:
            exact = ops.siren_forward(exact_pack, sample)
            chosen, self.auto_errors = "fp32", {}
            for mode in ("fp16", "fp16x3"):
                pack = self._build_pack(mode)
                try:
                    err = float((ops.siren_forward(pack, sample) - exact).abs().max())
                finally:
                    pack.free()
                self.auto_errors[mode] = err
                if err <= self.AUTO_TOLERANCE:
                    chosen = mode
                    break
        finally:
            exact_pack.free()
        self.precision_selected = chosen
        self._auto_key = self._weights_key()

    def _packed_as(self, precision: str) -> ops.PackedWeights:
        key = (precision,) + self._weights_key()
        if self._pack is None or key != self._pack_key:
            if not self.grid.is_cuda:
                raise RuntimeError("mri_inr_b200.ModulatedSiren runs on CUDA only: call .to('cuda') first "
                                   "(there is no CPU fallback)")
            # the same configuration with new VALUES (an optimizer step, load_state_dict): refresh the existing handle
            # in place -- no allocation, no synchronisation; anything else (precision, shapes, device): a new handle
            if not (self._pack is not None and self._pack.precision == precision
                    and self._pack.refresh(**self._pack_kwargs())):
                if self._pack is not None:
                    self._pack.free()
                self._pack = self._build_pack(precision)
            self._pack_key = key
        return self._pack

    def _check_inference(self) -> None:
        """The batched pipeline and the two halves ``modulations`` / ``synthesize`` are inference-only entry points
        (no dropout, no autograd graph); ``forward`` itself also serves training."""
        if self.training and self.dropout > 0:
            raise RuntimeError("this entry point is inference only: call .eval() (in train() mode the reference applies "
                               "dropout, modulated_siren.py:124,156; use forward(tiles) for training)")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("this entry point builds no autograd graph: call it under torch.no_grad() as "
                               "test_mod_siren.py:131-132 does (use forward(tiles) for training)")

    # ---- training (src/train/training.py:177-207)
    def _train_params(self):
        """Every parameter in the order of the C ABI's weights view (``None`` where ``use_bias=False``)."""
        ts = []
        for layer in self.net.layers:
            ts += [layer.weight, layer.bias]
        ts += [self.net.last_layer.weight, self.net.last_layer.bias]
        for seq in self.modulator.layers:
            ts += [seq[0].weight, seq[0].bias]
        ts += self.encoder.params()
        return ts

    def _weights_view(self, tensors=None):
        """``MrinrWeightsView`` over the parameters -- or over ``tensors``, a list shaped like ``_train_params()``
        (the gradient buffers; ``None`` only where the parameter itself does not exist)."""
        ts = self._train_params() if tensors is None else list(tensors)
        L = self.num_layers
        mod = ts[2 * L + 2: 4 * L + 2]
        return ops.weights_view(grid=self.grid, net_weights=ts[0:2 * L:2], net_biases=ts[1:2 * L:2],
                                last_weight=ts[2 * L], last_bias=ts[2 * L + 1], mod_weights=mod[0::2],
                                mod_biases=mod[1::2], w0=self.net.w0, w0_initial=self.net.w0_initial,
                                activation=self.activation, siren_patch_size=self.siren_patch_size,
                                encoder_params=ts[4 * L + 2:], outer_patch_size=self.outer_patch_size,
                                allow_none=tensors is not None)

    def _forward_train(self, tiles: torch.Tensor) -> torch.Tensor:
        if self.dim_hidden != 256:
            raise RuntimeError("the training path needs dim_hidden == 256")
        # dropout seed from torch's generator: reproducible under torch.manual_seed, different every call
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        mask = getattr(self, "_train_keep_mask", None)       # test hook: an explicit uint8 keep-mask [L, B*C, H]
        return _TrainFunction.apply(self, tiles, seed, mask, *self._train_params())

    # ---- the two halves of forward, exposed for the batched pipeline
    def modulations(self, tiles: torch.Tensor, out: Optional[torch.Tensor] = None,
                    workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``self.modulator(self.encoder(tiles))`` (modulated_siren.py:446) -> ``[L,B,H]``."""
        z = self.encoder(tiles, workspace=workspace)
        return ops.modulator_forward(self._packed(), z.contiguous(), out=out)

    def synthesize(self, mods: torch.Tensor, black: Optional[torch.Tensor] = None,
                   out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``self.net(coords, mods)`` over the module's grid (modulated_siren.py:448-455) -> ``[B,S,S]``."""
        s = self.siren_patch_size
        if self.precision == "auto" and self._auto_key != self._weights_key() and mods.shape[1] > 0:
            self._resolve_auto(mods)
        y = ops.siren_forward(self._packed(), mods, black=black, out=out, workspace=workspace)
        return y.view(-1, s, s)

    def forward(self, tiles: torch.Tensor) -> torch.Tensor:
        if not tiles.is_cuda:
            raise RuntimeError("tiles must be a CUDA tensor: mri_inr_b200 has no CPU path")
        if tiles.shape[0] == 0:
            return tiles.new_zeros((0, self.siren_patch_size, self.siren_patch_size))
        tiles = tiles.to(torch.float32).contiguous()
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad or (self.training and self.dropout > 0):
            return self._forward_train(tiles)
        return self.synthesize(self.modulations(tiles))
