"""Tensor-level wrappers of the C ABI (``include/mrinr.h``).  PyTorch is used only to own device memory
and streams; every operation below is one or a few launches of the hand-written sm_100a kernels in
``mri_inr_b200/csrc``.  All inputs must be CUDA tensors; nothing here falls back to torch ops.
"""
from __future__ import annotations

import ctypes
from ctypes import c_void_p
from typing import Optional, Sequence, Tuple

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def make_grid(siren_patch_size: int, device) -> torch.Tensor:
    """Coordinate grid ``[S*S, 2]`` (src/networks/modulated_siren.py:427-433), generated on the device."""
    lib = _lib.load()
    device = torch.device(device)
    out = torch.empty(siren_patch_size * siren_patch_size, 2, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(lib.mrinr_make_grid(siren_patch_size, out.data_ptr(), _lib.stream_ptr(device)), "make_grid")
    return out


def weights_view(*, grid: torch.Tensor, net_weights: Sequence[torch.Tensor],
                 net_biases: Sequence[Optional[torch.Tensor]], last_weight: torch.Tensor,
                 last_bias: Optional[torch.Tensor], mod_weights: Sequence[torch.Tensor],
                 mod_biases: Sequence[torch.Tensor], w0: float, w0_initial: float, activation: str,
                 siren_patch_size: int, encoder_params: Optional[Sequence[torch.Tensor]] = None,
                 outer_patch_size: int = 32, allow_none: bool = False):
    """A ``MrinrWeightsView`` (include/mrinr.h) over fp32 CUDA tensors: ``(view, keepalive)``.  The keepalive list owns
    the ctypes pointer arrays and any converted copies: it must outlive every use of ``view``.  The same struct carries
    gradient buffers for ``mrinr_train_backward`` (``allow_none``: missing tensors become null pointers)."""
    if activation not in _lib.ACTIVATIONS:
        # the reference treats anything but "morlet" as sine (modulated_siren.py:120-123)
        activation = "sine"
    L = len(net_weights)
    dev = grid.device
    keep = []

    def f32(t, name):
        t = _lib.require_cuda(t.detach(), name)
        if t.dtype != torch.float32 or not t.is_contiguous():
            if allow_none:
                raise RuntimeError(f"{name} must be a contiguous fp32 tensor")
            t = t.to(torch.float32).contiguous()
        if t.device != dev:
            raise RuntimeError(f"{name} is on {t.device}, expected {dev}")
        keep.append(t)
        return t

    def ptr_array(ts, name):
        arr = (c_void_p * L)()
        for i, t in enumerate(ts):
            arr[i] = None if t is None else f32(t, f"{name}[{i}]").data_ptr()
        keep.append(arr)
        return arr

    H = net_weights[0].shape[0]
    view = _lib.WeightsView()
    view.dim_in = net_weights[0].shape[1]
    view.dim_hidden = H
    view.dim_out = last_weight.shape[0]
    view.num_layers = L
    view.latent_dim = mod_weights[0].shape[1]
    view.siren_patch_size = siren_patch_size
    view.w0 = float(w0)
    view.w0_initial = float(w0_initial)
    view.activation = _lib.ACTIVATIONS[activation]
    view.d_grid = f32(grid, "grid").data_ptr()
    if tuple(grid.shape) != (siren_patch_size * siren_patch_size, 2):
        raise RuntimeError(f"grid has shape {tuple(grid.shape)}, expected {(siren_patch_size ** 2, 2)}")
    view.d_net_weight = ctypes.cast(ptr_array(net_weights, "net.layers.weight"), ctypes.POINTER(c_void_p))
    view.d_net_bias = ctypes.cast(ptr_array(net_biases, "net.layers.bias"), ctypes.POINTER(c_void_p))
    view.d_mod_weight = ctypes.cast(ptr_array(mod_weights, "modulator.layers.weight"), ctypes.POINTER(c_void_p))
    view.d_mod_bias = ctypes.cast(ptr_array(mod_biases, "modulator.layers.bias"), ctypes.POINTER(c_void_p))
    view.d_last_weight = f32(last_weight, "net.last_layer.weight").data_ptr()
    view.d_last_bias = None if last_bias is None else f32(last_bias, "net.last_layer.bias").data_ptr()
    # patch encoder (optional): conv1 w,b, conv2 w,b, conv3 w,b, fc w,b  (siren_encoder.py:503-512)
    view.outer_patch_size = int(outer_patch_size)
    if encoder_params is not None:
        names = ("conv1_weight", "conv1_bias", "conv2_weight", "conv2_bias", "conv3_weight", "conv3_bias",
                 "fc_weight", "fc_bias")
        shapes = ((16, 1, 3, 3), (16,), (32, 16, 3, 3), (32,), (64, 32, 8, 8), (64,), (view.latent_dim, 64),
                  (view.latent_dim,))
        if len(encoder_params) != 8:
            raise RuntimeError("encoder_params must hold 8 tensors (3 convolutions + 1 linear, weight and bias)")
        for n, shp, t in zip(names, shapes, encoder_params):
            if t is None and allow_none:
                continue
            if tuple(t.shape) != shp:
                raise RuntimeError(f"encoder {n} has shape {tuple(t.shape)}, expected {shp}")
            setattr(view, "d_enc_" + n, f32(t, "encoder." + n).data_ptr())
    return view, keep


class PackedWeights:
    """Owner of a ``MrinrPacked`` handle (re-tiled weights, layer-0 table)."""

    def __init__(self, *, grid: torch.Tensor, net_weights: Sequence[torch.Tensor],
                 net_biases: Sequence[Optional[torch.Tensor]], last_weight: torch.Tensor,
                 last_bias: Optional[torch.Tensor], mod_weights: Sequence[torch.Tensor],
                 mod_biases: Sequence[torch.Tensor], w0: float, w0_initial: float, activation: str,
                 precision: str, siren_patch_size: int,
                 encoder_params: Optional[Sequence[torch.Tensor]] = None, outer_patch_size: int = 32):
        lib = _lib.load()
        if activation not in _lib.ACTIVATIONS:
            activation = "sine"
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}, got {precision!r}")
        L = len(net_weights)
        dev = grid.device
        H = net_weights[0].shape[0]
        view, keep = weights_view(grid=grid, net_weights=net_weights, net_biases=net_biases, last_weight=last_weight,
                                  last_bias=last_bias, mod_weights=mod_weights, mod_biases=mod_biases, w0=w0,
                                  w0_initial=w0_initial, activation=activation, siren_patch_size=siren_patch_size,
                                  encoder_params=encoder_params, outer_patch_size=outer_patch_size)
        self.has_encoder = encoder_params is not None
        handle = c_void_p()
        with torch.cuda.device(dev):
            rc = lib.mrinr_pack_weights(ctypes.byref(view), _lib.PRECISIONS[precision], _lib.stream_ptr(dev),
                                        ctypes.byref(handle))
        _lib.check(rc, "pack_weights")
        self._handle = handle
        self.device = dev
        self.H, self.L, self.Z = H, L, view.latent_dim
        self.S = siren_patch_size
        self.C = siren_patch_size * siren_patch_size
        self.precision = precision
        self.activation = activation
        # what a refresh must leave unchanged (everything that sizes or specialises the packed buffers)
        self.signature = self._signature(grid, net_weights, mod_weights, w0, w0_initial, activation, siren_patch_size,
                                         encoder_params, net_biases, last_bias)

    @staticmethod
    def _signature(grid, net_weights, mod_weights, w0, w0_initial, activation, siren_patch_size, encoder_params,
                   net_biases, last_bias):
        return (str(grid.device), len(net_weights), tuple(net_weights[0].shape), tuple(mod_weights[0].shape),
                float(w0), float(w0_initial), activation if activation in _lib.ACTIVATIONS else "sine",
                int(siren_patch_size), encoder_params is not None, tuple(b is not None for b in net_biases),
                last_bias is not None)

    def refresh(self, *, grid, net_weights, net_biases, last_weight, last_bias, mod_weights, mod_biases, w0, w0_initial,
                activation, siren_patch_size, encoder_params=None, outer_patch_size: int = 32) -> bool:
        """Re-derive the packed copies from new parameter VALUES (``mrinr_refresh_weights``: no allocation, no
        synchronisation; on the current stream).  Returns ``False`` -- and does nothing -- when the configuration
        differs from the one this handle was packed for (the caller then builds a new handle)."""
        sig = self._signature(grid, net_weights, mod_weights, w0, w0_initial, activation, siren_patch_size,
                              encoder_params, net_biases, last_bias)
        if self._handle is None or sig != self.signature:
            return False
        view, keep = weights_view(grid=grid, net_weights=net_weights, net_biases=net_biases, last_weight=last_weight,
                                  last_bias=last_bias, mod_weights=mod_weights, mod_biases=mod_biases, w0=w0,
                                  w0_initial=w0_initial, activation=activation, siren_patch_size=siren_patch_size,
                                  encoder_params=encoder_params, outer_patch_size=outer_patch_size)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().mrinr_refresh_weights(self.handle, ctypes.byref(view), _lib.stream_ptr(self.device)),
                       "refresh_weights")
        del keep
        return True

    @property
    def handle(self) -> c_void_p:
        if self._handle is None:
            raise RuntimeError("PackedWeights already freed")
        return self._handle

    def free(self) -> None:
        if getattr(self, "_handle", None) is not None:
            _lib.load().mrinr_free_packed(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def layer0_table(self) -> torch.Tensor:
        out = torch.empty(self.C, self.H, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().mrinr_packed_layer0_table(self.handle, out.data_ptr(), _lib.stream_ptr(self.device)),
                       "packed_layer0_table")
        return out


def encoder_forward(packed: PackedWeights, patches: torch.Tensor, out: Optional[torch.Tensor] = None,
                    workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``Encoder.forward`` (modulated_siren.py:282-301 -> siren_encoder.py:565-577): patches ``[B,32,32]`` ->
    latent ``[B,Z]``.  ``workspace``: optional uint8 tensor of ``encoder_workspace_bytes(B)`` bytes (reused by the
    batched pipeline); allocated per call otherwise."""
    lib = _lib.load()
    _lib.require_cuda(patches, "patches", torch.float32)
    if not packed.has_encoder:
        raise RuntimeError("these PackedWeights were built without encoder parameters")
    if patches.dim() != 3 or tuple(patches.shape[1:]) != (32, 32) or not patches.is_contiguous():
        raise RuntimeError(f"patches must be a contiguous [B,32,32] tensor, got {tuple(patches.shape)}")
    B = patches.shape[0]
    if out is None:
        out = torch.empty(B, packed.Z, dtype=torch.float32, device=patches.device)
    else:
        _lib.require_cuda(out, "out", torch.float32)
        assert tuple(out.shape) == (B, packed.Z) and out.is_contiguous()
    need = encoder_workspace_bytes(B)
    if workspace is None:
        workspace = torch.empty(max(need, 16), dtype=torch.uint8, device=patches.device)
    elif workspace.numel() * workspace.element_size() < need:
        raise RuntimeError(f"encoder workspace too small: {workspace.numel() * workspace.element_size()} < {need}")
    with torch.cuda.device(patches.device):
        _lib.check(lib.mrinr_encoder_forward(packed.handle, patches.data_ptr(), B, out.data_ptr(), workspace.data_ptr(),
                                             workspace.numel() * workspace.element_size(),
                                             _lib.stream_ptr(patches.device)), "encoder_forward")
    return out


def encoder_workspace_bytes(B: int) -> int:
    return int(_lib.load().mrinr_encoder_workspace_bytes(B))


def modulator_forward(packed: PackedWeights, latent: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``Modulator.forward`` (modulated_siren.py:325-343): latent ``[B,Z]`` -> mods ``[L,B,H]`` in one launch."""
    lib = _lib.load()
    _lib.require_cuda(latent, "latent", torch.float32)
    if latent.dim() != 2 or latent.shape[1] != packed.Z:
        raise RuntimeError(f"latent must be [B,{packed.Z}], got {tuple(latent.shape)}")
    B = latent.shape[0]
    if out is None:
        out = torch.empty(packed.L, B, packed.H, dtype=torch.float32, device=latent.device)
    else:
        _lib.require_cuda(out, "out", torch.float32)
        assert tuple(out.shape) == (packed.L, B, packed.H)
    with torch.cuda.device(latent.device):
        _lib.check(lib.mrinr_modulator_forward(packed.handle, latent.data_ptr(), B, out.data_ptr(),
                                               _lib.stream_ptr(latent.device)), "modulator_forward")
    return out


def siren_forward(packed: PackedWeights, mods: torch.Tensor, black: Optional[torch.Tensor] = None,
                  out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``SirenNet.forward`` over the grid of every patch (modulated_siren.py:215-233, :446-455):
    mods ``[L,B,H]`` (+ optional black mask ``[B]`` uint8) -> ``[B, S*S]``."""
    lib = _lib.load()
    _lib.require_cuda(mods, "mods", torch.float32)
    if mods.dim() != 3 or mods.shape[0] != packed.L or mods.shape[2] != packed.H:
        raise RuntimeError(f"mods must be [{packed.L},B,{packed.H}], got {tuple(mods.shape)}")
    B = mods.shape[1]
    dev = mods.device
    if out is None:
        out = torch.empty(B, packed.C, dtype=torch.float32, device=dev)
    else:
        _lib.require_cuda(out, "out", torch.float32)
        assert out.numel() == B * packed.C
    ws_ptr, ws_bytes = None, 0
    if black is not None:
        _lib.require_cuda(black, "black", torch.uint8)
        assert black.numel() == B
        need = int(lib.mrinr_siren_workspace_bytes(B))
        if workspace is None or workspace.numel() * workspace.element_size() < need:
            workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        ws_ptr, ws_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    with torch.cuda.device(dev):
        _lib.check(lib.mrinr_siren_forward(packed.handle, mods.data_ptr(), _ptr(black), B, out.data_ptr(), ws_ptr,
                                           ws_bytes, _lib.stream_ptr(dev)), "siren_forward")
    return out


def image_to_patches(images: torch.Tensor, outer: int, inner: int, with_black_mask: bool = False,
                     out: Optional[torch.Tensor] = None, black_out: Optional[torch.Tensor] = None):
    """``image_to_patches`` (src/util/tiling.py:10-64) for a batch ``[N,H,W]`` of equally sized images.
    Returns ``(patches [N*nV*nH, outer, outer], (nV, nH), black_mask or None)``."""
    lib = _lib.load()
    _lib.require_cuda(images, "images", torch.float32)
    if images.dim() != 3:
        raise RuntimeError(f"images must be [N,H,W], got {tuple(images.shape)}")
    N, H, W = images.shape
    nV, nH = -(-H // inner), -(-W // inner)
    dev = images.device
    P = N * nV * nH
    if out is None:
        out = torch.empty(P, outer, outer, dtype=torch.float32, device=dev)
    black = None
    if with_black_mask:
        black = torch.empty(P, dtype=torch.uint8, device=dev) if black_out is None else black_out[:P]
    with torch.cuda.device(dev):
        _lib.check(lib.mrinr_image_to_patches(images.data_ptr(), N, H, W, outer, inner, out.data_ptr(), _ptr(black),
                                              _lib.stream_ptr(dev)), "image_to_patches")
    return out, (nV, nH), black


def classify_patches(patches: torch.Tensor) -> torch.Tensor:
    """``classify_patches`` (tiling.py:184-198) for every patch: uint8 mask, 1 = black (mean < 1e-10)."""
    lib = _lib.load()
    _lib.require_cuda(patches, "patches", torch.float32)
    n = patches.shape[0]
    elems = patches[0].numel() if n > 0 else 1
    black = torch.empty(n, dtype=torch.uint8, device=patches.device)
    with torch.cuda.device(patches.device):
        _lib.check(lib.mrinr_classify_patches(patches.data_ptr(), n, elems, black.data_ptr(),
                                              _lib.stream_ptr(patches.device)), "classify_patches")
    return black


def patches_to_image(tiles: torch.Tensor, n_images: int, grid_shape: Tuple[int, int], inner: int,
                     weights: Optional[torch.Tensor] = None, black: Optional[torch.Tensor] = None,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Overlap reassembly (tiling.py:91-181) of ``n_images`` images -> ``[N, nV*inner, nH*inner]``."""
    lib = _lib.load()
    _lib.require_cuda(tiles, "tiles", torch.float32)
    nV, nH = grid_shape
    K = tiles.shape[-1]
    if tiles.shape[0] != n_images * nV * nH or tiles.shape[-2] != K:
        raise RuntimeError(f"tiles {tuple(tiles.shape)} do not match {n_images} images of {nV}x{nH} patches")
    dev = tiles.device
    if weights is not None:
        _lib.require_cuda(weights, "weights", torch.float32)
        assert tuple(weights.shape) == (K, K)
    if black is not None:
        _lib.require_cuda(black, "black", torch.uint8)
    if out is None:
        out = torch.empty(n_images, nV * inner, nH * inner, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mrinr_patches_to_image(tiles.data_ptr(), _ptr(weights), _ptr(black), n_images, nV, nH, K, inner,
                                              out.data_ptr(), _lib.stream_ptr(dev)), "patches_to_image")
    return out


def complex_abs(data: torch.Tensor) -> torch.Tensor:
    """``fastmri.complex_abs`` as used at src/data/preprocessing.py:58: ``[...,2]`` -> ``[...]``."""
    lib = _lib.load()
    _lib.require_cuda(data, "data", torch.float32)
    if data.shape[-1] != 2:
        raise RuntimeError("last dimension must be 2 (real, imag)")
    out = torch.empty(data.shape[:-1], dtype=torch.float32, device=data.device)
    with torch.cuda.device(data.device):
        _lib.check(lib.mrinr_complex_abs(data.data_ptr(), out.numel(), out.data_ptr(), _lib.stream_ptr(data.device)),
                   "complex_abs")
    return out


def minmax_normalize(x: torch.Tensor, groups: int = 1) -> torch.Tensor:
    """``normalize_scan`` (src/util/visualization.py:113-126): per group ``(x-min)/(max-min)``; ``x`` is
    viewed as ``[groups, -1]`` (one group = one image or one volume, preprocessing.py:127-137)."""
    lib = _lib.load()
    _lib.require_cuda(x, "x", torch.float32)
    n = x.numel() // max(groups, 1)
    out = torch.empty_like(x)
    scratch = torch.empty(2 * max(groups, 1), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.mrinr_minmax_normalize(x.data_ptr(), groups, n, out.data_ptr(), scratch.data_ptr(),
                                              _lib.stream_ptr(x.device)), "minmax_normalize")
    return out


def image_metrics(original: torch.Tensor, predicted: torch.Tensor) -> torch.Tensor:
    """PSNR / SSIM / NRMSE of N image pairs on the device (src/util/error.py:23-84 as called by metrics_error,
    :256-269): ``original, predicted [N,H,W]`` (or ``[H,W]``) -> ``[N,3]`` float64 (psnr, ssim, nrmse), with the
    reference's data range ``max(both) - min(both)`` per pair."""
    lib = _lib.load()
    _lib.require_cuda(original, "original", torch.float32)
    _lib.require_cuda(predicted, "predicted", torch.float32)
    if original.shape != predicted.shape:
        raise RuntimeError(f"shape mismatch: {tuple(original.shape)} vs {tuple(predicted.shape)}")
    if original.dim() == 2:
        original, predicted = original[None], predicted[None]
    if original.dim() != 3:
        raise RuntimeError("expected [N,H,W] or [H,W]")
    original, predicted = original.contiguous(), predicted.contiguous()
    N, H, W = original.shape
    out = torch.empty(N, 3, dtype=torch.float64, device=original.device)
    if N == 0:
        return out
    need = int(lib.mrinr_image_metrics_scratch_bytes(N))
    scratch = torch.empty((need + 7) // 8, dtype=torch.float64, device=original.device)
    with torch.cuda.device(original.device):
        _lib.check(lib.mrinr_image_metrics(original.data_ptr(), predicted.data_ptr(), N, H, W, out.data_ptr(),
                                           scratch.data_ptr(), scratch.numel() * 8, _lib.stream_ptr(original.device)),
                   "image_metrics")
    return out


def _fft_args(x: torch.Tensor, name: str):
    _lib.require_cuda(x, name, torch.float32)
    if x.dim() == 3:
        x = x[None]
    if x.dim() != 4 or x.shape[-1] != 2:
        raise RuntimeError(f"{name} must be [N,H,W,2] or [H,W,2] (real, imag), got {tuple(x.shape)}")
    return x.contiguous()


def fft2c(x: torch.Tensor, inverse: bool = False) -> torch.Tensor:
    """Centred orthonormal 2-D FFT of ``[N,H,W,2]`` (``fastmri.fft2c`` / ``ifft2c``): ``fftshift(fft2(ifftshift(x),
    norm="ortho"))`` over H and W, on the hand-written Stockham kernels (sizes: products of 2, 3, 5 up to 1024)."""
    lib = _lib.load()
    squeeze = x.dim() == 3
    x = _fft_args(x, "x")
    N, H, W, _ = x.shape
    out = torch.empty_like(x)
    ws = torch.empty(max(1, int(lib.mrinr_fft2c_workspace_bytes(N, H, W)) // 4), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.mrinr_fft2c(x.data_ptr(), N, H, W, 1 if inverse else 0, out.data_ptr(), ws.data_ptr(),
                                   ws.numel() * 4, _lib.stream_ptr(x.device)), "fft2c")
    return out[0] if squeeze else out


def kspace_to_image(kspace: torch.Tensor, column_mask: Optional[torch.Tensor] = None,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``load_mri_scan`` after the file read (src/data/preprocessing.py:49-58): k-space ``[N,H,W,2]`` x column mask
    ``[W]`` (bool / uint8, None = fully sampled) -> ``fastmri.ifft2c`` -> ``fastmri.complex_abs`` -> ``[N,H,W]``."""
    lib = _lib.load()
    squeeze = kspace.dim() == 3
    k = _fft_args(kspace, "kspace")
    N, H, W, _ = k.shape
    mask = None
    if column_mask is not None:
        mask = _lib.require_cuda(column_mask, "column_mask").to(torch.uint8).contiguous()
        if mask.numel() != W:
            raise RuntimeError(f"column_mask must have {W} entries, got {mask.numel()}")
    if out is None:
        out = torch.empty(N, H, W, dtype=torch.float32, device=k.device)
    ws = torch.empty(max(1, int(lib.mrinr_fft2c_workspace_bytes(N, H, W)) // 4), dtype=torch.float32, device=k.device)
    with torch.cuda.device(k.device):
        _lib.check(lib.mrinr_kspace_to_image(k.data_ptr(), _ptr(mask), N, H, W, out.data_ptr(), ws.data_ptr(),
                                             ws.numel() * 4, _lib.stream_ptr(k.device)), "kspace_to_image")
    return out[0] if squeeze else out


def train_workspace_bytes(packed: PackedWeights, B: int) -> int:
    return int(_lib.load().mrinr_train_workspace_bytes(packed.handle, B))


def train_forward(packed: PackedWeights, view, tiles: torch.Tensor, dropout_p: float, seed: int,
                  keep_mask: Optional[torch.Tensor], workspace: torch.Tensor) -> torch.Tensor:
    """``ModulatedSiren.forward`` in ``train()`` mode (dropout after every hidden activation,
    modulated_siren.py:124,154-156): tiles ``[B,32,32]`` -> ``[B, S*S]``; the intermediates stay in ``workspace``
    (``train_workspace_bytes(packed, B)`` bytes) for :func:`train_backward`."""
    lib = _lib.load()
    _lib.require_cuda(tiles, "tiles", torch.float32)
    B = tiles.shape[0]
    out = torch.empty(B, packed.C, dtype=torch.float32, device=tiles.device)
    if keep_mask is not None:
        _lib.require_cuda(keep_mask, "keep_mask", torch.uint8)
        if keep_mask.numel() != packed.L * B * packed.C * packed.H:
            raise RuntimeError(f"keep_mask must have L*B*C*H = {packed.L * B * packed.C * packed.H} entries")
    with torch.cuda.device(tiles.device):
        _lib.check(lib.mrinr_train_forward(packed.handle, ctypes.byref(view), tiles.data_ptr(), B, float(dropout_p),
                                           int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(keep_mask), out.data_ptr(),
                                           workspace.data_ptr(), workspace.numel() * workspace.element_size(),
                                           _lib.stream_ptr(tiles.device)), "train_forward")
    return out


def train_backward(packed: PackedWeights, view, grads_view, tiles: torch.Tensor, dout: torch.Tensor, dropout_p: float,
                   seed: int, keep_mask: Optional[torch.Tensor], workspace: torch.Tensor) -> None:
    """Backward of :func:`train_forward`: ``dout [B, S*S]`` -> gradients accumulated into the zero-initialised
    buffers of ``grads_view`` (a ``weights_view`` over gradient tensors)."""
    import math

    lib = _lib.load()
    _lib.require_cuda(dout, "dout", torch.float32)
    B = tiles.shape[0]
    # exact power-of-two gradient scale (see include/mrinr.h): max|dout| * scale ~ 2^10.  One 4-byte read-back per
    # iteration (the reference's loop reads loss.item() every iteration anyway, training.py:204).
    amax = float(dout.abs().max())
    grad_scale = 1.0
    if math.isfinite(amax) and amax > 0.0:
        grad_scale = 2.0 ** max(-100, min(100, 10 - math.ceil(math.log2(amax))))
    with torch.cuda.device(tiles.device):
        _lib.check(lib.mrinr_train_backward(packed.handle, ctypes.byref(view), tiles.data_ptr(), dout.data_ptr(), B,
                                            float(dropout_p), int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(keep_mask),
                                            float(grad_scale), ctypes.byref(grads_view), workspace.data_ptr(),
                                            workspace.numel() * workspace.element_size(),
                                            _lib.stream_ptr(tiles.device)), "train_backward")
