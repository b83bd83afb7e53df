"""Batched reconstruction pipeline: the body of ``metrics_error`` (src/util/error.py:230-248) for many
slices per launch instead of one slice per Python iteration (test_mod_siren.py:196-234).

    images [N,H,W]  --image_to_patches + black mask-->  patches [N*P,O,O], black [N*P]
                    --encoder kernels + modulator (split-fp16 tcgen05 products)-->  mods [L,N*P,Hd]
                    --fused tcgen05 synthesis kernel (black patches skipped, zero-filled)-->  [N*P,S,S]
                    --weighted overlap reassembly-->  recon [N, nV*I, nH*I]

Slices are processed in chunks so that the intermediates (patches, modulations) stay a fixed, reusable
allocation; every chunk is a handful of launches on the caller's current stream.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .modulated_siren import ModulatedSiren
from .tiling import _weights_on


class ReconstructionPipeline:
    def __init__(self, model: ModulatedSiren, chunk_slices: int = 256):
        self.model = model
        self.chunk_slices = int(chunk_slices)
        self._buf = {}

    def _buffer(self, name: str, shape, dtype, device) -> torch.Tensor:
        key = (name, str(device))
        numel = 1
        for s in shape:
            numel *= int(s)
        t = self._buf.get(key)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(numel, dtype=dtype, device=device)
            self._buf[key] = t
        return t[:numel].view(*shape)

    @torch.no_grad()
    def reconstruct(self, images: torch.Tensor, out: Optional[torch.Tensor] = None,
                    skip_black: bool = True, kernel_events: Optional[list] = None) -> torch.Tensor:
        """``images [N,H,W]`` (undersampled, fp32, CUDA) -> reconstructed ``[N, nV*I, nH*I]``.

        ``kernel_events``: if a list is given, a ``(start, end, n_patches)`` CUDA-event pair bracketing every
        synthesis-kernel launch is appended (recorded on the launching stream; used by bench.py's roofline)."""
        m = self.model
        m._check_inference()
        if images.dim() != 3 or not images.is_cuda:
            raise RuntimeError("images must be a CUDA tensor [N,H,W]")
        images = images.to(torch.float32).contiguous()
        N, H, W = images.shape
        O, I, S = m.outer_patch_size, m.inner_patch_size, m.siren_patch_size
        nV, nH = -(-H // I), -(-W // I)
        P = nV * nH
        dev = images.device
        if out is None:
            out = torch.empty(N, nV * I, nH * I, dtype=torch.float32, device=dev)
        packed = m._packed()
        wts = _weights_on(S, dev)
        cs = max(1, min(self.chunk_slices, N))
        patches_buf = self._buffer("patches", (cs * P, O, O), torch.float32, dev)
        mods_buf = self._buffer("mods", (packed.L * cs * P * packed.H,), torch.float32, dev)
        tiles_buf = self._buffer("tiles", (cs * P, S, S), torch.float32, dev)
        ws_buf = self._buffer("ws", (int(ops._lib.load().mrinr_siren_workspace_bytes(cs * P)),), torch.uint8, dev)
        enc_ws = None
        if packed.has_encoder:
            enc_ws = self._buffer("enc_ws", (max(16, ops.encoder_workspace_bytes(cs * P)),), torch.uint8, dev)
        for s0 in range(0, N, cs):
            n = min(cs, N - s0)
            B = n * P
            patches, _, black = ops.image_to_patches(images[s0:s0 + n], O, I, with_black_mask=skip_black,
                                                     out=patches_buf[:B])
            z = m.encoder(patches, workspace=enc_ws)
            mods = ops.modulator_forward(packed, z.contiguous(), out=mods_buf[: packed.L * B * packed.H].view(packed.L, B, packed.H))
            if kernel_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            tiles = ops.siren_forward(packed, mods, black=black, out=tiles_buf[:B], workspace=ws_buf)
            if kernel_events is not None:
                e1.record()
                kernel_events.append((e0, e1, B))
            ops.patches_to_image(tiles.view(B, S, S), n, (nV, nH), I, weights=wts, black=black, out=out[s0:s0 + n])
        return out
