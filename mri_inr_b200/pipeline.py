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


def host_schedule(n: int, chunk: int):
    """``[(start, count), ...]`` for ``reconstruct_from_host``: with more than two chunks of work the first and the last
    chunk are ``chunk // 4`` slices -- the upload of the first and the download of the last chunk are the only transfers
    that nothing hides (3.2 ms of a 145 ms step at 8 GPUs with uniform 216-slice chunks) -- and the slices in between
    are split evenly into chunks of at most ``chunk``."""
    if n <= 0:
        return []
    edge = chunk // 4
    if edge < 1 or n <= 2 * chunk:
        return [(s0, min(chunk, n - s0)) for s0 in range(0, n, chunk)]
    middle = n - 2 * edge
    k = -(-middle // chunk)
    base, extra = divmod(middle, k)
    sizes = [edge] + [base + (1 if i < extra else 0) for i in range(k)] + [edge]
    out, s0 = [], 0
    for c in sizes:
        out.append((s0, c))
        s0 += c
    return out


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

    def _prepare(self, H: int, W: int, n_max: int, dev):
        """Chunk-sized, reusable intermediates for images of H x W (allocated once per shape)."""
        m = self.model
        O, I, S = m.outer_patch_size, m.inner_patch_size, m.siren_patch_size
        nV, nH = -(-H // I), -(-W // I)
        P = nV * nH
        packed = m._packed()
        cs = max(1, min(self.chunk_slices, n_max))
        front = []
        for k in range(1):
            front.append(dict(
                patches=self._buffer(f"patches{k}", (cs * P, O, O), torch.float32, dev),
                black=self._buffer(f"black{k}", (cs * P,), torch.uint8, dev),
                z=self._buffer(f"z{k}", (cs * P, packed.Z), torch.float32, dev),
                mods=self._buffer(f"mods{k}", (packed.L * cs * P * packed.H,), torch.float32, dev),
                enc_ws=(self._buffer(f"enc_ws{k}", (max(16, ops.encoder_workspace_bytes(cs * P)),), torch.uint8, dev)
                        if packed.has_encoder else None)))
        bufs = dict(
            front=front,
            tiles=self._buffer("tiles", (cs * P, S, S), torch.float32, dev),
            ws=self._buffer("ws", (int(ops._lib.load().mrinr_siren_workspace_bytes(cs * P)),), torch.uint8, dev),
            wts=_weights_on(S, dev),
        )
        return packed, cs, (nV, nH, P, O, I, S), bufs

    def _front(self, images: torch.Tensor, packed, geom, fb, skip_black: bool):
        """Front end of one chunk on the current stream: images [n,H,W] -> (modulations [L,B,H], black mask [B])."""
        m = self.model
        nV, nH, P, O, I, S = geom
        B = images.shape[0] * P
        patches, _, black = ops.image_to_patches(images, O, I, with_black_mask=skip_black, out=fb["patches"][:B],
                                                 black_out=fb["black"])
        z = ops.encoder_forward(packed, patches, out=fb["z"][:B], workspace=fb["enc_ws"])
        mods = ops.modulator_forward(packed, z, out=fb["mods"][: packed.L * B * packed.H].view(packed.L, B, packed.H))
        return mods, black

    def _back(self, mods, black, n: int, out: torch.Tensor, packed, geom, bufs, kernel_events: Optional[list]) -> None:
        """Synthesis + weighted reassembly of one chunk on the current stream."""
        nV, nH, P, O, I, S = geom
        B = n * P
        if kernel_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        tiles = ops.siren_forward(packed, mods, black=black, out=bufs["tiles"][:B], workspace=bufs["ws"])
        if kernel_events is not None:
            e1.record()
            # the kernel's own count of non-black patches (first word of its workspace, written by the compaction
            # launch): what the kernel actually computed, as opposed to the B patches it was handed
            nact = bufs["ws"][:4].view(torch.int32).clone() if black is not None else None
            kernel_events.append((e0, e1, B, nact))
        ops.patches_to_image(tiles.view(B, S, S), n, (nV, nH), I, weights=bufs["wts"], black=black, out=out)

    def _chunk(self, images: torch.Tensor, out: torch.Tensor, packed, geom, bufs, skip_black: bool,
               kernel_events: Optional[list]) -> None:
        """One chunk of slices, a handful of launches on the current stream: images [n,H,W] -> out [n,nV*I,nH*I]."""
        m = self.model
        if m.precision == "auto":
            packed = m._packed()        # the handle is rebuilt when the self-check below changes the mode
        mods, black = self._front(images, packed, geom, bufs["front"][0], skip_black)
        if m.precision == "auto":
            # precision="auto": the first chunk's modulations decide the mode (one-off self-check against the fp32
            # kernel, ModulatedSiren._resolve_auto); the handle may have been rebuilt, so fetch it again
            if m._auto_key != m._weights_key():
                m._resolve_auto(mods)
            packed = m._packed()
        self._back(mods, black, images.shape[0], out, packed, geom, bufs, kernel_events)

    @torch.no_grad()
    def reconstruct(self, images: torch.Tensor, out: Optional[torch.Tensor] = None,
                    skip_black: bool = True, kernel_events: Optional[list] = None) -> torch.Tensor:
        """``images [N,H,W]`` (undersampled, fp32, CUDA) -> reconstructed ``[N, nV*I, nH*I]``.

        ``kernel_events``: if a list is given, ``(start, end, n_patches, n_active)`` is appended for every
        synthesis-kernel launch: a CUDA-event pair recorded on the launching stream around it, the patches handed to
        it and a 1-element int32 device tensor with the non-black patches it computed (``None`` without black
        skipping).  Used by bench.py's roofline."""
        m = self.model
        m._check_inference()
        if images.dim() != 3 or not images.is_cuda:
            raise RuntimeError("images must be a CUDA tensor [N,H,W]")
        images = images.to(torch.float32).contiguous()
        N, H, W = images.shape
        dev = images.device
        packed, cs, geom, bufs = self._prepare(H, W, N, dev)
        nV, nH, P, O, I, S = geom
        if out is None:
            out = torch.empty(N, nV * I, nH * I, dtype=torch.float32, device=dev)
        for s0 in range(0, N, cs):
            n = min(cs, N - s0)
            self._chunk(images[s0:s0 + n], out[s0:s0 + n], packed, geom, bufs, skip_black, kernel_events)
        return out

    @torch.no_grad()
    def reconstruct_graphed(self, images: torch.Tensor, skip_black: bool = True) -> torch.Tensor:
        """``reconstruct`` replayed from a CUDA graph, for the launch-bound case: the reference's evaluation loop hands
        over ONE slice per call (test_mod_siren.py:196-234), a dozen kernels of a few microseconds each.  The launches of
        ``images``' shape are captured once (static input / output buffers owned by the pipeline; nothing in the chunk
        synchronises or allocates) and replayed on the current stream afterwards.  New parameter VALUES are picked up
        without re-capturing (the packed handle is refreshed in place before the replay); a new handle (other precision,
        shapes, device) re-captures.  Returns the pipeline's static output buffer: valid until the next call."""
        m = self.model
        m._check_inference()
        if images.dim() != 3 or not images.is_cuda:
            raise RuntimeError("images must be a CUDA tensor [N,H,W]")
        images = images.to(torch.float32)
        if m.precision == "auto" and m._auto_key != m._weights_key():
            self.reconstruct(images, skip_black=skip_black)          # the one-off self-check synchronises: not capturable
        packed = m._packed()                                          # refreshes the handle if the values changed
        key = (tuple(images.shape), str(images.device), bool(skip_black), id(packed))
        graphs = self._buf.setdefault("graphs", {})
        entry = graphs.get(key)
        if entry is None:
            graphs.clear()                                            # one captured shape at a time (buffers are shared)
            static_in = images.clone().contiguous()
            static_out = self.reconstruct(static_in, skip_black=skip_black).clone()      # warm-up: buffers, opt-ins
            torch.cuda.current_stream(images.device).synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.reconstruct(static_in, out=static_out, skip_black=skip_black)
            entry = graphs[key] = (graph, static_in, static_out)
        graph, static_in, static_out = entry
        static_in.copy_(images)
        graph.replay()
        return static_out

    @torch.no_grad()
    def reconstruct_from_host(self, host_images: torch.Tensor, host_out: Optional[torch.Tensor] = None,
                              device_out: Optional[torch.Tensor] = None, device=None,
                              skip_black: bool = True) -> torch.Tensor:
        """The same reconstruction for slices that live in (pinned) HOST memory: ``host_images [N,H,W]`` ->
        ``host_out [N, nV*I, nH*I]`` (pinned host) or, if ``device_out`` is given instead, a device tensor (the
        multi-GPU sweep gathers on the device first).

        Chunks are double-buffered over three streams -- upload of chunk i+1, compute of chunk i and download of
        chunk i-1 run concurrently (PCIe is full duplex and the copy engines are independent of the SMs) -- so the
        transfers cost one chunk of latency instead of two full passes; the first and the last chunk are a quarter of
        the others (``host_schedule``), because only THEIR upload / download is exposed.  On return all work has been enqueued and
        the CURRENT stream waits for it: synchronise that stream (or the device) before reading ``host_out``."""
        m = self.model
        m._check_inference()
        if host_images.dim() != 3 or host_images.is_cuda or host_images.dtype != torch.float32:
            raise RuntimeError("host_images must be a CPU fp32 tensor [N,H,W] (pinned for asynchronous copies)")
        if device is None:
            device = m.grid.device
        dev = torch.device(device)
        N, H, W = host_images.shape
        packed, cs, geom, bufs = self._prepare(H, W, N, dev)
        nV, nH, P, O, I, S = geom
        if device_out is None and host_out is None:
            host_out = torch.empty(N, nV * I, nH * I, dtype=torch.float32).pin_memory()
        comp = torch.cuda.current_stream(dev)
        if "streams" not in self._buf:
            self._buf["streams"] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        s_in, s_out = self._buf["streams"]
        in_buf = [self._buffer(f"h2d{i}", (cs, H, W), torch.float32, dev) for i in range(2)]
        out_buf = [self._buffer(f"d2h{i}", (cs, nV * I, nH * I), torch.float32, dev) for i in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_comp = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        start = torch.cuda.Event()
        start.record(comp)
        s_in.wait_event(start)          # nothing of this call may overtake earlier work on the caller's stream
        s_out.wait_event(start)
        for i, (s0, n) in enumerate(host_schedule(N, cs)):
            k = i & 1
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_comp[k])                 # chunk i-2 no longer reads in_buf[k]
                in_buf[k][:n].copy_(host_images[s0:s0 + n], non_blocking=True)
                ev_in[k].record(s_in)
            comp.wait_event(ev_in[k])
            if device_out is not None:
                self._chunk(in_buf[k][:n], device_out[s0:s0 + n], packed, geom, bufs, skip_black, None)
                ev_comp[k].record(comp)
            else:
                if i >= 2:
                    comp.wait_event(ev_out[k])                  # chunk i-2 has left out_buf[k]
                self._chunk(in_buf[k][:n], out_buf[k][:n], packed, geom, bufs, skip_black, None)
                ev_comp[k].record(comp)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_comp[k])
                    host_out[s0:s0 + n].copy_(out_buf[k][:n], non_blocking=True)
                    ev_out[k].record(s_out)
        done = torch.cuda.Event()
        done.record(s_out)
        comp.wait_event(done)
        return device_out if device_out is not None else host_out

    @torch.no_grad()
    def evaluate(self, undersampled: torch.Tensor, fully_sampled: torch.Tensor, skip_black: bool = True):
        """``metrics_error`` (src/util/error.py:200-271) for N slices at once, entirely on the device:
        reconstruct ``undersampled [N,H,W]`` and score it against ``fully_sampled [N,H,W]``.

        Returns ``(recon [N,nV*I,nH*I], metrics [N,3] float64 = psnr, ssim, nrmse)``.  The reference re-assembles the
        fully sampled image from its patches with ``patches_to_image`` (:244-249); for patches cut by
        ``image_to_patches`` that is the identity on the (padded) image, so the image itself is scored -- the image
        size must be a multiple of the inner patch size (320 = 20 x 16 in every configuration of the reference)."""
        m = self.model
        I = m.inner_patch_size
        if undersampled.shape != fully_sampled.shape:
            raise RuntimeError("undersampled and fully_sampled must have the same shape")
        if undersampled.shape[-1] % I or undersampled.shape[-2] % I:
            raise RuntimeError(f"image size {tuple(undersampled.shape[-2:])} is not a multiple of the inner patch size {I}")
        recon = self.reconstruct(undersampled, skip_black=skip_black)
        return recon, ops.image_metrics(fully_sampled.to(torch.float32).contiguous(), recon)
