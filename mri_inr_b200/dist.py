"""Multi-GPU plumbing: one process per GPU, slices block-partitioned, weights replicated, and exactly one
exchange step -- the final gather of reconstructed slices to rank 0.  The reference has no
distributed code; this follows north_star's sharding statement.  Works on NCCL (GPU) and gloo (CPU tests).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition: ranks ``< n_items % world`` get one extra item.  Whole slices only, so the
    overlap reassembly never needs a halo."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_counts(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def gather_slices(local: torch.Tensor, n_total: int, dst: int = 0, out: Optional[torch.Tensor] = None,
                  group=None) -> Optional[torch.Tensor]:
    """Gather the per-rank blocks ``local [n_local, ...]`` (block partition of ``n_total`` items, see
    ``shard_range``) into ``[n_total, ...]`` on rank ``dst``.  Point-to-point sends land directly in the final
    buffer (no staging copy, no padding to equal sizes).  Returns the gathered tensor on ``dst``, ``None`` elsewhere."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        if out is not None:
            out.copy_(local)
            return out
        return local
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    s, e = shard_range(n_total, rank, world)
    if local.shape[0] != e - s:
        raise RuntimeError(f"rank {rank}: local block has {local.shape[0]} items, expected {e - s}")
    local = local.contiguous()
    if rank == dst:
        if out is None:
            out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        ops = []
        for r in range(world):
            rs, re = shard_range(n_total, r, world)
            if r == dst:
                out[rs:re].copy_(local)
            elif re > rs:
                ops.append(dist.P2POp(dist.irecv, out[rs:re], r, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out
    if e > s:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, local, dst, group)]):
            req.wait()
    return None


class _RawCudaBuffer:
    """Minimal ``__cuda_array_interface__`` carrier: lets ``torch.as_tensor`` alias a raw device pointer."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 3, "strides": None}


class PeerGather:
    """The exchange step without a collective: rank ``dst`` owns the final ``[n_total, *item_shape]`` fp32 buffer,
    exports it with CUDA IPC, and every other rank maps it over NVLink / NVSwitch peer access.  ``local_view`` is
    this rank's block of the FINAL buffer (``shard_range``): pass it as ``out=`` to
    ``ReconstructionPipeline.reconstruct`` and the reassembly kernel stores the slices where they belong -- compute
    and exchange are one kernel, chunk by chunk.  ``finish()`` (stream sync + barrier) makes the buffer readable on
    ``dst``.  Raises ``RuntimeError`` when peer mapping is unavailable (callers fall back to ``gather_slices``).

    One node, one process per GPU, every GPU visible to every process (the torchrun default)."""

    def __init__(self, n_total: int, item_shape, device: torch.device, dst: int = 0, group=None):
        from . import _lib

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerGather needs an initialised process group")
        self.lib = _lib.load()
        self.group, self.dst = group, dst
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        self.shape = (int(n_total),) + tuple(int(x) for x in item_shape)
        nbytes = 4
        for x in self.shape:
            nbytes *= x
        self._ptr = ctypes.c_void_p()
        self._owner = self.rank == dst
        handle = torch.zeros(64, dtype=torch.uint8)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            if self._owner:
                buf = (ctypes.c_uint8 * 64)()
                rc = self.lib.mrinr_peer_alloc(nbytes, ctypes.byref(self._ptr), buf)
                if rc != 0:
                    status += 1
                handle = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
            hdev = handle.to(self.device)
            dist.broadcast(hdev, src=dst, group=group)
            if not self._owner:
                raw = bytes(hdev.cpu().numpy().tobytes())
                buf = (ctypes.c_uint8 * 64).from_buffer_copy(raw)
                rc = self.lib.mrinr_peer_open(buf, ctypes.byref(self._ptr))
                if rc != 0:
                    status += 1
            dist.all_reduce(status, group=group)
            if int(status.item()) != 0:
                self.close()
                raise RuntimeError("PeerGather: CUDA IPC / peer access unavailable: "
                                   + self.lib.mrinr_last_error().decode(errors="replace"))
            self.full = torch.as_tensor(_RawCudaBuffer(self._ptr.value, self.shape, "<f4"), device=self.device)
        s, e = shard_range(n_total, self.rank, self.world)
        self.local_view = self.full[s:e]

    def finish(self) -> Optional[torch.Tensor]:
        """Every rank's stores are complete and visible on ``dst``.  Returns the full buffer on ``dst``."""
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        return self.full if self._owner else None

    def close(self) -> None:
        if getattr(self, "_ptr", None) is not None and self._ptr.value:
            self.full = self.local_view = None
            with torch.cuda.device(self.device):
                torch.cuda.synchronize(self.device)
                if self._owner:
                    if dist.is_initialized():
                        dist.barrier(group=self.group)        # nobody still maps it
                    self.lib.mrinr_peer_free(self._ptr)
                else:
                    self.lib.mrinr_peer_close(self._ptr)
                    if dist.is_initialized():
                        dist.barrier(group=self.group)
            self._ptr = ctypes.c_void_p()
