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


class _LibPeerBackend:
    """CUDA IPC through the C ABI (``mrinr_peer_*``, include/mrinr.h).  Every method returns ``(rc, value)``; a
    non-zero ``rc`` leaves the message in ``mrinr_last_error()``."""

    def __init__(self):
        from . import _lib

        self.lib = _lib.load()

    def alloc(self, nbytes: int):
        ptr, buf = ctypes.c_void_p(), (ctypes.c_uint8 * 64)()
        rc = self.lib.mrinr_peer_alloc(nbytes, ctypes.byref(ptr), buf)
        return rc, (ptr.value or 0, bytes(buf))

    def open(self, handle: bytes):
        ptr = ctypes.c_void_p()
        rc = self.lib.mrinr_peer_open((ctypes.c_uint8 * 64).from_buffer_copy(handle), ctypes.byref(ptr))
        return rc, ptr.value or 0

    def close(self, ptr: int) -> int:
        return self.lib.mrinr_peer_close(ctypes.c_void_p(ptr))

    def free(self, ptr: int) -> int:
        return self.lib.mrinr_peer_free(ctypes.c_void_p(ptr))

    def last_error(self) -> str:
        return self.lib.mrinr_last_error().decode(errors="replace")

    def as_tensor(self, ptr: int, shape, device) -> torch.Tensor:
        return torch.as_tensor(_RawCudaBuffer(ptr, shape, "<f4"), device=device)


class PeerGather:
    """The exchange step without a collective: rank ``dst`` owns the final ``[n_total, *item_shape]`` fp32 buffer,
    exports it with CUDA IPC, and every other rank maps it over NVLink / NVSwitch peer access.  ``local_view`` is
    this rank's block of the FINAL buffer (``shard_range``): pass it as ``out=`` to
    ``ReconstructionPipeline.reconstruct`` and the reassembly kernel stores the slices where they belong -- compute
    and exchange are one kernel, chunk by chunk.  ``finish()`` (stream sync + barrier) makes the buffer readable on
    ``dst``.  Raises ``RuntimeError`` ON EVERY RANK when peer mapping is unavailable on any rank (callers fall back
    to ``gather_slices``).

    Collective discipline: construction, ``finish()`` and ``close()`` are collective calls, and every rank issues
    exactly the same sequence of collectives whatever its local state is (a rank whose mapping failed still takes
    part in the release barrier; otherwise the ranks' collective streams would be off by one from then on).

    One node, one process per GPU, every GPU visible to every process (the torchrun default)."""

    def __init__(self, n_total: int, item_shape, device: torch.device, dst: int = 0, group=None, backend=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerGather needs an initialised process group")
        self.backend = backend if backend is not None else _LibPeerBackend()
        self.group, self.dst = group, dst
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        self.shape = (int(n_total),) + tuple(int(x) for x in item_shape)
        nbytes = 4
        for x in self.shape:
            nbytes *= x
        self._ptr = 0
        self._owner = self.rank == dst
        self.full = self.local_view = None
        handle = torch.zeros(64, dtype=torch.uint8)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        err = ""
        with self._device_ctx():
            if self._owner:
                rc, (ptr, raw) = self.backend.alloc(nbytes)
                if rc != 0:
                    status += 1
                    err = self.backend.last_error()
                else:
                    self._ptr = ptr
                    handle = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
            hdev = handle.to(self.device)
            dist.broadcast(hdev, src=dst, group=group)
            dist.all_reduce(status, group=group)             # did the owner get its buffer?
            if int(status.item()) == 0 and not self._owner:
                rc, ptr = self.backend.open(bytes(hdev.cpu().numpy().tobytes()))
                if rc != 0:
                    status += 1
                    err = self.backend.last_error()
                else:
                    self._ptr = ptr
            dist.all_reduce(status, group=group)             # did every rank map it?
            if int(status.item()) != 0:
                self._release()
                raise RuntimeError(f"PeerGather: CUDA IPC / peer access unavailable on {int(status.item())} rank(s)"
                                   + (f": {err}" if err else ""))
            self.full = self.backend.as_tensor(self._ptr, self.shape, self.device)
        s, e = shard_range(n_total, self.rank, self.world)
        self.local_view = self.full[s:e]

    def _device_ctx(self):
        import contextlib

        return torch.cuda.device(self.device) if self.device.type == "cuda" else contextlib.nullcontext()

    def _sync(self) -> None:
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def finish(self) -> Optional[torch.Tensor]:
        """Every rank's stores are complete and visible on ``dst``.  Returns the full buffer on ``dst``."""
        self._sync()
        dist.barrier(group=self.group)
        return self.full if self._owner else None

    def _release(self, barrier: bool = True) -> None:
        """ONE barrier that every rank takes whatever its local state (nobody stores into the buffer any more), then
        the purely local part: non-owners unmap, the owner frees.  No collective runs AFTER a mapping has been closed
        (the lazily enabled peer access of the CUDA-IPC mapping is state it shares with NCCL's P2P transport towards
        the owner's GPU: nothing is torn down underneath a collective).  The collective sequence is the same on every
        rank even after a partial failure."""
        errors = []
        self.full = self.local_view = None
        with self._device_ctx():
            self._sync()
            if barrier and dist.is_initialized():
                dist.barrier(group=self.group)
                self._sync()
            if self._ptr:
                rc = self.backend.free(self._ptr) if self._owner else self.backend.close(self._ptr)
                if rc != 0:
                    errors.append(("mrinr_peer_free: " if self._owner else "mrinr_peer_close: ") + self.backend.last_error())
                self._ptr = 0
        if errors:
            raise RuntimeError("PeerGather release failed: " + "; ".join(errors))

    def close(self, barrier: bool = True) -> None:
        """Collective (one barrier) unless ``barrier=False``: then the caller guarantees that every rank has finished
        with the buffer (e.g. it has just run its own barrier) and the call is purely local."""
        if getattr(self, "_closed", False):
            return
        self._closed = True
        self._release(barrier)
