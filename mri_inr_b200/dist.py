"""Multi-GPU plumbing: one process per GPU, slices block-partitioned, weights replicated, and exactly one
exchange step -- the final gather of reconstructed slices to rank 0.  The reference has no
distributed code; this follows north_star's sharding statement.  Works on NCCL (GPU) and gloo (CPU tests).
"""
from __future__ import annotations

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
