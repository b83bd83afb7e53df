"""CPU, world_size 2, gloo: the only exchange step of the path -- the final gather of reconstructed slices."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mri_inr_b200.dist import gather_slices, shard_range

    s, e = shard_range(n_total, rank, world)
    local = torch.arange(s, e, dtype=torch.float32).reshape(-1, 1, 1).expand(-1, 4, 5).contiguous()
    out = gather_slices(local, n_total, dst=0)
    if rank == 0:
        want = torch.arange(n_total, dtype=torch.float32).reshape(-1, 1, 1).expand(-1, 4, 5)
        q.put(bool(torch.equal(out, want)))
    else:
        q.put(out is None)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 2, 1])
def test_gather_slices_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(results)


def test_gather_single_process_is_identity():
    from mri_inr_b200.dist import gather_slices

    x = torch.rand(3, 4, 4)
    assert gather_slices(x, 3) is x


class _FakePeerBackend:
    """Host-memory stand-in for the CUDA IPC calls (``mri_inr_b200.dist._LibPeerBackend``): lets the collective
    discipline of ``PeerGather`` run on gloo.  ``fail_open_on``: ranks whose mapping fails."""

    def __init__(self, rank, fail_open_on=(), fail_alloc=False):
        self.rank, self.fail_open_on, self.fail_alloc = rank, set(fail_open_on), fail_alloc
        self.calls = []
        self._keep = None

    def alloc(self, nbytes):
        self.calls.append("alloc")
        if self.fail_alloc:
            return 2, (0, bytes(64))
        self._keep = torch.zeros(nbytes // 4, dtype=torch.float32)
        return 0, (self._keep.data_ptr(), bytes([7] * 64))

    def open(self, handle):
        self.calls.append("open")
        assert handle == bytes([7] * 64)
        if self.rank in self.fail_open_on:
            return 201, 0
        self._keep = torch.zeros(16, dtype=torch.float32)
        return 0, self._keep.data_ptr()

    def close(self, ptr):
        self.calls.append("close")
        return 0

    def free(self, ptr):
        self.calls.append("free")
        return 0

    def last_error(self):
        return "injected failure"

    def as_tensor(self, ptr, shape, device):
        n = 1
        for x in shape:
            n *= x
        return torch.zeros(n, dtype=torch.float32).view(*shape)


def _peer_worker(rank, world, port, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mri_inr_b200.dist import PeerGather, gather_slices, shard_range

    n_total = 5
    be = _FakePeerBackend(rank, fail_open_on=(1,) if mode == "open_fails_on_rank1" else (),
                          fail_alloc=(mode == "alloc_fails"))
    raised = False
    try:
        peer = PeerGather(n_total, (2, 2), torch.device("cpu"), dst=0, backend=be)
    except RuntimeError:
        raised, peer = True, None
    # whatever happened above, the ranks' collective sequences must still line up: the documented fallback
    # (gather_slices) and a few more collectives complete with the right values
    s, e = shard_range(n_total, rank, world)
    local = torch.arange(s, e, dtype=torch.float32).reshape(-1, 1, 1).expand(-1, 2, 2).contiguous()
    out = gather_slices(local, n_total, dst=0)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t)
    ok = float(t.item()) == 3.0
    if rank == 0:
        ok = ok and torch.equal(out, torch.arange(n_total, dtype=torch.float32).reshape(-1, 1, 1).expand(-1, 2, 2))
    if peer is not None:
        peer.finish()
        peer.close()
        peer.close()          # idempotent
    dist.barrier()
    q.put((rank, raised, ok, be.calls))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["ok", "open_fails_on_rank1", "alloc_fails"])
def test_peer_gather_collective_discipline_world2(mode):
    """a mapping failure on ONE rank must raise on EVERY rank and leave the collective sequence
    aligned, so that the fallback gather completes; the owner frees only after the non-owners have unmapped."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, raised0, ok0, calls0), (r1, raised1, ok1, calls1) = results
    assert ok0 and ok1
    assert raised0 == raised1 == (mode != "ok")
    if mode == "ok":
        assert calls0 == ["alloc", "free"] and calls1 == ["open", "close"]
    elif mode == "open_fails_on_rank1":
        assert calls0 == ["alloc", "free"] and calls1 == ["open"]
    else:
        assert calls0 == ["alloc"] and calls1 == []
