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
