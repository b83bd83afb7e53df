"""torchrun check of dist.PeerGather: every rank reconstructs its block of a small synthetic set straight into rank 0's
buffer (peer stores from the reassembly kernel) and rank 0 compares with a single-GPU reconstruction of the whole set
and with the NCCL gather.   torchrun --nproc-per-node 2 tools/check_peer_gather.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from tools.diag_gpu import build
from mri_inr_b200.dist import PeerGather, gather_slices, shard_range
from mri_inr_b200.pipeline import ReconstructionPipeline
from mri_inr_b200.synthetic import synthetic_slices

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
import tools.diag_gpu as dg
dg.DEV = str(dev)
m, sd = build(dict(seed=12, mod_bias_shift=0.5), precision="fp16")
n_total = 37
s, e = shard_range(n_total, rank, world)
imgs_all = synthetic_slices(n_total, 320, 320, device=dev, seed=7)
pipe = ReconstructionPipeline(m, chunk_slices=8)
peer = PeerGather(n_total, (320, 320), dev, dst=0)
peer.local_view.fill_(-7.0)
pipe.reconstruct(imgs_all[s:e], out=peer.local_view)
full = peer.finish()
local = pipe.reconstruct(imgs_all[s:e]).clone()
gathered = gather_slices(local, n_total, dst=0)
ok = torch.ones(1, device=dev)
if rank == 0:
    want = pipe.reconstruct(imgs_all)
    d1 = float((full - want).abs().max()); d2 = float((gathered - want).abs().max())
    print(f"peer vs single-GPU: {d1:.3e}   nccl gather vs single-GPU: {d2:.3e}   (bit-identical expected)")
    if d1 != 0.0 or d2 != 0.0:
        ok.zero_()
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
full = None
peer.close()
dist.destroy_process_group()
if rank == 0:
    print("PEER_GATHER_OK" if float(ok.item()) == 1.0 else "PEER_GATHER_FAIL")
sys.exit(0 if float(ok.item()) == 1.0 else 1)
