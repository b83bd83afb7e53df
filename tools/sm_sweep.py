"""Sustained rate of the synthesis kernel as a function of the number of CTA pairs it may use (the kernel is limited by
the 1 kW power cap, not by SM count: how many SMs can be given to the front end of the next chunk for free?).
    python tools/sm_sweep.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools.diag_gpu import build, DEV
from mri_inr_b200 import ops, _lib

m, _ = build(dict(seed=12, mod_bias_shift=0.5), precision="fp16")
packed = m._packed()
lib = _lib.load()
L, Bp = 5, 400 * 235
mods = torch.rand(L, Bp, 256, device=DEV) * 0.5
out = torch.empty(Bp, 576, device=DEV)
for clusters in (74, 74, 72, 70, 68, 66, 64, 60, 74):
    lib.mrinr_set_synthesis_clusters(packed.handle, clusters)
    for _ in range(8):
        ops.siren_forward(packed, mods, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 40
    for _ in range(reps):
        ops.siren_forward(packed, mods, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{clusters:3d} CTA pairs ({2 * clusters:3d} SMs): {ms:7.3f} ms per 235 slices  "
          f"{(L - 1) * 2 * 256 * 256 * Bp * 576 / ms / 1e9:7.1f} TFLOP/s sustained", flush=True)
lib.mrinr_set_synthesis_clusters(packed.handle, 0)
