"""Sustained rate of the Morlet synthesis kernel (and a parity check against the CPU oracle) for the library selected by
MRINR_LIB: used to pick how the envelope exp(-x^2/2) is split between the special-function unit and the FMA pipe
(make -C mri_inr_b200/csrc morlet).   python tools/morlet_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import siren as osiren
from oracle.synth import synth_tiles
from tools.diag_gpu import build, DEV
from mri_inr_b200 import ops

kw = dict(seed=14, mod_bias_shift=0.5)
m, sd = build(kw, act="morlet", precision="fp16")
tiles_np = synth_tiles(77, 64)
with torch.no_grad():
    y = m(torch.from_numpy(tiles_np).to(DEV)).cpu().numpy()
err = float(np.abs(y - osiren.model_forward(sd, torch.from_numpy(tiles_np), activation="morlet").numpy()).max())
packed = m._packed()
L, Bp = 5, 400 * 235
mods = torch.rand(L, Bp, 256, device=DEV) * 0.5
out = torch.empty(Bp, 576, device=DEV)
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
print(f"lib={os.environ.get('MRINR_LIB', 'default (mask 0xF: all FMA)')}: {ms:7.3f} ms per 235 slices  "
      f"{(L - 1) * 2 * 256 * 256 * Bp * 576 / ms / 1e9:7.1f} TFLOP/s sustained   max-abs err vs oracle {err:.3e}", flush=True)
