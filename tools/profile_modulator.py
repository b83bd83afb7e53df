"""Times the modulator kernel alone: python tools/profile_modulator.py [patches] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools.diag_gpu import build, DEV
from mri_inr_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 94000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
m, sd = build(dict(seed=12, mod_bias_shift=0.5), precision="fp16")
packed = m._packed()
lat = torch.rand(n, 256, device=DEV)
out = torch.empty(5, n, 256, device=DEV)
for _ in range(2):
    ops.modulator_forward(packed, lat, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.modulator_forward(packed, lat, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"modulator: {ms:.3f} ms for {n} patches -> {n * 1.179648e6 / ms / 1e9:.1f} TFLOP/s fp32")
