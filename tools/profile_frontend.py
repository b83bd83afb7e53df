"""Times the front-end kernels alone: python tools/profile_frontend.py [patches] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools.diag_gpu import build, DEV
from mri_inr_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 94000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
m, sd = build(dict(seed=12, mod_bias_shift=0.5), precision="fp16")
packed = m._packed()
patches = torch.rand(n, 32, 32, device=DEV)
lat = torch.empty(n, 256, device=DEV)
mods = torch.empty(5, n, 256, device=DEV)
ws = torch.empty(ops.encoder_workspace_bytes(n), dtype=torch.uint8, device=DEV)


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_enc = timeit(lambda: ops.encoder_forward(packed, patches, out=lat, workspace=ws))
t_mod = timeit(lambda: ops.modulator_forward(packed, lat, out=mods))
print(f"encoder  : {t_enc:.3f} ms for {n} patches ({n * 958464 / t_enc / 1e9:.1f} TFLOP/s fp32-equivalent)")
print(f"modulator: {t_mod:.3f} ms for {n} patches ({n * 1179648 / t_mod / 1e9:.1f} TFLOP/s fp32-equivalent)")
