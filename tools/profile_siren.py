"""Runs only the synthesis kernel (for ncu captures / quick timing): python tools/profile_siren.py [slices] [reps] [activation] [precision]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools.diag_gpu import build, DEV
from mri_inr_b200 import ops

def main():
    nslices = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    act = sys.argv[3] if len(sys.argv) > 3 else "sine"
    prec = sys.argv[4] if len(sys.argv) > 4 else "fp16"
    L = int(sys.argv[5]) if len(sys.argv) > 5 else 5
    m, sd = build(dict(seed=12, mod_bias_shift=0.5, num_layers=L), act=act, precision=prec, num_layers=L)
    Bp = 400 * nslices
    packed = m._packed()
    torch.manual_seed(0)
    mods = torch.rand(L, Bp, 256, device=DEV) * 0.5
    out = torch.empty(Bp, 576, device=DEV)
    ops.siren_forward(packed, mods, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.siren_forward(packed, mods, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = (L - 1) * 2 * 256 * 256 * Bp * 576
    print(f"siren {act} {prec} L={L}: {ms:.3f} ms / {nslices} slices -> {nslices / ms * 1e3:.0f} slices/s, {flops / ms / 1e9:.1f} TFLOP/s")

if __name__ == "__main__":
    main()
