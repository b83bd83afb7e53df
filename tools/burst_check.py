"""Is the synthesis kernel power-limited?  Times single launches of increasing length after an idle pause (the power
limiter needs a few ms to react), then a long back-to-back run.   python tools/burst_check.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools.diag_gpu import build, DEV
from mri_inr_b200 import ops

ZERO_FRAC = float(os.environ.get("MRINR_ZERO_FRAC", "0"))


def _mods(L, Bp):
    """Modulations in [0, 0.5); MRINR_ZERO_FRAC of them exactly 0 (ReLU outputs: about half at random init)."""
    m = torch.rand(L, Bp, 256, device=DEV) * 0.5
    if ZERO_FRAC > 0:
        m = m * (torch.rand(L, Bp, 256, device=DEV) >= ZERO_FRAC)
    return m


m, _ = build(dict(seed=12, mod_bias_shift=0.5), precision="fp16")
packed = m._packed()
L = 5
for nsl in (4, 8, 16, 32, 64, 128, 256, 1024):
    Bp = 400 * nsl
    mods = _mods(L, Bp)
    out = torch.empty(Bp, 576, device=DEV)
    ops.siren_forward(packed, mods, out=out)
    torch.cuda.synchronize()
    time.sleep(0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.siren_forward(packed, mods, out=out); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{nsl:5d} slices, one launch after 0.5 s idle: {ms:8.3f} ms  {(L-1)*2*256*256*Bp*576/ms/1e9:7.1f} TFLOP/s", flush=True)
Bp = 400 * 256
mods = _mods(L, Bp)
out = torch.empty(Bp, 576, device=DEV)
for reps in (1, 4, 16, 64):
    torch.cuda.synchronize(); time.sleep(1.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): ops.siren_forward(packed, mods, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"256 slices x {reps:3d} back to back: {ms:8.3f} ms each  {(L-1)*2*256*256*Bp*576/ms/1e9:7.1f} TFLOP/s", flush=True)
