"""A/B check of one synthesis-kernel variant: MRINR_TC_VARIANT=<n> python tools/variant_check.py [slices] [reps]

Compares the 16-bit tensor-core kernel selected by MRINR_TC_VARIANT with the library's fp32 kernel on the same
modulations (sine and Morlet, 5 and 9 layers, ragged patch counts), then times it.  Development aid, not a test."""
import os, sys
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


def main():
    nslices = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    print("variant", os.environ.get("MRINR_TC_VARIANT", "default"), flush=True)
    worst = 0.0
    for act in ("sine", "morlet"):
        for L in (5, 9):
            kw = dict(seed=12, mod_bias_shift=0.5, num_layers=L)
            m16, _ = build(kw, act=act, precision="fp16", num_layers=L)
            m32, _ = build(kw, act=act, precision="fp32", num_layers=L)
            p16, p32 = m16._packed(), m32._packed()
            for Bp in (1, 3, 130, 401, 2500):
                torch.manual_seed(Bp)
                mods = _mods(L, Bp)
                y16 = torch.full((Bp, 576), 7.0, device=DEV)
                y32 = torch.empty(Bp, 576, device=DEV)
                ops.siren_forward(p16, mods, out=y16)
                ops.siren_forward(p32, mods, out=y32)
                torch.cuda.synchronize()
                err = float((y16 - y32).abs().max())
                worst = max(worst, err)
                print(f"{act} L={L} B={Bp}: max abs err vs fp32 kernel {err:.3e}", flush=True)
    print("WORST", worst, "OK" if worst <= 1e-3 else "FAIL", flush=True)
    for act, L in (("sine", 5), ("sine", 9), ("morlet", 5)):
        m, _ = build(dict(seed=12, mod_bias_shift=0.5, num_layers=L), act=act, precision="fp16", num_layers=L)
        Bp = 400 * nslices
        packed = m._packed()
        torch.manual_seed(0)
        mods = _mods(L, Bp)
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
        print(f"TIME {act} L={L}: {ms:.3f} ms / {nslices} slices -> {nslices / ms * 1e3:.0f} slices/s, "
              f"{flops / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
