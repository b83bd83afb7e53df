"""Development diagnostic for the GPU box: runs each stage against the oracle and prints the error, then times
the synthesis kernel.  Not a test (tests/ has the asserting versions); handy in a single gpurun call because it
keeps going after a mismatch and prints numbers."""
from __future__ import annotations

import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import siren as osiren  # noqa: E402
from oracle.synth import synth_tiles  # noqa: E402

DEV = "cuda:0"


def build(sd_kw, act="sine", precision="fp16", **model_kw):
    from mri_inr_b200.modulated_siren import ModulatedSiren

    sd = osiren.synth_state_dict(**sd_kw)
    m = ModulatedSiren(dim_in=2, dim_hidden=256, dim_out=1, num_layers=model_kw.get("num_layers", 5),
                       latent_dim=model_kw.get("latent_dim", 256), w0=model_kw.get("w0", 1.0), w0_initial=30.0,
                       use_bias=True, dropout=0.1, modulate=True, encoder_type="custom", encoder_path=None,
                       outer_patch_size=32, inner_patch_size=16, siren_patch_size=24, device=torch.device("cpu"),
                       activation=act)
    m.load_state_dict(sd, strict=True)
    m.to(DEV).eval()
    m.precision = precision
    return m, sd


def section(name, fn):
    print(f"--- {name}", flush=True)
    try:
        fn()
    except Exception:
        traceback.print_exc()
    sys.stdout.flush()


def main():
    print(torch.cuda.get_device_name(0), torch.version.cuda, flush=True)
    from mri_inr_b200 import ops

    def grid():
        g = ops.make_grid(24, DEV).cpu().numpy()
        print("grid bit-exact:", np.array_equal(g, osiren.make_grid(24)))

    section("grid", grid)

    B = 7
    tiles_np = synth_tiles(1, B)
    for act in ("sine", "morlet"):
        for prec in ("fp32", "fp16", "bf16"):
            def run(act=act, prec=prec):
                m, sd = build(dict(seed=12, mod_bias_shift=0.5), act=act, precision=prec)
                want, z, mods = osiren.model_forward(sd, torch.from_numpy(tiles_np), activation=act,
                                                     return_intermediates=True)
                with torch.no_grad():
                    zz = m.encoder(torch.from_numpy(tiles_np).to(DEV))
                    mm = m.modulations(torch.from_numpy(tiles_np).to(DEV))
                    y = m(torch.from_numpy(tiles_np).to(DEV))
                torch.cuda.synchronize()
                print(f"{act} {prec}: latent err {float((zz.cpu() - z).abs().max()):.2e}  mods err "
                      f"{float((mm.cpu() - torch.stack(mods)).abs().max()):.2e}  out err "
                      f"{float((y.cpu() - want).abs().max()):.3e}  (out range {float(want.min()):.3f}..{float(want.max()):.3f})")
            section(f"forward {act} {prec}", run)

    def timing():
        m, sd = build(dict(seed=12, mod_bias_shift=0.5), precision="fp16")
        nslices = 256
        Bp = 400 * nslices
        packed = m._packed()
        mods = torch.rand(5, Bp, 256, device=DEV) * 0.5
        out = torch.empty(Bp, 576, device=DEV)
        for _ in range(2):
            ops.siren_forward(packed, mods, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.siren_forward(packed, mods, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        flops = 524288.0 * Bp * 576
        print(f"siren_tc fp16: {ms:.2f} ms for {nslices} slices -> {nslices / ms * 1e3:.0f} slices/s, "
              f"{flops / ms / 1e9:.1f} TFLOP/s")
        lat = torch.rand(Bp, 256, device=DEV)
        e0.record()
        ops.modulator_forward(packed, lat)
        e1.record()
        torch.cuda.synchronize()
        print(f"modulator: {e0.elapsed_time(e1):.2f} ms for {Bp} patches")

    section("timing", timing)


if __name__ == "__main__":
    main()
