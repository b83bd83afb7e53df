"""Error of every precision mode of the synthesis kernel against the CPU fp32 oracle, as a function of the weight scale
("error-vs-scale table").   python tools/precision_table.py > profiles/r02_precision_table.txt

Rows: hidden-weight scale x modulator-bias shift (SURVEY H2's regimes: random-init, trained-like, W x 2, W x 3), sine and
Morlet.  Columns: max-abs error of fp32 (CUDA cores), fp16x3 (split operands), fp16, bf16, and what precision="auto"
picks.  north_star's bound is 1e-3 on outputs in [-1, 1]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import siren as osiren
from oracle.synth import synth_tiles
from tools.diag_gpu import build


def main():
    tiles_np = synth_tiles(4321, 48)
    tiles = torch.from_numpy(tiles_np).cuda()
    print(f"{'activation':8s} {'W scale':>7s} {'mod shift':>9s} {'out ptp':>8s} | {'fp32':>9s} {'fp16x3':>9s} {'fp16':>9s} {'bf16':>9s} | auto")
    for act in ("sine", "morlet"):
        for scale, shift in ((1.0, 0.0), (1.0, 0.5), (1.0, 1.0), (1.5, 1.0), (2.0, 0.5), (2.0, 1.0), (3.0, 1.0)):
            kw = dict(seed=90, mod_bias_shift=shift, hidden_weight_scale=scale)
            sd = osiren.synth_state_dict(**kw)
            want = osiren.model_forward(sd, torch.from_numpy(tiles_np), activation=act).numpy()
            errs = {}
            for prec in ("fp32", "fp16x3", "fp16", "bf16", "auto"):
                m, _ = build(kw, act=act, precision=prec)
                with torch.no_grad():
                    y = m(tiles).cpu().numpy()
                errs[prec] = float(np.abs(y - want).max())
                if prec == "auto":
                    chosen = m.precision_selected
            flag = lambda e: f"{e:9.2e}" + ("" if e <= 1e-3 else "!")
            print(f"{act:8s} {scale:7.1f} {shift:9.2f} {np.ptp(want):8.3f} | {flag(errs['fp32'])} {flag(errs['fp16x3'])} "
                  f"{flag(errs['fp16'])} {flag(errs['bf16'])} | {chosen} ({errs['auto']:.2e})")
    print("('!' = above north_star's 1e-3 bound)")


if __name__ == "__main__":
    main()
