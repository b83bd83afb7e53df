"""Development aid: phase timeline of the synthesis kernel.

    make -C mri_inr_b200/csrc timeline && MRINR_LIB=build/libmrinr_tl.so python tools/timeline.py
"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools.diag_gpu import build, DEV
from mri_inr_b200 import ops, _lib

m, sd = build(dict(seed=12, mod_bias_shift=0.5), precision="fp16")
nsl = 16
Bp = 400 * nsl
packed = m._packed()
mods = torch.rand(5, Bp, 256, device=DEV) * 0.5
out = torch.empty(Bp, 576, device=DEV)
lib = _lib.load()
buf = (ctypes.c_longlong * (2048 * 4))()
ops.siren_forward(packed, mods, out=out); torch.cuda.synchronize()
lib.mrinr_debug_timeline(buf, 2048)
ops.siren_forward(packed, mods, out=out); torch.cuda.synchronize()
n = lib.mrinr_debug_timeline(buf, 2048)
ev = [(buf[i*4], buf[i*4+1], buf[i*4+2]) for i in range(n)]
t0 = min(e[2] for e in ev)
for who in sorted(set(e[0] for e in ev)):
    print("== cta*100+warp", who)
    prev = None
    for w, tag, t in sorted([e for e in ev if e[0] == who], key=lambda x: x[2])[:150]:
        print(f"  {tag:5d}  t={t - t0:8d}  dt={(t - prev) if prev else 0:7d}")
        prev = t
