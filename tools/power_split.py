"""Where does the power go?  Runs the synthesis kernel back to back for a few seconds and samples nvidia-smi
(power draw, SM clock) meanwhile.  Use with instrumented builds (make -C mri_inr_b200/csrc power):
    python tools/power_split.py                                   # the product kernel
    MRINR_LIB=build/libmrinr_NO_MMA.so python tools/power_split.py # epilogue only (no tcgen05.mma issued)
    MRINR_LIB=build/libmrinr_NO_SIN.so python tools/power_split.py # MMAs + epilogue without MUFU.SIN
    MRINR_LIB=build/libmrinr_NO_BIAS.so ...                        # without the 17th K step (bias through the MMA)
    MRINR_LIB=build/libmrinr_NO_STS.so ...                         # without the operand stores to shared memory
The instrumented builds compute wrong results; they exist for this measurement only."""
import os, subprocess, sys, threading, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools.diag_gpu import build, DEV
from mri_inr_b200 import ops

m, _ = build(dict(seed=12, mod_bias_shift=0.5), precision="fp16")
packed = m._packed()
L, nsl = 5, 256
Bp = 400 * nsl
mods = torch.rand(L, Bp, 256, device=DEV) * 0.5
out = torch.empty(Bp, 576, device=DEV)
ops.siren_forward(packed, mods, out=out)
torch.cuda.synchronize()
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=power.draw,clocks.sm", "--format=csv,noheader,nounits", "-i", "0",
                         "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
threading.Thread(target=lambda: [rows.append((time.time(), l)) for l in proc.stdout], daemon=True).start()
time.sleep(1.0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 120
t_start = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.siren_forward(packed, mods, out=out)
e1.record()
torch.cuda.synchronize()
t_end = time.time()
time.sleep(0.3)
proc.terminate()
ms = e0.elapsed_time(e1) / reps
pw, ck = [], []
for t, l in rows:
    if t_start + 0.4 * (t_end - t_start) <= t <= t_end:
        try:
            a, b = l.split(",")
            pw.append(float(a)); ck.append(float(b))
        except Exception:
            pass
idle = []
for t, l in rows:
    if t < t_start - 0.2:
        try:
            idle.append(float(l.split(",")[0]))
        except Exception:
            pass
print(f"lib={os.environ.get('MRINR_LIB', 'default')}: {ms:.3f} ms per 256 slices "
      f"({(L-1)*2*256*256*Bp*576/ms/1e9:.0f} TFLOP/s-equivalent), median power {statistics.median(pw) if pw else float('nan'):.0f} W, "
      f"median SM clock {statistics.median(ck) if ck else float('nan'):.0f} MHz, idle before {statistics.median(idle) if idle else float('nan'):.0f} W, "
      f"{len(pw)} samples")
