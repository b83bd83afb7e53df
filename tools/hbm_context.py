"""Context for the HBM roofline fractions: what plain torch fill_ (write only), a reduction (read only) and copy_
(read + write) reach on the same box.   python tools/hbm_context.py"""
import torch
n = 1056 * 400 * 1024
x = torch.empty(n, device="cuda"); y = torch.empty(n, device="cuda")


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, fn, b in (("fill_ (write only)", lambda: x.fill_(1.0), n * 4), ("amax (read only)", lambda: x.amax(), n * 4),
                    ("copy_ (read + write)", lambda: y.copy_(x), n * 8)):
    ms = t(fn)
    print(f"{name:24s} {ms:7.3f} ms  {b / ms / 1e6:8.1f} GB/s")
