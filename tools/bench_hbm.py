"""HBM-roofline check of the memory-bound kernels: python tools/bench_hbm.py [slices]
Times every kernel with CUDA events on inputs larger than L2 (126 MB) and prints achieved GB/s of ALGORITHMIC bytes
 against the measured copy bandwidth in MEASURED_PEAKS.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mri_inr_b200 import ops
from mri_inr_b200.tiling import _weights_on

DEV = "cuda:0"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
H = W = 320
peak = 6539.9
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.isfile(p):
    peak = float(json.load(open(p))["hbm_gbs"])


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, bytes_):
    gbs = bytes_ / ms / 1e6
    print(f"{name:34s} {ms:8.3f} ms  {bytes_ / 1e6:9.1f} MB  {gbs:8.1f} GB/s  {gbs / peak:5.2f} of measured {peak:.0f}")


torch.manual_seed(0)
img = torch.rand(N, H, W, device=DEV)
px = N * H * W
patches = torch.empty(N * 400, 32, 32, device=DEV)
report("image_to_patches + black mask", timeit(lambda: ops.image_to_patches(img, 32, 16, with_black_mask=True, out=patches)),
       px * 4 + N * 400 * (4096 + 1))
tiles = torch.rand(N * 400, 24, 24, device=DEV)
rec = torch.empty(N, H, W, device=DEV)
wts = _weights_on(24, torch.device(DEV))
report("patches_to_image (weighted, k=24)", timeit(lambda: ops.patches_to_image(tiles, N, (20, 20), 16, weights=wts, out=rec)),
       N * 400 * 2304 + px * 4)
blk = torch.zeros(N * 400, dtype=torch.uint8, device=DEV)
blk[::17] = 1
report("  .. with a black mask (pipeline)", timeit(lambda: ops.patches_to_image(tiles, N, (20, 20), 16, weights=wts, black=blk, out=rec)),
       N * 400 * 2304 + px * 4 + N * 400)
report("patches_to_image (unit, k=32)", timeit(lambda: ops.patches_to_image(patches, N, (20, 20), 16, out=rec)),
       N * 400 * 4096 + px * 4)
report("minmax_normalize (per volume)", timeit(lambda: ops.minmax_normalize(img, groups=max(1, N // 11) if N % 11 == 0 else 1)),
       px * 12)
k = torch.randn(N, H, W, 2, device=DEV)
mask = (torch.rand(W, device=DEV) < 0.3)
report("kspace_to_image (mask+ifft2c+abs)", timeit(lambda: ops.kspace_to_image(k, mask, out=rec)), px * 28)
report("complex_abs", timeit(lambda: ops.complex_abs(k)), px * 12)
report("image_metrics (psnr/ssim/nrmse)", timeit(lambda: ops.image_metrics(img, rec)), px * 16)
