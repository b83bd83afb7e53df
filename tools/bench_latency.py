"""Per-slice latency of the drop-in flow (test_mod_siren.py:196-234 processes one slice per call):
python tools/bench_latency.py [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools.diag_gpu import build, DEV
from mri_inr_b200 import tiling
from mri_inr_b200.pipeline import ReconstructionPipeline
from mri_inr_b200.synthetic import synthetic_slices

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
m, sd = build(dict(seed=12, mod_bias_shift=0.5), precision="fp16")
img = synthetic_slices(11, device=DEV)
pipe = ReconstructionPipeline(m, chunk_slices=1)


def timed(fn, n):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


with torch.no_grad():
    patches, info = tiling.image_to_patches(img[:1], 32, 16)
    print(f"model(tiles[400,32,32])                       : {timed(lambda: m(patches), reps):7.3f} ms / slice")
    print(f"pipeline.reconstruct(1 slice)                 : {timed(lambda: pipe.reconstruct(img[:1]), reps):7.3f} ms / slice")

    def ref_flow():      # the body of metrics_error (error.py:230-248) with the reference's function names
        p, inf = tiling.image_to_patches(img[:1], 32, 16)
        kept, black, shape = tiling.filter_and_remember_black_patches(p)
        out = m(kept)
        full = tiling.reintegrate_black_patches(out, black, shape)
        return tiling.patches_to_image_weighted_average(full, inf, 24, 16, DEV)

    print(f"reference-named flow (filter/model/reintegrate) : {timed(ref_flow, reps):7.3f} ms / slice")
    if hasattr(pipe, "reconstruct_graphed"):
        print(f"pipeline.reconstruct_graphed(1 slice)         : {timed(lambda: pipe.reconstruct_graphed(img[:1]), reps):7.3f} ms / slice")
