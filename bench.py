#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): reconstructed 320x320 slices/s of the modulated-SIREN inference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--activation sine|morlet]
                    [--num-layers L --latent-dim Z] [--mods random-init|dense] [--precision fp16|fp16x3|bf16|fp32|auto]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

BASELINE.json configs: [1] default flags; [2] --activation morlet; [3] --num-layers 9 --latent-dim 128 (the
"residual-connection ablation shape": deeper MLP, reduced latent; SURVEY D4 -- the residual code itself is not in
the reference snapshot); [4] --gpus 2|4|8 under torchrun.

Workload (config[1] of BASELINE.json): baseline modulated SIREN, random-init weights (reference init ranges),
batched inference over a synthetic 940-volume-shaped validation set = 940 volumes x 11 slices = 10 340 slices of
320x320 (acc 6 / cf 0.05 synthetic single-coil k-space).  One step = one pass over the whole set.  With N GPUs the
set is block-partitioned by slice over the ranks (strong scaling) and the reconstructed slices are gathered to
rank 0 with point-to-point NCCL transfers inside the timed region.

`value`: slices/s with the inputs resident in HBM.  `e2e`: the same pass through the public host-buffer API
(`ReconstructionPipeline.reconstruct_from_host`): pinned HOST slices in, HOST reconstructions out, with the uploads
and downloads double-buffered against the compute inside the timed region (N > 1: every rank streams its block over
its own PCIe link into one shared, page-locked host buffer).
`roofline`: the fused tcgen05 synthesis kernel, timed with CUDA events around every launch in the timed region.
`cpu_baseline` / `--impl reference`: the reference's own CPU op sequence (oracle/flow.py; the reference is pure
Python and /root/reference does not exist on the GPU box, so the port is timed: kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_VOLUMES, SLICES_PER_VOLUME, IMG = 940, 11, 320
N_SLICES = N_VOLUMES * SLICES_PER_VOLUME                 # 10 340
PATCHES_PER_SLICE, COORDS_PER_PATCH = 400, 576
FLOP_PER_LAYER_COORD = 2 * 256 * 256                     # one tensor-eligible hidden contraction
METRIC = "reconstructed 320x320 slices/sec"
# dram__bytes_read.sum + dram__bytes_write.sum of the synthesis kernel from the ncu --set full capture in
# profiles/r02_ncu_siren_sine.txt (518.75 MB + 199.19 MB for one 235-slice launch of this bench);
# algorithmic: 400 x 5 KB of modulations read + 400 x 2304 B of outputs written per slice = 2.97 MB (ratio 1.03)
DRAM_TRAFFIC_BYTES_PER_SLICE = (518.75e6 + 199.19e6) / 235
MODEL_KW = dict(dim_in=2, dim_hidden=256, dim_out=1, num_layers=5, latent_dim=256, w0=1.0, w0_initial=30.0,
                use_bias=True, dropout=0.1, modulate=True, encoder_type="custom", encoder_path=None,
                outer_patch_size=32, inner_patch_size=16, siren_patch_size=24)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            try:
                pw.append(float(r[2]))
            except Exception:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons),
                "power_w": statistics.median(pw) if pw else None}


def build_state_dict(activation: str, seed: int = 0, num_layers: int = 5, latent_dim: int = 256,
                     mods: str = "random-init"):
    """Random-init weights of the baseline architecture with the reference's init ranges
    (Siren.init_, modulated_siren.py:126-142; nn.Linear / nn.Conv2d defaults).

    ``mods="dense"``: trained-like modulations -- every modulator bias is shifted by +0.5, so the ReLU outputs are
    dense and O(0.5) instead of ~50 % exact zeros of magnitude ~0.02; zero operands cost the
    tensor core less energy, so the random-init figure flatters a power-limited kernel slightly."""
    import torch

    from mri_inr_b200.modulated_siren import ModulatedSiren

    torch.manual_seed(seed)
    kw = dict(MODEL_KW, num_layers=num_layers, latent_dim=latent_dim)
    model = ModulatedSiren(device=torch.device("cpu"), activation=activation, **kw)
    if mods == "dense":
        with torch.no_grad():
            for seq in model.modulator.layers:
                seq[0].bias.add_(0.5)
    return model, {k: v.detach().clone() for k, v in model.state_dict().items()}


def cpu_reference_rate(sd, slices_cpu, activation, n_warm=1, budget_s=20.0, max_slices=24, num_layers=5):
    """slices/s of the reference's CPU op sequence (oracle/flow.py) on a bounded sample, all host threads."""
    import torch

    from oracle import flow

    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: use every host core
    for i in range(n_warm):
        flow.reconstruct_slice(sd, slices_cpu[i % len(slices_cpu)], activation=activation, num_layers=num_layers)
    t0, n = time.perf_counter(), 0
    while n < max_slices and (n < 2 or time.perf_counter() - t0 < budget_s):
        flow.reconstruct_slice(sd, slices_cpu[n % len(slices_cpu)], activation=activation, num_layers=num_layers)
        n += 1
    dt = time.perf_counter() - t0
    return n / dt, n, torch.get_num_threads()


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (port, see module docstring)."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)      # torchrun exports OMP_NUM_THREADS=1: use every host core
    from mri_inr_b200.synthetic import column_mask  # noqa: F401  (host-only helper; no GPU needed here)
    import numpy as np

    L = args.num_layers
    _, sd = build_state_dict(args.activation, num_layers=L, latent_dim=args.latent_dim, mods=args.mods)
    # the same kind of input as the GPU arm, generated on the host (no GPU work in this arm)
    rs = np.random.RandomState(1234)
    sample = max(1, args.ref_slices)
    slices = []
    yy, xx = np.meshgrid(np.linspace(-1, 1, IMG), np.linspace(-1, 1, IMG), indexing="ij")
    mask = column_mask(IMG, 6, 0.05, 1234)
    for i in range(sample):
        field = rs.uniform(size=(IMG, IMG))
        f = np.fft.fft2(field) * np.exp(-(np.fft.fftfreq(IMG)[:, None] ** 2 + np.fft.fftfreq(IMG)[None, :] ** 2) / (2 * 0.03 ** 2))
        ph = (np.fft.ifft2(f).real - np.fft.ifft2(f).real.min()) * ((xx / 0.7) ** 2 + (yy / 0.8) ** 2 <= 1)
        k = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(ph), norm="ortho")) * mask
        im = np.abs(np.fft.fftshift(np.fft.ifft2(np.fft.ifftshift(k), norm="ortho")))
        im = (im - im.min()) / (im.max() - im.min())
        slices.append(torch.from_numpy(im.astype(np.float32)))
    from oracle import flow

    for _ in range(args.warmup):
        flow.reconstruct_slice(sd, slices[0], activation=args.activation, num_layers=L)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for s in slices:
            flow.reconstruct_slice(sd, s, activation=args.activation, num_layers=L)
    dt = time.perf_counter() - t0
    rate = args.steps * sample / dt
    threads = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "slices/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{'baseline' if (L == 5 and args.latent_dim == 256) else f'num_layers={L}, latent_dim={args.latent_dim}'} "
                               f"modulated SIREN ({args.activation}) inference, {N_SLICES} synthetic 320x320 "
                               f"slices (940 volumes x 11), {args.mods} weights; CPU arm times a bounded sample of {sample} slice(s) per step",
                   "num_layers": L, "latent_dim": args.latent_dim, "activation": args.activation, "modulations": args.mods,
                   "coords_per_s": rate * PATCHES_PER_SLICE * COORDS_PER_PATCH},
        "cpu_baseline": {"value": rate, "unit": "slices/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} slice(s)/step x {args.steps} steps, torch {torch.__version__} CPU, "
                                   f"{threads} threads of {os.cpu_count()} cpus"},
        "e2e": {"value": rate, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _shared_cudart():
    """ctypes handle of the libcudart PyTorch itself uses (already loaded in this process): needed to read and CLEAR
    the runtime's last-error slot, which torch's own cudart binding does not expose."""
    import ctypes

    try:
        with open("/proc/self/maps") as f:
            for line in f:
                if "libcudart" in line:
                    return ctypes.CDLL(line.split()[-1])
    except OSError:
        pass
    return None


class SharedHostBuffer:
    """ONE result buffer in POSIX shared memory that every rank maps, so each rank streams its block
    host -> device -> host over its own PCIe link (funnelling 4.2 GB through rank 0's link would serialise the
    downloads).  Every rank page-locks ONLY ITS OWN block of the buffer, after touching it (round 1 registered the
    whole buffer on every rank before the pages existed: N processes faulting in and pinning the same untouched tmpfs
    pages at the same time made cudaHostRegister fail with cudaErrorOperatingSystem now and then, and the stale error
    it left in the CUDA runtime killed the next launch of the rank -- the rc = 1 of SCALE_r01's N = 4 point).
    Construction and ``close()`` are collective; if any rank fails, every rank ends up with ``ok == False`` (the callers
    then fall back to gather + one download on rank 0) and the collective sequence stays aligned."""

    def __init__(self, n_total, img, rank, world, dev, s0, s1):
        import ctypes

        import torch
        import torch.distributed as dist

        self.rank, self.tensor, self._registered = rank, None, None
        self.path = f"/dev/shm/mrinr_bench_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}.bin"
        nbytes = n_total * img * img * 4
        failed = 0
        if rank == 0:
            try:
                with open(self.path, "wb") as f:
                    f.truncate(nbytes)
            except Exception as e:                  # noqa: BLE001
                failed = 1
                print(f"[bench] rank {rank}: cannot create {self.path}: {e}", file=sys.stderr, flush=True)
        flag = torch.tensor([failed], dtype=torch.int32, device=dev)
        dist.all_reduce(flag)                       # also the barrier behind which the file exists
        failed = 0
        if int(flag.item()) == 0:
            rt = _shared_cudart()
            try:
                t = torch.from_file(self.path, shared=True, size=n_total * img * img, dtype=torch.float32)
                self.tensor = t.view(n_total, img, img)
                mine = self.tensor[s0:s1]
                mine.zero_()                        # my pages exist before they are pinned; nobody else touches them
                if os.environ.get("MRINR_BENCH_FAIL_SHM") == "odd" and rank % 2 == 1:
                    raise RuntimeError("failure injected by MRINR_BENCH_FAIL_SHM=odd (tests/test_gpu_multi.py)")
                if mine.numel() > 0:
                    if rt is None:
                        raise RuntimeError("libcudart not found in /proc/self/maps")
                    rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
                    rc = rt.cudaHostRegister(ctypes.c_void_p(mine.data_ptr()), mine.numel() * 4, 0)
                    if rc != 0:
                        rt.cudaGetLastError()       # clear the runtime's last-error slot: the failure is handled HERE
                        raise RuntimeError(f"cudaHostRegister failed with CUDA error {rc}")
                    self._registered = (rt, mine.data_ptr())
            except Exception as e:                  # noqa: BLE001 - any failure selects the fallback on every rank
                failed = 1
                print(f"[bench] rank {rank}: shared host buffer unavailable ({e})", file=sys.stderr, flush=True)
        flag = torch.tensor([failed + (1 if int(flag.item()) else 0)], dtype=torch.int32, device=dev)
        dist.all_reduce(flag)
        self.ok = int(flag.item()) == 0
        if not self.ok:
            self.close(barrier=False)              # the all_reduce above already lined the ranks up

    def close(self, barrier=True):
        """Collective (one barrier) unless ``barrier=False`` (the caller has just synchronised the ranks itself)."""
        import ctypes

        import torch
        import torch.distributed as dist

        if barrier:
            dist.barrier()                          # nobody writes into the buffer any more
        if self._registered is not None:
            rt, ptr = self._registered
            torch.cuda.synchronize()
            if rt.cudaHostUnregister(ctypes.c_void_p(ptr)) != 0:
                rt.cudaGetLastError()
            self._registered = None
        self.tensor = None
        if self.rank == 0:                          # unlinking while other ranks still map the file is fine on POSIX
            try:
                os.unlink(self.path)
            except OSError:
                pass


def auto_chunk(n_local: int, target: int = 235) -> int:
    """Chunk size near ``target`` that splits this rank's block evenly (no ragged last chunk: at N = 8 the block is
    1 293 slices = 5 x 235 + 118, and the half-empty last launch cost 0.8 % of scaling efficiency in round 1)."""
    if n_local <= 0:
        return target
    n_chunks = max(1, -(-n_local // target))
    return -(-n_local // n_chunks)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from mri_inr_b200 import _lib
    from mri_inr_b200.dist import PeerGather, gather_slices, shard_range
    from mri_inr_b200.pipeline import ReconstructionPipeline
    from mri_inr_b200.synthetic import synthetic_slices

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    saved_stdout = None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to fd 1 when the first communicator
        # is created, so fd 1 points at stderr until the line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime

        # a rank that dies must turn into an error on the others within minutes, not into a silent hang
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    n_total = args.slices
    s0, s1 = shard_range(n_total, rank, world)
    n_local = s1 - s0
    L = args.num_layers
    flop_per_coord = (L - 1) * FLOP_PER_LAYER_COORD
    # every rank uses the chunk size of the largest block, so the per-launch work is the same on every rank
    chunk = args.chunk if args.chunk > 0 else auto_chunk(shard_range(n_total, 0, world)[1])

    model, sd = build_state_dict(args.activation, num_layers=L, latent_dim=args.latent_dim, mods=args.mods)
    model.to(dev).eval()
    model.precision = args.precision
    pipe = ReconstructionPipeline(model, chunk_slices=chunk)

    # synthetic undersampled slices of this rank's block (set-up, untimed); seeds depend on the global index
    images = synthetic_slices(n_local, IMG, IMG, device=dev, seed=1234 + s0)
    # The one exchange step (reconstructed slices -> rank 0).  Default: rank 0's final [n_total,320,320] buffer is
    # mapped into every process over NVLink peer access and each rank's reassembly kernel stores its slices straight
    # into it (compute and exchange in one kernel, chunk by chunk).  --exchange nccl, or a box without peer access:
    # local buffer + point-to-point NCCL gather at the end of the step.
    peer, exchange = None, "single GPU"
    if world > 1:
        exchange = "NCCL gather (point-to-point) after the last chunk"
        if args.exchange == "peer":
            try:
                peer = PeerGather(n_total, (IMG, IMG), dev, dst=0)      # collective; raises on every rank or on none
                exchange = "fused into the reassembly kernel: stores into rank 0's buffer over NVLink peer memory (CUDA IPC)"
            except RuntimeError as e:
                print(f"[bench] rank {rank}: {e}; falling back to the NCCL gather", file=sys.stderr, flush=True)
    recon = peer.local_view if peer is not None else torch.empty(n_local, IMG, IMG, dtype=torch.float32, device=dev)
    gathered = (torch.empty(n_total, IMG, IMG, dtype=torch.float32, device=dev)
                if (world > 1 and rank == 0 and peer is None) else None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(events=None):
        pipe.reconstruct(images, out=recon, kernel_events=events)
        if world > 1 and peer is None:
            gather_slices(recon, n_total, dst=0, out=gathered)

    for _ in range(args.warmup):
        step()
    barrier()
    model.precision_selected = getattr(model._packed(), "precision", args.precision)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    events = []
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(events)
    e1.record()
    barrier()
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    kern_ms = sum(a.elapsed_time(b) for a, b, _, _ in events)
    kern_patches_handed = sum(n for _, _, n, _ in events)
    # patches the synthesis kernel actually computed: its own device-side count of non-black patches per launch
    kern_patches = int(sum(int(na.item()) if na is not None else n for _, _, n, na in events))
    t = torch.tensor([ms_total, float(kern_patches_handed), float(kern_patches)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms_total = float(tmax[0].item())
    black_frac = 1.0 - float(t[2].item()) / max(float(t[1].item()), 1.0)      # over all ranks
    value = args.steps * n_total / (ms_total * 1e-3)

    # ---- end to end: pinned host -> device -> pipeline -> pinned host, every step
    # (every rank uploads its block; the reconstructed set lands in host memory)
    host_in = torch.empty(n_local, IMG, IMG, dtype=torch.float32).pin_memory()
    host_in.copy_(images)
    shared, e2e_mode = None, "pinned"
    if world > 1:
        shared = SharedHostBuffer(n_total, IMG, rank, world, dev, s0, s1)
        if shared.ok:
            e2e_mode = "shared host buffer, one PCIe link per rank"
        else:
            shared, e2e_mode = None, "gather to rank 0 on the device, then one download"
    host_out = None
    if shared is None:
        host_out = torch.empty(n_total if rank == 0 else 1, IMG, IMG, dtype=torch.float32).pin_memory()

    def e2e_step():
        # public API for host-resident slices: chunked, double-buffered upload / compute / download
        if world == 1:
            pipe.reconstruct_from_host(host_in, host_out=host_out, device=dev)
        elif shared is not None:
            pipe.reconstruct_from_host(host_in, host_out=shared.tensor[s0:s1], device=dev)
        else:
            pipe.reconstruct_from_host(host_in, device_out=recon, device=dev)
            if peer is not None:
                full = peer.finish()
                if rank == 0:
                    host_out.copy_(full, non_blocking=True)
            else:
                gather_slices(recon, n_total, dst=0, out=gathered)
                if rank == 0:
                    host_out.copy_(gathered, non_blocking=True)

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = e2e_steps * n_total / (float(t2.item()) * 1e-3)

    # one launch of the synthesis kernel after an idle pause: the kernel's rate before the power limiter reacts
    # (tools/burst_check.py); reported next to the sustained figure, never as the headline
    burst = dense_tflops = None
    if rank == 0 and not args.no_burst:
        try:
            from mri_inr_b200 import ops
            nb = 256 * PATCHES_PER_SLICE
            packed = model._packed()
            bm = torch.rand(packed.L, nb, packed.H, device=dev) * 0.5
            bo = torch.empty(nb, COORDS_PER_PATCH, device=dev)
            ops.siren_forward(packed, bm, out=bo)
            torch.cuda.synchronize()
            time.sleep(0.5)
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record()
            ops.siren_forward(packed, bm, out=bo)
            b1.record()
            torch.cuda.synchronize()
            burst = nb * COORDS_PER_PATCH * flop_per_coord / (b0.elapsed_time(b1) * 1e-3) / 1e12
            # ... and the same kernel SUSTAINED on dense (trained-like) modulations: random-init modulations are ReLU
            # outputs, ~50 % exact zeros, and zero operands cost the tensor core less energy under the power cap
            #  -- so the dense figure is printed beside the headline one (40 launches back to back, ~1 s)
            n_dense = 40
            for _ in range(4):
                ops.siren_forward(packed, bm, out=bo)
            b0.record()
            for _ in range(n_dense):
                ops.siren_forward(packed, bm, out=bo)
            b1.record()
            torch.cuda.synchronize()
            dense_tflops = n_dense * nb * COORDS_PER_PATCH * flop_per_coord / (b0.elapsed_time(b1) * 1e-3) / 1e12
            del bm, bo
        except Exception as e:  # noqa: BLE001 - an extra, never fatal
            print(f"[bench] burst measurement skipped: {e}", file=sys.stderr, flush=True)
    if rank == 0:
        peaks, peak_src = measured_peaks()
        peak = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        achieved = (kern_patches * COORDS_PER_PATCH * flop_per_coord) / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else 0.0
        prec = getattr(model, "precision_selected", args.precision)
        mma_per_product = {"fp16x3": 3}.get(prec, 1)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, n, threads = cpu_reference_rate(sd, [images[i].cpu() for i in range(min(4, n_local))], args.activation,
                                                  num_layers=L)
            cpu = {"value": rate, "unit": "slices/s", "cores": threads, "kind": "port",
                   "sample": f"{n} slices of the same workload through oracle/flow.py (reference op sequence), "
                             f"{threads} torch threads of {os.cpu_count()} cpus"}
        shape = "baseline" if (L == 5 and args.latent_dim == 256) else f"num_layers={L}, latent_dim={args.latent_dim}"
        line = {
            "metric": METRIC, "value": value, "unit": "slices/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp16": "f16", "fp16x3": "f16x3 (split operands, fp32-class)"}.get(prec, prec),
            "data": "synthetic",
            "config": {"workload": f"{shape} modulated SIREN ({args.activation}) batched inference over a synthetic "
                                   f"940-volume-shaped set: {n_total} slices 320x320 (acc 6 / cf 0.05), "
                                   f"{args.mods} weights; patches -> encoder -> modulator -> fused tcgen05 MLP -> "
                                   f"weighted reassembly",
                       "slices": n_total, "chunk_slices": chunk, "num_layers": L, "latent_dim": args.latent_dim,
                       "activation": args.activation, "modulations": args.mods,
                       "parallelism": f"slices block-partitioned x{world}", "exchange": exchange,
                       "precision": f"{prec} operands, fp32 accumulate" + (f" (requested: {args.precision})" if prec != args.precision else ""),
                       "black_patch_fraction": black_frac,
                       "black_patch_note": "1 - (patches the synthesis kernel computed, its device-side count) / (patches handed to it)",
                       "coords_per_s": value * PATCHES_PER_SLICE * COORDS_PER_PATCH * (1.0 - black_frac),
                       "parity": "model + tiling path pinned against goldens of the unmodified reference; UNPINNED: the "
                                 "k-space front end that makes the inputs (fastmri absent: checked against numpy fp64) "
                                 "and PSNR/SSIM (scikit-image absent: checked against two independent fp64 "
                                 "formulations) -- neither is inside the timed region",
                       "l2": f"inputs larger than L2 ({n_local * IMG * IMG * 4 / 1e6:.0f} MB of slices per rank per step; "
                             f"intermediates {chunk * 400 * (1024 + L * 256 + 576) * 4 / 1e6:.0f} MB per chunk)"},
            "e2e": {"value": e2e_value, "unit": "slices/s", "h2d_bytes_per_step": n_total * IMG * IMG * 4,
                    "d2h_bytes_per_step": n_total * IMG * IMG * 4, "steps": e2e_steps, "result": e2e_mode},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None,
                         "traffic": DRAM_TRAFFIC_BYTES_PER_SLICE * kern_patches / PATCHES_PER_SLICE / max(len(events), 1)
                         if (L == 5) else None,
                         "traffic_note": "bytes per launch, scaled from the ncu capture in profiles/r02_ncu_siren_sine.txt (one 235-slice launch)",
                         "kernel": "siren_tc5_kernel (fused modulated-SIREN MLP, tcgen05 cta_group::2)", "peak_source": f"{peak_src}, sustained bf16",
                         "flop_per_coord": flop_per_coord, "patches_billed": kern_patches,
                         "patches_handed": kern_patches_handed,
                         "tensor_flops_issued_per_algorithmic_flop": mma_per_product,
                         "frac_of_burst_peak": achieved / float(peaks["bf16_tflops"]),
                         "kernel_ms_per_step": kern_ms / args.steps, "kernel_launches_timed": len(events),
                         "kernel_share_of_step": kern_ms / ms_total,
                         "single_launch_after_idle_tflops": burst,
                         "dense_modulations_sustained": (None if dense_tflops is None else {
                             "achieved": dense_tflops, "frac": dense_tflops / peak,
                             "note": "the synthesis kernel alone, 40 launches of 256 slices back to back on dense "
                                     "uniform(0, 0.5) modulations (no exact zeros)"}),
                         "note": "the sustained figure is limited by the 1 kW power cap (clocks.reasons), see profiles/r02_siren.md"},
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    # ---- teardown.  The measurement is complete and printed.  ONE barrier (every rank is done with the shared host
    # buffer and the peer-mapped device buffer), then only local clean-up: no collective runs after a mapping has been
    # closed (see PeerGather._release).  A failure here is reported on stderr but does not turn a finished measurement
    # into rc != 0.
    if world > 1:
        try:
            recon = None
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            if shared is not None:
                shared.close(barrier=False)
            if peer is not None:
                peer.close(barrier=False)
        except Exception as e:  # noqa: BLE001
            print(f"[bench] rank {rank}: teardown: {type(e).__name__}: {e}", file=sys.stderr, flush=True)
        try:
            dist.destroy_process_group()
        except Exception as e:  # noqa: BLE001
            print(f"[bench] rank {rank}: destroy_process_group: {type(e).__name__}: {e}", file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--activation", default="sine", choices=["sine", "morlet"])
    ap.add_argument("--num-layers", type=int, default=5, help="hidden layers of the synthesis MLP (BASELINE config 3: 9)")
    ap.add_argument("--latent-dim", type=int, default=256, choices=[64, 128, 256], help="BASELINE config 3: 128")
    ap.add_argument("--mods", default="random-init", choices=["random-init", "dense"],
                    help="dense: trained-like modulations (modulator biases + 0.5), see build_state_dict")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp16x3", "bf16", "fp32", "auto"])
    ap.add_argument("--slices", type=int, default=N_SLICES)
    ap.add_argument("--chunk", type=int, default=0, help="slices per launch; 0 = near 235, dividing the per-rank block evenly")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--ref-slices", type=int, default=4, help="slices per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-burst", action="store_true", help="skip the single-launch measurement of the synthesis kernel")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the reconstructed slices reach rank 0 (see run_ours)")
    args = ap.parse_args()
    rank = os.environ.get("RANK", "0")
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    except BaseException as e:  # noqa: BLE001 - leave one line per failing rank in the tail of stderr, then re-raise
        if not isinstance(e, SystemExit):
            import traceback

            tb = traceback.format_exc()
            print(f"[bench] rank {rank}: FAILED: {type(e).__name__}: {e}\n{tb}", file=sys.stderr, flush=True)
        raise


if __name__ == "__main__":
    try:
        from torch.distributed.elastic.multiprocessing.errors import record

        main = record(main)          # torchrun then prints the failing rank's traceback in its summary
    except Exception:  # noqa: BLE001
        pass
    main()
