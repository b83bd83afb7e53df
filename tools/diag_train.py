"""Development diagnostic: every gradient of the CUDA training path against the CPU restatement (oracle/train.py),
printed per parameter (keeps going after a mismatch).   python tools/diag_train.py [B]"""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from oracle import siren as osiren, train as otrain
from oracle.synth import synth_tiles, train_keep_mask

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3


def run(act, p, layers=5, latent=256):
    from mri_inr_b200.modulated_siren import ModulatedSiren
    from mri_inr_b200.tiling import extract_center_batch

    print(f"=== activation {act}, dropout {p}, L={layers}, Z={latent}, B={B}", flush=True)
    sd = osiren.synth_state_dict(seed=12, mod_bias_shift=0.5, num_layers=layers, latent_dim=latent)
    m = ModulatedSiren(dim_in=2, dim_hidden=256, dim_out=1, num_layers=layers, latent_dim=latent, w0=1.0, w0_initial=30.0,
                       use_bias=True, dropout=p, modulate=True, encoder_type="custom", encoder_path=None,
                       outer_patch_size=32, inner_patch_size=16, siren_patch_size=24, device=torch.device("cpu"),
                       activation=act)
    m.load_state_dict(sd, strict=True)
    m.to(DEV).train()
    keep_np = train_keep_mask(5, layers, B * 576, 256, p) if p > 0 else None
    under_np, full_np = synth_tiles(50, B), synth_tiles(60, B)
    m._train_keep_mask = torch.from_numpy(keep_np).to(DEV) if keep_np is not None else None
    out = m(torch.from_numpy(under_np).to(DEV))
    loss = torch.nn.functional.mse_loss(out, extract_center_batch(torch.from_numpy(full_np).to(DEV), 32, 24))
    loss.backward()
    torch.cuda.synchronize()
    want_out, want_loss, want = otrain.train_iteration(sd, torch.from_numpy(under_np), torch.from_numpy(full_np),
                                                       torch.from_numpy(keep_np) if keep_np is not None else None, p,
                                                       num_layers=layers, activation=act)
    print(f"forward max-abs err {float((out.detach().cpu() - want_out).abs().max()):.3e}   loss {float(loss):.6f} vs {want_loss:.6f}")
    for k, prm in m.named_parameters():
        g = prm.grad.detach().cpu() if prm.grad is not None else None
        w = want[k]
        if g is None:
            print(f"  {k:44s} NO GRAD")
            continue
        err = float((g - w).abs().max()) / max(float(w.abs().max()), 1e-30)
        nrm = float(g.norm()) / max(float(w.norm()), 1e-30)
        print(f"  {k:44s} rel max err {err:9.2e}   norm ratio {nrm:8.5f}   {'ok' if err <= 1e-3 else '<<<<<<'}")


for args in (("sine", 0.0), ("sine", 0.1), ("morlet", 0.1), ("sine", 0.1, 9, 128)):
    try:
        run(*args)
    except Exception:
        traceback.print_exc()
    sys.stdout.flush()
