"""Generate ``tests/golden/*.npz`` by executing the UNMODIFIED reference (``/root/reference``) in the build
container.  Test infrastructure only.  Run:  ``python -m oracle.make_golden``

The fixtures pin the oracle restatement (``oracle/siren.py``, ``oracle/tiling.py``) and, through it, the CUDA
path.  Weights are not stored: they are re-drawn from ``oracle.siren.synth_state_dict(seed, ...)`` (numpy legacy
``RandomState`` stream), loaded *strictly* into the reference module, and only inputs' seeds and the reference's
outputs are saved.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ref_import, siren
from .synth import (GRAD_SAMPLE, HARD_CASES, MODEL_CASES, TRAIN_BATCH, TRAIN_CASES, grad_sample_index, synth_image,
                    synth_tiles, train_keep_mask)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

def main() -> None:
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = ref_import.load_reference()
    torch.set_num_threads(1)

    # ---- grids (modulated_siren.py:427-433) and fold weights (tiling.py:67-88)
    out = {}
    for s in (8, 16, 24, 32, 48):
        lin = torch.linspace(-1, 1, steps=s)
        g = torch.stack(torch.meshgrid(lin, lin, indexing="ij"), dim=-1).reshape(-1, 2)
        out[f"grid_{s}"] = g.numpy()
    m = ref_import.build_reference_model()
    out["grid_buffer_24"] = m.grid.numpy().copy()
    for k in (16, 24, 32):
        out[f"weights_{k}"] = ref.tiling.generate_weight_matrix(k).numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "grid_weights.npz"), **out)

    # ---- model forward (modulated_siren.py:435-457) with strictly loaded synthetic state_dicts
    out = {}
    for name, sd_kw, act, model_kw in MODEL_CASES + HARD_CASES:
        sd = siren.synth_state_dict(**sd_kw)
        model = ref_import.build_reference_model(activation=act, **model_kw)
        missing = model.load_state_dict(sd, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        assert list(model.state_dict().keys()) == siren.state_dict_key_order(
            model_kw.get("num_layers", 5), model_kw.get("use_bias", True))
        model.eval()
        tiles = torch.from_numpy(synth_tiles(100 + sd_kw["seed"], 5))
        with torch.no_grad():
            z = model.encoder(tiles)
            mods = model.modulator(z)
            y = model(tiles)
        out[f"{name}_latent"] = z.numpy()
        out[f"{name}_mods"] = torch.stack(list(mods)).numpy()
        out[f"{name}_out"] = y.numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "model_forward.npz"), **out)

    # ---- training iteration (src/train/training.py:177-207): the reference module in train() mode, MSE against the
    # centre crop of the fully sampled patches (extract_center_batch, tiling.py:306-322), autograd for every parameter.
    # Dropout: nn.Dropout(p) of every hidden layer (modulated_siren.py:124) is swapped for a module that applies a
    # FIXED keep-mask with the same 1/(1-p) scaling, so that both sides see the same mask.
    class FixedDropout(torch.nn.Module):
        def __init__(self, keep, p):
            super().__init__()
            self.keep, self.p = keep, p

        def forward(self, x):
            return x * self.keep.view_as(x) / (1.0 - self.p)

    out = {}
    for name, sd_kw, act, p in TRAIN_CASES:
        sd = siren.synth_state_dict(**sd_kw)
        model = ref_import.build_reference_model(activation=act)
        model.load_state_dict(sd, strict=True)
        model.train()
        B, S, H, L = TRAIN_BATCH, 24, 256, 5
        if p > 0:
            keep = torch.from_numpy(train_keep_mask(1000 + sd_kw["seed"], L, B * S * S, H, p)).float()
            for l, layer in enumerate(model.net.layers):
                layer.dropout = FixedDropout(keep[l], p)
        else:
            for layer in model.net.layers:
                layer.dropout = torch.nn.Identity()
        under = torch.from_numpy(synth_tiles(300 + sd_kw["seed"], B))
        full = torch.from_numpy(synth_tiles(400 + sd_kw["seed"], B))
        target = ref.tiling.extract_center_batch(full, 32, 24).float()          # training.py:190-196
        outputs = model(under)                                                   # :199
        loss = torch.nn.functional.mse_loss(outputs, target)                     # criterion "MSE", :108-114, :200
        loss.backward()                                                          # :201 (no GradScaler on the CPU)
        out[f"{name}_out"] = outputs.detach().numpy()
        out[f"{name}_loss"] = np.array(loss.item(), dtype=np.float64)
        for k, prm in model.named_parameters():
            g = prm.grad.detach().numpy().reshape(-1)
            out[f"{name}_gnorm_{k}"] = np.array(np.sqrt((g.astype(np.float64) ** 2).sum()))
            out[f"{name}_gsample_{k}"] = g[grad_sample_index(g.size)].copy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "training.npz"), **out)

    # ---- tiling (src/util/tiling.py) on odd-sized and baseline-sized images
    out = {}
    for tag, (h, w) in {"a": (50, 37), "b": (64, 48), "c": (320, 320)}.items():
        img = synth_image(7 + h, h, w)
        t = torch.from_numpy(img)[None]
        patches, info = ref.tiling.image_to_patches(t, 32, 16)
        kept, black, shape = ref.tiling.filter_and_remember_black_patches(patches)
        rs = np.random.RandomState(h * w)
        small = torch.from_numpy(rs.uniform(-1, 1, size=(patches.shape[0], 24, 24)).astype(np.float32))
        small_kept = small[[i for i in range(patches.shape[0]) if i not in black]]
        reint = ref.tiling.reintegrate_black_patches(small_kept, black, shape)
        wavg = ref.tiling.patches_to_image_weighted_average(reint, info, 24, 16, torch.device("cpu"))
        plain = ref.tiling.patches_to_image(patches, info, 32, 16)
        out[f"{tag}_info"] = np.array(info[0])
        out[f"{tag}_black"] = np.array(black, dtype=np.int64)
        out[f"{tag}_wavg"] = wavg.numpy()
        out[f"{tag}_plain"] = plain.numpy()
        if tag != "c":
            out[f"{tag}_patches"] = patches.numpy()
        else:
            out[f"{tag}_patches_sum"] = patches.double().sum(dim=(1, 2)).numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "tiling.npz"), **out)

    # ---- normalize_scan (visualization.py:113-126)
    x = torch.from_numpy(np.random.RandomState(5).normal(size=(3, 40, 40)).astype(np.float32))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "normalize.npz"), out=ref.visualization.normalize_scan(x).numpy())
    print("golden fixtures written to", GOLDEN_DIR)
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)))


if __name__ == "__main__":
    main()
