"""GPU: training-mode forward and backward (csrc/train.cu through the C ABI and torch.autograd) against the goldens of
the unmodified reference's autograd (tests/golden/training.npz) and against the CPU restatement (oracle/train.py).
Tolerance: 1e-3 relative (of the largest element of each gradient / its norm) -- VERDICT r1 #5."""
import numpy as np
import pytest
import torch

from oracle import siren as osiren
from oracle import train as otrain
from oracle.synth import TRAIN_BATCH, TRAIN_CASES, grad_sample_index, synth_tiles, train_keep_mask

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

PARAM_KEYS = None


def _model(sd_kw, act, p, num_layers=5, latent_dim=256):
    from mri_inr_b200.modulated_siren import ModulatedSiren

    sd = osiren.synth_state_dict(**sd_kw)
    m = ModulatedSiren(dim_in=2, dim_hidden=256, dim_out=1, num_layers=num_layers, latent_dim=latent_dim, w0=1.0,
                       w0_initial=30.0, use_bias=True, dropout=p, modulate=True, encoder_type="custom",
                       encoder_path=None, outer_patch_size=32, inner_patch_size=16, siren_patch_size=24,
                       device=torch.device("cpu"), activation=act)
    m.load_state_dict(sd, strict=True)
    m.to(DEV).train()
    return m, sd


def _iteration(m, under, full, keep):
    """Trainer._train_iteration (training.py:177-207) without the optimizer step."""
    m._train_keep_mask = keep
    m.zero_grad(set_to_none=True)
    out = m(under)
    from mri_inr_b200.tiling import extract_center_batch

    target = extract_center_batch(full, 32, 24).float()                   # training.py:190-196
    loss = torch.nn.functional.mse_loss(out, target)
    loss.backward()
    return out.detach(), float(loss.item()), {k: p.grad.detach().cpu() for k, p in m.named_parameters()}


@pytest.mark.parametrize("case", TRAIN_CASES, ids=[c[0] for c in TRAIN_CASES])
def test_training_iteration_matches_reference_autograd(golden, case):
    name, sd_kw, act, p = case
    g = golden["training"]
    m, sd = _model(sd_kw, act, p)
    B = TRAIN_BATCH
    keep = (torch.from_numpy(train_keep_mask(1000 + sd_kw["seed"], 5, B * 576, 256, p)).to(DEV) if p > 0 else None)
    under = torch.from_numpy(synth_tiles(300 + sd_kw["seed"], B)).to(DEV)
    full = torch.from_numpy(synth_tiles(400 + sd_kw["seed"], B)).to(DEV)
    out, loss, grads = _iteration(m, under, full, keep)
    assert np.abs(out.cpu().numpy() - g[f"{name}_out"]).max() <= 5e-5
    assert abs(loss - float(g[f"{name}_loss"])) <= 1e-5
    assert len(grads) == 30
    worst = 0.0
    for k, gr in grads.items():
        flat = gr.numpy().reshape(-1)
        want = g[f"{name}_gsample_{k}"]
        norm = float(g[f"{name}_gnorm_{k}"])
        rel_norm = abs(np.sqrt((flat.astype(np.float64) ** 2).sum()) - norm) / norm
        rel_max = np.abs(flat[grad_sample_index(flat.size)] - want).max() / np.abs(want).max()
        worst = max(worst, rel_norm, rel_max)
        assert rel_norm <= 1e-3 and rel_max <= 1e-3, (k, rel_norm, rel_max)
    print(f"{name}: worst relative gradient error {worst:.2e}")


@pytest.mark.parametrize("B,layers,latent,act", [(1, 5, 256, "sine"), (7, 9, 128, "sine"), (4, 3, 64, "morlet")])
def test_training_matches_oracle_other_shapes(B, layers, latent, act):
    """Other batch sizes / depths / latent widths (the 'residual shape' L=9, Z=128 of BASELINE config 3) against the
    CPU restatement, every gradient compared in full."""
    p = 0.1
    sd_kw = dict(seed=40 + B, num_layers=layers, latent_dim=latent, mod_bias_shift=0.5)
    m, sd = _model(sd_kw, act, p, num_layers=layers, latent_dim=latent)
    keep_np = train_keep_mask(7 + B, layers, B * 576, 256, p)
    under_np, full_np = synth_tiles(50 + B, B), synth_tiles(60 + B, B)
    out, loss, grads = _iteration(m, torch.from_numpy(under_np).to(DEV), torch.from_numpy(full_np).to(DEV),
                                  torch.from_numpy(keep_np).to(DEV))
    torch.set_num_threads(4)
    want_out, want_loss, want = otrain.train_iteration(sd, torch.from_numpy(under_np), torch.from_numpy(full_np),
                                                       torch.from_numpy(keep_np), p, num_layers=layers, activation=act)
    assert float((out.cpu() - want_out).abs().max()) <= 5e-5
    assert abs(loss - want_loss) <= 1e-5
    for k, gr in grads.items():
        w = want[k]
        err = float((gr - w).abs().max()) / max(float(w.abs().max()), 1e-20)
        assert err <= 1e-3, (k, err)


def test_hash_dropout_statistics_and_determinism():
    """Without an explicit mask the keep decision is a hash of (seed, layer, element): the kept fraction is 1 - p, a
    call is reproducible under torch.manual_seed and differs between calls; eval() + no_grad stays on the inference path."""
    name, sd_kw, act, _ = TRAIN_CASES[0]
    p = 0.25
    m, sd = _model(sd_kw, act, p)
    m._train_keep_mask = None
    tiles = torch.from_numpy(synth_tiles(5, 6)).to(DEV)
    torch.manual_seed(3)
    with torch.no_grad():
        a = m(tiles)
        b = m(tiles)
    torch.manual_seed(3)
    with torch.no_grad():
        a2 = m(tiles)
    assert torch.equal(a, a2) and not torch.equal(a, b)
    m.eval()
    with torch.no_grad():
        e = m(tiles)
    want = osiren.model_forward(sd, tiles.cpu(), activation=act)
    assert float((e.cpu() - want).abs().max()) <= 1e-3
    assert float((a - e).abs().mean()) > 1e-3            # dropout actually did something
    # gradients enabled with p = 0: the training path without dropout is the eval forward in fp32-class arithmetic
    m0, _ = _model(sd_kw, act, 0.0)
    t0 = m0(tiles)
    assert t0.requires_grad
    assert float((t0.detach().cpu() - want).abs().max()) <= 5e-5


def test_optimizer_steps_reduce_the_loss():
    """Adam on the CUDA path for a few iterations of Trainer._train_iteration (training.py:177-207): the loss goes
    down, the packed weights follow the parameter updates (no stale handle), GradScaler composes with the op."""
    name, sd_kw, act, _ = TRAIN_CASES[0]
    m, sd = _model(sd_kw, act, 0.1)
    m._train_keep_mask = None
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    scaler = torch.amp.GradScaler("cuda")
    under = torch.from_numpy(synth_tiles(11, 16)).to(DEV)
    target = torch.from_numpy(synth_tiles(11, 16)).to(DEV)[:, 4:28, 4:28]      # learn to reproduce the centre crop
    losses = []
    torch.manual_seed(0)
    criterion = torch.nn.MSELoss()                                        # training.py:108-114 ("MSE")
    for it in range(12):
        with torch.amp.autocast("cuda", enabled=True):                    # training.py:197-204, line for line
            opt.zero_grad()
            outputs = m(under)
            loss = criterion(outputs, target)
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
            loss_item = loss.item()
        assert outputs.dtype == torch.float32
        losses.append(float(loss_item))
    print("losses", [round(x, 5) for x in losses])
    assert all(np.isfinite(losses)) and losses[-1] < 0.8 * losses[0]


def test_training_side_tiling_helpers():
    """filter_black_patches / filter_black_patches_indices / extract_center_batch (tiling.py:201-241,306-322), the
    helpers src/train/training.py:18-25 and the dataset import."""
    from mri_inr_b200 import tiling

    u = torch.from_numpy(synth_tiles(3, 9)).to(DEV)        # patch 1 is black
    u[4] = 0.0
    f = torch.from_numpy(synth_tiles(4, 9)).to(DEV)
    assert tiling.filter_black_patches_indices(u) == [0, 2, 3, 5, 6, 7, 8]
    (fu,), (ff,) = tiling.filter_black_patches([u.clone()], [f.clone()])
    assert fu.shape[0] == 7 and torch.equal(ff, f[[0, 2, 3, 5, 6, 7, 8]])
    c = tiling.extract_center_batch(f, 32, 24)
    assert c.shape == (9, 24, 24) and torch.equal(c, f[:, 4:28, 4:28])


def test_training_without_biases_and_w0_2():
    """use_bias=False (no bias parameters: the gradient view carries null pointers) and a hidden w0 of 2 (the activation
    derivative is w0 cos(w0 z)), against the CPU restatement."""
    from mri_inr_b200.modulated_siren import ModulatedSiren

    p, B = 0.1, 3
    sd = osiren.synth_state_dict(seed=16, use_bias=False, w0=2.0, mod_bias_shift=0.5)
    m = ModulatedSiren(dim_in=2, dim_hidden=256, dim_out=1, num_layers=5, latent_dim=256, w0=2.0, w0_initial=30.0,
                       use_bias=False, dropout=p, modulate=True, encoder_type="custom", encoder_path=None,
                       outer_patch_size=32, inner_patch_size=16, siren_patch_size=24, device=torch.device("cpu"),
                       activation="sine")
    m.load_state_dict(sd, strict=True)
    m.to(DEV).train()
    keep_np = train_keep_mask(3, 5, B * 576, 256, p)
    under_np, full_np = synth_tiles(81, B), synth_tiles(82, B)
    out, loss, grads = _iteration(m, torch.from_numpy(under_np).to(DEV), torch.from_numpy(full_np).to(DEV),
                                  torch.from_numpy(keep_np).to(DEV))
    want_out, want_loss, want = otrain.train_iteration(sd, torch.from_numpy(under_np), torch.from_numpy(full_np),
                                                       torch.from_numpy(keep_np), p, w0=2.0)
    assert float((out.cpu() - want_out).abs().max()) <= 5e-5 and abs(loss - want_loss) <= 1e-5
    assert set(grads) == set(want) and not any("net.layers" in k and k.endswith("bias") for k in grads)
    for k, gr in grads.items():
        err = float((gr - want[k]).abs().max()) / max(float(want[k].abs().max()), 1e-20)
        assert err <= 1e-3, (k, err)
