"""CPU: the training-iteration restatement (oracle/train.py) against the goldens produced by the unmodified
reference's own autograd (tests/golden/training.npz)."""
import numpy as np
import pytest
import torch

from oracle import siren, train
from oracle.synth import TRAIN_BATCH, TRAIN_CASES, grad_sample_index, synth_tiles, train_keep_mask


@pytest.mark.parametrize("case", TRAIN_CASES, ids=[c[0] for c in TRAIN_CASES])
def test_train_iteration_matches_reference_autograd(golden, case):
    name, sd_kw, act, p = case
    g = golden["training"]
    sd = siren.synth_state_dict(**sd_kw)
    B = TRAIN_BATCH
    keep = torch.from_numpy(train_keep_mask(1000 + sd_kw["seed"], 5, B * 576, 256, p)) if p > 0 else None
    under = torch.from_numpy(synth_tiles(300 + sd_kw["seed"], B))
    full = torch.from_numpy(synth_tiles(400 + sd_kw["seed"], B))
    torch.set_num_threads(1)
    out, loss, grads = train.train_iteration(sd, under, full, keep, p, activation=act)
    np.testing.assert_allclose(out.numpy(), g[f"{name}_out"], rtol=0, atol=2e-5)
    assert abs(loss - float(g[f"{name}_loss"])) <= 1e-6
    assert len(grads) == 30
    for k, gr in grads.items():
        flat = gr.numpy().reshape(-1)
        want = g[f"{name}_gsample_{k}"]
        norm = float(g[f"{name}_gnorm_{k}"])
        assert abs(np.sqrt((flat.astype(np.float64) ** 2).sum()) - norm) <= 1e-4 * norm, k
        err = np.abs(flat[grad_sample_index(flat.size)] - want).max()
        assert err <= 1e-4 * max(np.abs(want).max(), 1e-12), (k, err)
