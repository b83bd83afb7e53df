"""SURVEY H6 / VERDICT r1 #10: the reference's evaluation script, replayed line for line through the import overlay
(tools/replay_test_mod_siren.py), on synthetic slices; every CSV row is compared with the CPU oracle flow
(oracle/flow.py + oracle/metrics.py).  Tolerances: north_star's 0.05 dB PSNR / 1e-3 SSIM (and 1e-3 NRMSE)."""
import csv
import os

import numpy as np
import pytest
import torch

from oracle import flow, metrics
from oracle import siren as osiren


@pytest.mark.gpu
@pytest.mark.parametrize("activation,seed", [("sine", 12), ("morlet", 14)])
def test_replay_of_test_mod_siren(tmp_path, activation, seed):
    from mri_inr_b200.synthetic import synthetic_slices
    from tools.replay_test_mod_siren import SyntheticSampler, make_config, replay

    sd = osiren.synth_state_dict(seed, mod_bias_shift=0.5)               # trained-like: the output actually varies
    ckpt = tmp_path / "modulated_siren.pth"
    torch.save(sd, ckpt)                                                 # what test_mod_siren.py:116-118 loads
    n = 3
    under = synthetic_slices(n, 320, 320, device="cuda:0", seed=77).cpu()
    full = synthetic_slices(n, 320, 320, device="cuda:0", seed=77, undersampled=False).cpu()
    sampler = SyntheticSampler(full, under)
    config = make_config(model=dict(activation=activation),
                         testing=dict(output_dir=str(tmp_path), output_name="modulated_sired", model_path=str(ckpt)))
    path = replay(config, sampler)
    assert path.endswith(os.path.join("modulated_sired", "test", "metrics_error.csv"))
    rows = list(csv.reader(open(path)))
    assert rows[0] == ["FILENAME", "PSNR", "SSIM", "NRMSE"]              # test_mod_siren.py:238
    assert len(rows) == 1 + n and [r[0] for r in rows[1:]] == sampler.names
    for i, row in enumerate(rows[1:]):
        want = flow.reconstruct_slice(sd, under[i], activation=activation).numpy()
        ref = full[i].numpy()
        psnr, ssim, nrmse = float(row[1]), float(row[2]), float(row[3])
        assert abs(psnr - metrics.psnr(ref, want)) <= 0.05, (i, psnr, metrics.psnr(ref, want))
        assert abs(ssim - metrics.ssim(ref, want)) <= 1e-3
        assert abs(nrmse - metrics.nrmse(ref, want)) <= 1e-3
    assert np.isfinite([float(r[1]) for r in rows[1:]]).all()
