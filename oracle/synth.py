"""Deterministic synthetic inputs shared by the fixture generator (``oracle/make_golden.py``) and the tests.
Test infrastructure only (see ``oracle/__init__.py``).  numpy legacy ``RandomState`` streams are frozen across
numpy versions, so the same seeds reproduce the same arrays on the GPU box."""
from __future__ import annotations

import numpy as np

# (name, kwargs for synth_state_dict, activation, extra model kwargs)
MODEL_CASES = [
    ("sine_init", dict(seed=11), "sine", {}),
    ("sine_trained", dict(seed=12, mod_bias_shift=0.5), "sine", {}),
    ("morlet_init", dict(seed=13), "morlet", {}),
    ("morlet_trained", dict(seed=14, mod_bias_shift=0.5), "morlet", {}),
    # deeper MLP (the reference's FixedAutoencoder is hard-wired to a 256-d latent, siren_encoder.py:498-500,541,
    # so the reference itself can only run latent_dim=256; the reduced-latent "residual shape" is oracle-only)
    ("deep9", dict(seed=15, num_layers=9, mod_bias_shift=0.25), "sine", dict(num_layers=9)),
    ("nobias_w0_2", dict(seed=16, use_bias=False, w0=2.0, mod_bias_shift=0.5), "sine", dict(use_bias=False, w0=2.0)),
]

# Weight scales at which an 11-bit significand no longer meets the 1e-3 bound (hidden W x 2, dense
# modulations ~1): the cases the fp16x3 tensor-core mode exists for.  Same tuple layout as MODEL_CASES.
HARD_CASES = [
    ("sine_w2_dense", dict(seed=17, mod_bias_shift=1.0, hidden_weight_scale=2.0), "sine", {}),
    ("morlet_w2_dense", dict(seed=18, mod_bias_shift=1.0, hidden_weight_scale=2.0), "morlet", {}),
    ("sine_w3_dense", dict(seed=19, mod_bias_shift=1.0, hidden_weight_scale=3.0), "sine", {}),
]


def synth_tiles(seed: int, n: int, outer: int = 32) -> np.ndarray:
    """Smooth-ish non-negative patches in [0,1] (like normalised MRI magnitudes), with patch 1 black."""
    rs = np.random.RandomState(seed)
    base = rs.uniform(0.0, 1.0, size=(n, outer, outer)).astype(np.float32)
    ramp = np.linspace(0, 1, outer, dtype=np.float32)
    t = 0.5 * base + 0.5 * ramp[None, :, None] * ramp[None, None, :]
    if n > 1:
        t[1] = 0.0
    return t.astype(np.float32)


def synth_image(seed: int, h: int, w: int) -> np.ndarray:
    rs = np.random.RandomState(seed)
    img = rs.uniform(0.0, 1.0, size=(h, w)).astype(np.float32)
    img[: max(24, h // 4), : max(24, w // 4)] = 0.0   # a black corner (air around the head): >= 1 black patch
    return img




# Training parity cases (src/train/training.py:177-207): (name, synth_state_dict kwargs, activation, dropout p).
# p = 0: dropout off; p > 0: a FIXED keep-mask (train_keep_mask) replaces torch's dropout RNG on both sides.
TRAIN_CASES = [
    ("train_sine_nodrop", dict(seed=12, mod_bias_shift=0.5), "sine", 0.0),
    ("train_sine_drop", dict(seed=12, mod_bias_shift=0.5), "sine", 0.1),
    ("train_morlet_drop", dict(seed=14, mod_bias_shift=0.5), "morlet", 0.1),
]
TRAIN_BATCH = 3
GRAD_SAMPLE = 2048


def train_keep_mask(seed: int, num_layers: int, rows: int, hidden: int, p: float) -> np.ndarray:
    """uint8 keep-mask ``[L, rows, H]`` (1 = keep, probability 1 - p), rows = B * S * S."""
    rs = np.random.RandomState(seed)
    return (rs.random_sample((num_layers, rows, hidden)) >= p).astype(np.uint8)


def grad_sample_index(numel: int) -> np.ndarray:
    """Fixed sample of a flattened gradient (all of it when small): what the golden files keep per parameter."""
    if numel <= GRAD_SAMPLE:
        return np.arange(numel)
    return (np.arange(GRAD_SAMPLE, dtype=np.int64) * numel) // GRAD_SAMPLE
