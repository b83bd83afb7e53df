"""CPU: the oracle restatement (oracle/) against the golden vectors produced by the unmodified reference
(tests/golden/*.npz, generator oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import siren, tiling
from oracle.synth import HARD_CASES, MODEL_CASES, synth_image, synth_tiles


@pytest.mark.parametrize("s", [8, 16, 24, 32, 48])
def test_grid_bit_exact(golden, s):
    g = siren.make_grid(s)
    assert g.dtype == np.float32
    assert np.array_equal(g.view(np.uint32), golden["grid_weights"][f"grid_{s}"].view(np.uint32))


def test_grid_buffer_is_linspace_grid(golden):
    assert np.array_equal(golden["grid_weights"]["grid_buffer_24"], golden["grid_weights"]["grid_24"])
    g = siren.make_grid(24)
    assert np.array_equal(g[::-1] * -1.0, g)          # antisymmetric
    assert g[1, 1] == np.float32(float.fromhex("-0x1.d37a70p-1"))


@pytest.mark.parametrize("k", [16, 24, 32])
def test_weight_matrix_bit_exact(golden, k):
    w = tiling.weight_matrix(k)
    assert np.array_equal(w.view(np.uint32), golden["grid_weights"][f"weights_{k}"].view(np.uint32))
    assert w.max() == 1.0


@pytest.mark.parametrize("case", MODEL_CASES + HARD_CASES, ids=[c[0] for c in MODEL_CASES + HARD_CASES])
def test_model_forward_matches_reference(golden, case):
    name, sd_kw, act, model_kw = case
    sd = siren.synth_state_dict(**sd_kw)
    tiles = torch.from_numpy(synth_tiles(100 + sd_kw["seed"], 5))
    torch.set_num_threads(1)
    out, z, mods = siren.model_forward(sd, tiles, num_layers=model_kw.get("num_layers", 5),
                                       w0=model_kw.get("w0", 1.0), activation=act, return_intermediates=True)
    g = golden["model_forward"]
    np.testing.assert_allclose(z.numpy(), g[f"{name}_latent"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(torch.stack(mods).numpy(), g[f"{name}_mods"], rtol=0, atol=5e-6)
    # same ATen ops in the same order as the reference: agreement to fp32 rounding noise
    np.testing.assert_allclose(out.numpy(), g[f"{name}_out"], rtol=0, atol=2e-5)
    assert out.shape == (5, 24, 24)


def test_hard_cases_span_the_output_range(golden):
    """W x 2 / W x 3 with dense modulations: the regime where operand rounding shows."""
    g = golden["model_forward"]
    for name, *_ in HARD_CASES:
        assert np.ptp(g[f"{name}_out"]) > 1.9


def test_trained_like_case_is_sensitive(golden):
    """The trained-like weights must produce outputs that actually vary (random init is nearly constant)."""
    g = golden["model_forward"]
    assert np.ptp(g["sine_trained_out"]) > 0.5
    assert np.ptp(g["sine_init_out"]) < 0.05


@pytest.mark.parametrize("tag,hw", [("a", (50, 37)), ("b", (64, 48)), ("c", (320, 320))])
def test_tiling_matches_reference(golden, tag, hw):
    g = golden["tiling"]
    h, w = hw
    img = synth_image(7 + h, h, w)
    patches, info = tiling.image_to_patches(img[None], 32, 16)
    assert tuple(info[0]) == tuple(g[f"{tag}_info"])
    if tag != "c":
        assert np.array_equal(patches, g[f"{tag}_patches"])
    else:
        np.testing.assert_allclose(patches.astype(np.float64).sum(axis=(1, 2)), g["c_patches_sum"], rtol=1e-12)
    kept, black, shape = tiling.filter_and_remember_black_patches(patches)
    assert black == list(g[f"{tag}_black"])
    assert len(black) > 0
    rs = np.random.RandomState(h * w)
    small = rs.uniform(-1, 1, size=(patches.shape[0], 24, 24)).astype(np.float32)
    keep = np.ones(patches.shape[0], bool)
    keep[black] = False
    reint = tiling.reintegrate_black_patches(small[keep], black, shape)
    assert np.all(reint[black] == 0)
    wavg = tiling.patches_to_image_weighted_average(reint, info, 24, 16)
    plain = tiling.patches_to_image(patches, info, 32, 16)
    assert wavg.shape == g[f"{tag}_wavg"].shape
    # the restatement accumulates in F.fold's order (decreasing patch row, then column): bit-identical to the reference
    assert np.array_equal(wavg.view(np.uint32), g[f"{tag}_wavg"].view(np.uint32))
    assert np.array_equal(plain.view(np.uint32), g[f"{tag}_plain"].view(np.uint32))


def test_normalize_scan(golden):
    x = np.random.RandomState(5).normal(size=(3, 40, 40)).astype(np.float32)
    assert np.array_equal(tiling.normalize_scan(x), golden["normalize"]["out"])


def test_state_dict_layout():
    keys = siren.state_dict_key_order()
    assert len(keys) == 31 and keys[0] == "grid"
    sd = siren.synth_state_dict(0)
    assert list(sd.keys()) == keys
    assert sum(v.numel() for k, v in sd.items() if k != "grid") == 1007873
