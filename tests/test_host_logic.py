"""CPU: host-side mirror of the reference interface (no GPU, no kernels)."""
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import siren as osiren

KW = dict(dim_in=2, dim_hidden=256, dim_out=1, num_layers=5, latent_dim=256, w0=1.0, w0_initial=30.0, use_bias=True,
          dropout=0.1, modulate=True, encoder_type="custom", encoder_path=None, outer_patch_size=32,
          inner_patch_size=16, siren_patch_size=24, device=torch.device("cpu"), activation="sine")


def make(**over):
    from mri_inr_b200.modulated_siren import ModulatedSiren

    kw = dict(KW)
    kw.update(over)
    return ModulatedSiren(**kw)


def test_ctor_takes_the_17_reference_kwargs_and_sets_attributes():
    m = make()
    for name in ("dim_hidden", "dim_out", "num_layers", "latent_dim", "modulate", "encoder_type", "outer_patch_size",
                 "inner_patch_size", "siren_patch_size", "activation", "net", "modulator", "encoder", "grid"):
        assert hasattr(m, name), name
    with pytest.raises(TypeError):
        make(bogus=1)


def test_state_dict_layout_matches_reference():
    m = make()
    sd = m.state_dict()
    assert list(sd.keys()) == osiren.state_dict_key_order()
    assert len(sd) == 31
    assert sum(p.numel() for p in m.parameters()) == 1007873
    assert tuple(sd["grid"].shape) == (576, 2)
    assert tuple(sd["modulator.layers.1.0.weight"].shape) == (256, 512)
    assert tuple(sd["encoder.encoder.encoder.4.weight"].shape) == (64, 32, 8, 8)
    nb = make(use_bias=False).state_dict()
    assert list(nb.keys()) == osiren.state_dict_key_order(use_bias=False)


def test_strict_load_of_reference_state_dict():
    m = make()
    sd = osiren.synth_state_dict(3)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m.net.layers[2].weight, sd["net.layers.2.weight"])
    bad = dict(sd)
    bad.pop("grid")
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad, strict=True)


def test_grid_buffer_bit_exact(golden):
    from mri_inr_b200.modulated_siren import make_grid_host

    for s in (8, 16, 24, 32, 48):
        assert np.array_equal(make_grid_host(s).numpy().view(np.uint32), golden["grid_weights"][f"grid_{s}"].view(np.uint32))
    assert np.array_equal(make().grid.numpy(), golden["grid_weights"]["grid_buffer_24"])


def test_init_ranges_follow_reference():
    torch.manual_seed(0)
    m = make()
    assert float(m.net.layers[0].weight.abs().max()) <= 0.5              # U(+-1/dim_in)
    b = np.sqrt(6.0 / 256.0)
    assert float(m.net.layers[1].weight.abs().max()) <= b + 1e-6         # U(+-sqrt(6/dim)/w0)
    assert float(m.net.layers[1].weight.abs().max()) > 0.9 * b
    assert float(m.net.last_layer.bias.abs().max()) <= b + 1e-6


def test_weight_matrix_bit_exact(golden):
    from mri_inr_b200.tiling import generate_weight_matrix

    for k in (16, 24, 32):
        assert np.array_equal(generate_weight_matrix(k).numpy().view(np.uint32),
                              golden["grid_weights"][f"weights_{k}"].view(np.uint32))


def test_no_cpu_path_and_inference_only():
    m = make().eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(2, 32, 32))
    m.train()
    with pytest.raises(RuntimeError, match="CUDA"):          # the training path is CUDA-only as well
        m(torch.zeros(2, 32, 32))
    # the batched pipeline and the two halves of forward are inference-only entry points
    with torch.no_grad(), pytest.raises(RuntimeError, match="eval"):
        m._check_inference()
    m.eval()
    with pytest.raises(RuntimeError, match="no_grad"):
        m._check_inference()
    from mri_inr_b200 import ops

    with pytest.raises(RuntimeError, match="CUDA"):
        ops.image_to_patches(torch.zeros(1, 64, 64), 32, 16)


def test_out_of_scope_encoder_is_refused():
    with pytest.raises(NotImplementedError):
        make(encoder_type="vgg")


def test_fixed_encoder_loads_reference_checkpoint_format():
    """FixedEncoder reads {"state_dict": FixedAutoencoder.state_dict()} (siren_encoder.py:544-549): encoder.* keys are
    used, decoder.* keys ignored."""
    sd = osiren.synth_state_dict(5)
    p = "encoder.encoder.encoder."
    ckpt = {"encoder." + k[len(p):]: v for k, v in sd.items() if k.startswith(p)}
    ckpt["decoder.0.weight"] = torch.zeros(64, 256)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "custom_encoder.pth")
        torch.save({"state_dict": ckpt}, path)
        m = make(encoder_path=path)
    assert torch.equal(m.encoder.encoder.encoder[7].weight, sd[p + "7.weight"])
    # the loaded layers (evaluated through PyTorch, test-side only) equal the oracle restatement ...
    tiles = torch.rand(3, 32, 32)
    with torch.no_grad():
        z = m.encoder.fc(m.encoder.encoder(tiles))      # the PyTorch submodules, as the reference evaluates them
    assert torch.allclose(z, osiren.encoder_forward(sd, tiles), atol=1e-6)
    # ... and the product path refuses CPU tensors instead of falling back
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            m.encoder(tiles)


def test_shard_range_partitions_every_item_once():
    from mri_inr_b200.dist import shard_counts, shard_range

    for n in (0, 1, 7, 10340):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
            assert shard_counts(n, world) == [e - s for s, e in spans]
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_column_mask_statistics():
    from mri_inr_b200.synthetic import column_mask

    m = column_mask(320, 6, 0.05, 1234)
    assert m.dtype == bool and m.shape == (320,)
    assert m[152:168].all()                       # 16 centre columns (round(320*0.05))
    assert 30 <= m.sum() <= 80                    # ~320/6 = 53 columns expected


def test_import_overlay_redirects_reference_imports():
    """`from src.networks.modulated_siren import ModulatedSiren` (test_mod_siren.py:14) resolves to this package."""
    import subprocess
    import sys

    code = ("import mri_inr_b200.compat as c; c.install();"
            "from src.networks.modulated_siren import ModulatedSiren;"
            "from src.util.tiling import image_to_patches, patches_to_image_weighted_average, patches_to_image, "
            "filter_and_remember_black_patches, reintegrate_black_patches;"
            "import mri_inr_b200.modulated_siren as m; assert ModulatedSiren is m.ModulatedSiren; print('ok')")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_fma_pipe_gaussian_coefficients():
    """csrc/tc_ptx.cuh gauss2: exp(-x^2/2) = 2^y by magic-number rounding + a degree-4 polynomial of the fraction +
    an integer add into the exponent.  Emulated here in numpy fp32 (FMAs as fp64 products rounded once), with the
    coefficient bit patterns parsed from the CUDA source: relative error <= 1e-5 where the envelope matters."""
    import re
    import struct

    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mri_inr_b200", "csrc",
                            "tc_ptx.cuh")).read()
    body = src[src.index("uint64_t gauss2("):]
    body = body[:body.index("return pk2(g0, g1);")]
    def lits(text):
        return [struct.unpack("<f", struct.pack("<I", int(x, 16) & 0xffffffff))[0]
                for x in re.findall(r"0x([0-9A-F]{16})ULL", text)]

    negc, magic, magic2, c4, c3, c2, c1, c0 = lits(body)
    assert magic == magic2 == 12582912.0 and abs(negc + 0.5 / np.log(2)) < 1e-7
    x = np.concatenate([np.linspace(-14, 14, 400001), np.random.RandomState(0).normal(size=200000) * 2]).astype(np.float32)
    t = np.minimum((x * x).astype(np.float32), np.float32(174.0))
    y = t.astype(np.float64) * np.float64(np.float32(negc))
    z = (y + magic).astype(np.float32)
    nneg = (np.float32(magic) - z).astype(np.float32)
    fr = (y + nneg.astype(np.float64)).astype(np.float32)
    assert fr.min() >= -0.5001 and fr.max() <= 0.5001
    p = np.full_like(fr, np.float32(c4))
    for c in (c3, c2, c1, c0):
        p = (p.astype(np.float64) * fr.astype(np.float64) + np.float64(np.float32(c))).astype(np.float32)
    g = (p.view(np.int32) + (z.view(np.int32) << 23)).view(np.float32)
    ref = np.exp(-0.5 * x.astype(np.float64) ** 2)
    assert np.abs(g - ref).max() <= 5e-6
    near = np.abs(x) < 9
    assert np.abs(g[near] / ref[near] - 1).max() <= 1e-5
    assert np.all(np.isfinite(g)) and np.all(g >= 0) and g[np.abs(x) > 13.5].max() < 1e-37


def test_bench_chunking_and_reference_arm_line():
    """bench.py host logic: the chunk size divides every rank's block without a ragged tail worth mentioning, and the
    `--impl reference` arm prints one JSON line with the contract's keys (CPU only, bounded sample)."""
    import json
    import subprocess
    import sys

    import bench
    from mri_inr_b200.dist import shard_range

    for world in (1, 2, 4, 8):
        n_local = shard_range(10340, 0, world)[1]
        c = bench.auto_chunk(n_local)
        n_chunks = -(-n_local // c)
        assert 200 <= c <= 235 and n_chunks * c - n_local < n_chunks          # last chunk short by < 1 slice per chunk
    assert bench.auto_chunk(0) == 235 and bench.auto_chunk(7) == 7
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-slices", "1",
                        "--num-layers", "3"], cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in line, k
    assert line["impl"] == "reference" and line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["num_layers"] == 3
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1


def test_host_schedule_covers_every_slice_once_with_short_edge_chunks():
    """pipeline.host_schedule: contiguous cover of [0, n), no chunk above the buffer size, and -- with more than two
    chunks of work -- a first and a last chunk of chunk // 4 slices (the only exposed transfers of the e2e leg)."""
    from mri_inr_b200.pipeline import host_schedule

    for n, c in [(10340, 235), (1293, 216), (1, 235), (470, 235), (471, 235), (0, 235), (7, 2), (100, 3), (5170, 235)]:
        sch = host_schedule(n, c)
        assert sum(k for _, k in sch) == n
        assert all(0 < k <= c for _, k in sch)
        assert all(sch[i][0] + sch[i][1] == sch[i + 1][0] for i in range(len(sch) - 1))
        if sch:
            assert sch[0][0] == 0
        if n > 2 * c and c >= 4:
            assert sch[0][1] == c // 4 and sch[-1][1] == c // 4
            mid = [k for _, k in sch[1:-1]]
            assert max(mid) - min(mid) <= 1
