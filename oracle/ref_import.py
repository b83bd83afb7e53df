"""Import the UNMODIFIED reference (``/root/reference``) in the build container.

Test infrastructure only (see ``oracle/__init__.py``).  ``/root/reference`` does
not exist on the GPU box, so nothing that runs there may call into this module;
it is used by ``oracle/make_golden.py`` (fixture generation) and by the CPU
tests that are skipped when the reference tree is absent.

The reference imports, at module import time, packages that are unrelated to
the hot path and absent from this image:

* ``src/networks/encoding/siren_encoder.py:11-15`` -> ``src/data/mri_sampler.py:8,10``
  (``polars``, ``fastmri``) and ``src/util/error.py:10-12`` (``skimage.metrics``),
  ``src/util/visualization.py:8-10`` (``matplotlib``, ``seaborn``).

They are replaced by empty stub modules before the import; no reference source
is modified or copied.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("MRINR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "networks", "modulated_siren.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _install_stubs() -> None:
    def _missing(*_a, **_k):
        raise RuntimeError("stubbed third-party function called (not part of the hot path)")

    if "polars" not in sys.modules:
        cfg = types.SimpleNamespace(set_tbl_rows=lambda *a, **k: None)
        _stub("polars", Config=cfg, LazyFrame=object, DataFrame=object,
              scan_csv=_missing, read_csv=_missing, col=_missing)
    if "fastmri" not in sys.modules:
        fm = _stub("fastmri", ifft2c=_missing, complex_abs=_missing)
        data = _stub("fastmri.data")
        tr = _stub("fastmri.data.transforms", to_tensor=_missing, apply_mask=_missing)
        sub = _stub("fastmri.data.subsample", RandomMaskFunc=_missing)
        fm.data = data
        data.transforms = tr
        data.subsample = sub
    if "skimage" not in sys.modules:
        sk = _stub("skimage")
        sk.metrics = _stub("skimage.metrics", normalized_root_mse=_missing,
                           peak_signal_noise_ratio=_missing, structural_similarity=_missing)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib", use=lambda *a, **k: None)
        mpl.pyplot = _stub("matplotlib.pyplot")
    if "seaborn" not in sys.modules:
        _stub("seaborn")
    if "h5py" not in sys.modules:
        _stub("h5py", File=_missing)


_cached = None


def load_reference():
    """Return a namespace with the reference's own modules (imported unmodified)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    # The reference uses the top-level package name ``src``.  If some other ``src``
    # is already imported (e.g. this repo's overlay), refuse rather than mix them.
    if "src" in sys.modules and not getattr(sys.modules["src"], "__file__", "").startswith(REFERENCE_ROOT):
        raise RuntimeError("another top-level package named 'src' is already imported")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    ns.modulated_siren = importlib.import_module("src.networks.modulated_siren")
    ns.siren_encoder = importlib.import_module("src.networks.encoding.siren_encoder")
    ns.tiling = importlib.import_module("src.util.tiling")
    ns.configuration = importlib.import_module("src.configuration.configuration")
    ns.visualization = importlib.import_module("src.util.visualization")
    _cached = ns
    return ns


def synth_encoder_checkpoint(path: str, seed: int = 0) -> str:
    """Write ``{"state_dict": FixedAutoencoder().state_dict()}`` (siren_encoder.py:544-549)."""
    import torch

    ref = load_reference()
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    ae = ref.siren_encoder.FixedAutoencoder()
    torch.random.set_rng_state(g)
    torch.save({"state_dict": ae.state_dict()}, path)
    return path


def build_reference_model(activation: str = "sine", seed: int = 0, num_layers: int = 5,
                          latent_dim: int = 256, dim_hidden: int = 256, use_bias: bool = True,
                          siren_patch_size: int = 24, w0: float = 1.0, w0_initial: float = 30.0):
    """Instantiate the reference ``ModulatedSiren`` with the baseline kwargs
    (configuration/test_modulated_siren.yaml:17-32; call site test_mod_siren.py:96-114)."""
    import torch

    ref = load_reference()
    tmp = tempfile.mkdtemp(prefix="mrinr_ref_")
    enc_path = synth_encoder_checkpoint(os.path.join(tmp, "custom_encoder.pth"), seed=seed)
    torch.manual_seed(seed)
    model = ref.modulated_siren.ModulatedSiren(
        dim_in=2, dim_hidden=dim_hidden, dim_out=1, num_layers=num_layers, latent_dim=latent_dim,
        w0=w0, w0_initial=w0_initial, use_bias=use_bias, dropout=0.1, modulate=True,
        encoder_type="custom", encoder_path=enc_path, outer_patch_size=32, inner_patch_size=16,
        siren_patch_size=siren_patch_size, device=torch.device("cpu"), activation=activation,
    )
    model.eval()
    return model
