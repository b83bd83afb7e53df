"""Import overlay: make the reference's own scripts (`test_mod_siren.py`, `src/util/error.py`) pick up this package
without editing them.

    import mri_inr_b200.compat as compat
    compat.install()                      # before `import test_mod_siren` / `from src... import ...`

After `install()`:

* ``from src.networks.modulated_siren import ModulatedSiren`` (test_mod_siren.py:14, train_mod_siren.py:18) resolves to
  :class:`mri_inr_b200.modulated_siren.ModulatedSiren`;
* ``from src.util.tiling import image_to_patches, patches_to_image_weighted_average, ...`` (test_mod_siren.py:15-19,
  src/util/error.py:14-19) resolves to :mod:`mri_inr_b200.tiling`.

Everything else under ``src`` (configuration, data, error, visualization) keeps coming from the reference tree, which
must be importable (its root on ``sys.path``) — or, when it is not, empty ``src`` / ``src.networks`` / ``src.util``
packages are created so that the two imports above still work stand-alone.
"""
from __future__ import annotations

import importlib
import sys
import types


def _ensure_package(name: str) -> types.ModuleType:
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        mod = types.ModuleType(name)
        mod.__path__ = []  # type: ignore[attr-defined]
        sys.modules[name] = mod
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(_ensure_package(parent), child, mod)
        return mod


def install(gpu_metrics: bool = False) -> None:
    """``gpu_metrics=True`` additionally replaces ``src.util.error.metrics_error`` and the ``calculate_*`` helpers by
    the device versions of :mod:`mri_inr_b200.error` (the reference's own module must then be importable, i.e.
    scikit-image / matplotlib installed; its ``visual_error`` stays untouched)."""
    from . import modulated_siren, tiling

    for pkg in ("src", "src.networks", "src.util"):
        _ensure_package(pkg)
    sys.modules["src.networks.modulated_siren"] = modulated_siren
    sys.modules["src.util.tiling"] = tiling
    setattr(sys.modules["src.networks"], "modulated_siren", modulated_siren)
    setattr(sys.modules["src.util"], "tiling", tiling)
    if gpu_metrics:
        from . import error as our_error

        try:
            ref_error = importlib.import_module("src.util.error")
        except Exception:
            # the reference tree (or scikit-image / matplotlib, which its error.py imports) is not there: this package's
            # module IS src.util.error then (metrics_error + the calculate_* helpers; no visual_error)
            sys.modules["src.util.error"] = our_error
            setattr(sys.modules["src.util"], "error", our_error)
        else:
            for name in our_error.__all__:
                setattr(ref_error, name, getattr(our_error, name))


def uninstall() -> None:
    for name in ("src.networks.modulated_siren", "src.util.tiling", "src.util.error"):
        mod = sys.modules.get(name)
        if mod is not None and mod.__name__.startswith("mri_inr_b200."):
            del sys.modules[name]
