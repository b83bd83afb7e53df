"""mri_inr_b200 -- B200-native (sm_100a) implementation of the mri-inr hot path: the modulated-SIREN dense
forward over every patch's coordinate grid, and the tiling either side of it.

Public surface mirrors the reference (``src.networks.modulated_siren.ModulatedSiren``, ``src.util.tiling``);
the compute lives in ``libmrinr.so`` (C ABI: ``include/mrinr.h``).  No CPU fallback.
"""
from . import _lib  # noqa: F401
from .modulated_siren import ModulatedSiren  # noqa: F401

__all__ = ["ModulatedSiren"]
__version__ = "0.1.0"
