"""CPU oracle for the mri-inr modulated-SIREN hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU arm), never as the thing shipped.  The product
path (``mri_inr_b200``) fails loudly when its CUDA library is missing; it never
routes through this package.

Parity status: PINNED.  The reference repository ships no golden vectors or
tests of its own, so the restatement here is pinned against
outputs of the *unmodified* reference code executed in the build container
(``oracle/ref_import.py`` imports ``/root/reference`` with import-time stubs for
its unrelated missing dependencies; ``oracle/make_golden.py`` writes the
fixtures under ``tests/golden/``).  The third-party arithmetic that is *not*
under ``/root/reference`` (scikit-image metrics, fastmri iFFT/mask) is restated
from the published algorithms and is marked "parity unpinned" in the modules
that hold it (``oracle/metrics.py``, ``oracle/kspace.py``).
"""
