"""Import the UNMODIFIED upstream parllel from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so
nothing that runs there (``-m gpu`` tests, ``smoke()``, ``bench.py``) may call
this; it is used by ``tests/golden/gen_golden.py`` (fixture generation) and by
the ``-m "not gpu"`` tests that pin the oracle against the live reference.

The only import the reference needs that this image lacks is ``gymnasium``
(type names + ``gymnasium.utils.colorize``, parllel/logger/logger.py:11); the
stub package in ``baseline/gymnasium_stub`` satisfies it.
"""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = os.environ.get("PLB_REFERENCE_ROOT", "/root/reference")
_STUB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "gymnasium_stub")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "parllel"))


def load_reference():
    """Return the imported upstream ``parllel`` module (raises if unavailable)."""
    if not reference_available():
        raise ImportError(f"reference not mounted at {REFERENCE_ROOT}")
    try:
        import gymnasium  # noqa: F401
    except ImportError:
        if _STUB not in sys.path:
            sys.path.insert(0, _STUB)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import parllel  # noqa: E402

    return parllel
