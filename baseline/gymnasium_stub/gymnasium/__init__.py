"""Minimal stand-in for the `gymnasium` package (TEST SCAFFOLDING, not product code).

The upstream parllel package imports gymnasium only for type names and for
`gymnasium.utils.colorize`; none of that is on the batch post-processing path.
This stub lets `oracle/ref_loader.py` import the unmodified reference in the
build container to validate the oracle and to generate golden vectors.
"""
from . import spaces, utils  # noqa: F401


class Env:  # pragma: no cover - type name only
    pass


class Wrapper(Env):  # pragma: no cover - type name only
    def __init__(self, env=None):
        self.env = env


Space = spaces.Space
