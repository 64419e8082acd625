"""``BatchSpec`` (parllel/types/batch_spec.py:4-10)."""
from typing import NamedTuple


class BatchSpec(NamedTuple):
    T: int  # time steps per sampler batch
    B: int  # environment instances

    @property
    def size(self) -> int:
        return self.T * self.B
