"""Shape of one sampler batch: ``T`` time steps of ``B`` environments (the reference's
``parllel.types.BatchSpec``, types/batch_spec.py:4-10), plus the column sharding used when the env
axis is split across GPUs."""
from __future__ import annotations

import collections

_Base = collections.namedtuple("BatchSpec", ["T", "B"])


class BatchSpec(_Base):
    """``BatchSpec(T, B)``: unpacks like a tuple; ``size`` is the number of transitions."""

    __slots__ = ()

    def __new__(cls, T: int, B: int):
        if int(T) < 1 or int(B) < 1:
            raise ValueError(f"BatchSpec needs T >= 1 and B >= 1, got T={T}, B={B}")
        return super().__new__(cls, int(T), int(B))

    @property
    def size(self) -> int:
        return self.T * self.B

    def shard(self, world_size: int, rank: int) -> "BatchSpec":
        """The spec of one rank's column block when ``B`` is split over ``world_size`` ranks."""
        base, extra = divmod(self.B, world_size)
        return BatchSpec(self.T, base + (1 if rank < extra else 0))
