"""``Transform`` interface (parllel/transforms/transform.py:8-11): ``__call__(tree) -> tree``;
implementations mutate leaves in place and return the same tree object."""
from __future__ import annotations

from abc import ABC, abstractmethod


class Transform(ABC):
    @abstractmethod
    def __call__(self, batch_samples):
        raise NotImplementedError
