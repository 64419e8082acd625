"""Common base of the batch post-processing operators.

Every operator is a callable ``op(tree) -> tree`` that mutates leaves of the sample tree in place
and hands the very same tree object back, which is the contract the samplers rely on when they
chain operators (``sample_tree = transform(sample_tree)``, parllel/transforms/transform.py:8-11 and
parllel/samplers/basic.py:124-125).
"""
from __future__ import annotations

import abc


class Transform(abc.ABC):
    """Interface of a sample-tree operator; see the module docstring for the in-place contract."""

    #: tree fields the operator reads / writes; informative, used by ``describe()``
    reads: tuple = ()
    writes: tuple = ()

    @abc.abstractmethod
    def __call__(self, batch_samples):
        """Apply the operator to ``batch_samples`` in place and return the same object."""

    def describe(self) -> str:
        """One-line summary, handy when logging a transform chain."""
        io = f"reads {', '.join(self.reads) or '-'}; writes {', '.join(self.writes) or '-'}"
        return f"{type(self).__name__}({io})"
