"""``ArrayDict``: the container type at the drop-in boundary.

Behavioural mirror of parllel's ``ArrayDict`` (parllel/tree/array_dict.py:15-162): a
string-keyed mapping of array-like leaves (numpy arrays, torch tensors, parllel
``Array``s, :class:`parllel_b200.arrays.DeviceArray`) that share leading batch
dimensions.  ``tree["k"]`` returns a leaf; any other key indexes every leaf and
returns a new tree; assignment with a non-string key fans out; attribute access fans
out (``tree.rotate()``, ``tree.to(device=...)``).  The transforms and the replay
buffer accept either this class or parllel's own (duck-typed through ``Mapping``).
"""
from __future__ import annotations

import dataclasses
from collections.abc import Mapping, MutableMapping
from typing import Any, Callable, Iterator

import numpy as np


def _wrap_children(items: dict) -> dict:
    for name, child in items.items():
        if isinstance(child, Mapping) and not isinstance(child, ArrayDict):
            items[name] = ArrayDict(child)
    return items


class ArrayDict(MutableMapping):
    def __init__(self, items=None, _run_checks: bool = True) -> None:
        store = dict(items) if items is not None else {}
        self._dict = _wrap_children(store) if _run_checks else store

    # -- mapping protocol -------------------------------------------------------------
    def __getitem__(self, key: Any):
        if isinstance(key, str):
            return self._dict[key]
        picked = {}
        for name, leaf in self._dict.items():
            try:
                picked[name] = leaf[key]
            except IndexError as err:
                raise IndexError(f"Index error in field '{name}' for index '{key}'") from err
        return ArrayDict(picked, _run_checks=False)

    def __setitem__(self, key: Any, value: Any) -> None:
        if isinstance(key, str):
            self._dict[key] = value
            return
        if isinstance(value, Mapping):
            sources = {name: value[name] for name in self._dict.keys() & value.keys()}
        elif dataclasses.is_dataclass(value):
            names = {f.name for f in dataclasses.fields(value)}
            sources = {name: getattr(value, name) for name in self._dict.keys() & names}
        else:  # one value (scalar or array) broadcast to every field
            sources = {name: value for name in self._dict}
        for name, sub in sources.items():
            try:
                self._dict[name][key] = sub
            except IndexError as err:
                raise IndexError(f"Index error in field '{name}' for index '{key}'") from err

    def __delitem__(self, key: str) -> None:
        if not isinstance(key, str):
            raise IndexError(f"Cannot delete index {key}")
        del self._dict[key]

    def __iter__(self) -> Iterator[str]:
        return iter(self._dict)

    def __len__(self) -> int:
        return len(self._dict)

    def __contains__(self, key) -> bool:
        return key in self._dict

    def get(self, key: str, default=None):
        return self._dict.get(key, default)

    def __repr__(self) -> str:
        return f"ArrayDict({self._dict!r})"

    # -- attribute fan-out ------------------------------------------------------------
    def __getattr__(self, name: str):
        if name.startswith("__") or name == "_dict":
            raise AttributeError(name)
        gathered = {}
        for field, leaf in self._dict.items():
            try:
                gathered[field] = getattr(leaf, name)
            except AttributeError as err:
                raise AttributeError(f"Attribute error in field '{field}' for attribute '{name}'") from err
        return _FannedAttribute(gathered, name)

    def __getstate__(self):
        return {"_dict": self._dict}

    def __setstate__(self, state) -> None:
        self.__dict__.update(state)

    # -- functional helpers -----------------------------------------------------------
    def apply(self, fn: Callable, **kwargs) -> "ArrayDict":
        out = {}
        for field, leaf in self._dict.items():
            out[field] = leaf.apply(fn, **kwargs) if isinstance(leaf, ArrayDict) else fn(leaf, **kwargs)
        return ArrayDict(out, _run_checks=False)

    map = apply

    def to_ndarray(self) -> "ArrayDict":
        return self.apply(_leaf_to_ndarray)


def _leaf_to_ndarray(leaf):
    if hasattr(leaf, "to_ndarray"):
        return leaf.to_ndarray()
    if hasattr(leaf, "detach") and hasattr(leaf, "cpu"):  # torch tensor, possibly on a GPU
        return leaf.detach().cpu().numpy()
    return np.asarray(leaf)


class _FannedAttribute(ArrayDict):
    """Result of ``tree.<attr>``: a tree of per-leaf attributes; calling it calls each."""

    def __init__(self, items, name: str) -> None:
        super().__init__(items, _run_checks=False)
        self.__dict__["name"] = name

    def __call__(self, *args, **kwargs) -> ArrayDict:
        results = {}
        for field, member in self._dict.items():
            if isinstance(member, _FannedAttribute):
                results[field] = member(*args, **kwargs)
                continue
            if not callable(member):
                raise RuntimeError(f"Attribute '{self.name}' of field '{field}' is not callable!")
            try:
                results[field] = member(*args, **kwargs)
            except Exception as err:
                raise RuntimeError(f"Exception from calling method '{self.name}' of field '{field}'") from err
        return ArrayDict(results)


def dict_map(func: Callable, tree, *args, **kwargs):
    """parllel/tree/utils.py:11 -- apply ``func`` to every leaf of a (nested) mapping."""
    if isinstance(tree, Mapping):
        return ArrayDict({k: dict_map(func, v, *args, **kwargs) for k, v in tree.items()})
    return func(tree, *args, **kwargs)
