"""Nested-structure helpers with the reference's `vima.utils` surface, re-implemented without dm-tree.

Mirrors (behaviour, not code) /root/reference/vima/utils.py: `__all__` (:13-24) + `DataDict` (:228-647).  Host-side
glue only -- nothing here is on the GPU hot path.  Mappings are walked in sorted-key order (dm-tree's rule),
which is what fixes the `front`,`top` view order downstream.
"""
from __future__ import annotations

import collections.abc as cabc
from typing import Any, Callable

import numpy as np
import torch

__all__ = [
    "any_concat",
    "any_stack",
    "any_to_torch_tensor",
    "any_to_numpy",
    "any_to_datadict",
    "stack_sequence_fields",
    "get_batch_size",
    "any_slice",
    "any_transpose_first_two_axes",
    "add_batch_dim",
]


# ------------------------------------------------------------------------------------------------------------
# structure walking
# ------------------------------------------------------------------------------------------------------------
def _is_map(x) -> bool:
    return isinstance(x, cabc.Mapping)


def _is_seq(x) -> bool:
    return isinstance(x, (list, tuple))


def _is_nest(x) -> bool:
    return _is_map(x) or _is_seq(x)


def _kids(x):
    if _is_map(x):
        return [(k, x[k]) for k in sorted(x.keys())]
    return list(enumerate(x))


def _rebuild(proto, values):
    if _is_map(proto):
        d = dict(zip(sorted(proto.keys()), values))
        return DataDict(d) if isinstance(proto, DataDict) else type(proto)(d)
    if isinstance(proto, tuple) and hasattr(proto, "_fields"):
        return type(proto)(*values)
    return type(proto)(values)


def map_structure_with_path(fn: Callable, *structs):
    def rec(path, *xs):
        if not _is_nest(xs[0]):
            return fn(path, *xs)
        return _rebuild(xs[0], [rec(path + (k,), *[x[k] for x in xs]) for k, _ in _kids(xs[0])])

    return rec((), *structs)


def map_structure(fn: Callable, *structs):
    return map_structure_with_path(lambda _p, *xs: fn(*xs), *structs)


def flatten(s):
    if not _is_nest(s):
        return [s]
    out = []
    for _, c in _kids(s):
        out.extend(flatten(c))
    return out


def unflatten_as(proto, flat):
    it = iter(list(flat))

    def rec(p):
        if not _is_nest(p):
            return next(it)
        return _rebuild(p, [rec(c) for _, c in _kids(p)])

    return rec(proto)


# ------------------------------------------------------------------------------------------------------------
# DataDict
# ------------------------------------------------------------------------------------------------------------
def _wrap(v):
    if isinstance(v, DataDict):
        return v
    if _is_map(v):
        return DataDict(v)
    if isinstance(v, list):
        return [_wrap(e) for e in v]
    if isinstance(v, tuple) and not hasattr(v, "_fields"):
        return tuple(_wrap(e) for e in v)
    return v


class DataDict(cabc.MutableMapping):
    """Nested dict with attribute / dotted-key access and whole-structure array slicing (reference :228-647)."""

    def __init__(self, _data_=None, **kwargs):
        object.__setattr__(self, "_data_", {})
        if _data_ is not None:
            assert not kwargs, "DataDict takes either one mapping / iterable of pairs or **kwargs, not both"
            items = _data_.items() if _is_map(_data_) else dict(_data_).items()
        else:
            items = kwargs.items()
        for k, v in items:
            self[k] = v

    # attribute access -----------------------------------------------------------------------------
    def __getattr__(self, name):
        data = object.__getattribute__(self, "_data_")
        if name not in data:
            raise AttributeError(f'Missing key-attribute "{name}"')
        return data[name]

    def __setattr__(self, name, value):
        self[name] = value

    # mapping protocol -----------------------------------------------------------------------------
    def __getitem__(self, key):
        if isinstance(key, str):
            if "." in key:
                head, rest = key.split(".", 1)
                if head not in self._data_:
                    raise KeyError(f'Missing parent key "{head}" in "{key}"')
                return self._data_[head][rest]
            if key not in self._data_:
                raise KeyError(f'Missing key "{key}"')
            return self._data_[key]
        return map_structure(lambda x: x[key] if isinstance(x, (np.ndarray, torch.Tensor, list, tuple)) else x, self)

    def __setitem__(self, key, value):
        if isinstance(key, str):
            if "." in key:
                head, rest = key.split(".", 1)
                if head not in self._data_:
                    self._data_[head] = DataDict()
                self._data_[head][rest] = value
            else:
                self._data_[key] = _wrap(value)
            return
        value = _wrap(value)

        def assign(path, ours):
            v = value
            if _is_map(value):
                for k in path:
                    v = v[k]
            ours[key] = v
            return ours

        map_structure_with_path(assign, self)

    def __delitem__(self, key):
        if isinstance(key, str) and "." in key:
            head, rest = key.split(".", 1)
            del self._data_[head][rest]
        else:
            del self._data_[key]

    def __iter__(self):
        return iter(self._data_)

    def __len__(self):
        return len(self._data_)

    def __contains__(self, key):
        if isinstance(key, str) and "." in key:
            head, rest = key.split(".", 1)
            return head in self._data_ and rest in self._data_[head]
        return key in self._data_

    def __repr__(self):
        return f"DataDict({self.to_container()!r})"

    def keys(self):
        return self._data_.keys()

    def values(self):
        return self._data_.values()

    def items(self):
        return self._data_.items()

    def to_container(self):
        def rec(x):
            if isinstance(x, DataDict):
                return {k: rec(v) for k, v in x._data_.items()}
            if isinstance(x, list):
                return [rec(e) for e in x]
            if isinstance(x, tuple) and not hasattr(x, "_fields"):
                return tuple(rec(e) for e in x)
            return x

        return rec(self)

    def copy(self):
        return DataDict(self._data_)

    # structure ops --------------------------------------------------------------------------------
    def map_structure(self, func: Callable, *other, with_path: bool = False, inplace: bool = False):
        mapper = map_structure_with_path if with_path else map_structure
        out = mapper(func, self, *other)
        if inplace:
            object.__setattr__(self, "_data_", out._data_)
            return self
        return out

    def to_torch_tensor(self, dtype=None, device=None, copy=False, non_blocking=False, inplace: bool = False, **_):
        return self.map_structure(
            lambda x: any_to_torch_tensor(x, dtype=dtype, device=device, copy=copy, non_blocking=non_blocking), inplace=inplace
        )

    def to_numpy(self, dtype=None, copy=False, non_blocking=False, inplace: bool = False, **_):
        return self.map_structure(lambda x: any_to_numpy(x, dtype=dtype, copy=copy), inplace=inplace)


# ------------------------------------------------------------------------------------------------------------
# the ten public helpers
# ------------------------------------------------------------------------------------------------------------
def _join(fn_np, fn_t, xs, dim):
    def leaf(*vals):
        v0 = vals[0]
        if isinstance(v0, np.ndarray):
            return fn_np(vals, axis=dim)
        if torch.is_tensor(v0):
            return fn_t(vals, dim=dim)
        if isinstance(v0, float):
            return np.array(vals, dtype=np.float32)
        return np.array(vals)

    return map_structure(leaf, *xs)


def any_concat(xs: list, *, dim: int = 0):
    return _join(np.concatenate, torch.cat, xs, dim)


def any_stack(xs: list, *, dim: int = 0):
    return _join(np.stack, torch.stack, xs, dim)


def _torch_dtype(d):
    if d is None or isinstance(d, torch.dtype):
        return d
    if isinstance(d, str):
        return getattr(torch, d)
    return torch.from_numpy(np.zeros((), dtype=d)).dtype


def any_to_torch_tensor(x, dtype=None, device=None, copy=False, non_blocking=False, smart_optimize: bool = True):
    dtype = _torch_dtype(dtype)
    if isinstance(device, int):
        device = torch.device("cuda", device)
    elif isinstance(device, str):
        device = torch.device(device)
    if not isinstance(x, (torch.Tensor, np.ndarray)):
        x = torch.tensor(x, dtype=dtype)
        copy = False
    x = torch.as_tensor(x)
    dtype = dtype or x.dtype
    device = device or x.device
    if x.dtype == dtype and x.device == device:
        return x.clone() if copy else x
    # move the narrower representation across the bus (reference :96-122)
    if x.element_size() > torch.empty((), dtype=dtype).element_size():
        return x.to(dtype=dtype).to(device=device, non_blocking=non_blocking)
    return x.to(device=device, non_blocking=non_blocking).to(dtype=dtype)


def any_to_numpy(x, dtype=None, copy: bool = False, non_blocking: bool = False, smart_optimize: bool = True):
    if torch.is_tensor(x):
        x = x.detach().cpu().numpy()
    elif not isinstance(x, np.ndarray):
        x = np.array(x)
    if dtype is not None:
        x = x.astype(dtype, copy=copy)
    elif copy:
        x = x.copy()
    return x


def any_to_datadict(x) -> DataDict:
    if isinstance(x, DataDict):
        return x
    if _is_map(x):
        return DataDict(x)
    raise NotImplementedError(f"cannot convert {type(x)} to DataDict")


def stack_sequence_fields(sequence):
    if not sequence:
        raise ValueError("Input sequence must not be empty")
    flats = [flatten(s) for s in sequence]

    def join(vals):
        try:
            return np.stack(vals)
        except ValueError:
            return np.asarray(vals)

    return unflatten_as(sequence[-1], [join(v) for v in zip(*flats)])


def get_batch_size(x, strict: bool = False) -> int:
    def one(v):
        if isinstance(v, np.ndarray):
            return v.shape[0]
        if torch.is_tensor(v):
            return v.size(0)
        return len(v)

    xs = flatten(x)
    if strict:
        sizes = [one(v) for v in xs]
        assert all(s == sizes[0] for s in sizes), f"batch sizes must all be the same in nested structure: {sizes}"
        return sizes[0]
    return one(xs[0])


def any_slice(x, slice):
    return map_structure(lambda v: v[slice] if isinstance(v, (np.ndarray, torch.Tensor)) else v, x)


def any_transpose_first_two_axes(x):
    def f(v):
        if isinstance(v, np.ndarray):
            return np.swapaxes(v, 0, 1)
        if torch.is_tensor(v):
            return torch.swapaxes(v, 0, 1)
        raise ValueError(f"Input ({type(v)}) must be either a numpy array or a tensor.")

    return map_structure(f, x)


def add_batch_dim(x):
    def f(v):
        if isinstance(v, np.ndarray):
            return np.expand_dims(v, axis=0)
        if torch.is_tensor(v):
            return v.unsqueeze(0)
        raise NotImplementedError(f"Unsupported data structure: {type(v)}")

    return map_structure(f, x)
