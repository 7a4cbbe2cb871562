"""Minimal stand-in for the `tensordict` package (generator-only; see README.md here).

Container semantics only, as used by the reference at agents/agent.py:63-73,
:106-111, :153, :162, :328-331 — a flat {dotted-name: tensor} mapping that can be
lifted from / re-attached to an tnn.Module, stacked over modules, and is a pytree so
``torch.vmap`` can batch over it.
"""
from __future__ import annotations

import torch
from torch import nn as tnn
from torch.utils import _pytree as pytree


def _named_tensors(module: tnn.Module):
    for n, p in module.named_parameters():
        yield n, p
    for n, b in module.named_buffers():
        yield n, b


def _owner(module: tnn.Module, dotted: str):
    *path, leaf = dotted.split(".")
    for part in path:
        module = getattr(module, part)
    return module, leaf


class _Swap:
    """Installs the tensors now; restores the previous ones on ``__exit__`` if used as a context."""

    def __init__(self, td, module):
        self._saved = []
        for name, t in td._d.items():
            owner, leaf = _owner(module, name)
            slot = owner._parameters if leaf in owner._parameters else owner._buffers
            self._saved.append((slot, leaf, slot[leaf]))
            slot[leaf] = t

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        for slot, leaf, old in self._saved:
            slot[leaf] = old
        return False


class TensorDict:
    def __init__(self, source=None, batch_size=None, device=None, **_):
        self._d = dict(source or {})

    # -- construction -----------------------------------------------------
    @classmethod
    def from_module(cls, module: tnn.Module, as_module: bool = False):
        return cls({n: t for n, t in _named_tensors(module)})

    @classmethod
    def from_modules(cls, *modules: tnn.Module, as_module: bool = False):
        keys = [n for n, _ in _named_tensors(modules[0])]
        out = {}
        for k in keys:
            ts = [dict(_named_tensors(m))[k] for m in modules]
            stacked = torch.stack([t.detach() for t in ts])
            if isinstance(ts[0], tnn.Parameter):
                stacked = tnn.Parameter(stacked, requires_grad=ts[0].requires_grad)
            out[k] = stacked
        return cls(out)

    # -- views ---------------------------------------------------------------
    @property
    def data(self):
        return TensorDict({k: v.data for k, v in self._d.items()})

    def clone(self):
        return TensorDict({k: v.clone() for k, v in self._d.items()})

    def detach(self):
        return TensorDict({k: v.detach() for k, v in self._d.items()})

    def to_module(self, module: tnn.Module):
        return _Swap(self, module)

    def lerp_(self, end: "TensorDict", weight: float):
        torch._foreach_lerp_(list(self._d.values()), [end._d[k] for k in self._d], weight)
        return self

    # -- mapping ---------------------------------------------------------------
    def __getitem__(self, k):
        return self._d[k]

    def __setitem__(self, k, v):
        self._d[k] = v

    def keys(self):
        return self._d.keys()

    def values(self):
        return self._d.values()

    def items(self):
        return self._d.items()

    def update(self, other):
        self._d.update(other._d if isinstance(other, TensorDict) else other)
        return self

    def to_dict(self):
        return dict(self._d)

    def clear(self):
        self._d.clear()


pytree.register_pytree_node(
    TensorDict,
    lambda td: (list(td._d.values()), list(td._d.keys())),
    lambda values, keys: TensorDict(dict(zip(keys, values))),
)
