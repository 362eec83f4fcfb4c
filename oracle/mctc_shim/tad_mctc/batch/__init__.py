"""Padding masks and ``pack`` (restated)."""
from __future__ import annotations

import torch

from . import mask  # noqa: F401
from .mask import real_atoms, real_pairs, real_triples

__all__ = ["real_atoms", "real_pairs", "real_triples", "pack"]


def pack(tensors, axis=0, value=0, size=None):
    tensors = list(tensors)
    if size is None:
        size = tuple(max(t.shape[d] for t in tensors) for d in range(tensors[0].dim()))
    out = torch.full(
        (len(tensors), *size), value, dtype=tensors[0].dtype, device=tensors[0].device
    )
    for n, t in enumerate(tensors):
        out[(n, *[slice(0, s) for s in t.shape])] = t
    if axis != 0:
        out = out.movedim(0, axis)
    return out
