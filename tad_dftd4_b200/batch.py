"""``pack``: right-pad tensors with zeros and stack them (the padding
convention of the reference's inputs, ``tad_mctc.batch.pack``)."""

from __future__ import annotations

import torch

__all__ = ["pack"]


def pack(tensors, value=0):
    tensors = list(tensors)
    size = tuple(max(t.shape[d] for t in tensors) for d in range(tensors[0].dim()))
    out = torch.full((len(tensors), *size), value, dtype=tensors[0].dtype, device=tensors[0].device)
    for n, t in enumerate(tensors):
        out[(n, *[slice(0, s) for s in t.shape])] = t
    return out
