"""Safe torch ops (restated from upstream semantics)."""
from __future__ import annotations

import torch

__all__ = ["cdist", "sqrt", "divide"]


def sqrt(x, *, eps=None):
    e = torch.finfo(x.dtype).eps if eps is None else eps
    return torch.sqrt(torch.clamp(x, min=e))


def divide(x, y, *, eps=None):
    e = torch.finfo(y.dtype).eps if eps is None else eps
    y_safe = torch.where(y == 0, torch.full_like(y, e), y)
    return torch.divide(x, y_safe)


def cdist(x, y=None, p=2):
    if y is None:
        y = x
    if p != 2:
        diff = torch.abs(x.unsqueeze(-2) - y.unsqueeze(-3))
        d = torch.sum(torch.pow(diff, p), -1)
        return torch.pow(torch.clamp(d, min=torch.finfo(x.dtype).eps), 1.0 / p)
    # quadratic expansion |x|^2 + |y|^2 - 2 x.y, clamped before the root
    xnorm = torch.einsum("...ij,...ij->...i", x, x)
    ynorm = torch.einsum("...ij,...ij->...i", y, y)
    n = xnorm.unsqueeze(-1) + ynorm.unsqueeze(-2)
    prod = torch.einsum("...ik,...jk->...ij", x, y)
    return sqrt(n - 2.0 * prod)
