"""``einsum`` wrapper (upstream: opt_einsum if installed, else torch.einsum)."""
import torch

__all__ = ["einsum"]


def einsum(*args, optimize=None):
    return torch.einsum(*args)
