"""Masks for padded batches (``numbers == 0`` is padding)."""
import torch

__all__ = ["real_atoms", "real_pairs", "real_triples"]


def real_atoms(numbers):
    return numbers != 0


def real_pairs(numbers, mask_diagonal=True):
    real = real_atoms(numbers)
    mask = real.unsqueeze(-2) * real.unsqueeze(-1)
    if mask_diagonal:
        mask = mask * ~torch.diag_embed(torch.ones_like(real))
    return mask


def real_triples(numbers, mask_diagonal=True, mask_self=True):
    real = real_pairs(numbers, mask_diagonal=False)
    mask = real.unsqueeze(-3) * real.unsqueeze(-2) * real.unsqueeze(-1)
    if mask_diagonal:
        mask = mask * ~torch.diag_embed(torch.ones_like(real))
    if mask_self:
        mask = mask * ~torch.diag_embed(torch.ones_like(real), offset=0, dim1=-3, dim2=-2)
        mask = mask * ~torch.diag_embed(torch.ones_like(real), offset=0, dim1=-3, dim2=-1)
    return mask
