"""DFT-D4 coordination number on device (``tad_mctc.ncoord.cn_d4`` with ``erf_count``;
call site ``/root/reference/src/tad_dftd4/dispersion/base.py:390``)."""

from __future__ import annotations

import torch

__all__ = ["cn_d4", "erf_count"]


def erf_count(r, r0, kcn: float = 7.5):
    """``0.5 * (1 + erf(-kcn (r/r0 - 1)))`` -- the counting function fused into the kernels."""
    return 0.5 * (1.0 + torch.special.erf(-kcn * (r / r0 - 1.0)))


def cn_d4(numbers: torch.Tensor, positions: torch.Tensor, **kwargs) -> torch.Tensor:
    """EN-weighted erf coordination numbers, cutoff 30 Bohr (CUDA kernel)."""
    if kwargs.get("counting_function", erf_count) is not erf_count or any(
        kwargs.get(k) is not None for k in ("rcov", "en", "cutoff")
    ):
        raise NotImplementedError("only the default cn_d4 (erf_count, default radii) is accelerated")
    from .disp import get_properties

    # plain values (not on the autograd tape): the fused kernels differentiate the coordination number
    # themselves; a term that receives this tensor through Disp.calculate ignores it (dispersion.py)
    q = torch.zeros(numbers.shape, dtype=positions.dtype, device=positions.device)
    return get_properties(numbers, positions.detach(), q=q)[0]
