"""DFT-D4 coordination number (restated)."""
from __future__ import annotations

import math

import torch

from .. import storch
from ..batch import real_pairs
from ..data import COV_D3, PAULING

__all__ = ["cn_d4", "erf_count", "cn_d3"]

# same numbers as /root/reference/src/tad_dftd4/defaults.py:26,41-51
D4_CN_CUTOFF = 30.0
D4_KCN = 7.5
D4_K4 = 4.10451
D4_K5 = 19.08857
D4_K6 = 2 * 11.28174**2


def erf_count(r, r0, kcn=D4_KCN):
    return 0.5 * (1.0 + torch.special.erf(-kcn * (r / r0 - 1.0)))


def cn_d4(
    numbers,
    positions,
    *,
    counting_function=erf_count,
    rcov=None,
    en=None,
    cutoff=None,
    kcn=D4_KCN,
    **kwargs,
):
    dd = {"device": positions.device, "dtype": positions.dtype}
    if cutoff is None:
        cutoff = torch.tensor(D4_CN_CUTOFF, **dd)
    if rcov is None:
        rcov = COV_D3(**dd)[numbers]
    if en is None:
        en = PAULING(**dd)[numbers]
    if numbers.shape != rcov.shape:
        raise ValueError("Shape of covalent radii is not consistent with numbers.")
    if numbers.shape != positions.shape[:-1]:
        raise ValueError("Shape of positions is not consistent with numbers.")

    mask = real_pairs(numbers, mask_diagonal=True)
    distances = torch.where(
        mask,
        storch.cdist(positions, positions, p=2),
        torch.tensor(torch.finfo(positions.dtype).eps, **dd),
    )
    endiff = torch.abs(en.unsqueeze(-2) - en.unsqueeze(-1))
    den = D4_K4 * torch.exp(-((endiff + D4_K5) ** 2.0) / D4_K6)
    rc = rcov.unsqueeze(-2) + rcov.unsqueeze(-1)
    cf = torch.where(
        mask * (distances <= cutoff),
        den * counting_function(distances, rc, kcn, **kwargs),
        torch.tensor(0.0, **dd),
    )
    return torch.sum(cf, dim=-1)


def cn_d3(*args, **kwargs):  # pragma: no cover - D3 is out of scope
    raise NotImplementedError("cn_d3 is not part of the D4 hot path shim")
