"""
TEST INFRASTRUCTURE ONLY -- CPU restatement of ``tad_multicharge.get_eeq_charges``
(``tad-multicharge==0.5.0``, ``/root/reference/setup.cfg:37``), the third-party
step immediately before the D4 hot path.  The dependency is absent from
``/root/reference`` and from this image; call sites in the reference:
``dispersion/base.py:401-407`` and ``disp.py:190``
(``get_eeq_charges(numbers, positions, charge, cutoff=cutoff.cn_eeq)``).

Published algorithm (EEQ-2019, Caldeweyher et al., J. Chem. Phys. 150, 154122):

* ``cn_i = cut(sum_{j != i, r_ij <= cutoff} 1/2 (1 + erf(-7.5 (r_ij/(rcov_i+rcov_j) - 1))))`` with
  ``rcov = COV_D3`` and ``cut(cn) = log(1+e^cn_max) - log(1+e^(cn_max-cn))``, ``cn_max = 8``
  (``defaults.py:29-33`` of the reference name the same cutoff 25 and cn_max 8);
* ``A_ij = erf(r_ij / sqrt(a_i^2 + a_j^2)) / r_ij``, ``A_ii = eta_i + sqrt(2/pi)/a_i``,
  ``x_i = -chi_i + kappa_i sqrt(cn_i)``; solve ``[[A, 1], [1^T, 0]] (q, lambda) = (x, Q)``
  (padding rows get a unit diagonal and a zero constraint entry, so their charge is 0).

Dense, float64-capable, differentiable by ``torch.autograd`` (forces of the
reference run through the EEQ charges when ``q`` is not given).

Pinning: with this restatement plugged into the UNMODIFIED reference
(``oracle/mctc_shim/tad_multicharge``) the reference's own full-path known-answer
vectors are reproduced -- SiH4 ``s10`` golden to 3e-16, LiH / SiH4 autograd
gradients to 2e-15, README / ``examples/single.py`` energies to 5e-10 (their
print precision), formamide-dimer doctest to its 10 digits, EEQ charges of the
test samples to 2e-7 (Fortran values) -- see ``tests/test_oracle_kat.py``.
That pins the element parameters of H, Li, C, N, O, Si, S; the others are
restated from the publication and unpinned upstream-wise.
"""

from __future__ import annotations

import importlib.util
import math
from functools import lru_cache

import torch

import d4_oracle as orc

Tensor = torch.Tensor

EEQ_CN_CUTOFF = 25.0  # defaults.py:29 (D4_CN_EEQ_CUTOFF)
EEQ_CN_MAX = 8.0  # defaults.py:32 (D4_CN_EEQ_MAX)
EEQ_KCN = 7.5


@lru_cache(maxsize=None)
def _param() -> dict[str, Tensor]:
    spec = importlib.util.spec_from_file_location("_eeq2019", orc._DATA / "eeq2019.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # data only, shared with the product like the other element tables
    return {
        "chi": torch.tensor(mod.EEQ_CHI, dtype=torch.float64),
        "eta": torch.tensor(mod.EEQ_ETA, dtype=torch.float64),
        "kcn": torch.tensor(mod.EEQ_KCN, dtype=torch.float64),
        "rad": torch.tensor(mod.EEQ_RAD, dtype=torch.float64),
    }


def cn_eeq(numbers: Tensor, positions: Tensor, cutoff: float = EEQ_CN_CUTOFF, cn_max: float | None = EEQ_CN_MAX):
    """tad_mctc.ncoord.cn_eeq: erf counting, cut smoothly at ``cn_max``."""
    dtype = positions.dtype
    d, mask = orc.masked_distances(numbers, positions)
    rc = orc._tables()["rcov"].to(dtype)[numbers]
    r0 = rc.unsqueeze(-1) + rc.unsqueeze(-2)
    count = 0.5 * (1.0 + torch.erf(-EEQ_KCN * (d / r0 - 1.0)))
    cn = torch.where(mask & (d <= cutoff), count, torch.zeros((), dtype=dtype)).sum(-1)
    if cn_max is None:
        return cn
    cmax = torch.tensor(float(cn_max), dtype=dtype)
    return torch.log(1.0 + torch.exp(cmax)) - torch.log(1.0 + torch.exp(cmax - cn))


def get_eeq_charges(numbers: Tensor, positions: Tensor, chrg, *, cutoff=None, return_energy: bool = False, **_):
    """tad_multicharge.get_eeq_charges (EEQModel.param2019 + solve)."""
    dtype = positions.dtype
    if int(numbers.max()) > 86:
        raise ValueError("EEQ-2019 is parameterised for Z <= 86")
    p = _param()
    chi, eta, kappa, rad = (p[k].to(dtype)[numbers] for k in ("chi", "eta", "kcn", "rad"))
    chrg = torch.as_tensor(chrg, dtype=dtype)
    cut = EEQ_CN_CUTOFF if cutoff is None else float(cutoff)

    cn = cn_eeq(numbers, positions, cut)
    eps = torch.finfo(dtype).eps
    zero = torch.zeros((), dtype=dtype)
    one = torch.ones((), dtype=dtype)
    real = numbers != 0
    d, mask = orc.masked_distances(numbers, positions)

    cc = torch.where(real, -chi + torch.sqrt(torch.clamp(cn, min=eps)) * kappa, zero)
    rhs = torch.cat((cc, chrg.expand(numbers.shape[:-1]).unsqueeze(-1)), dim=-1)

    rads = torch.where(mask, rad.unsqueeze(-1) ** 2 + rad.unsqueeze(-2) ** 2, one)
    gamma = torch.where(mask, 1.0 / torch.sqrt(rads), zero)
    diag = torch.where(real, eta + math.sqrt(2.0 / math.pi) / torch.where(real, rad, one), one)
    coulomb = torch.where(mask, torch.erf(d * gamma) / d, zero) + torch.diag_embed(diag)

    constraint = real.to(dtype)
    top = torch.cat((coulomb, constraint.unsqueeze(-1)), dim=-1)
    bottom = torch.cat((constraint, torch.zeros_like(constraint[..., :1])), dim=-1).unsqueeze(-2)
    matrix = torch.cat((top, bottom), dim=-2)

    x = torch.linalg.solve(matrix, rhs)
    if not return_energy:
        return x[..., :-1]
    e = x * (0.5 * torch.einsum("...ij,...j->...i", matrix, x) - rhs)
    return e[..., :-1], x[..., :-1]
