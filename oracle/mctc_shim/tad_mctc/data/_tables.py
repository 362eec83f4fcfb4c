"""Loads the element lists from ``tad_dftd4_b200/data/elements.py`` BY PATH (no
package import, so the oracle stays independent of the product code)."""
from __future__ import annotations

import importlib.util
from pathlib import Path

import torch

_EL = Path(__file__).resolve().parents[4] / "tad_dftd4_b200" / "data" / "elements.py"
_spec = importlib.util.spec_from_file_location("_d4b200_elements", _EL)
_el = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_el)


def PAULING(device=None, dtype=None):
    return torch.tensor(_el.PAULING, device=device, dtype=dtype)


def GAM(device=None, dtype=None):
    return torch.tensor(_el.GAM, device=device, dtype=dtype)


def ZEFF(device=None, dtype=None):
    # integer table upstream (indexed, then added to float charges)
    return torch.tensor(_el.ZEFF, device=device)


def COV_D3(device=None, dtype=None):
    # 4/3 * r_cov(Angstrom) * AA2AU, evaluated in float64 then cast
    t = torch.tensor(_el.COV_2009, dtype=torch.float64) * _el.AA2AU * 4.0 / 3.0
    return t.to(device=device, dtype=dtype if dtype is not None else torch.get_default_dtype())


def VDW_PAIRWISE(device=None, dtype=None):
    # Only gathered and shape-checked by the D4/BJ path
    # (/root/reference/src/tad_dftd4/dispersion/base.py:381-387); values unused.
    n = _el.MAX_ELEMENT
    return torch.zeros((n, n), device=device, dtype=dtype)
