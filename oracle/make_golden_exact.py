#!/usr/bin/env python
"""
TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/exact/*.npz``: energies and autograd gradients of the
UNMODIFIED reference's ``DispD4Exact`` (``/root/reference/src/tad_dftd4/dispersion/d4.py:67-84``: rational
two-body term + ATM term with the exact Casimir-Polder C9, ``threebody.py:276-302``) on top of
``oracle/mctc_shim``, for inputs taken from the committed fixtures ``tests/golden/*.npz``, plus the ATM term
alone (``D4ATMExact`` registered on a bare ``Disp``).  In the same run ``oracle/d4_oracle.py`` (``c9="exact"``)
must reproduce every stored number.

Run in the build container only:  ``python oracle/make_golden_exact.py``
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "mctc_shim"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, str(HERE))

from tad_dftd4.damping import ZeroDamping  # noqa: E402  (the real reference)
from tad_dftd4.dispersion.base import Disp  # noqa: E402
from tad_dftd4.dispersion.d4 import D4ATMExact, DispD4Exact  # noqa: E402

import d4_oracle as orc  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"
OUT = GOLDEN / "exact"
F64 = torch.float64
CASES = ("single_pbe0", "sih4_tpssh", "organic_33", "ragged_batch", "tight_cutoffs", "all_elements")


def main() -> None:
    OUT.mkdir(parents=True, exist_ok=True)
    for name in CASES:
        raw = np.load(GOLDEN / f"{name}.npz")
        n, p, q = (torch.from_numpy(raw[k]) for k in ("numbers", "positions", "q"))
        param = {str(k): float(v) for k, v in zip(raw["param_keys"], raw["param_vals"])}
        par = {k: torch.tensor(v, dtype=F64) for k, v in param.items()}
        kw, okw = {}, {}
        okw = {str(k): float(v) for k, v in zip(raw["cutoff_keys"], raw["cutoff_vals"])}
        if okw:
            from tad_dftd4.cutoff import Cutoff

            kw["cutoff"] = Cutoff(dtype=F64, **okw)
        charge = torch.zeros(n.shape[:-1], dtype=F64)
        pos = p.clone().requires_grad_(True)
        e = DispD4Exact(dtype=F64).calculate(n, pos, charge, par, q=q, **kw)
        (g,) = torch.autograd.grad(e.sum(), pos)
        eo, go = orc.energy_and_gradient(n, p, param, q, c9="exact", **okw)
        de = ((e.detach() - eo).abs().max() / e.detach().abs().max()).item()
        dg = (g - go).abs().max().item()
        # the ATM term alone
        atm = Disp(model="d4", dtype=F64)
        atm.register(D4ATMExact(damping_fn=ZeroDamping(), charge_dependent=False))
        pos3 = p.clone().requires_grad_(True)
        e3 = atm.calculate(n, pos3, charge, par, **kw)
        (g3,) = torch.autograd.grad(e3.sum(), pos3)
        _, e3o, *_ = orc.dftd4(n, p, param, q, parts=True, c9="exact", **okw)
        d3 = ((e3.detach() - e3o).abs().max() / e3.detach().abs().max().clamp(min=1e-300)).item()
        e_apx = orc.dftd4(n, p, param, q, **okw)
        print(f"{name:14s} sum E = {e.sum().item():+.12e}  oracle-ref dE {de:.1e} dG {dg:.1e} dE3 {d3:.1e}   "
              f"(exact vs approximate C9: {((e.detach() - e_apx).abs().max() / e_apx.abs().max()).item():.2e})")
        assert de < 1e-12 and dg < 1e-14 and d3 < 1e-12
        np.savez_compressed(OUT / f"{name}.npz", energy=e.detach().numpy(), gradient=g.numpy(),
                            energy_atm=e3.detach().numpy(), gradient_atm=g3.numpy())


if __name__ == "__main__":
    main()
