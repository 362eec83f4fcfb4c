#!/usr/bin/env python
"""
TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/gfn2/*.npz``: energies and autograd gradients
of the UNMODIFIED reference (``/root/reference/src/tad_dftd4`` on top of ``oracle/mctc_shim``) with the
GFN2-xTB reference charges, ``D4Model(numbers, ref_charges="gfn2")`` / ``D4SModel(...)``
(model/base.py:388-399, model/d4.py:142-149), for inputs taken from the committed fixtures
``tests/golden/*.npz``.  In the same run ``oracle/d4_oracle.py`` must reproduce every stored number.

Run in the build container only:  ``python oracle/make_golden_gfn2.py``
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

import tad_dftd4 as ref  # noqa: E402  (the real reference)
from tad_dftd4.model.d4 import D4Model  # noqa: E402
from tad_dftd4.model.d4s import D4SModel  # noqa: E402

import d4_oracle as orc  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"
OUT = GOLDEN / "gfn2"
F64 = torch.float64
CASES = ("single_pbe0", "sih4_tpssh", "organic_33", "ragged_batch", "all_elements")


def main() -> None:
    OUT.mkdir(parents=True, exist_ok=True)
    for name in CASES:
        raw = np.load(GOLDEN / f"{name}.npz")
        n, p, q = (torch.from_numpy(raw[k]) for k in ("numbers", "positions", "q"))
        if name == "all_elements":  # the GFN2 reference charges are tabulated up to Rn: atoms beyond become padding
            keep = n <= 86
            n, p, q = n * keep, p * keep.unsqueeze(-1), q * keep
        param = {str(k): float(v) for k, v in zip(raw["param_keys"], raw["param_vals"])}
        store = {}
        for key, cls in (("d4", D4Model), ("d4s", D4SModel)):
            pos = p.clone().requires_grad_(True)
            par = {k: torch.tensor(v, dtype=F64) for k, v in param.items()}
            model = cls(n, ref_charges="gfn2", dtype=F64)
            e = ref.dftd4(n, pos, torch.zeros(n.shape[:-1], dtype=F64), par, q=q, model=model)
            (g,) = torch.autograd.grad(e.sum(), pos)
            eo, go = orc.energy_and_gradient(n, p, param, q, model=key, ref_charges="gfn2")
            de = ((e.detach() - eo).abs().max() / e.detach().abs().max()).item()
            dg = (g - go).abs().max().item()
            e_eeq = orc.dftd4(n, p, param, q, model=key)
            print(f"{name:14s} {key:4s} sum E = {e.sum().item():+.12e}  oracle-ref dE {de:.1e} dG {dg:.1e}   "
                  f"(gfn2 vs eeq: {((e.detach() - e_eeq).abs().max() / e_eeq.abs().max()).item():.2e})")
            assert de < 1e-13 and dg < 1e-15
            store[f"energy_{key}"] = e.detach().numpy()
            store[f"gradient_{key}"] = g.numpy()
        np.savez_compressed(OUT / f"{name}.npz", **store)


if __name__ == "__main__":
    main()
