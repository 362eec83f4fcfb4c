#!/usr/bin/env python
"""
TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/hessian/*.npz``: nuclear Hessians
``d^2 (sum_i E_i) / d positions^2`` from the UNMODIFIED reference (``/root/reference/src/tad_dftd4`` on
top of ``oracle/mctc_shim``) by double backward through its dense tape -- what
``test/test_grad/test_hessian.py:75-112`` does with TPSS0-D4-ATM parameters -- for inputs taken from
the committed fixtures ``tests/golden/*.npz``:

* ``hess_q``     explicit (constant) charges of the fixture, D4 and D4S;
* ``hess_eeq``   the reference's default ``q=None`` path (EEQ charges on the tape, through
                 ``oracle/mctc_shim/tad_multicharge``), D4.

Run in the build container only:  ``python oracle/make_golden_hessian.py``
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

GOLDEN = HERE.parent / "tests" / "golden"
OUT = GOLDEN / "hessian"
F64 = torch.float64
# test/test_grad/test_hessian.py:84-90 (TPSS0-D4-ATM)
TPSS0 = dict(s6=1.0, s8=1.62438102, s9=1.0, a1=0.40329022, a2=4.80537871)
CASES = ("lih_tpssh", "sih4_tpssh", "organic_5", "organic_20")


def hessian(numbers, positions, q, model):
    pos = positions.clone().requires_grad_(True)
    param = ref.damping.Param(**{k: torch.tensor(v, dtype=F64) for k, v in TPSS0.items()})
    e = ref.dftd4(numbers, pos, torch.tensor(0.0, dtype=F64), param, q=q, model=model).sum()
    (g,) = torch.autograd.grad(e, pos, create_graph=True)
    rows = [torch.autograd.grad(gi, pos, retain_graph=True)[0] for gi in g.reshape(-1)]
    return torch.stack(rows).reshape(*positions.shape, *positions.shape).numpy()


def main() -> None:
    OUT.mkdir(parents=True, exist_ok=True)
    for name in CASES:
        raw = np.load(GOLDEN / f"{name}.npz")
        n, p, q = (torch.from_numpy(raw[k]) for k in ("numbers", "positions", "q"))
        store = {"param_keys": np.array(list(TPSS0)), "param_vals": np.array(list(TPSS0.values()))}
        store["hess_q_d4"] = hessian(n, p, q, "d4")
        store["hess_q_d4s"] = hessian(n, p, q, "d4s")
        store["hess_eeq_d4"] = hessian(n, p, None, "d4")
        h = store["hess_eeq_d4"].reshape(p.numel(), p.numel())
        print(f"{name:12s} |H|max {np.abs(h).max():.3e}  asym {np.abs(h - h.T).max():.1e}  "
              f"translation rows {np.abs(store['hess_eeq_d4'].sum(2)).max():.1e}")
        np.savez_compressed(OUT / f"{name}.npz", **store)


if __name__ == "__main__":
    main()
