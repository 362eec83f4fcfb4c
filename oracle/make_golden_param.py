#!/usr/bin/env python
"""
TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/param/*.npz``: gradients of
``sum_i g_i E_i`` with respect to the damping parameters (s6, s8, s9, s10, a1, a2, alp) from the
UNMODIFIED reference (``/root/reference/src/tad_dftd4`` on top of ``oracle/mctc_shim``; what
``test/test_grad/test_param.py:40-100`` differentiates), for inputs taken from the committed
fixtures ``tests/golden/*.npz``, D4 and D4S, float64.  In the same run the self-contained
restatement ``oracle/d4_oracle.py`` must reproduce every stored number to 1e-11 relative.

Run in the build container only:  ``python oracle/make_golden_param.py``
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

import d4_oracle as orc  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"
OUT = GOLDEN / "param"
F64 = torch.float64
KEYS = ("s6", "s8", "s9", "s10", "a1", "a2", "alp")
# test/test_grad/test_param.py:54-62 (TPSS0-D4-ATM, s10 = 0.0 present)
TPSS0 = dict(s6=1.0, s8=0.78981345, s9=1.0, s10=0.0, a1=0.49484001, a2=5.73083694, alp=16.0)
CASES = {
    "lih_tpssh": TPSS0,
    "sih4_tpssh": TPSS0,
    "nan17": TPSS0,
    "organic_20": TPSS0,
    "ragged_batch": TPSS0,
    "holes": TPSS0,
    "tight_cutoffs": dict(TPSS0, s9=1.3, alp=14.0, s10=0.4),
}


def load(name):
    raw = np.load(GOLDEN / f"{name}.npz")
    cut = {str(k): float(v) for k, v in zip(raw["cutoff_keys"], raw["cutoff_vals"])}
    return raw["numbers"], raw["positions"], raw["q"], cut


def grads(fn, numbers, positions, q, values, g, model, cut):
    tp = {k: torch.tensor(v, dtype=F64, requires_grad=True) for k, v in values.items()}
    e = fn(numbers, positions, q, tp, model, cut)
    out = torch.autograd.grad((e * g).sum(), [tp[k] for k in KEYS])
    return np.array([o.item() for o in out])


def run_reference(numbers, positions, q, tp, model, cut):
    cutoff = ref.Cutoff(dtype=F64, **cut) if cut else None
    return ref.dftd4(numbers, positions, torch.zeros(numbers.shape[:-1], dtype=F64), ref.damping.Param(**tp),
                     q=q, model=model, cutoff=cutoff)  # fmt: skip


def run_oracle(numbers, positions, q, tp, model, cut):
    return orc.dftd4(numbers, positions, tp, q, model=model, **cut)


def main() -> None:
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    worst = 0.0
    for name, values in CASES.items():
        numbers, positions, q, cut = load(name)
        n, p, qq = torch.from_numpy(numbers), torch.from_numpy(positions), torch.from_numpy(q)
        g = np.random.default_rng(7).normal(size=numbers.shape)
        gt = torch.from_numpy(g)
        store = {"param_keys": np.array(KEYS), "param_vals": np.array([values[k] for k in KEYS]), "g": g}
        for model in ("d4", "d4s"):
            r = grads(run_reference, n, p, qq, values, gt, model, cut)
            o = grads(run_oracle, n, p, qq, values, gt, model, cut)
            rel = np.abs(r - o).max() / np.abs(r).max()
            worst = max(worst, rel)
            print(f"{name:16s} {model:4s} " + " ".join(f"{v:+.6e}" for v in r) + f"   oracle-ref rel {rel:.1e}")
            assert np.all(np.abs(r - o) <= 1e-11 * np.abs(r) + 1e-18), (name, model, r, o)
            store[f"grad_param_{model}"] = r
        np.savez_compressed(OUT / f"{name}.npz", **store)
    print(f"worst oracle-vs-reference (relative to the largest component): {worst:.2e}")


if __name__ == "__main__":
    main()
