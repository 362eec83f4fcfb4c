"""The algebra used by the CUDA kernels (tests/kernel_model.py) vs the oracle."""
from __future__ import annotations

import numpy as np
import pytest
import torch

import d4_oracle as orc
from helpers import as_torch, load_golden
from kernel_model import energy_and_gradient

CASES = ["sih4_tpssh", "lih_tpssh", "single_pbe0", "single_tpssh_s10", "nan17", "organic_5",
         "organic_20", "tight_cutoffs", "big_charges", "chain150"]  # fmt: skip


@pytest.mark.parametrize("name", CASES)
def test_model_matches_golden(name):
    case = load_golden(name)
    e, grad, dq, cn = energy_and_gradient(
        case["numbers"], case["positions"], case["q"], case["param"], **case["cutoff"]
    )
    eref, gref = case["energy_d4"], case["grad_d4"]
    assert np.abs(cn - case["cn"]).max() < 1e-12
    assert np.abs(e - eref).max() / np.abs(eref).max() < 1e-10
    assert np.abs(grad - gref).max() < 1e-9


def test_dq_and_weighted_upstream():
    """dL/dq and arbitrary upstream weights g_i vs oracle autograd."""
    case = load_golden("organic_20")
    numbers, positions, q = as_torch(case)
    rng = np.random.default_rng(5)
    g = rng.normal(size=len(case["numbers"]))
    pos = positions.clone().requires_grad_(True)
    qq = q.clone().requires_grad_(True)
    e = orc.dftd4(numbers, pos, case["param"], qq)
    gp, gq = torch.autograd.grad((e * torch.from_numpy(g)).sum(), (pos, qq))
    _, grad, dq, _ = energy_and_gradient(case["numbers"], case["positions"], case["q"], case["param"], g=g)
    assert np.abs(grad - gp.numpy()).max() < 1e-9
    assert np.abs(dq - gq.numpy()).max() < 1e-9


PARAM_KEYS = ("s6", "s8", "s9", "s10", "a1", "a2", "alp")


@pytest.mark.parametrize("name,with_s10", [("sih4_tpssh", True), ("organic_20", False), ("tight_cutoffs", True),
                                           ("single_pbe0", True)])  # fmt: skip
def test_param_gradient_model(name, with_s10):
    """dL/d(damping parameters) (test/test_grad/test_param.py of the reference differentiates
    the same seven) vs oracle autograd, with random upstream weights."""
    from kernel_model import param_gradient

    case = load_golden(name)
    numbers, positions, q = as_torch(case)
    base = {"s6": 1.0, "s8": 0.78981345, "s9": 1.0, "a1": 0.49484001, "a2": 5.73083694, "alp": 16.0}
    base.update({k: float(v) for k, v in case["param"].items()})
    if with_s10:
        base.setdefault("s10", 0.0)
    else:
        base.pop("s10", None)
    rng = np.random.default_rng(11)
    g = rng.normal(size=len(case["numbers"]))
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in base.items()}
    e = orc.dftd4(numbers, positions, tp, q, **case["cutoff"])
    keys = [k for k in PARAM_KEYS if k in tp]
    ref = torch.autograd.grad((e * torch.from_numpy(g)).sum(), [tp[k] for k in keys])
    got = param_gradient(case["numbers"], case["positions"], case["q"], base, g=g, **case["cutoff"])
    for k, r in zip(keys, ref):
        v = got[PARAM_KEYS.index(k)]
        assert abs(v - r.item()) <= 1e-12 + 1e-9 * abs(r.item()), (k, v, r.item())


def test_gradient_visit_factorisation():
    """The 21-instruction gradient visit of the kernel equals the textbook expression
    (random triangles, damping arguments and alp)."""
    from kernel_model import grad_visit_kernel, grad_visit_reference

    rng = np.random.default_rng(3)
    for _ in range(2000):
        x = rng.normal(size=(3, 3)) * 3.0
        a, b, c = ((x[0] - x[1]) ** 2).sum(), ((x[1] - x[2]) ** 2).sum(), ((x[0] - x[2]) ** 2).sum()
        P = rng.uniform(1e-6, 1e-2, size=3)
        u = rng.uniform(1e-3, 30.0, size=3)
        alp = rng.choice([16.0, 14.0, 12.5])
        e_ref, de_ref = grad_visit_reference(a, b, c, *P, *u, alp)
        _, (e, de) = grad_visit_kernel(a, b, c, *P, *u, alp)
        # scale: the two terms of 0.375 s + abc can cancel (flat triangles)
        f = 1.0 / (1.0 + 6.0 * u.prod())
        scale = P.prod() * f * a * b * c
        assert abs(e - e_ref) <= 1e-12 * scale
        assert abs(de - de_ref) <= 1e-11 * scale * (alp + 10.0) / min(a, b, c)
