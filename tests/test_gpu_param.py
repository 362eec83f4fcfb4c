"""
Gradients with respect to the damping parameters through the public API:
``Param`` values that are tensors with ``requires_grad`` (the reference's
``test/test_grad/test_param.py:40-100`` differentiates s6, s8, s9, s10, a1, a2, alp of
TPSS0-D4-ATM) -> ``d4b200_param_vjp_*``; compared with autograd of the float64 oracle on
the CPU.  Tolerance: 1e-9 relative (+1e-12 absolute) in FP64, 1e-4 relative in FP32.
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

import d4_oracle as orc
from helpers import as_torch, load_golden

pytestmark = pytest.mark.gpu

KEYS = ("s6", "s8", "s9", "s10", "a1", "a2", "alp")
# test/test_grad/test_param.py:54-62
TPSS0 = {"s6": 1.0, "s8": 0.78981345, "s9": 1.0, "s10": 0.0, "a1": 0.49484001, "a2": 5.73083694, "alp": 16.0}


def _params(case, use_case_param: bool):
    base = dict(TPSS0)
    if use_case_param:
        base = {"s6": 1.0, "s8": 1.0, "s9": 1.0, "alp": 16.0}
        base.update(case["param"])
    return base


def _oracle(case, base, model, g, with_pos=False):
    numbers, positions, q = as_torch(case)
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in base.items()}
    pos = positions.clone().requires_grad_(with_pos)
    e = orc.dftd4(numbers, pos, tp, q, model=model, **case["cutoff"])
    keys = [k for k in KEYS if k in tp]
    L = (e * torch.from_numpy(g)).sum()
    out = torch.autograd.grad(L, [tp[k] for k in keys] + ([pos] if with_pos else []))
    return keys, [o.detach().numpy() for o in out]


def _cuda(case, base, model, g, dtype=torch.float64, with_pos=False):
    import tad_dftd4_b200 as d4

    dev = torch.device("cuda:0")
    numbers, positions, q = as_torch(case, dev, dtype)
    tp = {k: torch.tensor(v, dtype=dtype, device=dev, requires_grad=True) for k, v in base.items()}
    pos = positions.clone().requires_grad_(with_pos)
    cut = d4.Cutoff(**case["cutoff"], device=dev, dtype=dtype) if case["cutoff"] else None
    e = d4.dftd4(numbers, pos, 0.0, d4.Param(**tp), q=q, model=model, cutoff=cut)
    keys = [k for k in KEYS if k in tp]
    L = (e * torch.from_numpy(g).to(dev, dtype)).sum()
    out = torch.autograd.grad(L, [tp[k] for k in keys] + ([pos] if with_pos else []))
    return keys, [o.detach().cpu().numpy().astype(np.float64) for o in out]


def _upstream(case, seed=7):
    return np.random.default_rng(seed).normal(size=case["numbers"].shape)


@pytest.mark.parametrize("model", ["d4", "d4s"])
@pytest.mark.parametrize("name", ["lih_tpssh", "sih4_tpssh", "nan17", "organic_33", "ragged_batch", "holes"])
def test_param_gradient_tpss0(name, model):
    """The reference's parameter set (s10 = 0.0 present: its derivative is not zero)."""
    case = load_golden(name)
    g = _upstream(case)
    keys, ref = _oracle(case, TPSS0, model, g)
    keys2, got = _cuda(case, TPSS0, model, g)
    assert keys == keys2 == list(KEYS)
    for k, r, v in zip(keys, ref, got):
        assert v.shape == r.shape == ()
        assert abs(v - r) <= 1e-12 + 1e-9 * abs(r), (k, float(v), float(r))


@pytest.mark.parametrize("name", ["single_tpssh_s10", "tight_cutoffs", "organic_100", "big_charges"])
def test_param_gradient_case_parameters(name):
    """Non-default alp (general power path), s10 != 0, tight cutoffs (open triples), 100 atoms."""
    case = load_golden(name)
    base = _params(case, True)
    g = _upstream(case, 3)
    keys, ref = _oracle(case, base, "d4", g)
    _, got = _cuda(case, base, "d4", g)
    for k, r, v in zip(keys, ref, got):
        assert abs(v - r) <= 1e-12 + 1e-9 * abs(r), (k, float(v), float(r))


def test_param_and_position_gradients_together():
    """One backward pass for positions and parameters (fused forward + parameter VJP)."""
    case = load_golden("organic_20")
    g = np.ones(case["numbers"].shape)
    keys, ref = _oracle(case, TPSS0, "d4", g, with_pos=True)
    _, got = _cuda(case, TPSS0, "d4", g, with_pos=True)
    for k, r, v in zip(keys, ref[:-1], got[:-1]):
        assert abs(v - r) <= 1e-12 + 1e-9 * abs(r), (k, float(v), float(r))
    assert np.abs(got[-1] - ref[-1]).max() < 1e-9


def test_subset_of_parameters_and_defaults():
    """Only a1 and s8 are differentiated; s6/s9/alp are absent (defaults), a2 is a float."""
    import tad_dftd4_b200 as d4

    case = load_golden("single_pbe0")
    dev = torch.device("cuda:0")
    numbers, positions, q = as_torch(case, dev)
    a1 = torch.tensor(0.40085597, dtype=torch.float64, device=dev, requires_grad=True)
    s8 = torch.tensor(1.20065498, dtype=torch.float64, requires_grad=True)  # a CPU tensor, as from get_params
    e = d4.dftd4(numbers, positions, 0.0, {"a1": a1, "a2": 5.02928789, "s8": s8}, q=q)
    ga1, gs8 = torch.autograd.grad(e.sum(), (a1, s8))
    assert ga1.device == a1.device and gs8.device == s8.device
    n_, p_, q_ = as_torch(case)
    ta1 = torch.tensor(0.40085597, dtype=torch.float64, requires_grad=True)
    ts8 = torch.tensor(1.20065498, dtype=torch.float64, requires_grad=True)
    er = orc.dftd4(n_, p_, {"a1": ta1, "a2": 5.02928789, "s8": ts8}, q_)
    ra1, rs8 = torch.autograd.grad(er.sum(), (ta1, ts8))
    assert abs(ga1.item() - ra1.item()) <= 1e-9 * abs(ra1.item())
    assert abs(gs8.item() - rs8.item()) <= 1e-9 * abs(rs8.item())


def test_param_gradient_default_charges():
    """q=None (EEQ charges on device): the charges do not depend on the damping parameters."""
    import eeq_oracle as eeq
    import tad_dftd4_b200 as d4

    case = load_golden("nan17")
    dev = torch.device("cuda:0")
    numbers, positions, _ = as_torch(case)
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in TPSS0.items()}
    er = orc.dftd4(numbers, positions, tp, eeq.get_eeq_charges(numbers, positions, 0.0))
    ref = torch.autograd.grad(er.sum(), [tp[k] for k in KEYS])
    cp = {k: torch.tensor(v, dtype=torch.float64, device=dev, requires_grad=True) for k, v in TPSS0.items()}
    e = d4.dftd4(numbers.to(dev), positions.to(dev), torch.tensor(0.0, device=dev, dtype=torch.float64), cp)
    got = torch.autograd.grad(e.sum(), [cp[k] for k in KEYS])
    for k, r, v in zip(KEYS, ref, got):
        assert abs(v.item() - r.item()) <= 1e-12 + 1e-8 * abs(r.item()), (k, v.item(), r.item())


def test_param_gradient_f32():
    case = load_golden("organic_33")
    g = _upstream(case)
    keys, ref = _oracle(case, TPSS0, "d4", g)
    _, got = _cuda(case, TPSS0, "d4", g, dtype=torch.float32)
    for k, r, v in zip(keys, ref, got):
        assert abs(v - r) <= 1e-4 * abs(r) + 1e-9, (k, float(v), float(r))


def test_param_gradient_optimizer_step_changes_energy():
    """The reference's use case: fit the damping parameters with a torch optimizer."""
    import tad_dftd4_b200 as d4

    case = load_golden("organic_20")
    dev = torch.device("cuda:0")
    numbers, positions, q = as_torch(case, dev)
    p = {k: torch.tensor(v, dtype=torch.float64, device=dev, requires_grad=True) for k, v in
         (("s8", 1.2), ("a1", 0.4), ("a2", 5.0))}  # fmt: skip
    opt = torch.optim.SGD(list(p.values()), lr=1.0)
    target = -0.05
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = (d4.dftd4(numbers, positions, 0.0, p, q=q).sum() - target) ** 2
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[2] < losses[1] < losses[0]


from helpers import PARAM_GOLDEN_CASES, load_param_golden  # noqa: E402


@pytest.mark.parametrize("model", ["d4", "d4s"])
@pytest.mark.parametrize("name", PARAM_GOLDEN_CASES)
def test_param_gradient_reference_golden(name, model):
    """CUDA path vs parameter gradients produced by the UNMODIFIED reference
    (tests/golden/param, oracle/make_golden_param.py)."""
    case = load_param_golden(name)
    keys, got = _cuda(case, case["pvalues"], model, case["g"])
    assert keys == list(KEYS)
    for k, r, v in zip(keys, case[f"grad_param_{model}"], got):
        assert abs(v - r) <= 1e-12 + 1e-9 * abs(r), (k, float(v), float(r))


def test_param_gradients_in_the_tiled_large_system_family():
    """A structure beyond the one-CTA kernels (150 atoms): parameter gradients from re-run tiled energy kernels
    (large.large_param_vjp: exact parts for the scaling factors, 4th-order differences for a1 / a2 / alp), together
    with the position gradient of the same weighted sum, against autograd of the float64 oracle."""
    numbers, positions, q = orc.organic_batch([150], seed=43)
    case = {"numbers": numbers[0].numpy(), "positions": positions[0].numpy(), "q": q[0].numpy(), "cutoff": {}}
    base = dict(TPSS0, s10=0.3)
    g = _upstream(case, seed=9)
    keys, want = _oracle(case, base, "d4", g, with_pos=True)
    keys2, got = _cuda(case, base, "d4", g, with_pos=True)
    assert keys == keys2
    for k, w, v in zip(keys + ["positions"], want, got):
        if k == "positions":
            assert np.abs(v - w).max() < 1e-9, k
        else:
            assert abs(v - w) <= 1e-8 * abs(w) + 1e-12, (k, float(v), float(w))
