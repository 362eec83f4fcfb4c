"""
Second derivatives (second half): double backward through ``dftd4`` on the B200 --
the analytic VJP kernels differentiated semi-numerically (``disp._D4Vjp`` / ``eeq._EeqVjp``) --
against nuclear Hessians of the UNMODIFIED reference (``tests/golden/hessian``, made by
``oracle/make_golden_hessian.py`` the way ``test/test_grad/test_hessian.py:75-112`` obtains them),
with the reference's tolerance 1e-7 (measured: ~1e-10), and ``gradgradcheck`` after
``test/test_grad/test_pos.py:83-89``.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest
import torch

import tad_dftd4_b200 as d4

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = ("lih_tpssh", "sih4_tpssh", "organic_5", "organic_20")
TOL = 1e-7  # test/test_grad/test_hessian.py:47 of the reference


def _load(name):
    raw, hes = np.load(GOLDEN / f"{name}.npz"), np.load(GOLDEN / "hessian" / f"{name}.npz")
    dev = torch.device("cuda:0")
    n, p, q = (torch.from_numpy(raw[k]).to(dev) for k in ("numbers", "positions", "q"))
    param = {str(k): float(v) for k, v in zip(hes["param_keys"], hes["param_vals"])}
    return n, p, q, param, hes


def _hessian(numbers, positions, param, **kw):
    pos = positions.clone().requires_grad_(True)
    e = d4.dftd4(numbers, pos, 0.0, param, **kw).sum()
    (g,) = torch.autograd.grad(e, pos, create_graph=True)
    rows = [torch.autograd.grad(gi, pos, retain_graph=True)[0] for gi in g.reshape(-1)]
    return torch.stack(rows).reshape(*positions.shape, *positions.shape).cpu().numpy()


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("variant", ["q_d4", "q_d4s", "eeq_d4"])
def test_hessian_against_reference_autograd(name, variant):
    n, p, q, param, hes = _load(name)
    kw = dict(model=variant.split("_")[1])
    if variant.startswith("q_"):
        kw["q"] = q
    h = _hessian(n, p, param, **kw)
    ref = hes[f"hess_{variant}"]
    assert np.abs(h - ref).max() <= TOL * max(1.0, np.abs(ref).max()), np.abs(h - ref).max()
    assert np.abs(h - ref).max() <= 2e-9  # what the semi-numerical route actually delivers
    hm = h.reshape(p.numel(), p.numel())
    assert np.abs(hm - hm.T).max() < 2e-9


def test_gradgradcheck_positions():
    n, p, q, param, _ = _load("sih4_tpssh")
    pos = p.clone().requires_grad_(True)
    assert torch.autograd.gradgradcheck(lambda x: d4.dftd4(n, x, 0.0, param), (pos,), atol=TOL, nondet_tol=1e-9)
    assert torch.autograd.gradgradcheck(lambda x: d4.dftd4(n, x, 0.0, param, q=q), (pos,), atol=TOL, nondet_tol=1e-9)


def test_hessian_of_a_padded_batch_is_block_diagonal():
    n1, p1, q1, param, h1 = _load("lih_tpssh")
    n2, p2, q2, _, h2 = _load("sih4_tpssh")
    numbers = torch.zeros((2, 5), dtype=torch.int64, device=p1.device)
    positions = torch.zeros((2, 5, 3), dtype=torch.float64, device=p1.device)
    numbers[0, :2], positions[0, :2], numbers[1], positions[1] = n1, p1, n2, p2
    h = _hessian(numbers, positions, param)  # (2, 5, 3, 2, 5, 3)
    assert np.abs(h[0, :, :, 1]).max() == 0.0 and np.abs(h[1, :, :, 0]).max() == 0.0
    assert np.abs(h[0, :2, :, 0, :2] - h1["hess_eeq_d4"]).max() < 2e-9
    assert np.abs(h[1, :, :, 1] - h2["hess_eeq_d4"]).max() < 2e-9
    assert np.abs(h[0, 2:]).max() == 0.0  # padding rows


def test_mixed_second_derivative_charge_position():
    """d/dq of the forces with explicit charges (force matching on charges): against finite differences
    of the analytic gradient."""
    n, p, q, param, _ = _load("organic_5")
    pos = p.clone().requires_grad_(True)
    qq = q.clone().requires_grad_(True)
    e = d4.dftd4(n, pos, 0.0, param, q=qq).sum()
    (g,) = torch.autograd.grad(e, pos, create_graph=True)
    w = torch.randn(g.shape, dtype=g.dtype, device=g.device, generator=torch.Generator(device=g.device).manual_seed(3))
    (mixed,) = torch.autograd.grad((g * w).sum(), qq)
    h = 1e-4
    for k in range(q.numel()):
        dq = torch.zeros_like(q)
        dq[k] = h
        gs = []
        for s in (1.0, -1.0):
            x = p.clone().requires_grad_(True)
            gs.append(torch.autograd.grad(d4.dftd4(n, x, 0.0, param, q=q + s * dq).sum(), x)[0])
        fd = ((gs[0] - gs[1]) * w).sum() / (2 * h)
        assert abs(fd.item() - mixed[k].item()) < 1e-9
