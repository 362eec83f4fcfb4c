"""DispD4Exact / D4ATMExact (dispersion/d4.py:67-84, threebody.py:276-302 of the reference): golden
vectors from the UNMODIFIED reference (tests/golden/exact, oracle/make_golden_exact.py) against the oracle and the
single-node tables on the CPU, and against the kernels (23 ATM-only launches + one two-body launch) on the B200."""
from __future__ import annotations

import numpy as np
import pytest
import torch

import d4_oracle as orc
from helpers import GOLDEN, as_torch, load_golden

CASES = ("single_pbe0", "sih4_tpssh", "organic_33", "ragged_batch", "tight_cutoffs", "all_elements")


def _load(name):
    case = load_golden(name)
    return case, np.load(GOLDEN / "exact" / f"{name}.npz")


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_reference_exact_c9(name):
    case, gold = _load(name)
    n, p, q = as_torch(case)
    e, g = orc.energy_and_gradient(n, p, case["param"], q, c9="exact", **case["cutoff"])
    assert np.abs(e.numpy() - gold["energy"]).max() <= 1e-13 * np.abs(gold["energy"]).max()
    assert np.abs(g.numpy() - gold["gradient"]).max() < 1e-15
    _, e3, *_ = orc.dftd4(n, p, case["param"], q, parts=True, c9="exact", **case["cutoff"])
    assert np.abs(e3.numpy() - gold["energy_atm"]).max() <= 1e-12 * np.abs(gold["energy_atm"]).max()


def test_single_node_tables_sum_to_the_exact_c9():
    """What the 23 launches rely on: with the table of node w the kernels' pair quantity c_ij = A_i . A_j is
    b_i b_j, sqrt(c_ij c_jk c_ik) = b_i b_j b_k, and the sum over the nodes is the Casimir-Polder integral."""
    from tad_dftd4_b200.tables import NFREQ, build_tables

    case = load_golden("organic_33")
    n, p, q = as_torch(case)
    cn = orc.cn_d4(n, p)
    w0 = orc.weight_references_d4(n, cn, None).numpy()
    aiw = np.einsum("nr,nrw->nw", w0, orc.reference_alpha(n).numpy())
    tw = np.asarray(orc._CP_WEIGHTS)
    exact = 3.0 / np.pi * np.einsum("w,iw,jw,kw->ijk", tw, aiw, aiw, aiw)
    total = np.zeros_like(exact)
    z = n.numpy()
    for w in range(NFREQ):
        tab = build_tables(3.0, 2.0, "eeq", w)
        assert np.count_nonzero(np.abs(tab.alpha_w).sum((0, 1))) == 1
        a0 = np.einsum("nr,nrw->nw", w0, tab.alpha_w[z])  # the kernels' weighted-polarizability vectors
        c = a0 @ a0.T
        total += np.sqrt(np.abs(c[:, :, None] * c[None, :, :] * c[:, None, :]))
    assert np.abs(total - exact).max() <= 1e-13 * np.abs(exact).max()
    with pytest.raises(ValueError):
        build_tables(3.0, 2.0, "eeq", 23)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernels_exact_c9(name):
    import tad_dftd4_b200 as d4
    from tad_dftd4_b200.dispersion import D4ATMExact, Disp, DispD4Exact

    case, gold = _load(name)
    dev = torch.device("cuda:0")
    n, p, q = as_torch(case, dev)
    cutoff = d4.Cutoff(**case["cutoff"]) if case["cutoff"] else None
    pos = p.clone().requires_grad_(True)
    e = DispD4Exact().calculate(n, pos, 0.0, case["param"], q=q, cutoff=cutoff)
    (g,) = torch.autograd.grad(e.sum(), pos)
    assert np.abs(e.detach().cpu().numpy() - gold["energy"]).max() <= 1e-10 * np.abs(gold["energy"]).max()
    assert np.abs(g.cpu().numpy() - gold["gradient"]).max() < 1e-9
    # the ATM term alone, registered on a bare Disp (test/test_disp/test_variants.py:64-86 of the reference)
    atm = Disp(model="d4")
    atm.register(D4ATMExact())
    pos3 = p.clone().requires_grad_(True)
    e3 = atm.calculate(n, pos3, 0.0, case["param"], cutoff=cutoff)
    (g3,) = torch.autograd.grad(e3.sum(), pos3)
    scale = max(np.abs(gold["energy_atm"]).max(), 1e-300)
    assert np.abs(e3.detach().cpu().numpy() - gold["energy_atm"]).max() <= 1e-10 * scale
    assert np.abs(g3.cpu().numpy() - gold["gradient_atm"]).max() < 1e-9


@pytest.mark.gpu
def test_exact_c9_is_for_the_d4_model():
    from tad_dftd4_b200.dispersion import DispD4Exact

    dev = torch.device("cuda:0")
    case = load_golden("single_pbe0")
    n, p, q = as_torch(case, dev)
    with pytest.raises(NotImplementedError):
        DispD4Exact(model="d4s").calculate(n, p, 0.0, case["param"], q=q)


@pytest.mark.gpu
def test_registered_term_combinations_are_sums_of_launches():
    """Disp.calculate beyond the fused default (dispersion/base.py:409-431): every registered term is one launch
    (or 23) with the other part switched off; the oracle evaluates the same combinations."""
    from tad_dftd4_b200.dispersion import D4ATMApprox, D4ATMExact, Disp, TwoBodyTerm

    dev = torch.device("cuda:0")
    case = load_golden("ragged_batch")
    n, p, q = as_torch(case)
    par = case["param"]
    e2q, e3, *_ = orc.dftd4(n, p, par, q, parts=True)
    e20, *_ = orc.dftd4(n, p, par, torch.zeros_like(q), parts=True)
    _, e3x, *_ = orc.dftd4(n, p, par, q, parts=True, c9="exact")
    nd, pd, qd = n.to(dev), p.to(dev), q.to(dev)

    def run(terms, **kw):
        disp = Disp(model="d4")
        for t in terms:
            disp.register(t)
        return disp.calculate(nd, pd, 0.0, par, **kw).cpu()

    def close(a, b):
        return (a - b).abs().max().item() <= 1e-10 * b.abs().max().item()

    assert close(run([TwoBodyTerm(charge_dependent=False)]), e20)
    assert close(run([TwoBodyTerm(charge_dependent=False), D4ATMApprox()]), e20 + e3)
    assert close(run([TwoBodyTerm(), D4ATMExact()], q=qd), e2q + e3x)
    assert close(run([D4ATMApprox(), D4ATMExact()]), e3 + e3x)
    # the ATM terms take their BJ radii from the defaults when a1/a2 are not given (threebody.py:244-256)
    nopar = {k: v for k, v in par.items() if k not in ("a1", "a2")}
    _, e3d, *_ = orc.dftd4(n, p, dict(nopar, a1=0.4, a2=5.0), q, parts=True)
    disp = Disp(model="d4")
    disp.register(D4ATMApprox())
    assert close(disp.calculate(nd, pd, 0.0, nopar).cpu(), e3d)
    with pytest.raises(NotImplementedError):
        run([TwoBodyTerm(), D4ATMApprox(charge_dependent=True)], q=qd)


@pytest.mark.gpu
def test_exact_c9_and_gfn2_charges_in_the_tiled_large_system_family():
    """Structures beyond the one-CTA-per-structure kernels take the tiled family; the single-node tables (exact C9)
    and the GFN2 reference charges reach it through the model tuple (large.py)."""
    import tad_dftd4_b200 as d4
    from tad_dftd4_b200.dispersion import DispD4Exact

    dev = torch.device("cuda:0")
    n, p, q = orc.organic_batch([150], seed=41)
    n, p, q = n[0], p[0], q[0]
    par = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)
    for kind in ("exact", "gfn2"):
        okw = dict(c9="exact") if kind == "exact" else dict(ref_charges="gfn2")
        e_ref, g_ref = orc.energy_and_gradient(n, p, par, q, **okw)
        pos = p.to(dev).requires_grad_(True)
        if kind == "exact":
            e = DispD4Exact().calculate(n.to(dev), pos, 0.0, par, q=q.to(dev))
        else:
            e = d4.dftd4(n.to(dev), pos, 0.0, par, q=q.to(dev), model=d4.D4Model(n.to(dev), ref_charges="gfn2"))
        (g,) = torch.autograd.grad(e.sum(), pos)
        assert ((e.detach().cpu() - e_ref).abs().max() / e_ref.abs().max()).item() < 1e-10, kind
        assert (g.cpu() - g_ref).abs().max().item() < 1e-9, kind
