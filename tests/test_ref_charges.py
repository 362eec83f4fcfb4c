"""ref_charges="gfn2" (model/base.py:388-399, model/d4.py:142-149 of the reference):
golden vectors from the UNMODIFIED reference (tests/golden/gfn2, oracle/make_golden_gfn2.py) against the
oracle and the compiled element tables on the CPU, and against the kernels on the B200."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest
import torch

import d4_oracle as orc

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = ("single_pbe0", "sih4_tpssh", "organic_33", "ragged_batch", "all_elements")


def _load(name):
    raw, gold = np.load(GOLDEN / f"{name}.npz"), np.load(GOLDEN / "gfn2" / f"{name}.npz")
    n, p, q = (torch.from_numpy(raw[k]) for k in ("numbers", "positions", "q"))
    if name == "all_elements":  # tabulated up to Rn: atoms beyond are padding in this fixture (make_golden_gfn2.py)
        keep = n <= 86
        n, p, q = n * keep, p * keep.unsqueeze(-1), q * keep
    param = {str(k): float(v) for k, v in zip(raw["param_keys"], raw["param_vals"])}
    return n, p, q, param, gold


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("model", ["d4", "d4s"])
def test_oracle_reproduces_reference_with_gfn2_charges(name, model):
    n, p, q, param, gold = _load(name)
    e, g = orc.energy_and_gradient(n, p, param, q, model=model, ref_charges="gfn2")
    assert np.abs(e.numpy() - gold[f"energy_{model}"]).max() <= 1e-13 * np.abs(gold[f"energy_{model}"]).max()
    assert np.abs(g.numpy() - gold[f"gradient_{model}"]).max() < 1e-15


def test_compiled_tables_follow_the_reference_charges():
    from tad_dftd4_b200.tables import build_tables

    z = torch.tensor([1, 6, 7, 8, 14, 16, 35, 53, 82])
    for ref in ("eeq", "gfn2"):
        tab = build_tables(3.0, 2.0, ref)
        rc6 = orc.reference_c6(z, ref_charges=ref).numpy()
        assert np.abs(tab.rc6[z.numpy()][:, z.numpy()] - rc6).max() <= 1e-12 * np.abs(rc6).max()
        refq, _ = orc._ref_charge_tables(ref)
        assert np.allclose(tab.refq[z.numpy()] - tab.zeff[z.numpy(), None], refq.numpy()[z.numpy()], rtol=0, atol=1e-13)
    assert np.abs(build_tables(3.0, 2.0, "gfn2").rc6 - build_tables(3.0, 2.0, "eeq").rc6).max() > 1e-3
    with pytest.raises(ValueError):
        build_tables(3.0, 2.0, "mulliken")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("model", ["d4", "d4s"])
def test_kernels_with_gfn2_reference_charges(name, model):
    import tad_dftd4_b200 as d4

    n, p, q, param, gold = _load(name)
    dev = torch.device("cuda:0")
    cls = d4.D4Model if model == "d4" else d4.D4SModel
    pos = p.to(dev).requires_grad_(True)
    e = d4.dftd4(n.to(dev), pos, 0.0, param, q=q.to(dev), model=cls(n.to(dev), ref_charges="gfn2"))
    (g,) = torch.autograd.grad(e.sum(), pos)
    e_ref, g_ref = gold[f"energy_{model}"], gold[f"gradient_{model}"]
    assert np.abs(e.detach().cpu().numpy() - e_ref).max() <= 1e-10 * np.abs(e_ref).max()
    assert np.abs(g.cpu().numpy() - g_ref).max() < 1e-9
    # and the default charges still give the default result (separate table sets per engine)
    e0 = d4.dftd4(n.to(dev), p.to(dev), 0.0, param, q=q.to(dev), model=model)
    ref0 = orc.dftd4(n, p, param, q, model=model)
    assert ((e0.cpu() - ref0).abs().max() / ref0.abs().max()) < 1e-10


@pytest.mark.gpu
def test_gfn2_charges_beyond_the_table_raise():
    import tad_dftd4_b200 as d4

    dev = torch.device("cuda:0")
    n = torch.tensor([90, 1], device=dev)
    p = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 4.0]], dtype=torch.float64, device=dev)
    with pytest.raises(IndexError):
        d4.dftd4(n, p, 0.0, dict(a1=0.4, a2=5.0), q=torch.zeros(2, dtype=torch.float64, device=dev),
                 model=d4.D4Model(n, ref_charges="gfn2"))
