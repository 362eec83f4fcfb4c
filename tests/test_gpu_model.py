"""
Model-level API on device (csrc/d4b200_model.cu): D4Model / D4SModel.weight_references,
get_atomic_c6, get_weighted_pols, get_polarizabilities against the oracle and against the
reference's in-tree known-answer vectors K1/K3 (test/test_model/samples.py:63-124,
test_weights.py:94-101) -- the tests read like test/test_model/test_{weights,c6}.py.
"""
from __future__ import annotations

import pytest
import torch

import d4_oracle as orc
from test_oracle_kat import LIH_Q, LIH_XYZ, LIH_Z, SIH4_Q, SIH4_XYZ, SIH4_Z

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda:0")
F64 = torch.float64


def _d4():
    import tad_dftd4_b200 as d4

    return d4


def _batch(seed=21):
    numbers, positions, q = orc.organic_batch([9, 17, 5], seed=seed)
    cn = orc.cn_d4(numbers, positions)
    return numbers, cn, q


@pytest.mark.parametrize("with_q", [True, False])
def test_weight_references_d4(with_q):
    d4 = _d4()
    numbers, cn, q = _batch()
    model = d4.D4Model(numbers.to(DEV), dtype=F64)
    gw = model.weight_references(cn.to(DEV), q.to(DEV) if with_q else None)
    ref = orc.weight_references_d4(numbers, cn, q if with_q else None)
    assert gw.shape == (*numbers.shape, 7) and gw.dtype == F64
    assert (gw.cpu() - ref).abs().max().item() < 1e-13
    assert (gw.cpu()[numbers == 0] == 0).all()


def test_weight_reference_derivatives_d4():
    d4 = _d4()
    numbers, cn, q = _batch(22)
    model = d4.D4Model(numbers.to(DEV), dtype=F64)
    gw, dcn, dq = model.weight_references(cn.to(DEV), q.to(DEV), with_dgwdq=True, with_dgwdcn=True)
    cnr, qr = cn.clone().requires_grad_(True), q.clone().requires_grad_(True)
    ref = orc.weight_references_d4(numbers, cnr, qr)
    for a in range(7):  # the weights of an atom depend on its own cn and q only
        gcn, gq = torch.autograd.grad(ref[..., a].sum(), (cnr, qr), retain_graph=True)
        assert (dcn.cpu()[..., a] - gcn).abs().max().item() < 1e-12
        assert (dq.cpu()[..., a] - gq).abs().max().item() < 1e-12
    only_q = model.weight_references(cn.to(DEV), q.to(DEV), with_dgwdq=True)
    assert isinstance(only_q, tuple) and len(only_q) == 2 and torch.equal(only_q[1], dq)


def test_weight_references_d4s_and_c6():
    d4 = _d4()
    numbers, cn, q = _batch(23)
    model = d4.D4SModel(numbers.to(DEV), dtype=F64)
    gw = model.weight_references(cn.to(DEV), q.to(DEV))
    ref = orc.weight_references_d4s(numbers, cn, q)
    assert gw.shape == (*numbers.shape, numbers.shape[-1], 7)
    real = (numbers != 0).unsqueeze(-1) & (numbers != 0).unsqueeze(-2)
    assert ((gw.cpu() - ref).abs().amax(-1)[real]).max().item() < 1e-13
    c6 = model.get_atomic_c6(gw).cpu()
    c6_ref = orc.atomic_c6_d4s(orc.reference_c6(numbers), ref)
    assert ((c6 - c6_ref).abs()[real]).max().item() / c6_ref.abs().max().item() < 1e-13


def test_atomic_c6_and_polarizabilities_d4():
    d4 = _d4()
    numbers, cn, q = _batch(24)
    model = d4.D4Model(numbers.to(DEV), dtype=F64)
    gw = model.weight_references(cn.to(DEV), q.to(DEV))
    ref_w = orc.weight_references_d4(numbers, cn, q)
    c6 = model.get_atomic_c6(gw).cpu()
    c6_ref = orc.atomic_c6_d4(orc.reference_c6(numbers), ref_w)
    assert (c6 - c6_ref).abs().max().item() / c6_ref.abs().max().item() < 1e-13
    alpha_ref = orc.reference_alpha(numbers)  # (..., nat, 7, 23)
    pols = model.get_weighted_pols(gw).cpu()
    pols_ref = torch.einsum("...nr,...nrw->...nw", ref_w, alpha_ref)
    assert pols.shape == (*numbers.shape, 23)
    assert (pols - pols_ref).abs().max().item() / pols_ref.abs().max().item() < 1e-13
    a0 = model.get_polarizabilities(gw).cpu()
    assert a0.shape == numbers.shape
    assert (a0 - pols_ref[..., 0]).abs().max().item() / pols_ref.abs().max().item() < 1e-13


def test_k1_lih_c6_from_golden_weights():
    # test/test_model/samples.py:72-85 (gw) and :104-112 (c6); tolerance of test_c6.py:41 is 1e-5
    d4 = _d4()
    gw = torch.zeros(2, 7, dtype=F64)
    gw[0, :3] = torch.tensor([1.8699287753787968e-02, 9.7889292523075033e-01, 1.8718044551687104e-37])
    gw[1, :3] = torch.tensor([7.9608926855620182e-02, 3.5225968112617356e00, 0.0])
    model = d4.D4Model(LIH_Z.to(DEV), dtype=F64)
    c6 = model.get_atomic_c6(gw.to(DEV)).cpu()
    ref = torch.tensor([[4.1059628873073926e01, 2.9129176877403175e01],
                        [2.9129176877403175e01, 4.0408036338319796e01]], dtype=F64)  # fmt: skip
    assert ((c6 - ref).abs() / ref).max().item() < 1e-7


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_k3_lih_weights(dtype):
    # test/test_model/test_weights.py:54-101: cn, q -> gw (tolerance 1e-6 there, Fortran goldens)
    d4 = _d4()
    cn = orc.cn_d4(LIH_Z, LIH_XYZ)
    model = d4.D4Model(LIH_Z.to(DEV), dtype=dtype)
    gw = model.weight_references(cn.to(DEV, dtype), LIH_Q.to(DEV, dtype)).cpu()
    assert gw.dtype == dtype
    ref0 = torch.tensor([1.8699287753787968e-02, 9.7889292523075033e-01, 1.8718044551687104e-37], dtype=F64)
    ref1 = torch.tensor([7.9608926855620182e-02, 3.5225968112617356e00, 0.0], dtype=F64)
    tol = 1e-6 if dtype == F64 else 1e-5
    assert (gw[0, :3].double() - ref0).abs().max().item() < tol
    assert (gw[1, :3].double() - ref1).abs().max().item() < tol


def test_sih4_d4s_equals_d4_for_uniform_weighting_factor():
    # K10 (test/test_model/test_models.py:195-207): with the default wf for every pair the D4S
    # weights reduce to the D4 weights; checked here through the element pair H-H (wf = 6)
    d4 = _d4()
    z = torch.tensor([1, 1])
    cn = torch.tensor([0.9, 1.1], dtype=F64)
    gws = d4.D4SModel(z.to(DEV), dtype=F64).weight_references(cn.to(DEV)).cpu()
    gw = orc.weight_references_d4s(z, cn, None)
    assert (gws - gw).abs().max().item() < 1e-14


def test_shape_errors():
    d4 = _d4()
    model = d4.D4Model(SIH4_Z.to(DEV), dtype=F64)
    with pytest.raises(ValueError):
        model.weight_references(torch.zeros(4, dtype=F64, device=DEV))
    with pytest.raises(ValueError):
        model.get_atomic_c6(torch.zeros(5, 6, dtype=F64, device=DEV))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d4.D4Model(SIH4_Z, dtype=F64).weight_references()
