"""
EEQ charges on device (csrc/d4b200_eeq.cu) and the default ``q=None`` path of
``dftd4`` / ``get_properties`` against the oracle (oracle/eeq_oracle.py + d4_oracle.py)
and against the reference's own full-path known-answer vectors.

Tolerances: charges <= 1e-12 absolute (FP64), energies <= 1e-10 relative, gradients
<= 1e-9 absolute (north_star); FP32 I/O <= 1e-5 relative.
"""
from __future__ import annotations

import pytest
import torch

import d4_oracle as orc
import eeq_oracle as eeq
from test_oracle_kat import (FORMAMIDE2_XYZ, FORMAMIDE2_Z, FORMAMIDE_XYZ, FORMAMIDE_Z, LIH_XYZ, LIH_Z, SIH4_XYZ,
                             SIH4_Z, SINGLE_XYZ, SINGLE_Z, TPSS0, TPSSH)  # fmt: skip

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda:0")
F64 = torch.float64
PBE0 = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)


def _d4():
    import tad_dftd4_b200 as d4

    return d4


@pytest.mark.parametrize("sizes,seed", [([2], 1), ([12, 24, 7, 60, 33], 2), ([100] * 3, 3), ([160, 97], 4)])
def test_charges_and_vjp_match_oracle(sizes, seed):
    d4 = _d4()
    numbers, positions, _ = orc.organic_batch(sizes, seed=seed)
    charge = torch.arange(len(sizes), dtype=F64) % 3 - 1.0
    pos_ref = positions.clone().requires_grad_(True)
    q_ref = eeq.get_eeq_charges(numbers, pos_ref, charge)
    w = torch.randn(q_ref.shape, dtype=F64, generator=torch.Generator().manual_seed(seed))
    (g_ref,) = torch.autograd.grad((q_ref * w).sum(), pos_ref)

    pos = positions.to(DEV).requires_grad_(True)
    q = d4.get_eeq_charges(numbers.to(DEV), pos, charge.to(DEV))
    assert q.shape == numbers.shape and q.dtype == F64
    assert (q.detach().cpu() - q_ref.detach()).abs().max().item() < 1e-12
    assert (q.detach().cpu()[numbers == 0] == 0).all()
    (g,) = torch.autograd.grad((q * w.to(DEV)).sum(), pos)
    assert (g.cpu() - g_ref).abs().max().item() < 1e-11
    # bitwise reproducible
    q2 = d4.get_eeq_charges(numbers.to(DEV), positions.to(DEV), charge.to(DEV))
    assert torch.equal(q2, q.detach())


def test_padding_inside_the_atom_axis_and_empty_structure():
    d4 = _d4()
    numbers, positions, _ = orc.organic_batch([9, 14], seed=5)
    z = torch.zeros(3, 20, dtype=numbers.dtype)
    xyz = torch.zeros(3, 20, 3, dtype=F64)
    slots0 = torch.tensor([0, 2, 3, 5, 8, 9, 13, 17, 19])
    z[0, slots0], xyz[0, slots0] = numbers[0, :9], positions[0, :9]
    z[2, 3:17], xyz[2, 3:17] = numbers[1, :14], positions[1, :14]
    q = d4.get_eeq_charges(z.to(DEV), xyz.to(DEV), 0.0).cpu()
    q_ref = eeq.get_eeq_charges(numbers, positions, 0.0)
    assert (q[0, slots0] - q_ref[0, :9]).abs().max().item() < 1e-12
    assert (q[2, 3:17] - q_ref[1, :14]).abs().max().item() < 1e-12
    assert (q[1] == 0).all() and (q[z == 0] == 0).all()


def test_float32_io():
    d4 = _d4()
    numbers, positions, _ = orc.organic_batch([30, 45], seed=6)
    q_ref = eeq.get_eeq_charges(numbers, positions, 0.0)
    q = d4.get_eeq_charges(numbers.to(DEV), positions.float().to(DEV), 0.0)
    assert q.dtype == torch.float32
    assert (q.double().cpu() - q_ref).abs().max().item() < 2e-6


def test_dense_route_beyond_the_kernel_limit():
    d4 = _d4()
    n = d4.eeq.eeq_limit() + 12
    numbers, positions, _ = orc.organic_batch([n], seed=7)
    pos_ref = positions.clone().requires_grad_(True)
    q_ref = eeq.get_eeq_charges(numbers, pos_ref, 0.0)
    (g_ref,) = torch.autograd.grad((q_ref**2).sum(), pos_ref)
    pos = positions.to(DEV).requires_grad_(True)
    q = d4.get_eeq_charges(numbers.to(DEV), pos, 0.0)
    (g,) = torch.autograd.grad((q**2).sum(), pos)
    assert (q.detach().cpu() - q_ref.detach()).abs().max().item() < 1e-11
    assert (g.cpu() - g_ref).abs().max().item() < 1e-10


def test_bad_atomic_number():
    d4 = _d4()
    with pytest.raises(ValueError, match="outside 1..86"):
        d4.get_eeq_charges(torch.tensor([90, 1], device=DEV), torch.randn(2, 3, dtype=F64, device=DEV), 0.0)


@pytest.mark.parametrize("model", ["d4", "d4s"])
def test_dftd4_default_path_energy_and_forces(model):
    """dftd4(numbers, positions, charge, param) without q: EEQ on the tape like the reference."""
    d4 = _d4()
    numbers, positions, _ = orc.organic_batch([12, 40, 25], seed=8)
    charge = torch.tensor([0.0, 1.0, 0.0], dtype=F64)
    pos_ref = positions.clone().requires_grad_(True)
    e_ref = orc.dftd4(numbers, pos_ref, PBE0, eeq.get_eeq_charges(numbers, pos_ref, charge), model=model)
    (g_ref,) = torch.autograd.grad(e_ref.sum(), pos_ref)
    pos = positions.to(DEV).requires_grad_(True)
    e = d4.dftd4(numbers.to(DEV), pos, charge.to(DEV), PBE0, model=model)
    (g,) = torch.autograd.grad(e.sum(), pos)
    scale = e_ref.abs().max().item()
    assert (e.detach().cpu() - e_ref.detach()).abs().max().item() / scale < 1e-10
    assert (g.cpu() - g_ref).abs().max().item() < 1e-9


def test_k5_sih4_s10_golden_through_the_product():
    # /root/reference/test/test_d4/test_twobody.py:172-205
    d4 = _d4()
    par = dict(s8=1.85897750, s9=0.0, s10=1.0, a1=0.44286966, a2=4.60230534)
    e = d4.dftd4(SIH4_Z.to(DEV), SIH4_XYZ.to(DEV), torch.tensor(0.0, dtype=F64, device=DEV), par).cpu()
    ref = torch.tensor([-8.8928018057670788e-04] + [-3.3765541880036940e-04] * 4, dtype=F64)
    assert (e - ref).abs().max().item() < 1e-13


@pytest.mark.parametrize("name", ["LiH", "SiH4"])
def test_k6_reference_gradients_through_the_product(name):
    # /root/reference/test/test_grad/samples_grad.py:40-92 (autograd of the reference, EEQ on the tape)
    d4 = _d4()
    z, xyz = (LIH_Z, LIH_XYZ) if name == "LiH" else (SIH4_Z, SIH4_XYZ)
    pos = xyz.to(DEV).requires_grad_(True)
    e = d4.dftd4(z.to(DEV), pos, 0.0, TPSS0)
    (g,) = torch.autograd.grad(e.sum(), pos)
    if name == "LiH":
        ref = torch.tensor([[0, 0, -6.8677584018156501e-05], [0, 0, +6.8677584018156501e-05]], dtype=F64)
    else:
        s = 3.5863777807514914e-06
        ref = s * torch.tensor([[0, 0, 0], [-1, -1, 1], [1, 1, 1], [-1, 1, -1], [1, -1, -1]], dtype=F64)
    assert (g.cpu() - ref).abs().max().item() < 1e-13


def test_k7_examples_through_the_product():
    # examples/single.py:52-71 (atol 1e-8 there) and examples/batch.py:60-65 / README.md:302
    d4 = _d4()
    e = d4.dftd4(SINGLE_Z.to(DEV), SINGLE_XYZ.to(DEV), 0.0, TPSSH).cpu()
    ref = torch.tensor([-0.0020841344, -0.0018971195, -0.0018107513, -0.0018305695, -0.0021737693, -0.0019484236,
                        -0.0022788253, -0.0004080658, -0.0004261866, -0.0004199839, -0.0004280768, -0.0005108935],
                       dtype=F64)  # fmt: skip
    assert (e - ref).abs().max().item() < 1e-9
    e32 = d4.dftd4(SINGLE_Z.to(DEV), SINGLE_XYZ.float().to(DEV), 0.0, TPSSH).cpu()
    assert e32.dtype == torch.float32 and (e32.double() - ref).abs().max().item() < 1e-8
    z = torch.stack([FORMAMIDE2_Z, FORMAMIDE_Z]).to(DEV)
    xyz = torch.stack([FORMAMIDE2_XYZ, FORMAMIDE_XYZ]).to(DEV)
    tot = d4.dftd4(z, xyz, torch.zeros(2, dtype=F64, device=DEV), TPSSH).sum(-1).cpu()
    assert (tot - torch.tensor([-0.0088341432, -0.0027013607], dtype=F64)).abs().max().item() < 2e-9


def test_get_properties_default_charges():
    d4 = _d4()
    numbers, positions, _ = orc.organic_batch([15], seed=9)
    cn, q, c6, alpha = d4.get_properties(numbers.to(DEV), positions.to(DEV))
    q_ref = eeq.get_eeq_charges(numbers, positions, 0.0)
    cn_ref, _, c6_ref, alpha_ref = orc.get_properties(numbers, positions, q_ref)
    assert (q.cpu() - q_ref).abs().max().item() < 1e-12
    assert (c6.cpu() - c6_ref).abs().max().item() / c6_ref.abs().max().item() < 1e-10
    assert (alpha.cpu() - alpha_ref).abs().max().item() / alpha_ref.abs().max().item() < 1e-10


def test_total_charge_argument_types():
    """The total charge may be a Python int, a Python float or a tensor (what
    test/test_d4/test_general.py:32-52 of the reference pins): same energies for all three."""
    d4 = _d4()
    numbers = torch.tensor([1, 1], device=DEV)
    positions = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 1.0]], dtype=F64, device=DEV)
    param = d4.Param(s6=torch.tensor(1.0), s8=torch.tensor(1.0), a1=torch.tensor(0.4), a2=torch.tensor(5.0))
    results = [d4.dftd4(numbers, positions, chrg, param) for chrg in (0, 0.0, torch.tensor(0.0))]
    for e in results:
        assert e.shape == numbers.shape and bool((e < 0).all())
        assert torch.allclose(e, results[-1])


@pytest.mark.parametrize("dtype,tol", [(F64, 1e-12), (torch.float32, 2e-5)])
def test_vjp_from_the_saved_factor_equals_the_plain_vjp(dtype, tol):
    """d4b200_eeq_charges_factor_* / d4b200_eeq_vjp_factor_*: the backward pass substitutes its right-hand side
    into the matrix the forward pass eliminated; same result as eliminating again."""
    from tad_dftd4_b200 import eeq as deeq

    numbers, positions, _ = orc.organic_batch([1, 2, 31, 64, 33, 100, 128, 7], seed=11)
    # padding inside the atom axis of one structure
    numbers[2, 5] = 0
    n, p = numbers.to(DEV), positions.to(DEV, dtype)
    charge = torch.tensor([0.0, 1.0, -1.0, 0.0, 2.0, 0.0, 0.0, -1.0], dtype=dtype, device=DEV)
    eng = deeq._EeqEngine.get(DEV)
    q_plain = eng.charges(n, p, charge, 25.0)
    q, factor = eng.charges(n, p, charge, 25.0, keep_factor=True)
    assert factor is not None and factor.dtype == F64 and torch.equal(q, q_plain)
    gq = torch.randn(q.shape, dtype=dtype, device=DEV, generator=torch.Generator(device=DEV).manual_seed(3))
    g_plain = eng.vjp(n, p, 25.0, q, gq)
    g_fac = eng.vjp(n, p, 25.0, q, gq, factor)
    scale = max(1.0, g_plain.abs().max().item())
    assert (g_fac - g_plain).abs().max().item() < tol * scale
    assert (g_fac[numbers.to(DEV) == 0] == 0).all()
    # over the memory limit the forward pass keeps nothing and the backward pass eliminates again
    old = deeq._EeqEngine.FACTOR_LIMIT
    deeq._EeqEngine.FACTOR_LIMIT = 1024
    try:
        q2, none = eng.charges(n, p, charge, 25.0, keep_factor=True)
        assert none is None and torch.equal(q2, q_plain)
        pos = p.clone().requires_grad_(True)
        (g_auto,) = torch.autograd.grad((deeq.get_eeq_charges(n, pos, charge) * gq).sum(), pos)
        assert (g_auto - g_plain).abs().max().item() < tol * scale
    finally:
        deeq._EeqEngine.FACTOR_LIMIT = old
