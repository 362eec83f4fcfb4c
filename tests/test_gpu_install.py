"""
Plugin boundary on the B200 (SURVEY.md 8b secondary): the call sequences of the reference's
``examples/single.py``, ``batch.py``, ``d4s.py`` and ``forces.py`` -- CPU tensors, default
``q=None`` (EEQ), functional and class-based interface -- through the module that
``tad_dftd4_b200.install(device=...)`` registers as ``tad_dftd4``, checked against the numbers
those scripts assert / document; plus the term-level plugin interface
(``FusedD4Term.calculate``) and ``cn_d4`` as a ``CNFunc``.
"""
from __future__ import annotations

import pytest
import torch

import bench_inputs
import d4_oracle as orc
import tad_dftd4_b200 as pkg
from tad_dftd4_b200 import dispersion, ncoord

pytestmark = pytest.mark.gpu

# examples/single.py:52-69 of the reference (TPSSh-D4-ATM, float32 run)
SINGLE_REF = [-0.0020841344, -0.0018971195, -0.0018107513, -0.0018305695, -0.0021737693, -0.0019484236,
              -0.0022788253, -0.0004080658, -0.0004261866, -0.0004199839, -0.0004280768, -0.0005108935]  # fmt: skip
TPSSH = dict(s6=1.0, s8=1.85897750, s9=1.0, a1=0.44286966, a2=4.60230534)


@pytest.fixture()
def d4():
    assert pkg.install(device="cuda:0") == "alias"
    import tad_dftd4

    yield tad_dftd4
    pkg.uninstall()


def _single(dtype):
    return torch.tensor(bench_inputs.SINGLE_Z), torch.tensor(bench_inputs.SINGLE_XYZ, dtype=dtype)


def test_single_example_sequence(d4):
    numbers, positions = _single(torch.float32)  # the script's default dtype, tensors on the CPU
    charge = torch.tensor(0.0)
    param = d4.damping.Param(**{k: positions.new_tensor(v) for k, v in TPSSH.items()})
    energy1 = d4.dftd4(numbers, positions, charge, param)
    energy2 = d4.dispersion.DispD4().calculate(numbers, positions, charge, param)
    ref = torch.tensor(SINGLE_REF)
    assert energy1.device.type == "cpu" and energy1.dtype == torch.float32
    assert torch.allclose(energy1, ref, atol=1e-8), (energy1 - ref).abs().max()
    assert torch.allclose(energy2, ref, atol=1e-8), (energy2 - ref).abs().max()
    assert d4.get_params(method="d4", functional="tpssh") == {k: v for k, v in TPSSH.items() if k in ("s8", "a1", "a2")}


def test_forces_example_sequence(d4):
    numbers, positions = _single(torch.float64)
    charge = torch.tensor(0.0)
    param = d4.get_params(method="d4", functional="tpssh")
    pos = positions.clone().requires_grad_(True)
    energy = d4.dftd4(numbers, pos, charge, param)
    (grad,) = torch.autograd.grad(energy.sum(), pos)
    assert grad.device.type == "cpu"
    num_grad = torch.zeros_like(positions)
    step = 1e-5
    for i in range(numbers.shape[-1]):
        for j in range(3):
            positions[i, j] += step
            e1 = d4.dftd4(numbers, positions, charge, param).sum()
            positions[i, j] -= 2 * step
            e2 = d4.dftd4(numbers, positions, charge, param).sum()
            positions[i, j] += step
            num_grad[i, j] = (e1 - e2) / (2 * step)
    assert torch.allclose(grad, num_grad, atol=1e-8), (grad - num_grad).abs().max()


def test_d4s_example_sequence(d4):
    import eeq_oracle

    numbers, positions = _single(torch.float64)
    model = d4.model.D4SModel(numbers)
    param = d4.get_params(method="d4", functional="tpssh")
    energy = d4.dftd4(numbers, positions, torch.tensor(0.0), param, model=model)
    q = eeq_oracle.get_eeq_charges(numbers, positions, 0.0)
    want = orc.dftd4(numbers, positions, dict(param), q, model="d4s")
    assert torch.allclose(energy, want, rtol=1e-10, atol=1e-16)


def test_batch_example_sequence(d4):
    # S22 system 4 (formamide dimer / monomer): energies documented in examples/batch.py:60-63
    z = [[6, 6, 7, 7, 1, 1, 1, 1, 1, 1, 8, 8], [6, 8, 7, 1, 1, 1, 0, 0, 0, 0, 0, 0]]
    xyz = torch.zeros(2, 12, 3)
    xyz[0] = torch.tensor([
        [-3.81469488143921, +0.09993441402912, 0.0], [+3.81469488143921, -0.09993441402912, 0.0],
        [-2.66030049324036, -2.15898251533508, 0.0], [+2.66030049324036, +2.15898251533508, 0.0],
        [-0.73178529739380, -2.28237795829773, 0.0], [-5.89039325714111, -0.02589114569128, 0.0],
        [-3.71254944801331, -3.73605775833130, 0.0], [+3.71254944801331, +3.73605775833130, 0.0],
        [+0.73178529739380, +2.28237795829773, 0.0], [+5.89039325714111, +0.02589114569128, 0.0],
        [-2.74426102638245, +2.16115570068359, 0.0], [+2.74426102638245, -2.16115570068359, 0.0]])  # fmt: skip
    xyz[1, :6] = torch.tensor([
        [-0.55569743203406, +1.09030425468557, 0.0], [+0.51473634678469, +3.15152550263611, 0.0],
        [+0.59869690244446, -1.16861263789477, 0.0], [-0.45355203669134, -2.74568780438064, 0.0],
        [+2.52721209544999, -1.29200800956867, 0.0], [-2.63139587595376, +0.96447869452240, 0.0]])  # fmt: skip
    numbers = torch.tensor(z)
    param = d4.Param(**{k: xyz.new_tensor(v) for k, v in TPSSH.items()})
    energy = torch.sum(d4.dftd4(numbers, xyz, torch.tensor([0.0, 0.0]), param), -1)
    assert torch.allclose(energy, torch.tensor([-0.0088341432, -0.0027013607]), atol=2e-9), energy
    assert abs(float(energy[0] - 2 * energy[1]) + 0.0034314217) < 3e-9


def test_fused_term_plugin_interface_on_device():
    dev = torch.device("cuda:0")
    numbers, positions, q = (t.to(dev) for t in orc.organic_batch([9, 14, 30], seed=23))
    param = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)
    want = pkg.dftd4(numbers, positions, 0.0, param, q=q)
    cn = ncoord.cn_d4(numbers, positions)
    for terms in ([dispersion.FusedD4Term()], [dispersion.TwoBodyTerm(), dispersion.D4ATMApprox()]):
        energy = torch.zeros_like(q)
        for term in terms:  # the delegate loop of the reference's Disp.calculate (dispersion/base.py:413-431)
            energy = energy + term.calculate(numbers=numbers, positions=positions, param=param, cn=cn, model="d4",
                                             q=q, r4r2=None, rvdw=None, cutoff=None)  # fmt: skip
        assert torch.allclose(energy, want, rtol=1e-12, atol=1e-18)
    # cn_d4 as the CNFunc of Disp(cn_fn=...) (disp.py:117-122 of the reference): against the oracle
    cn_ref = orc.cn_d4(numbers.cpu(), positions.cpu())
    assert torch.allclose(cn.cpu(), cn_ref, rtol=1e-12, atol=1e-14)
    disp = dispersion.DispD4(cn_fn=ncoord.cn_d4)
    assert torch.equal(disp.calculate(numbers, positions, 0.0, param, q=q), want)
