"""
BASELINE.json's configurations at FULL size (C2: 4096 structures of 20-60 atoms, energy;
C3: 1024 structures of 100 atoms, energy + forces; C5: the C3 inputs with the D4S model,
FP32 against FP64), checked through size-independent properties of the dispersion energy
-- the dense oracle needs minutes for a whole batch, so it checks a sample:

* a sample of the batch against the float64 oracle (energy <= 1e-10 relative, gradient
  <= 1e-9 absolute: north_star tolerances);
* a structure's result does not depend on its position in the batch, on the batch it is
  in, or on the padded width (bitwise);
* rigid translation + rotation leaves atomic energies unchanged and rotates the forces;
* forces sum to zero and exert no torque (translational / rotational invariance of E);
* the analytic gradient agrees with a central difference of the ENERGY kernel along a random
  direction (two independent kernels);
* the host-buffer entry point returns the same bits as the device entry point;
* FP32 mode stays within 1e-5 relative of FP64 (energies) at full size.
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

import d4_oracle as orc

pytestmark = pytest.mark.gpu

PBE0 = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)  # d4.toml:269 of the reference


def _batch(nbatch, lo, hi, seed):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(lo, hi + 1, size=nbatch)
    return orc.organic_batch_parallel(sizes, seed=seed)


@pytest.fixture(scope="module")
def c2():
    return _batch(4096, 20, 60, 2)


@pytest.fixture(scope="module")
def c3():
    return _batch(1024, 100, 100, 3)


def _d4():
    import tad_dftd4_b200 as d4

    return d4


def _rotation(seed):
    g = torch.Generator().manual_seed(seed)
    qm, r = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64, generator=g))
    qm = qm * torch.sign(torch.diagonal(r))
    if torch.det(qm) < 0:
        qm[:, 0] = -qm[:, 0]
    return qm


def test_c2_full_batch_sample_against_oracle_and_batch_independence(c2):
    d4 = _d4()
    dev = torch.device("cuda:0")
    numbers_h, positions_h, q_h = c2
    assert numbers_h.shape == (4096, 60)
    numbers, positions, q = numbers_h.to(dev), positions_h.to(dev), q_h.to(dev)
    e = d4.dftd4(numbers, positions, 0.0, PBE0, q=q)
    assert e.shape == numbers.shape and e.dtype == torch.float64
    assert torch.all(e[numbers == 0] == 0.0)
    assert torch.isfinite(e).all() and (e.sum(-1) < 0).all()

    # oracle on a spread sample of the batch
    idx = torch.arange(0, 4096, 4096 // 24)
    e_ref = orc.dftd4(numbers_h[idx], positions_h[idx], PBE0, q_h[idx])
    scale = e_ref.abs().amax(-1, keepdim=True)
    assert ((e[idx.to(dev)].cpu() - e_ref).abs() / scale).max() < 1e-10
    tot, rtot = e[idx.to(dev)].sum(-1).cpu(), e_ref.sum(-1)
    assert ((tot - rtot).abs() <= 1e-10 * rtot.abs()).all()

    # the same structures alone, in another order, and with a wider padding: same bits
    sub = torch.randperm(4096, generator=torch.Generator().manual_seed(5))[:333].to(dev)
    assert torch.equal(d4.dftd4(numbers[sub], positions[sub], 0.0, PBE0, q=q[sub]), e[sub])
    perm = torch.randperm(4096, generator=torch.Generator().manual_seed(6)).to(dev)
    assert torch.equal(d4.dftd4(numbers[perm], positions[perm], 0.0, PBE0, q=q[perm]), e[perm])
    pad = 7
    nw = torch.nn.functional.pad(numbers, (0, pad))
    pw = torch.nn.functional.pad(positions, (0, 0, 0, pad))
    qw = torch.nn.functional.pad(q, (0, pad))
    ew = d4.dftd4(nw, pw, 0.0, PBE0, q=qw)
    assert torch.equal(ew[:, :60], e) and torch.all(ew[:, 60:] == 0.0)


def test_c2_full_batch_rigid_motion_invariance_and_host_entry(c2):
    d4 = _d4()
    dev = torch.device("cuda:0")
    numbers_h, positions_h, q_h = c2
    numbers, positions, q = numbers_h.to(dev), positions_h.to(dev), q_h.to(dev)
    e = d4.dftd4(numbers, positions, 0.0, PBE0, q=q)
    rot = _rotation(1).to(dev)
    shift = torch.tensor([3.1, -7.2, 11.3], dtype=torch.float64, device=dev)
    moved = torch.where((numbers != 0).unsqueeze(-1), positions @ rot.T + shift, torch.zeros_like(positions))
    e2 = d4.dftd4(numbers, moved, 0.0, PBE0, q=q)
    scale = e.abs().amax(-1, keepdim=True)
    assert ((e2 - e).abs() / scale).max() < 1e-11

    out = d4.dftd4_host(numbers_h.pin_memory(), positions_h.pin_memory(), 0.0, PBE0, q=q_h.pin_memory())
    assert torch.equal(out, e.cpu())


def test_c3_full_batch_forces(c3):
    d4 = _d4()
    dev = torch.device("cuda:0")
    numbers_h, positions_h, q_h = c3
    assert numbers_h.shape == (1024, 100)
    numbers, positions, q = numbers_h.to(dev), positions_h.to(dev), q_h.to(dev)
    pos = positions.clone().requires_grad_(True)
    e = d4.dftd4(numbers, pos, 0.0, PBE0, q=q)
    (g,) = torch.autograd.grad(e.sum(), pos)
    assert torch.isfinite(g).all()

    # sample against oracle autograd
    idx = torch.tensor([0, 341, 682, 1023])
    e_ref, g_ref = orc.energy_and_gradient(numbers_h[idx], positions_h[idx], PBE0, q_h[idx])
    scale = e_ref.abs().amax(-1, keepdim=True)
    assert ((e[idx.to(dev)].detach().cpu() - e_ref).abs() / scale).max() < 1e-10
    assert (g[idx.to(dev)].cpu() - g_ref).abs().max() < 1e-9

    # no net force, no net torque on any of the 1024 structures
    gmax = g.abs().amax((-2, -1))
    assert (g.sum(1).abs().amax(-1) / gmax).max() < 1e-11
    torque = torch.linalg.cross(positions, g, dim=-1).sum(1)
    lever = positions.norm(dim=-1).amax(-1)
    assert (torque.abs().amax(-1) / (gmax * lever)).max() < 1e-10

    # forces rotate with the structure
    rot = _rotation(2).to(dev)
    pos_r = (positions @ rot.T).requires_grad_(True)
    (g_r,) = torch.autograd.grad(d4.dftd4(numbers, pos_r, 0.0, PBE0, q=q).sum(), pos_r)
    assert ((g_r - g @ rot.T).abs().amax((-2, -1)) / gmax).max() < 1e-10

    # directional derivative from the ENERGY kernel (central difference) vs the gradient kernel
    gen = torch.Generator().manual_seed(9)
    d = torch.randn(positions.shape, dtype=torch.float64, generator=gen).to(dev)
    d = d / d.norm(dim=(-2, -1), keepdim=True)
    h = 1e-4
    ep = d4.dftd4(numbers, positions + h * d, 0.0, PBE0, q=q).sum(-1)
    em = d4.dftd4(numbers, positions - h * d, 0.0, PBE0, q=q).sum(-1)
    fd = (ep - em) / (2 * h)
    an = (g * d).sum((-2, -1))
    assert ((fd - an).abs() / (g.norm(dim=(-2, -1)) + 1e-300)).max() < 1e-5

    # host-buffer forces entry (chunked H2D / kernels / D2H pipeline): same bits
    e_h, g_h = d4.dftd4_host(numbers_h.pin_memory(), positions_h.pin_memory(), 0.0, PBE0, q=q_h.pin_memory(),
                             with_gradient=True)
    assert torch.equal(e_h, e.detach().cpu()) and torch.equal(g_h, g.cpu())

    # the fused forward (energy + gradient in one launch) and the energy-only kernel agree
    e_only = d4.dftd4(numbers, positions, 0.0, PBE0, q=q)
    assert ((e.detach() - e_only).abs() / e_only.abs().amax(-1, keepdim=True)).max() < 1e-13


@pytest.mark.parametrize("model", ["d4", "d4s"])
def test_c5_full_batch_fp32_against_fp64(c3, model):
    """C5: FP32 mode against FP64 on the full C3 inputs, tolerance reported by north_star
    (<= 1e-5 relative in the energy); gradients <= 1e-5 of the largest component."""
    d4 = _d4()
    dev = torch.device("cuda:0")
    numbers_h, positions_h, q_h = c3
    numbers, positions, q = numbers_h.to(dev), positions_h.to(dev), q_h.to(dev)
    p64 = positions.clone().requires_grad_(True)
    e64 = d4.dftd4(numbers, p64, 0.0, PBE0, q=q, model=model)
    (g64,) = torch.autograd.grad(e64.sum(), p64)
    p32 = positions.float().requires_grad_(True)
    e32 = d4.dftd4(numbers, p32, 0.0, PBE0, q=q.float(), model=model)
    (g32,) = torch.autograd.grad(e32.sum(), p32)
    assert e32.dtype == torch.float32 and g32.dtype == torch.float32
    rel = (e32.double().sum(-1) - e64.detach().sum(-1)).abs() / e64.detach().sum(-1).abs()
    assert rel.max() < 1e-5
    assert ((g32.double() - g64).abs().amax((-2, -1)) / g64.abs().amax((-2, -1))).max() < 1e-4

    if model == "d4s":  # sample of the D4S batch against the oracle
        idx = torch.tensor([5, 700])
        e_ref, g_ref = orc.energy_and_gradient(numbers_h[idx], positions_h[idx], PBE0, q_h[idx], model="d4s")
        scale = e_ref.abs().amax(-1, keepdim=True)
        assert ((e64[idx.to(dev)].detach().cpu() - e_ref).abs() / scale).max() < 1e-10
        assert (g64[idx.to(dev)].cpu() - g_ref).abs().max() < 1e-9
