"""Tiled large-system path (csrc/d4b200_large.cu) vs the oracle, evaluated live on the
host cores at sizes the dense formulation still handles."""
from __future__ import annotations

import numpy as np
import pytest
import torch

import d4_oracle as orc

pytestmark = pytest.mark.gpu
PBE0 = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)


def _cluster(nat, seed, spacing=5.0):
    """Jittered lattice of H/C/N/O atoms (bulk-like density, physical CNs)."""
    rng = np.random.default_rng(seed)
    m = int(np.ceil(nat ** (1 / 3))) + 1
    grid = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)
    grid = grid[np.argsort(np.linalg.norm(grid - m / 2, axis=1))][:nat]
    xyz = grid * spacing + rng.normal(scale=0.35, size=(nat, 3))
    z = rng.choice([1, 6, 7, 8], size=nat, p=[0.5, 0.3, 0.1, 0.1])
    q = 0.1 * rng.normal(size=nat)
    q -= q.mean()
    return torch.from_numpy(z.astype(np.int64)), torch.from_numpy(xyz), torch.from_numpy(q)


@pytest.mark.parametrize("nat,cut", [(150, {}), (260, {}), (200, dict(disp2=22.0, disp3=13.0)),
                                     (330, dict(disp2=30.0, disp3=17.0))])  # fmt: skip
def test_large_energy_f64(nat, cut):
    import tad_dftd4_b200 as d4
    from tad_dftd4_b200.large import dftd4_large

    numbers, positions, q = _cluster(nat, seed=nat)
    ref = orc.dftd4(numbers, positions, PBE0, q, **cut)
    dev = torch.device("cuda:0")
    cutoff = d4.Cutoff(**cut, device=dev, dtype=torch.float64) if cut else None
    e = dftd4_large(numbers.to(dev), positions.to(dev), PBE0, q.to(dev), cutoff=cutoff).cpu()
    assert (e - ref).abs().max() / ref.abs().max() < 1e-10
    assert abs(e.sum() - ref.sum()) <= 1e-10 * abs(ref.sum())


def test_dftd4_routes_large_structures():
    """A padded batch mixing one structure beyond the one-CTA kernels with small ones."""
    import tad_dftd4_b200 as d4

    nb, pb, qb = _cluster(170, seed=3)
    ns, ps, qs = orc.organic_batch([20, 33], seed=9)
    width = 180
    numbers = torch.zeros((3, width), dtype=torch.int64)
    positions = torch.zeros((3, width, 3), dtype=torch.float64)
    q = torch.zeros((3, width), dtype=torch.float64)
    numbers[0, 5:175], positions[0, 5:175], q[0, 5:175] = nb, pb, qb  # holes at both ends
    numbers[1:, : ns.shape[1]], positions[1:, : ns.shape[1]], q[1:, : ns.shape[1]] = ns, ps, qs
    ref = orc.dftd4(numbers, positions, PBE0, q)
    dev = torch.device("cuda:0")
    e = d4.dftd4(numbers.to(dev), positions.to(dev), 0.0, PBE0, q=q.to(dev)).cpu()
    assert (e - ref).abs().max() / ref.abs().max() < 1e-10
    assert torch.all(e[numbers == 0] == 0)


@pytest.mark.parametrize("nat,cut", [(150, {}), (230, dict(disp2=22.0, disp3=13.0))])
def test_large_gradient_f64(nat, cut):
    """Tiled analytic gradient (two stages) vs oracle autograd, incl. dL/dq and weights g."""
    import tad_dftd4_b200 as d4
    from tad_dftd4_b200.large import dftd4_large

    numbers, positions, q = _cluster(nat, seed=100 + nat)
    g = torch.from_numpy(np.random.default_rng(1).normal(size=nat))
    pos = positions.clone().requires_grad_(True)
    qq = q.clone().requires_grad_(True)
    e_ref = orc.dftd4(numbers, pos, PBE0, qq, **cut)
    gp_ref, gq_ref = torch.autograd.grad((e_ref * g).sum(), (pos, qq))

    dev = torch.device("cuda:0")
    cutoff = d4.Cutoff(**cut, device=dev, dtype=torch.float64) if cut else None
    pd = positions.to(dev).requires_grad_(True)
    qd = q.to(dev).requires_grad_(True)
    e = dftd4_large(numbers.to(dev), pd, PBE0, qd, cutoff=cutoff)
    gp, gq = torch.autograd.grad((e * g.to(dev)).sum(), (pd, qd))
    assert (e.detach().cpu() - e_ref.detach()).abs().max() / e_ref.detach().abs().max() < 1e-10
    assert (gp.cpu() - gp_ref).abs().max() < 1e-9
    assert (gq.cpu() - gq_ref).abs().max() < 1e-9


def test_dftd4_large_gradient_through_public_api():
    import tad_dftd4_b200 as d4

    numbers, positions, q = _cluster(140, seed=8)
    pos = positions.clone().requires_grad_(True)
    (g_ref,) = torch.autograd.grad(orc.dftd4(numbers, pos, PBE0, q).sum(), pos)
    dev = torch.device("cuda:0")
    pd = positions.to(dev).requires_grad_(True)
    e = d4.dftd4(numbers.to(dev), pd, 0.0, PBE0, q=q.to(dev))
    (g,) = torch.autograd.grad(e.sum(), pd)
    assert (g.cpu() - g_ref).abs().max() < 1e-9
