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
