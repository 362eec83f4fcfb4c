"""The blocked / pruned large-structure oracle (oracle/d4_oracle_blocked.py) against the dense
oracle (bit-identical to the unmodified reference on the golden cases) on structures both can
hold: default cutoffs and tight cutoffs (open triples, pairs beyond the two-body cutoff)."""
from __future__ import annotations

import numpy as np
import pytest
import torch

import bench_inputs
import d4_oracle as orc
import d4_oracle_blocked as blk

PBE0 = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)


def _organic(n, seed):
    z, xyz, q = bench_inputs.organic_blob(n, np.random.default_rng(seed))
    return torch.from_numpy(z), torch.from_numpy(xyz), torch.from_numpy(q)


@pytest.mark.parametrize("case", ["organic", "water"])
@pytest.mark.parametrize("cut", [(60.0, 40.0), (14.0, 9.0)])
def test_blocked_matches_dense(case, cut):
    numbers, positions, q = _organic(47, 5) if case == "organic" else bench_inputs.water_cluster(30, 8)
    param = dict(PBE0, s10=0.3) if case == "organic" else PBE0
    e_ref, g_ref = orc.energy_and_gradient(numbers, positions, param, q, disp2=cut[0], disp3=cut[1])
    e, g = blk.dftd4_blocked(numbers, positions, param, q, disp2=cut[0], disp3=cut[1], block=16, block3=5,
                             gradient=True)  # fmt: skip
    assert torch.allclose(e, e_ref, rtol=1e-12, atol=1e-18), ((e - e_ref).abs() / e_ref.abs()).max()
    assert (g - g_ref).abs().max() < 1e-15
    rows = torch.tensor([3, 17, 40])
    e_rows = blk.dftd4_blocked(numbers, positions, param, q, rows=rows, disp2=cut[0], disp3=cut[1])
    assert torch.allclose(e_rows, e_ref[rows], rtol=1e-12, atol=1e-18)
