"""Host-side checks of the EEQ stand-in (no GPU): the dense torch route of
``tad_dftd4_b200.eeq`` (used on device for structures beyond the shared-memory kernels)
reproduces the oracle restatement including its autograd gradient, and CPU tensors are
rejected by the public entry point (no CPU fallback)."""
from __future__ import annotations

import pytest
import torch

import d4_oracle as orc
import eeq_oracle as eeq
from tad_dftd4_b200 import eeq as prod


def test_dense_route_matches_oracle():
    numbers, positions, _ = orc.organic_batch([7, 19, 12], seed=11)
    charge = torch.tensor([0.0, 1.0, -1.0], dtype=torch.float64)
    pos = positions.clone().requires_grad_(True)
    q = prod._dense(numbers, pos, charge, 25.0)
    pos_ref = positions.clone().requires_grad_(True)
    q_ref = eeq.get_eeq_charges(numbers, pos_ref, charge)
    assert (q - q_ref).abs().max().item() < 1e-13
    assert (q.sum(-1) - charge).abs().max().item() < 1e-13
    assert (q[numbers == 0] == 0).all()
    w = torch.randn(q.shape, dtype=q.dtype, generator=torch.Generator().manual_seed(1))
    (g,) = torch.autograd.grad((q * w).sum(), pos)
    (g_ref,) = torch.autograd.grad((q_ref * w).sum(), pos_ref)
    assert (g - g_ref).abs().max().item() < 1e-12


def test_parameter_blob_layout():
    blob = prod._param_blob()
    assert blob.shape == (5, 87)
    assert blob[0, 1] == pytest.approx(1.23695041) and blob[3, 6] == pytest.approx(1.88862966)
    rcov = orc._tables()["rcov"][:87].numpy()
    assert (blob[4] == rcov).all()


def test_cpu_tensors_are_rejected():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        prod.get_eeq_charges(torch.tensor([3, 1]), torch.zeros(2, 3, dtype=torch.float64), 0.0)
    with pytest.raises(ValueError, match="Shape of positions"):
        prod.get_eeq_charges(torch.tensor([3, 1]), torch.zeros(3, 3, dtype=torch.float64), 0.0)
    with pytest.raises(ValueError, match="outside 1..86"):
        prod._dense(torch.tensor([[90, 1]]), torch.zeros(1, 2, 3, dtype=torch.float64), torch.zeros(1, dtype=torch.float64), 25.0)
