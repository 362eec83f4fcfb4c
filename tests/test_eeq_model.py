"""CPU check of the EEQ kernel model (tests/eeq_model.py) against the dense oracle
(oracle/eeq_oracle.py, autograd for the vector-Jacobian product)."""
from __future__ import annotations

import numpy as np
import pytest
import torch

import d4_oracle as orc
import eeq_model as km
import eeq_oracle as eeq


def _par():
    p = eeq._param()
    d = {k: v.numpy() for k, v in p.items()}
    d["rcov"] = orc._tables()["rcov"].numpy()
    return d


@pytest.mark.parametrize("nat,charge,seed", [(2, 0.0, 1), (9, 1.0, 2), (33, -1.0, 3), (60, 0.0, 4)])
def test_forward_and_vjp(nat, charge, seed):
    numbers, positions, _ = orc.organic_batch([nat], seed=seed)
    z, xyz = numbers[0], positions[0]
    # padding in the middle of the atom axis must not matter
    z = torch.cat([z[:1], torch.zeros(2, dtype=z.dtype), z[1:]])
    xyz = torch.cat([xyz[:1], torch.zeros(2, 3, dtype=xyz.dtype), xyz[1:]])
    pos = xyz.clone().requires_grad_(True)
    q_ref = eeq.get_eeq_charges(z, pos, charge)
    gq = torch.from_numpy(np.random.default_rng(seed).standard_normal(len(z)))
    (g_ref,) = torch.autograd.grad((q_ref * gq).sum(), pos)

    par = _par()
    q = km.eeq_forward(z.numpy(), xyz.numpy(), charge, par)
    assert np.abs(q - q_ref.detach().numpy()).max() < 1e-13
    assert abs(q.sum() - charge) < 1e-13
    g = km.eeq_backward(z.numpy(), xyz.numpy(), q, gq.numpy(), par)
    assert np.abs(g - g_ref.numpy()).max() < 1e-12
    assert (g[1:3] == 0).all()
