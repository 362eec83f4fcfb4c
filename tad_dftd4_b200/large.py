"""
Single large structures (hundreds to tens of thousands of atoms): host side of the
tiled kernels in ``csrc/d4b200_large.cu``.

The reference materialises ``(N, N, 7, 7)`` and ``(N, N, N)`` tensors
(``/root/reference/README.md:355-357``) and cannot run such systems at all.  Here the
atoms are sorted along a Morton curve, two-body rows and ATM centre groups are split
over the ranks of a ``torch.distributed`` group (cost-balanced by the neighbour counts
of the groups) and the per-atom energies are combined with ONE all-reduce.
"""

from __future__ import annotations

import ctypes as C
from typing import Callable, Sequence

import torch
import torch.distributed as dist

__all__ = ["dftd4_large", "morton_order", "balanced_ranges"]

Tensor = torch.Tensor


def morton_order(positions: Tensor, bits: int = 10) -> Tensor:
    """Permutation that sorts atoms along a Z-order curve (spatial coherence for the
    tiled kernels; any order is correct)."""
    p = positions.detach().to(torch.float64)
    lo = p.min(dim=0).values
    span = (p.max(dim=0).values - lo).clamp_min(1e-9)
    cells = ((p - lo) / span * (2**bits - 1)).round().to(torch.int64)
    code = torch.zeros(p.shape[0], dtype=torch.int64, device=p.device)
    for b in range(bits):
        for d in range(3):
            code |= ((cells[:, d] >> b) & 1) << (3 * b + d)
    return torch.argsort(code)


def balanced_ranges(cost: Sequence[float] | Tensor, world: int) -> list[tuple[int, int]]:
    """Contiguous ranges of groups with (nearly) equal summed cost for ``world`` ranks."""
    c = torch.as_tensor(cost, dtype=torch.float64).flatten().cpu()
    n = c.numel()
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * max(0, world - 1)
    csum = torch.cumsum(c, 0)
    total = float(csum[-1]) if n else 0.0
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        cut = int(torch.searchsorted(csum, torch.tensor(target, dtype=torch.float64)).item())
        bounds.append(max(bounds[-1], min(n, cut + 1 if total > 0 else 0)))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def _kernel_compute(engine, par, numbers, positions, q, rows, groups, want_cost):
    """Call the C ABI for one rank's share; returns (partial energy | None, group cost | None)."""
    lib = engine.lib
    nat = numbers.shape[0]
    fp32 = positions.dtype == torch.float32
    need = int(lib.d4b200_large_workspace_bytes(engine.handle, nat, int(fp32)))
    ws = engine.large_workspace(need)
    gs = int(lib.d4b200_large_group_size())
    ng = (nat + gs - 1) // gs
    energy = None if want_cost else torch.zeros(nat, dtype=positions.dtype, device=positions.device)
    cost = torch.empty(ng, dtype=torch.int32, device=positions.device) if want_cost else None
    fn = lib.d4b200_large_energy_f32 if fp32 else lib.d4b200_large_energy_f64
    stream = torch.cuda.current_stream(positions.device).cuda_stream
    from . import _lib

    _lib.check(
        fn(engine.handle, C.byref(par), nat, numbers.data_ptr(), positions.data_ptr(), q.data_ptr(),
           rows[0], rows[1], groups[0], groups[1],
           energy.data_ptr() if energy is not None else None, None,
           cost.data_ptr() if cost is not None else None, ws.data_ptr(), ws.numel(), stream),
        "d4b200_large_energy",
    )  # fmt: skip
    return energy, cost


def dftd4_large(
    numbers: Tensor,
    positions: Tensor,
    param,
    q: Tensor,
    *,
    cutoff=None,
    group=None,
    compute: Callable | None = None,
    group_size: int | None = None,
) -> Tensor:
    """Atom-resolved D4 energy of ONE structure ``(nat,)`` with the tiled kernels.

    With an initialised ``torch.distributed`` process group every rank passes the same
    (replicated) inputs, evaluates its share of two-body rows / ATM centre groups and
    the result is all-reduced, so every rank returns the full energy vector.

    ``compute(numbers, positions, q, rows, groups, want_cost)`` is injectable for CPU
    tests of this host logic.
    """
    if numbers.dim() != 1 or positions.shape != (numbers.shape[0], 3) or q.shape != numbers.shape:
        raise ValueError("dftd4_large expects numbers (nat,), positions (nat, 3), q (nat,)")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0

    real = numbers != 0
    keep = torch.nonzero(real).flatten()
    order = keep[morton_order(positions[keep])]
    num_s = numbers[order].to(torch.int64).contiguous()
    pos_s = positions[order].detach().contiguous()
    q_s = q[order].detach().to(positions.dtype).contiguous()
    nat = int(order.numel())
    out = torch.zeros(numbers.shape[0], dtype=positions.dtype, device=positions.device)
    if nat == 0:
        return out

    if compute is None:
        from . import defaults
        from .disp import _Engine, _flatten_param

        engine = _Engine.get(positions.device, defaults.GA_DEFAULT, defaults.GC_DEFAULT)
        par = _flatten_param(param, cutoff, 0, defaults.WF_DEFAULT)
        gs = int(engine.lib.d4b200_large_group_size())

        def compute(n, p, qq, rows, groups, want_cost):
            with torch.cuda.device(p.device):
                return _kernel_compute(engine, par, n, p, qq, rows, groups, want_cost)
    else:
        gs = group_size or 16
    ng = (nat + gs - 1) // gs

    if world > 1:
        # cost model: a centre group with m neighbours evaluates ~m^2/2 pairs x group size
        _, cost = compute(num_s, pos_s, q_s, (0, 0), (0, 0), True)
        granges = balanced_ranges(cost.to(torch.float64) ** 2, world)
        g0, g1 = granges[rank]
        # two-body rows follow the same blocks (cost ~ rows)
        r0, r1 = min(nat, g0 * gs), min(nat, g1 * gs)
        if rank == world - 1:
            r1 = nat
    else:
        g0, g1, r0, r1 = 0, ng, 0, nat
    e_s, _ = compute(num_s, pos_s, q_s, (r0, r1), (g0, g1), False)
    if world > 1:
        dist.all_reduce(e_s, op=dist.ReduceOp.SUM, group=group)
    out[order] = e_s
    return out
