"""
Multi-GPU host logic (one process per GPU, ``torch.distributed``).

* Padded batches shard structure-wise: structures are independent (no term of
  the reference couples them, masks are per structure), so every rank evaluates
  a contiguous slice with **no data-path collective**; :func:`dftd4_sharded`
  only gathers the per-atom energies at the end if asked to.
* Single large systems use the row-block partition of :mod:`tad_dftd4_b200.large`
  (centre atoms split over ranks, one all-reduce of the energies).

The compute callable is injectable so that the host logic is testable on CPU
with the ``gloo`` backend (``tests/test_parallel_gloo.py``).
"""

from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "dftd4_sharded"]


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced slice ``[begin, end)`` of ``n`` items for ``rank``."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"invalid rank/world: {rank}/{world}")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def dftd4_sharded(
    numbers: torch.Tensor,
    positions: torch.Tensor,
    charge,
    param,
    *,
    q: torch.Tensor,
    gather: bool = True,
    group=None,
    compute: Callable | None = None,
    **kwargs,
) -> torch.Tensor:
    """Evaluate a padded batch ``(B, N)`` with the structures sharded over the ranks
    of ``group``.  Returns the full ``(B, N)`` energy on every rank (``gather=True``)
    or only this rank's slice."""
    if compute is None:
        from .disp import dftd4 as compute  # the CUDA path
    if numbers.dim() != 2:
        raise ValueError("dftd4_sharded expects a padded batch of shape (B, N)")
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(numbers.shape[0], rank, world)
    chg = charge[lo:hi] if isinstance(charge, torch.Tensor) and charge.dim() > 0 else charge
    local = compute(numbers[lo:hi], positions[lo:hi], chg, param, q=q[lo:hi], **kwargs)
    if not gather or world == 1:
        return local
    # all_gather needs equal shapes: pad every shard to the largest one
    width = -(-numbers.shape[0] // world)
    buf = local.new_zeros((width, numbers.shape[1]))
    buf[: hi - lo] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = [parts[r][: shard_bounds(numbers.shape[0], r, world)[1] - shard_bounds(numbers.shape[0], r, world)[0]]
           for r in range(world)]  # fmt: skip
    return torch.cat(out, dim=0)
