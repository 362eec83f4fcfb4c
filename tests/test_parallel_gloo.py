"""Host-side multi-rank logic on CPU (gloo, world_size 2): structure-wise sharding of a
padded batch reproduces the single-process result.  The compute callable is the oracle
here (no GPU in this container); on the GPU box the same wrapper drives the CUDA path."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank: int, world: int, port: int, out_dir: str):
    for p in (ROOT, ROOT / "oracle"):
        sys.path.insert(0, str(p))
    import d4_oracle as orc
    from tad_dftd4_b200.parallel import dftd4_sharded, shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        numbers, positions, q = orc.organic_batch([9, 14, 5, 11, 7], seed=17)
        param = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)

        def compute(n, p, charge, par, q=None):
            return orc.dftd4(n, p, par, q)

        full = dftd4_sharded(numbers, positions, 0.0, param, q=q, compute=compute)
        local = dftd4_sharded(numbers, positions, 0.0, param, q=q, compute=compute, gather=False)
        ref = orc.dftd4(numbers, positions, param, q)
        lo, hi = shard_bounds(5, rank, world)
        assert local.shape[0] == hi - lo
        assert torch.equal(local, ref[lo:hi])
        assert torch.equal(full, ref), (full - ref).abs().max()
        Path(out_dir, f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    from tad_dftd4_b200.parallel import shard_bounds

    for n in (0, 1, 5, 8, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_sharded_batch_two_ranks(tmp_path):
    world = 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
