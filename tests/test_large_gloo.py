"""Row-block partition of a single large structure over two ranks (CPU, gloo): the host
logic of tad_dftd4_b200.large (Morton sort, cost-balanced ranges, all-reduce) with the
oracle standing in for the kernels."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
PARAM = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)
GS = 4  # centre-group size used by the stand-in


def _oracle_compute(orc):
    def compute(n, p, q, rows, groups, want_cost):
        nat = n.shape[0]
        ng = (nat + GS - 1) // GS
        if want_cost:
            d = torch.cdist(p, p)
            near = (d <= 40.0).sum(-1).to(torch.float64)
            cost = torch.stack([near[g * GS : (g + 1) * GS].sum() for g in range(ng)])
            return None, cost
        centres = torch.zeros(nat, dtype=torch.bool)
        centres[groups[0] * GS : min(nat, groups[1] * GS)] = True
        e2, e3, *_ = orc.dftd4(n, p, PARAM, q, parts=True, centres=centres)
        rowmask = torch.zeros(nat, dtype=torch.bool)
        rowmask[rows[0] : rows[1]] = True
        return torch.where(rowmask, e2, torch.zeros_like(e2)) + e3, None

    return compute


def _worker(rank, world, port, out_dir):
    for p in (ROOT, ROOT / "oracle"):
        sys.path.insert(0, str(p))
    import d4_oracle as orc
    from tad_dftd4_b200.large import dftd4_large

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z, xyz, q = orc.organic_blob(26, np.random.default_rng(5))
        numbers = torch.from_numpy(np.concatenate([z[:10], [0, 0], z[10:]]))  # padding holes
        positions = torch.from_numpy(np.concatenate([xyz[:10], np.zeros((2, 3)), xyz[10:]]))
        qq = torch.from_numpy(np.concatenate([q[:10], [0, 0], q[10:]]))
        e = dftd4_large(numbers, positions, PARAM, qq, compute=_oracle_compute(orc), group_size=GS)
        ref = orc.dftd4(numbers, positions, PARAM, qq)
        assert (e - ref).abs().max() < 1e-15, (e - ref).abs().max()
        Path(out_dir, f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_balanced_ranges():
    from tad_dftd4_b200.large import balanced_ranges

    r = balanced_ranges([1, 1, 1, 1, 4, 4], 2)
    assert r[0][0] == 0 and r[-1][1] == 6 and r[0][1] == r[1][0]
    assert abs(sum([1, 1, 1, 1, 4, 4][r[0][0] : r[0][1]]) - 6) <= 4
    assert balanced_ranges([], 3) == [(0, 0), (0, 0), (0, 0)]
    spans = balanced_ranges(np.ones(10), 4)
    assert spans[0][0] == 0 and spans[-1][1] == 10
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_morton_order_is_permutation():
    from tad_dftd4_b200.large import morton_order

    p = torch.rand(100, 3, dtype=torch.float64) * 50
    o = morton_order(p)
    assert sorted(o.tolist()) == list(range(100))


def test_rowblock_two_ranks(tmp_path):
    world = 2
    port = 31500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
