#!/usr/bin/env python
"""
Cross-rank consistency of the row-block partitioned large-system path on hardware (run under
torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/c4_crossrank.py [nmol]

Every rank evaluates energy + gradient of the same water cluster (a) partitioned over all N ranks
(NCCL all-reduces) and (b) alone (a one-rank group); rank 0 prints the largest differences.
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench_inputs  # noqa: E402
from tad_dftd4_b200.large import clear_plan_cache, dftd4_large  # noqa: E402

PBE0 = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)


def main():
    nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 6667
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    solo = [dist.new_group([r]) for r in range(world)][rank]  # every rank creates every group
    numbers, positions, q = (t.to(dev) for t in bench_inputs.water_cluster(nmol, 4))

    def run(group):
        clear_plan_cache()
        pos = positions.detach().requires_grad_(True)
        e = dftd4_large(numbers, pos, PBE0, q, group=group)
        (g,) = torch.autograd.grad(e.sum(), pos)
        return e.detach(), g

    e_all, g_all = run(None)
    e_one, g_one = run(solo)
    de = ((e_all - e_one).abs().max() / e_one.abs().max()).item()
    dg = (g_all - g_one).abs().max().item()
    stats = torch.tensor([de, dg], dtype=torch.float64, device=dev)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"atoms": int(numbers.shape[0]), "ranks": world, "max_rel_energy_diff": stats[0].item(),
                          "max_abs_gradient_diff": stats[1].item(), "energy_sum": e_all.sum().item()}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    assert stats[0].item() < 1e-12 and stats[1].item() < 1e-12, stats.tolist()


if __name__ == "__main__":
    main()
