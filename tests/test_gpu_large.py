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


@pytest.mark.parametrize("nat,cut", [(200, {}), (230, dict(disp2=22.0, disp3=13.0))])
def test_large_gradient_f32(nat, cut):
    """FP32 mode of the tiled family (float instantiations of the gradient kernels incl. the bulk-async tile
    pipeline) against the float64 oracle: north_star's 1e-5 relative; fused energies and a weighted upstream."""
    import tad_dftd4_b200 as d4
    from tad_dftd4_b200.large import dftd4_large

    numbers, positions, q = _cluster(nat, seed=300 + nat)
    g = torch.from_numpy(np.random.default_rng(2).normal(size=nat))
    pos = positions.clone().requires_grad_(True)
    e_ref = orc.dftd4(numbers, pos, PBE0, q, **cut)
    (gp_ref,) = torch.autograd.grad((e_ref * g).sum(), pos)
    (gs_ref,) = torch.autograd.grad(orc.dftd4(numbers, pos, PBE0, q, **cut).sum(), pos)

    dev = torch.device("cuda:0")
    f32 = torch.float32
    cutoff = d4.Cutoff(**cut, device=dev, dtype=f32) if cut else None
    pd = positions.to(dev, f32).requires_grad_(True)
    e = dftd4_large(numbers.to(dev), pd, PBE0, q.to(dev, f32), cutoff=cutoff)
    assert e.dtype == f32
    (gp,) = torch.autograd.grad((e * g.to(dev, f32)).sum(), pd, retain_graph=True)
    (gs,) = torch.autograd.grad(e.sum(), pd)
    er = e_ref.detach()
    assert abs(e.double().sum().item() - er.sum().item()) <= 1e-5 * abs(er.sum().item())
    assert (e.detach().cpu().double() - er).abs().max() / er.abs().max() < 5e-5
    assert (gp.cpu().double() - gp_ref).abs().max() < 5e-5 * max(gp_ref.abs().max().item(), 1e-3)
    assert (gs.cpu().double() - gs_ref).abs().max() < 5e-5 * max(gs_ref.abs().max().item(), 1e-3)


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


# --------------------------------------------------------------------------------------------
# Parity at scale with the reference's DEFAULT cutoffs (60 / 40 / 30 Bohr): the blocked oracle
# (oracle/d4_oracle_blocked.py, pinned against the dense oracle in tests/test_oracle_blocked.py)
# live at 1.2 k atoms incl. the gradient, golden vectors (oracle/make_golden_large.py) at 5 k atoms
# and on sampled atoms of the full 20 001-atom C4 cluster.
# --------------------------------------------------------------------------------------------
GOLDEN_LARGE = __import__("pathlib").Path(__file__).resolve().parent / "golden" / "large"


def _gpu_energy_gradient(numbers, positions, q, with_gradient=True):
    from tad_dftd4_b200.large import dftd4_large

    dev = torch.device("cuda:0")
    pos = positions.to(dev).requires_grad_(with_gradient)
    e = dftd4_large(numbers.to(dev), pos, PBE0, q.to(dev))
    g = torch.autograd.grad(e.sum(), pos)[0].cpu() if with_gradient else None
    return e.detach().cpu(), g


def test_rod_1200_energy_and_gradient_default_cutoffs():
    """400 H2O in a 3 x 3 column (260 Bohr long): pairs beyond every cutoff, open triples at 40 Bohr."""
    import bench_inputs
    import d4_oracle_blocked as blk

    numbers, positions, q = bench_inputs.water_cluster(400, 11, rod=(3, 3))
    e_ref, g_ref = blk.dftd4_blocked(numbers, positions, PBE0, q, gradient=True, block3=8)
    e, g = _gpu_energy_gradient(numbers, positions, q)
    assert (e - e_ref).abs().max() / e_ref.abs().max() < 1e-10
    assert abs(e.sum() - e_ref.sum()) <= 1e-10 * abs(e_ref.sum())
    assert (g - g_ref).abs().max() < 1e-9


def test_rod_5001_energy_golden():
    import bench_inputs

    gold = np.load(GOLDEN_LARGE / "rod5001.npz")
    numbers, positions, q = bench_inputs.water_cluster(int(gold["nmol"]), int(gold["seed"]), rod=tuple(gold["rod"]))
    e_ref = torch.from_numpy(gold["energy"])
    e, g = _gpu_energy_gradient(numbers, positions, q)
    assert (e - e_ref).abs().max() / e_ref.abs().max() < 1e-10
    assert abs(e.sum() - e_ref.sum()) <= 1e-10 * abs(e_ref.sum())
    # the gradient at this size: translation / rotation invariance and one finite difference of the energy
    assert g.sum(0).abs().max() < 1e-9 and torch.linalg.cross(positions, g).sum(0).abs().max() < 1e-7
    i, h = 2500, 1e-4
    ep = [_gpu_energy_gradient(numbers, positions + s * h * torch.eye(3)[0] * (torch.arange(numbers.shape[0]) == i)[:, None],
                               q, False)[0].sum().item() for s in (1.0, -1.0)]  # fmt: skip
    assert abs((ep[0] - ep[1]) / (2 * h) - g[i, 0].item()) < 1e-8


def test_c4_full_size_sampled_atoms_golden():
    """BASELINE config C4 itself (20 001 atoms, default cutoffs): atom-resolved energies of 12 sampled
    atoms against the blocked oracle; energy and gradient consistency at full size."""
    import bench_inputs

    gold = np.load(GOLDEN_LARGE / "c4_samples.npz")
    numbers, positions, q = bench_inputs.water_cluster(int(gold["nmol"]), int(gold["seed"]))
    rows, e_ref = torch.from_numpy(gold["rows"]), torch.from_numpy(gold["energy"])
    e, g = _gpu_energy_gradient(numbers, positions, q)
    assert ((e[rows] - e_ref).abs() / e_ref.abs()).max() < 1e-10
    assert g.sum(0).abs().max() < 1e-8 and torch.isfinite(g).all()
    i, h = int(rows[5]), 1e-4
    ep = [_gpu_energy_gradient(numbers, positions + s * h * torch.eye(3)[1] * (torch.arange(numbers.shape[0]) == i)[:, None],
                               q, False)[0].sum().item() for s in (1.0, -1.0)]  # fmt: skip
    assert abs((ep[0] - ep[1]) / (2 * h) - g[i, 1].item()) < 1e-7


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_cross_rank_equals_single_rank():
    """Energy and gradient of a 6 000-atom cluster partitioned over all GPUs of the box (NCCL
    all-reduces) against the single-rank evaluation: <= 1e-12 (tools/c4_crossrank.py)."""
    import json
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    n = min(torch.cuda.device_count(), 8)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(root / "tools" / "c4_crossrank.py"),
                          "2000"], capture_output=True, text=True, timeout=900, cwd=str(root))  # fmt: skip
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert res["ranks"] == n and res["max_rel_energy_diff"] < 1e-12 and res["max_abs_gradient_diff"] < 1e-12


def test_plan_cache_is_confirmed_against_the_atomic_numbers():
    """Two different structures of the same size behind the same address / version (what the caching allocator
    produces when the first one is freed): the cached plan of the first must not be used for the second."""
    from tad_dftd4_b200 import large

    dev = torch.device("cuda:0")
    large.clear_plan_cache()
    n1, p1, q1 = _cluster(150, seed=1)
    n2, p2, q2 = _cluster(150, seed=2)
    assert not torch.equal(n1, n2)
    numbers = n1.to(dev)
    e1 = large.dftd4_large(numbers, p1.to(dev), PBE0, q1.to(dev)).cpu()
    numbers.data.copy_(n2.to(dev))  # same address, same version counter, other structure
    e2 = large.dftd4_large(numbers, p2.to(dev), PBE0, q2.to(dev)).cpu()
    for e, (n, p, q) in ((e1, (n1, p1, q1)), (e2, (n2, p2, q2))):
        ref = orc.dftd4(n, p, PBE0, q)
        assert (e - ref).abs().max() / ref.abs().max() < 1e-10
    # and an unchanged tensor still hits
    before = len(large._PLAN_CACHE)
    large.dftd4_large(numbers, p2.to(dev) + 0.01, PBE0, q2.to(dev))
    assert len(large._PLAN_CACHE) == before


def test_previous_gradient_kernel_stays_selectable():
    """D4B200_LARGE_PIPE=0 (read once per process) selects large_atm_grad instead of the bulk-async pipeline: kept for
    same-box A/B timing, so it has to stay correct."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    code = (
        "import sys, torch, numpy as np\n"
        f"sys.path[:0] = [{str(root)!r}, {str(root / 'oracle')!r}, {str(root / 'tests')!r}]\n"
        "import d4_oracle as orc\n"
        "from test_gpu_large import _cluster, PBE0\n"
        "from tad_dftd4_b200.large import dftd4_large\n"
        "n, p, q = _cluster(150, seed=77)\n"
        "pr = p.clone().requires_grad_(True)\n"
        "er = orc.dftd4(n, pr, PBE0, q)\n"
        "(gr,) = torch.autograd.grad(er.sum(), pr)\n"
        "dev = torch.device('cuda:0')\n"
        "pd = p.to(dev).requires_grad_(True)\n"
        "e = dftd4_large(n.to(dev), pd, PBE0, q.to(dev))\n"
        "(g,) = torch.autograd.grad(e.sum(), pd)\n"
        "de = ((e.detach().cpu() - er.detach()).abs().max() / er.detach().abs().max()).item()\n"
        "dg = (g.cpu() - gr).abs().max().item()\n"
        "print('RESULT', de, dg)\n"
        "assert de < 1e-10 and dg < 1e-9\n"
    )
    env = dict(os.environ, D4B200_LARGE_PIPE="0")
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "RESULT" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
