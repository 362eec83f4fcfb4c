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


def _oracle_backend(orc):
    """Oracle stand-in for the kernel backend of tad_dftd4_b200.large (energy, the two
    gradient stages) with the same partial-sum semantics."""

    def partial(n, p, q, g, rows, groups, cn=None):
        nat = n.shape[0]
        centres = torch.zeros(nat, dtype=torch.bool)
        centres[groups[0] * GS : min(nat, groups[1] * GS)] = True
        e2, e3, *_ = orc.dftd4(n, p, PARAM, q, parts=True, centres=centres, cn=cn)
        rowmask = torch.zeros(nat, dtype=torch.bool)
        rowmask[rows[0] : rows[1]] = True
        e = torch.where(rowmask, e2, torch.zeros_like(e2)) + e3
        return e if g is None else (e * g)

    def grad1(n, p, q, g, rows, groups, energy=None):
        if energy is not None:  # fused energy + gradient call: this rank's partial atomic energies
            energy += partial(n, p, q, None, rows, groups).detach()
        pp = p.clone().requires_grad_(True)
        qq = q.clone().requires_grad_(True)
        cn = orc.cn_d4(n, p).requires_grad_(True)  # independent variable: direct terms only
        gg = torch.ones_like(q) if g is None else g
        # NB: the two-body rows of the kernels carry -1/2 (g_i + g_j) per ordered pair;
        # with a row mask on E2 that is reproduced by symmetrising the upstream weights
        nat = n.shape[0]
        rowmask = torch.zeros(nat, dtype=q.dtype)
        rowmask[rows[0] : rows[1]] = 1.0
        centres = torch.zeros(nat, dtype=torch.bool)
        centres[groups[0] * GS : min(nat, groups[1] * GS)] = True
        _, e3, *_ = orc.dftd4(n, pp, PARAM, qq, parts=True, centres=centres, cn=cn)
        # row-owner form of the pair terms: atom i's row holds 1/2 of every pair (i,j)
        L = (e3 * gg).sum() + _rowowner_twobody(orc, n, pp, qq, cn, gg, rowmask)
        f, dcn, dq = torch.autograd.grad(L, (pp, cn, qq), allow_unused=True)
        return f, dcn, dq

    def grad2(n, p, dcn_total, rows, force):
        pp = p.clone().requires_grad_(True)
        (f,) = torch.autograd.grad((orc.cn_d4(n, pp) * dcn_total).sum(), pp)
        out = force.clone()
        out[rows[0] : rows[1]] += f[rows[0] : rows[1]]
        return out

    def energy(n, p, q, rows, groups, want_cost):
        if want_cost:
            nat = n.shape[0]
            ng = (nat + GS - 1) // GS
            near = (torch.cdist(p, p) <= 40.0).sum(-1).to(torch.float64)
            return None, torch.stack([near[g * GS : (g + 1) * GS].sum() for g in range(ng)])
        return partial(n, p, q, None, rows, groups), None

    return dict(energy=energy, grad1=grad1, grad2=grad2, group_size=GS)


def _rowowner_twobody(orc, n, pos, q, cn, g, rowmask):
    """sum over ordered pairs (i in rows, j): -1/4 (g_i + g_j) c6_ij F_ij ... such that the
    derivative with respect to R_i, cn_i, q_i of the rows equals the full derivative."""
    # The kernels give, for every row atom i, the COMPLETE derivative of
    # L2 = sum_{i<j} -1/2 (g_i+g_j) c6 F with respect to R_i, cn_i and q_i.  Emulate it
    # with a stop-gradient on the partner atom: L_i = sum_j -1/2 (g_i+g_j) c6(i, sg(j)) F(i, sg(j)).
    t = orc._tables()
    dtype = pos.dtype
    r4r2 = t["r4r2"].to(dtype)[n]

    def wfun(cn_, q_):
        return orc.weight_references_d4(n, cn_, q_)

    rc6 = orc.reference_c6(n, dtype=dtype)
    w_live, w_stop = wfun(cn, q), wfun(cn.detach(), q.detach())
    c6 = torch.einsum("ijab,ia,jb->ij", rc6, w_live, w_stop)  # row i live, column j frozen
    d = (pos.unsqueeze(1) - pos.detach().unsqueeze(0)).norm(dim=-1)
    nat = n.shape[0]
    off = ~torch.eye(nat, dtype=torch.bool)
    d = torch.where(off, d, torch.ones_like(d))
    qq = 3 * r4r2.unsqueeze(-1) * r4r2.unsqueeze(-2)
    R0 = PARAM["a1"] * torch.sqrt(qq) + PARAM["a2"]
    F = 1.0 / (d**6 + R0**6) + PARAM["s8"] * qq / (d**8 + R0**8)
    F = torch.where(off & (d <= 60.0), F, torch.zeros_like(F))
    G2 = -0.5 * (g.unsqueeze(-1) + g.unsqueeze(-2))
    return ((G2 * c6 * F).sum(-1) * rowmask).sum()


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

        # two-stage gradient (direct terms -> all-reduce dL/dcn, dL/dq -> CN chain -> all-reduce)
        from tad_dftd4_b200.large import dftd4_large_vjp

        g = torch.from_numpy(np.random.default_rng(2).normal(size=numbers.shape[0]))
        pos = positions.clone().requires_grad_(True)
        qv = qq.clone().requires_grad_(True)
        gp_ref, gq_ref = torch.autograd.grad((orc.dftd4(numbers, pos, PARAM, qv) * g).sum(), (pos, qv))
        gp, gq = dftd4_large_vjp(numbers, positions, PARAM, qq, g, backend=_oracle_backend(orc))
        assert (gp - gp_ref).abs().max() < 1e-14, (gp - gp_ref).abs().max()
        assert (gq - gq_ref).abs().max() < 1e-14, (gq - gq_ref).abs().max()
        # fused energy + gradient (what the autograd function's forward runs when inputs are on the tape)
        gp1, gq1, e1 = dftd4_large_vjp(numbers, positions, PARAM, qq, None, backend=_oracle_backend(orc),
                                       with_energy=True)  # fmt: skip
        gp1_ref, gq1_ref = torch.autograd.grad(orc.dftd4(numbers, pos, PARAM, qv).sum(), (pos, qv))
        assert (e1 - ref).abs().max() < 1e-15 and (gp1 - gp1_ref).abs().max() < 1e-14
        assert (gq1 - gq1_ref).abs().max() < 1e-14
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


def test_large_param_vjp_host_logic_against_oracle_autograd():
    """large.large_param_vjp (exact parts for s6 / s8 / s9 / s10, fourth-order differences for a1 / a2 / alp) with the
    oracle standing in for the tiled energy kernels, against autograd of the oracle through the parameters."""
    sys.path[:0] = [str(ROOT), str(ROOT / "oracle")]
    import d4_oracle as orc
    from tad_dftd4_b200 import large

    numbers, positions, q = orc.organic_batch([28], seed=13)
    numbers, positions, q = numbers[0], positions[0], q[0]
    base = {"s6": 1.0, "s8": 0.78981345, "s9": 1.0, "s10": 0.3, "a1": 0.49484001, "a2": 5.73083694, "alp": 16.0}
    g = torch.randn(28, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in base.items()}
    keys = ("s6", "s8", "s9", "s10", "a1", "a2", "alp")
    want = torch.autograd.grad((orc.dftd4(numbers, positions, tp, q) * g).sum(), [tp[k] for k in keys])

    def factory(par):
        def energy(n, p, qq, rows, groups, want_cost):
            if want_cost:
                return None, torch.ones((n.shape[0] + GS - 1) // GS, dtype=torch.float64)
            return orc.dftd4(n, p, {k: v for k, v in par.items() if v is not None}, qq), None

        return dict(energy=energy, group_size=GS)

    large.clear_plan_cache()
    got = large.large_param_vjp(numbers, positions, base, q, g, [True] * 7, backend_factory=factory)
    for k, w, v in zip(keys, want, got):
        assert abs(float(v) - float(w)) <= 1e-8 * abs(float(w)) + 1e-13, (k, float(v), float(w))
    # parameters that are not asked for cost nothing and come back as None
    some = large.large_param_vjp(numbers, positions, base, q, None, [False, True, False, False, False, False, False],
                                 backend_factory=factory)
    assert [x is not None for x in some] == [False, True, False, False, False, False, False]
