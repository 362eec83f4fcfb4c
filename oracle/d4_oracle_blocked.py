"""
TEST INFRASTRUCTURE ONLY -- blocked / pruned evaluation of the CPU oracle for ONE large structure.

``d4_oracle.dftd4`` restates the reference's dense masked tensors; at a few thousand atoms its
``(N, N, 7, 7)`` and ``(N, N, N)`` temporaries no longer fit (the reference has the same limit,
``/root/reference/README.md:355-357``).  This module evaluates THE SAME SUMS -- the same
per-term formulas, masks and summation index roles -- for blocks of the first index ``i``:

* coordination numbers and the two-body rows: dense over ``j`` per block of ``i``
  (``tad_mctc.ncoord.cn_d4``, ``dispersion/twobody.py:134-201`` of the reference);
* ATM: ``E_i = 1/6 sum_{j,k} e[i,j,k]`` with the reference's mask ``real_triples & (r_ij <=
  cutoff) & (r_jk <= cutoff)`` (``dispersion/threebody.py:153-157``: ``j`` is the centre, ``r_ik``
  is NOT tested).  Terms the mask zeroes are not formed: ``j`` runs over the atoms within the
  cutoff of ``i``, ``k`` over the atoms within the cutoff of any such ``j``; inside that candidate
  set the dense masked expression of ``d4_oracle.atm_dispersion`` is evaluated unchanged;
* pair C6: ``einsum('ijab,ia,jb->ij')`` (``model/d4.py:268-289``) with the reference C6 gathered
  per block from the table of the structure's distinct elements.

It is pinned against ``d4_oracle.dftd4`` (itself bit-identical to the unmodified reference on the
golden cases) in ``tests/test_oracle_blocked.py`` on structures both can hold, including open
triples (tight cutoffs).  Only ``tests/`` import it.
"""

from __future__ import annotations

import torch

import d4_oracle as orc

Tensor = torch.Tensor


class Blocked:
    """Per-structure context: coordination numbers, reference weights, element tables."""

    def __init__(self, numbers: Tensor, positions: Tensor, param: dict, q: Tensor, *, ga=orc.GA_DEFAULT,
                 gc=orc.GC_DEFAULT, wf=orc.WF_DEFAULT, disp2=orc.DISP2_CUTOFF, disp3=orc.DISP3_CUTOFF,
                 block: int = 512):  # fmt: skip
        if numbers.dim() != 1 or bool((numbers == 0).any()):
            raise ValueError("one structure without padding expected")
        self.numbers, self.pos, self.param, self.q = numbers, positions, param, q
        self.disp2, self.disp3, self.block = disp2, disp3, block
        self.jchunk = 1 << 30  # e3_rows: candidate centres per piece (set lower for very large structures)
        t = orc._tables()
        dtype = positions.dtype
        self.n = numbers.shape[0]
        self.sq = (positions * positions).sum(-1)
        self.r4r2 = t["r4r2"].to(dtype)[numbers]
        # reference C6 of the distinct elements: (U, U, 7, 7), gathered per block
        self.uniq, self.inv = torch.unique(numbers, return_inverse=True)
        self.rc6 = orc.reference_c6(self.uniq, ga, gc, dtype)
        self.cn = torch.cat([self._cn_rows(torch.arange(b, min(self.n, b + block)))
                             for b in range(0, self.n, block)])  # fmt: skip
        self.wq = orc.weight_references_d4(numbers, self.cn, q, ga, gc, wf)
        self.w0 = orc.weight_references_d4(numbers, self.cn, None, ga, gc, wf)

    # tad_mctc.storch.cdist(p=2): quadratic expansion, clamped at eps (d4_oracle.cdist)
    def dist(self, rows: Tensor, cols: Tensor) -> Tensor:
        eps = torch.finfo(self.pos.dtype).eps
        d2 = self.sq[rows].unsqueeze(-1) + self.sq[cols].unsqueeze(-2) - 2.0 * (self.pos[rows] @ self.pos[cols].T)
        return torch.sqrt(torch.clamp(d2, min=eps))

    def _cn_rows(self, rows: Tensor) -> Tensor:
        """d4_oracle.cn_d4 for the rows of one block (all columns)."""
        t = orc._tables()
        dtype = self.pos.dtype
        cols = torch.arange(self.n)
        d = self.dist(rows, cols)
        mask = rows.unsqueeze(-1) != cols.unsqueeze(-2)
        d = torch.where(mask, d, torch.full((), torch.finfo(dtype).eps, dtype=dtype))
        en = t["pauling"].to(dtype)[self.numbers]
        rc = t["rcov"].to(dtype)[self.numbers]
        den = orc.K4 * torch.exp(-((en[rows].unsqueeze(-1) - en.unsqueeze(-2)).abs() + orc.K5) ** 2 / orc.K6)
        r0 = rc[rows].unsqueeze(-1) + rc.unsqueeze(-2)
        count = 0.5 * (1.0 + torch.erf(-orc.KCN * (d / r0 - 1.0)))
        cf = torch.where(mask & (d <= orc.CN_CUTOFF), den * count, torch.zeros((), dtype=dtype))
        return cf.sum(-1)

    def c6(self, rows: Tensor, cols: Tensor, w: Tensor, chunk: int = 128) -> Tensor:
        """model/d4.py:268-289 for the (rows x cols) block of pairs (row chunks bound the size of the
        gathered reference-C6 block)."""
        out = []
        for b in range(0, rows.shape[0], chunk):
            r = rows[b : b + chunk]
            rc6 = self.rc6[self.inv[r]][:, self.inv[cols]]  # (R, C, 7, 7)
            out.append(torch.einsum("ijab,ia,jb->ij", rc6, w[r], w[cols]))
        return torch.cat(out) if out else w.new_zeros((0, cols.shape[0]))

    def e2_rows(self, rows: Tensor) -> Tensor:
        """d4_oracle.dispersion2 for the rows of one block."""
        p, dtype = self.param, self.pos.dtype
        cols = torch.arange(self.n)
        mask = rows.unsqueeze(-1) != cols.unsqueeze(-2)
        d = torch.where(mask, self.dist(rows, cols), torch.full((), torch.finfo(dtype).eps, dtype=dtype))
        zero = torch.zeros((), dtype=dtype)
        c6 = self.c6(rows, cols, self.wq)
        qq = 3 * self.r4r2[rows].unsqueeze(-1) * self.r4r2.unsqueeze(-2)
        radius = p["a1"] * torch.sqrt(qq) + p["a2"]
        inside = mask & (d <= self.disp2)

        def damp(n: int) -> Tensor:
            return torch.where(inside, 1.0 / (d.pow(n) + radius.pow(n)), zero)

        e = p.get("s6", orc.S6_DEFAULT) * (c6 * damp(6)).sum(-1) + p.get("s8", orc.S8_DEFAULT) * (c6 * qq * damp(8)).sum(-1)
        if "s10" in p:
            e = e + p["s10"] * (c6 * qq.pow(2) * 49.0 / 40.0 * damp(10)).sum(-1)
        return -0.5 * e

    def e3_rows(self, rows: Tensor) -> Tensor:
        """d4_oracle.atm_dispersion for the rows ``i`` of one block, restricted to the index sets
        outside which the reference's mask is identically zero."""
        p, dtype = self.param, self.pos.dtype
        eps = torch.finfo(dtype).eps
        zero = torch.zeros((), dtype=dtype)
        c2 = self.disp3 * self.disp3
        alln = torch.arange(self.n)
        with torch.no_grad():
            reach = self.disp3 * (1.0 + 1e-12)  # candidate sets only have to be supersets of the mask
            near_i = (self.dist(rows, alln) <= reach).any(0)  # j: within the cutoff of some i of the block
            jset = torch.cat([torch.nonzero(near_i).flatten(), rows]).unique()
            near_j = (self.dist(jset, alln) <= reach).any(0)  # k: within the cutoff of some candidate j
            kset = torch.cat([torch.nonzero(near_j).flatten(), rows, jset]).unique()
        s9, alp = p.get("s9", orc.S9_DEFAULT), p.get("alp", orc.ALP_DEFAULT)

        def pair(a: Tensor, b: Tensor):
            m = a.unsqueeze(-1) != b.unsqueeze(-2)
            d = torch.where(m, self.dist(a, b), torch.full((), eps, dtype=dtype))
            rad = p["a1"] * torch.sqrt(torch.clamp(3.0 * self.r4r2[a].unsqueeze(-1) * self.r4r2[b].unsqueeze(-2), min=eps)) + p["a2"]
            return m, d.pow(2.0), rad, self.c6(a, b, self.w0)

        mik, d2ik, radik, cik = pair(rows, kset)
        total = torch.zeros(rows.shape[0], dtype=dtype)
        for b in range(0, jset.shape[0], self.jchunk):  # the sum over j in pieces (bounds the (i, j, k) temporaries)
            jc = jset[b : b + self.jchunk]
            mij, d2ij, radij, cij = pair(rows, jc)
            mjk, d2jk, radjk, cjk = pair(jc, kset)
            m3 = mij.unsqueeze(-1) & mik.unsqueeze(-2) & mjk.unsqueeze(-3)
            r2ij, r2ik, r2jk = d2ij.unsqueeze(-1), d2ik.unsqueeze(-2), d2jk.unsqueeze(-3)
            c9 = torch.sqrt(torch.clamp(torch.abs(cij.unsqueeze(-1) * cik.unsqueeze(-2) * cjk.unsqueeze(-3)), min=eps))
            r0 = radij.unsqueeze(-1) * radik.unsqueeze(-2) * radjk.unsqueeze(-3)
            r2 = r2ij * r2ik * r2jk
            r1 = torch.sqrt(r2)
            r3 = torch.where(m3, r1 * r2, torch.full((), eps, dtype=dtype))
            r5 = torch.where(m3, r2 * r3, torch.full((), eps, dtype=dtype))
            rr = torch.where(m3, r1, torch.ones((), dtype=dtype))
            tn = orc.RS9_DEFAULT * r0 / rr
            fdamp = torch.where(m3, 1.0 / (1.0 + 6.0 * tn ** (alp / 3.0)), zero)
            s = torch.where(m3, (r2ij + r2jk - r2ik) * (r2ij - r2jk + r2ik) * (-r2ij + r2jk + r2ik), zero)
            ang = torch.where(m3 & (r2ij <= c2) & (r2jk <= c2), 0.375 * s / r5 + 1.0 / r3, zero)
            total = total + (ang * fdamp * s9 * c9).sum((-2, -1))
        return total / 6.0


def dftd4_blocked(numbers: Tensor, positions: Tensor, param: dict, q: Tensor, *, rows: Tensor | None = None,
                  block: int = 512, block3: int = 4, jchunk: int | None = None, gradient: bool = False, **kw):  # fmt: skip
    """Atom-resolved D4 energy of one structure (all atoms, or the atoms ``rows`` only) and,
    with ``gradient=True``, ``d(sum_i E_i)/d positions`` by autograd, block by block."""
    pos = positions.detach().clone().requires_grad_(gradient)
    with torch.set_grad_enabled(gradient):
        ctx = Blocked(numbers, pos, param, q, block=block, **kw)
        if jchunk:
            ctx.jchunk = jchunk
        want = torch.arange(ctx.n) if rows is None else rows
        energy = torch.zeros(want.shape[0], dtype=pos.dtype)
        for b in range(0, want.shape[0], block):
            r = want[b : b + block]
            e2 = ctx.e2_rows(r)
            energy[b : b + block] += e2.detach()
            if gradient:
                e2.sum().backward(retain_graph=True)
        for b in range(0, want.shape[0], block3):
            r = want[b : b + block3]
            e3 = ctx.e3_rows(r)
            energy[b : b + block3] += e3.detach()
            if gradient:
                e3.sum().backward(retain_graph=True)
    if gradient:
        if rows is not None:
            raise ValueError("the gradient needs all rows")
        return energy, pos.grad.detach()
    return energy
