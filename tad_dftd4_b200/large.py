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

__all__ = ["dftd4_large", "dftd4_large_vjp", "large_energy", "morton_order", "balanced_ranges",
           "work_counts", "profile_begin", "profile_end", "clear_plan_cache"]  # fmt: skip

Tensor = torch.Tensor


# --------------------------------------------------------------------------
# optional section timing (bench.py: all-reduce time split out of the step)
# --------------------------------------------------------------------------
class _Profile:
    """CUDA-event pairs per named section, recorded on the caller's stream."""

    def __init__(self):
        self.events: list[tuple[str, torch.cuda.Event, torch.cuda.Event]] = []

    def section(self, name: str):
        return _Section(self, name)

    def totals_ms(self) -> dict[str, float]:
        out: dict[str, float] = {}
        for name, a, b in self.events:
            b.synchronize()
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out


class _Section:
    def __init__(self, prof, name):
        self.prof, self.name = prof, name

    def __enter__(self):
        if self.prof is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.prof is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            self.prof.events.append((self.name, self.a, b))


_PROFILE: _Profile | None = None


def profile_begin() -> _Profile:
    """Start timing the sections ``plan`` and ``all_reduce`` of every following call."""
    global _PROFILE
    _PROFILE = _Profile()
    return _PROFILE


def profile_end() -> None:
    global _PROFILE
    _PROFILE = None


def _section(name: str) -> _Section:
    return _Section(_PROFILE, name)


def work_counts(numbers: Tensor, positions: Tensor, *, cn_cutoff=30.0, disp2_cutoff=60.0, disp3_cutoff=40.0,
                block: int = 2048) -> dict[str, float]:  # fmt: skip
    """Exact work of ONE structure in SURVEY.md 8(d)'s units (device tensors, blocked):
    ``pairs_cn`` (r <= 30), ``pairs_disp2`` (r <= 60), ``centre_triples = sum_j C(n_j, 2)`` with
    ``n_j`` the neighbours of j within 40, and ``triples = centre_triples - 2 #closed`` -- the
    unordered triples with at least two distances inside the ATM cutoff (the reference's
    two-distance mask, dispersion/threebody.py:153-157), ``#closed`` those with all three."""
    keep = numbers != 0
    p = positions[keep].to(torch.float64)
    n = p.shape[0]
    adj = torch.empty((n, n), dtype=torch.float32, device=p.device)
    pcn = p2 = 0
    for b0 in range(0, n, block):
        d = torch.cdist(p[b0 : b0 + block], p)
        rows = torch.arange(b0, min(n, b0 + block), device=p.device)
        d[torch.arange(rows.numel(), device=p.device), rows] = float("inf")  # no self pairs
        pcn += int((d <= cn_cutoff).sum().item())
        p2 += int((d <= disp2_cutoff).sum().item())
        adj[b0 : b0 + block] = (d <= disp3_cutoff).to(torch.float32)
    nb = adj.sum(-1, dtype=torch.float64)
    ctrip = float((nb * (nb - 1) / 2).sum().item())
    closed6 = 0.0  # trace(A^3) = 6 #closed; entries of A.A are exact integers < 2^24 in float32
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for b0 in range(0, n, block):
            blk = adj[b0 : b0 + block]
            closed6 += float(((blk @ adj) * blk).sum(dtype=torch.float64).item())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    closed = closed6 / 6.0
    return {"atoms": n, "pairs_cn": pcn / 2, "pairs_disp2": p2 / 2, "centre_triples": ctrip,
            "closed_triples": closed, "triples": ctrip - 2.0 * closed}  # fmt: skip


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


# Topology of the most recent structures: atom order along the Morton curve and this rank's
# row / centre-group ranges.  Any atom order is correct and the ranges only balance the load,
# so both are reused for as long as the SAME ``numbers`` tensor comes back (a geometry
# optimisation or MD run moves the atoms a little per step); a call with another tensor,
# another world size or an in-place change of ``numbers`` builds a new plan.  The key is the
# tensor's address + version (no reference is kept), so a hit is confirmed against a copy of the
# atomic numbers the plan was made for: the caching allocator hands the address of a freed
# tensor to the next structure of the same size.
_PLAN_CACHE: dict[tuple, tuple] = {}
_PLAN_CACHE_SIZE = 4


def clear_plan_cache() -> None:
    _PLAN_CACHE.clear()


class _Plan:
    """Sorted/compacted view of one structure + this rank's share of the work."""

    def __init__(self, numbers, positions, q, group, compute_cost, gs):
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.group = group
        key = (numbers.data_ptr(), numbers._version, tuple(numbers.shape), str(numbers.device), self.world,
               self.rank, id(group), gs)  # fmt: skip
        hit = _PLAN_CACHE.get(key)
        if hit is not None and not torch.equal(hit[4], numbers):
            # the address of a freed tensor came back with another structure of the same size (or the tensor was
            # changed behind the version counter): the key matches, the plan does not
            hit = None
        with _section("plan"):
            if hit is not None:
                self.order, self.numbers, self.rows, self.groups = hit[:4]
            else:
                keep = torch.nonzero(numbers != 0).flatten()
                self.order = keep[morton_order(positions[keep])]
                self.numbers = numbers[self.order].to(torch.int64).contiguous()
            self.positions = positions[self.order].detach().contiguous()
            self.q = q[self.order].detach().to(positions.dtype).contiguous()
            self.nat = int(self.order.numel())
            if hit is None:
                ng = (self.nat + gs - 1) // gs
                if self.world > 1 and self.nat > 0:
                    # cost model: a centre group with m neighbours evaluates ~m^2/2 pairs x group size
                    cost = compute_cost(self.numbers, self.positions, self.q)
                    g0, g1 = balanced_ranges(cost.to(torch.float64) ** 2, self.world)[self.rank]
                    r0, r1 = min(self.nat, g0 * gs), min(self.nat, g1 * gs)  # rows follow the same blocks
                    if self.rank == self.world - 1:
                        r1 = self.nat
                else:
                    g0, g1, r0, r1 = 0, ng, 0, self.nat
                self.rows, self.groups = (r0, r1), (g0, g1)
                while len(_PLAN_CACHE) >= _PLAN_CACHE_SIZE:
                    _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
                _PLAN_CACHE.pop(key, None)
                _PLAN_CACHE[key] = (self.order, self.numbers, self.rows, self.groups, numbers.detach().clone())

    def all_reduce(self, t: Tensor) -> Tensor:
        if self.world > 1:
            with _section("all_reduce"):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


def _kernel_backend(positions: Tensor, param, cutoff, model=None):
    """``model``: (ga, gc, wf[, ref_charges[, c9_frequency]]) of the caller's D4Model (defaults when None)."""
    from . import _lib, defaults
    from .disp import _Engine, _flatten_param

    ga, gc, wf, *rest = model if model is not None else (defaults.GA_DEFAULT, defaults.GC_DEFAULT, defaults.WF_DEFAULT)
    ref_charges = rest[0] if rest else "eeq"
    c9_frequency = rest[1] if len(rest) > 1 else None

    engine = _Engine.get(positions.device, ga, gc, ref_charges, c9_frequency)
    par = _flatten_param(param, cutoff, 0, wf)
    lib = engine.lib
    gs = int(lib.d4b200_large_group_size())
    fp32 = positions.dtype == torch.float32

    def ws_for(nat):
        return engine.large_workspace(int(lib.d4b200_large_workspace_bytes(engine.handle, nat, int(fp32))))

    def stream():
        return torch.cuda.current_stream(positions.device).cuda_stream

    def energy(n, p, qq, rows, groups, want_cost):
        with torch.cuda.device(p.device):
            return _kernel_compute(engine, par, n, p, qq, rows, groups, want_cost)

    def grad1(n, p, qq, g, rows, groups, energy=None):
        """Stage 1 of the gradient; ``energy`` (zeroed tensor [nat]) additionally receives this rank's
        partial atomic energies from the same launches (fused energy + gradient call)."""
        nat = n.shape[0]
        force = torch.zeros((nat, 3), dtype=p.dtype, device=p.device)
        dcn = torch.zeros(nat, dtype=p.dtype, device=p.device)
        dq = torch.zeros(nat, dtype=p.dtype, device=p.device)
        ws = ws_for(nat)
        with torch.cuda.device(p.device):
            if energy is None:
                fn = lib.d4b200_large_gradient_f32 if fp32 else lib.d4b200_large_gradient_f64
                code = fn(engine.handle, C.byref(par), nat, n.data_ptr(), p.data_ptr(), qq.data_ptr(),
                          g.data_ptr() if g is not None else None, rows[0], rows[1], groups[0], groups[1],
                          force.data_ptr(), dcn.data_ptr(), dq.data_ptr(), ws.data_ptr(), ws.numel(), stream())  # fmt: skip
            else:
                fn = lib.d4b200_large_energy_gradient_f32 if fp32 else lib.d4b200_large_energy_gradient_f64
                code = fn(engine.handle, C.byref(par), nat, n.data_ptr(), p.data_ptr(), qq.data_ptr(),
                          g.data_ptr() if g is not None else None, rows[0], rows[1], groups[0], groups[1],
                          energy.data_ptr(), force.data_ptr(), dcn.data_ptr(), dq.data_ptr(), ws.data_ptr(),
                          ws.numel(), stream())  # fmt: skip
            _lib.check(code, "d4b200_large_gradient")
        return force, dcn, dq

    def grad2(n, p, dcn_total, rows, force):
        fn = lib.d4b200_large_cn_chain_f32 if fp32 else lib.d4b200_large_cn_chain_f64
        with torch.cuda.device(p.device):
            _lib.check(
                fn(engine.handle, C.byref(par), n.shape[0], n.data_ptr(), p.data_ptr(),
                   dcn_total.data_ptr(), rows[0], rows[1], force.data_ptr(), stream()),
                "d4b200_large_cn_chain",
            )  # fmt: skip
        return force

    return dict(energy=energy, grad1=grad1, grad2=grad2, group_size=gs)


def _check_single(numbers, positions, q):
    if numbers.dim() != 1 or positions.shape != (numbers.shape[0], 3) or q.shape != numbers.shape:
        raise ValueError("expected numbers (nat,), positions (nat, 3), q (nat,)")


def large_energy(numbers, positions, param, q, *, cutoff=None, group=None, backend=None, model=None) -> Tensor:
    """Atom-resolved D4 energy of ONE structure ``(nat,)`` with the tiled kernels
    (no autograd; see :func:`dftd4_large`)."""
    _check_single(numbers, positions, q)
    be = backend or _kernel_backend(positions, param, cutoff, model)
    plan = _Plan(numbers, positions, q, group,
                 lambda n, p, qq: be["energy"](n, p, qq, (0, 0), (0, 0), True)[1], be["group_size"])  # fmt: skip
    out = torch.zeros(numbers.shape[0], dtype=positions.dtype, device=positions.device)
    if plan.nat == 0:
        return out
    e_s, _ = be["energy"](plan.numbers, plan.positions, plan.q, plan.rows, plan.groups, False)
    out[plan.order] = plan.all_reduce(e_s)
    return out


def dftd4_large_vjp(numbers, positions, param, q, gout, *, cutoff=None, group=None, backend=None, model=None,
                    with_energy: bool = False):
    """``(dL/dpositions, dL/dq)`` for ``L = sum_i gout_i E_i`` of ONE structure (``gout=None``: ones).

    Two stages with one all-reduce each (plus the final one for the forces):
    direct two-body/ATM terms -> all-reduce(dL/dcn, dL/dq) -> CN chain rule on the
    rank's rows -> all-reduce(dL/dpositions).  ``with_energy=True`` additionally returns the
    atom-resolved energies, accumulated by the same kernel launches (one more all-reduce)."""
    _check_single(numbers, positions, q)
    be = backend or _kernel_backend(positions, param, cutoff, model)
    plan = _Plan(numbers, positions, q, group,
                 lambda n, p, qq: be["energy"](n, p, qq, (0, 0), (0, 0), True)[1], be["group_size"])  # fmt: skip
    gpos = torch.zeros_like(positions)
    gq = torch.zeros(numbers.shape[0], dtype=positions.dtype, device=positions.device)
    energy = torch.zeros(numbers.shape[0], dtype=positions.dtype, device=positions.device) if with_energy else None
    if plan.nat == 0:
        return (gpos, gq, energy) if with_energy else (gpos, gq)
    g_s = None if gout is None else gout[plan.order].to(positions.dtype).contiguous()
    e_s = torch.zeros(plan.nat, dtype=positions.dtype, device=positions.device) if with_energy else None
    if with_energy:
        force, dcn, dq = be["grad1"](plan.numbers, plan.positions, plan.q, g_s, plan.rows, plan.groups, energy=e_s)
    else:
        force, dcn, dq = be["grad1"](plan.numbers, plan.positions, plan.q, g_s, plan.rows, plan.groups)
    if plan.world > 1:
        both = torch.stack([dcn, dq])
        plan.all_reduce(both)
        dcn, dq = both[0].contiguous(), both[1].contiguous()
        if with_energy:
            plan.all_reduce(e_s)
    force = be["grad2"](plan.numbers, plan.positions, dcn, plan.rows, force)
    plan.all_reduce(force)
    gpos[plan.order] = force
    gq[plan.order] = dq
    if with_energy:
        energy[plan.order] = e_s
        return gpos, gq, energy
    return gpos, gq


def _plain_param(param) -> dict:
    """The damping parameters as plain floats (the kernels take them by value)."""
    return {k: (float(v.detach()) if isinstance(v, Tensor) else v) for k, v in param.items()}


# fourth-order central differences (multiples of the step, weights)
_FD4 = ((-2.0, 1.0 / 12.0), (-1.0, -8.0 / 12.0), (1.0, 8.0 / 12.0), (2.0, -1.0 / 12.0))


def large_param_vjp(numbers, positions, param, q, gout, need, *, cutoff=None, group=None, model=None,
                    backend_factory: Callable | None = None) -> list:  # fmt: skip
    """``d(sum_i gout_i E_i)/d(s6, s8, s9, s10, a1, a2, alp)`` of ONE large structure (entries of ``need``).

    The reference differentiates its dense tape (``test/test_grad/test_param.py:40-100``); the one-CTA family has
    analytic kernels for this (``csrc/d4b200_param.cu``).  Here the tiled energy kernels are re-run instead: the
    energy is LINEAR in the scaling factors, so ``dE/ds6``, ``dE/ds8``, ``dE/ds10`` and ``dE/ds9`` are the energies of
    the corresponding part alone (the other factors switched off -- exact), and ``a1``, ``a2``, ``alp`` enter smoothly
    through the damping radii / exponent, where fourth-order central differences of the weighted energy (relative step
    1e-3) are accurate to ~1e-11 relative.  Up to 3 two-body + 1 ATM launches for the factors, 8 full and 4 ATM-only
    energy evaluations for the rest; every evaluation is all-reduced like the energy itself, so all ranks agree."""
    from . import defaults

    base = _plain_param(param)
    g = None if gout is None else gout.detach()

    def weighted(par) -> Tensor:
        be = backend_factory(par) if backend_factory is not None else None  # CPU tests: oracle in place of the kernels
        e = large_energy(numbers, positions.detach(), par, q.detach(), cutoff=cutoff, group=group, model=model,
                         backend=be)  # fmt: skip
        return (e.sum() if g is None else (e * g).sum()).double()

    def part(**on) -> Tensor:  # one linear part: its factor = 1, the others = 0
        par = dict(base, s6=0.0, s8=0.0, s9=0.0)
        par.pop("s10", None)
        par.update(on)
        return weighted(par)

    def fd(key: str, default: float, atm_only: bool) -> Tensor:
        x = float(base.get(key, default) if base.get(key) is not None else default)
        h = 1e-3 * max(abs(x), 1.0)
        total = torch.zeros((), dtype=torch.float64, device=positions.device)
        for mult, coef in _FD4:
            par = dict(base)
            par[key] = x + mult * h
            if atm_only:  # the two-body part does not depend on this parameter
                par.update(s6=0.0, s8=0.0)
                par.pop("s10", None)
            total = total + coef * weighted(par)
        return total / h

    s9 = float(base.get("s9", defaults.S9) if base.get("s9") is not None else defaults.S9)
    out: list = [None] * 7
    if need[0]:
        out[0] = part(s6=1.0)
    if need[1]:
        out[1] = part(s8=1.0)
    if need[2]:
        out[2] = part(s9=1.0)
    if need[3]:
        out[3] = part(s10=1.0)
    if need[4]:
        out[4] = fd("a1", defaults.A1, False)
    if need[5]:
        out[5] = fd("a2", defaults.A2, False)
    if need[6]:
        out[6] = fd("alp", defaults.ALP, True) if s9 != 0.0 else torch.zeros((), dtype=torch.float64, device=positions.device)
    return out


class _LargeFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, positions, q, numbers, param, cutoff, group, model, *ptens):
        ctx.save_for_backward(positions, q, numbers)
        ctx.param, ctx.cutoff, ctx.group, ctx.model = param, cutoff, group, model
        ctx.ptens = ptens
        ctx.cached = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            # inputs on the tape: energies AND the gradient of sum(E) from one set of launches (the
            # gradient kernels evaluate every term anyway); backward only scales the cached gradient
            # when the upstream gradient is a broadcast scalar (energy.sum().backward(), forces)
            gpos, gq, energy = dftd4_large_vjp(numbers, positions, param, q, None, cutoff=cutoff, group=group,
                                               model=model, with_energy=True)  # fmt: skip
            ctx.cached = (gpos, gq)
            return energy
        return large_energy(numbers, positions, param, q, cutoff=cutoff, group=group, model=model)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        positions, q, numbers = ctx.saved_tensors
        need_in = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        gpos = gq = None
        if need_in:
            if ctx.cached is not None and all(s == 0 for s in gout.stride()):
                c = gout.reshape(-1)[0]
                gpos, gq = ctx.cached[0] * c, ctx.cached[1] * c
            else:
                gpos, gq = dftd4_large_vjp(numbers, positions, ctx.param, q, gout.contiguous(),
                                           cutoff=ctx.cutoff, group=ctx.group, model=ctx.model)  # fmt: skip
        gpar: list = [None] * len(ctx.ptens)
        need = [k < len(ctx.ptens) and ctx.ptens[k] is not None and ctx.needs_input_grad[7 + k] for k in range(7)]
        if any(need):
            vals = large_param_vjp(numbers, positions, ctx.param, q, gout.contiguous(), need, cutoff=ctx.cutoff,
                                   group=ctx.group, model=ctx.model)  # fmt: skip
            for k, v in enumerate(vals):
                if v is not None:
                    gpar[k] = v.to(ctx.ptens[k].dtype).reshape(ctx.ptens[k].shape)
        return (gpos, gq, None, None, None, None, None, *gpar)


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
    model: tuple[float, float, float] | None = None,
) -> Tensor:
    """Atom-resolved D4 energy of ONE structure ``(nat,)`` with the tiled kernels;
    differentiable with respect to ``positions`` and ``q``.

    With an initialised ``torch.distributed`` process group every rank passes the same
    (replicated) inputs, evaluates its share of two-body rows / ATM centre groups and
    the result is all-reduced, so every rank returns the full energy vector.

    ``model = (ga, gc, wf)`` carries the hyper-parameters of a caller-supplied ``D4Model``
    (defaults otherwise).  ``compute(numbers, positions, q, rows, groups, want_cost)`` is
    injectable for CPU tests of this host logic (energy only).
    """
    if compute is not None:
        be = dict(energy=compute, group_size=group_size or 16)
        return large_energy(numbers, positions, param, q, cutoff=cutoff, group=group, backend=be)
    from .disp import _param_tensors

    ptens = _param_tensors(param)  # damping parameters on the tape (tensors with requires_grad)
    return _LargeFunction.apply(positions, q, numbers, _plain_param(param) if ptens else param, cutoff, group, model,
                                *ptens)  # fmt: skip
