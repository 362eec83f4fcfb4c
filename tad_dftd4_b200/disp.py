"""
Functional API: ``dftd4`` and ``get_properties`` with the reference's
signatures (``/root/reference/src/tad_dftd4/disp.py:44-197``), executed by the
sm_100a kernels behind the C ABI of ``include/d4b200.h``.

What the reference does in ``Disp.calculate`` (``dispersion/base.py:285-431``)
-- validation, defaults, model construction, CN, two-body + ATM term loop --
collapses into input validation (same exception types) plus ONE C call; forces
come from a custom ``torch.autograd.Function`` whose backward is the fused
analytic-gradient kernel (the reference differentiates its dense tape,
``examples/forces.py:47-50``).

There is no CPU path and no PyTorch fallback: tensors must live on a B200.
"""

from __future__ import annotations

import ctypes as C
import weakref
from typing import Any

import numpy as np
import torch

from . import _lib, defaults
from .cutoff import Cutoff
from .damping import Param, RationalDamping
from .tables import build_tables

__all__ = ["dftd4", "dftd4_host", "get_properties", "set_checks", "set_fused_forward", "last_launch_count"]

Tensor = torch.Tensor

_ALLOWED_MODELS = ("d3", "d4", "d4s", "d5")
_CHECKS = True


def set_checks(enabled: bool) -> None:
    """Enable/disable the per-call device status read-back (one stream sync).

    With checks on (default) an atomic number outside 1..103 or a structure the
    kernels cannot handle raises; with checks off the call is fully
    asynchronous."""
    global _CHECKS
    _CHECKS = bool(enabled)


def last_launch_count() -> int:
    """Kernel launches issued by the last energy/gradient call of this thread."""
    return int(_lib.load().d4b200_last_launch_count())


# ---------------------------------------------------------------------------
# per-device engine: tables handle + workspace cache
# ---------------------------------------------------------------------------
class _Engine:
    _cache: dict[tuple[int, float, float, str, int | None], "_Engine"] = {}

    def __init__(self, device: torch.device, ga: float, gc: float, ref_charges: str = "eeq",
                 c9_frequency: int | None = None):  # fmt: skip
        lib = _lib.load()
        tab = build_tables(ga, gc, ref_charges, c9_frequency)
        f64 = np.ascontiguousarray(tab.f64_blob())
        i32 = np.ascontiguousarray(tab.i32_blob())
        handle = C.c_void_p()
        index = device.index if device.index is not None else torch.cuda.current_device()
        _lib.check(
            lib.d4b200_tables_create(
                index, f64.ctypes.data, f64.size, i32.ctypes.data, i32.size, ga, gc, C.byref(handle)
            ),
            "d4b200_tables_create",
        )
        self.lib = lib
        self.handle = handle
        self.ga, self.gc = float(ga), float(gc)
        self.limits: dict[tuple[bool, bool, int], int] = {}
        self.device = torch.device("cuda", index)
        self._ws_by_stream: dict[int, Tensor] = {}

    @classmethod
    def get(cls, device: torch.device, ga: float, gc: float, ref_charges: str = "eeq",
            c9_frequency: int | None = None) -> "_Engine":  # fmt: skip
        index = device.index if device.index is not None else torch.cuda.current_device()
        key = (index, float(ga), float(gc), ref_charges, c9_frequency)
        eng = cls._cache.get(key)
        if eng is None:
            eng = cls._cache[key] = cls(device, ga, gc, ref_charges, c9_frequency)
        return eng

    def large_workspace(self, need: int) -> Tensor:
        key = ("large", torch.cuda.current_stream(self.device).cuda_stream)
        ws = self._ws_by_stream.get(key)
        if ws is None or ws.numel() < need:
            ws = self._ws_by_stream[key] = torch.empty(need, dtype=torch.uint8, device=self.device)
        return ws

    def workspace(self, nbatch: int, nat: int) -> Tensor:
        """Workspace of the CURRENT CUDA stream: calls issued on different streams (or from different
        host threads on their own streams) never share the planning arrays, work queues and scratch
        planes a running call still reads."""
        need = int(self.lib.d4b200_workspace_bytes(nbatch, nat))
        key = torch.cuda.current_stream(self.device).cuda_stream
        ws = self._ws_by_stream.get(key)
        if ws is None or ws.numel() < need:
            if len(self._ws_by_stream) > 16:  # streams come and go: do not grow without bound
                self._ws_by_stream.clear()
            ws = self._ws_by_stream[key] = torch.empty(need, dtype=torch.uint8, device=self.device)
        return ws

    def _status(self, ws: Tensor, stream: int) -> None:
        bits = C.c_int(0)
        _lib.check(self.lib.d4b200_status(ws.data_ptr(), stream, C.byref(bits)), "d4b200_status")
        if bits.value & 1:
            raise ValueError("numbers contains an atomic number outside 1..103 (0 = padding).")
        if bits.value & 2:
            raise NotImplementedError(
                "a structure is larger than the one-CTA-per-structure kernels support "
                "for this dtype; the tiled large-system path handles it"
            )

    def energy(self, par: _lib.Params, numbers: Tensor, positions: Tensor, q: Tensor,
               want_cn: bool = False) -> tuple[Tensor, Tensor | None]:  # fmt: skip
        nbatch, nat = numbers.shape
        energy = torch.empty_like(q)
        cn = torch.empty_like(q) if want_cn else None
        ws = self.workspace(nbatch, nat)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        fn = self.lib.d4b200_energy_f64 if positions.dtype == torch.float64 else self.lib.d4b200_energy_f32
        _lib.check(
            fn(self.handle, C.byref(par), nbatch, nat, numbers.data_ptr(), positions.data_ptr(),
               q.data_ptr(), energy.data_ptr(), cn.data_ptr() if cn is not None else None,
               ws.data_ptr(), ws.numel(), stream),
            "d4b200_energy",
        )  # fmt: skip
        if _CHECKS:
            self._status(ws, stream)
        return energy, cn

    def properties(self, par: _lib.Params, numbers: Tensor, positions: Tensor, q: Tensor):
        nbatch, nat = numbers.shape
        cn = torch.empty_like(q)
        alpha = torch.empty_like(q)
        escr = torch.empty_like(q)
        c6 = torch.empty((nbatch, nat, nat), dtype=q.dtype, device=q.device)
        ws = self.workspace(nbatch, nat)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        fn = (self.lib.d4b200_properties_f64 if positions.dtype == torch.float64
              else self.lib.d4b200_properties_f32)  # fmt: skip
        _lib.check(
            fn(self.handle, C.byref(par), nbatch, nat, numbers.data_ptr(), positions.data_ptr(),
               q.data_ptr(), cn.data_ptr(), c6.data_ptr(), alpha.data_ptr(), escr.data_ptr(),
               ws.data_ptr(), ws.numel(), stream),
            "d4b200_properties",
        )  # fmt: skip
        if _CHECKS:
            self._status(ws, stream)
        return cn, c6, alpha

    def gradient(self, par: _lib.Params, numbers: Tensor, positions: Tensor, q: Tensor,
                 gout: Tensor | None, want_pos: bool, want_q: bool, with_energy: bool = False):  # fmt: skip
        """VJP of the energy; ``with_energy`` additionally returns the energies from the
        same launch (fused energy + gradient call)."""
        nbatch, nat = numbers.shape
        gpos = torch.empty_like(positions) if want_pos else None
        gq = torch.empty_like(q) if want_q else None
        energy = torch.empty_like(q) if with_energy else None
        ws = self.workspace(nbatch, nat)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        f64 = positions.dtype == torch.float64
        common = (self.handle, C.byref(par), nbatch, nat, numbers.data_ptr(), positions.data_ptr(),
                  q.data_ptr(), gout.data_ptr() if gout is not None else None)  # fmt: skip
        tail = (gpos.data_ptr() if gpos is not None else None, gq.data_ptr() if gq is not None else None,
                ws.data_ptr(), ws.numel(), stream)  # fmt: skip
        if with_energy:
            fn = self.lib.d4b200_energy_gradient_f64 if f64 else self.lib.d4b200_energy_gradient_f32
            _lib.check(fn(*common, energy.data_ptr(), *tail), "d4b200_energy_gradient")
        else:
            fn = self.lib.d4b200_gradient_f64 if f64 else self.lib.d4b200_gradient_f32
            _lib.check(fn(*common, *tail), "d4b200_gradient")
        if _CHECKS:
            self._status(ws, stream)
        return (gpos, gq, energy) if with_energy else (gpos, gq)


_FUSED_FORWARD = True


def set_fused_forward(enabled: bool) -> None:
    """When inputs require grad, evaluate energies AND the gradient of ``sum(E)`` in the
    forward launch (default).  ``backward`` then only scales the cached gradient if the
    upstream gradient is a broadcast scalar (``E.sum().backward()``,
    ``autograd.grad(E.sum(), positions)``); any other upstream gradient runs the general
    VJP kernel.  Disable to make the forward pass energy-only."""
    global _FUSED_FORWARD
    _FUSED_FORWARD = bool(enabled)


class _D4Function(torch.autograd.Function):
    """energy[b, i] = D4(numbers, positions, q); backward = fused analytic VJP."""

    @staticmethod
    def forward(ctx, positions: Tensor, q: Tensor, numbers: Tensor, par, engine: _Engine, *ptens):
        # ptens: the seven damping parameters (s6, s8, s9, s10, a1, a2, alp) as tensors where the
        # caller differentiates them (None otherwise); their values are already in ``par``
        need_pos, need_q = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        ctx.ptens = tuple((t.dtype, t.device, t.shape) if t is not None else None for t in ptens)
        ctx.cached = None
        if _FUSED_FORWARD and (need_pos or need_q):
            gpos, gq, energy = engine.gradient(par, numbers, positions, q, None, need_pos, need_q,
                                               with_energy=True)  # fmt: skip
            ctx.cached = (gpos, gq)
        else:
            energy, _ = engine.energy(par, numbers, positions, q)
        ctx.save_for_backward(positions, q, numbers)
        ctx.par = par
        ctx.engine = engine
        return energy

    @staticmethod
    def backward(ctx, gout: Tensor):
        positions, q, numbers = ctx.saved_tensors
        need_pos, need_q = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if torch.is_grad_enabled() and (need_pos or need_q):
            # backward under create_graph=True (Hessians, gradgradcheck, force matching): the VJP
            # itself has to stay on the tape -> second-order node, see _D4Vjp
            if ctx.ptens and any(ctx.needs_input_grad[5 + k] for k in range(len(ctx.ptens))):
                raise NotImplementedError(
                    "second derivatives are provided with respect to positions and charges; detach the "
                    "damping parameters (or differentiate them to first order only)"
                )
            gpos, gq = _D4Vjp.apply(gout, positions, q, numbers, ctx.par, ctx.engine)
            return (gpos if need_pos else None, gq if need_q else None, None, None, None,
                    *((None,) * len(ctx.ptens)))  # fmt: skip
        with torch.no_grad():
            return _D4Function._first_order(ctx, gout, positions, q, numbers, need_pos, need_q)

    @staticmethod
    def _first_order(ctx, gout, positions, q, numbers, need_pos, need_q):
        gpar: tuple = ()
        if ctx.ptens:
            want = [ctx.needs_input_grad[5 + k] for k in range(len(ctx.ptens))]
            gpar = (None,) * len(ctx.ptens)
            if any(want):
                vec = _param_vjp(ctx.engine, ctx.par, numbers, positions, q, gout.contiguous())
                gpar = tuple(
                    vec[k].to(device=meta[1], dtype=meta[0]).reshape(meta[2]) if (w and meta is not None) else None
                    for k, (w, meta) in enumerate(zip(want, ctx.ptens))
                )
        if not (need_pos or need_q):
            return (None, None, None, None, None, *gpar)
        if ctx.cached is not None and all(s == 0 for s in gout.stride()):
            # upstream gradient is one broadcast scalar c: dL/dx = c * d(sum E)/dx (no sync)
            c = gout.reshape(-1)[0]
            gpos, gq = ctx.cached
            return ((gpos * c if need_pos else None), (gq * c if need_q else None), None, None, None, *gpar)
        gpos, gq = ctx.engine.gradient(
            ctx.par, numbers, positions, q, gout.contiguous(), need_pos, need_q
        )
        return (gpos, gq, None, None, None, *gpar)


# Central-difference step of the second-order path (Bohr for positions, e for charges; applied to
# the direction normalised to unit maximum component) and its fourth-order stencil
_FD_STEP = 5.0e-4
_FD_STENCIL = ((1.0, 8.0 / 12.0), (-1.0, -8.0 / 12.0), (2.0, -1.0 / 12.0), (-2.0, 1.0 / 12.0))


def _fd_direction(*parts: Tensor | None) -> Tensor:
    """Per-structure step ``t`` (shape [B]) such that the largest component of ``t * direction`` is
    ``_FD_STEP``; structures whose direction vanishes get ``t = 1`` (their derivative is zero)."""
    amax = None
    for w in parts:
        if w is None:
            continue
        a = w.detach().abs().reshape(w.shape[0], -1).amax(-1)
        amax = a if amax is None else torch.maximum(amax, a)
    return torch.where(amax > 0, _FD_STEP / amax.clamp_min(1e-300), torch.ones_like(amax))


class _D4Vjp(torch.autograd.Function):
    """``(dL/dpositions, dL/dq)`` for ``L = sum gout * E`` -- the analytic VJP kernels -- as a node that
    can be differentiated once more.

    Second derivatives are SEMI-NUMERICAL, the standard route for dispersion Hessians: with the
    upstream direction ``w = (w_pos, w_q)`` the node needs ``d/dx (g . w) = D_w g`` (the Hessian is
    symmetric), the directional derivative of the ANALYTIC gradient, and ``d/dgout (g . w) = D_w E``.
    Both are fourth-order central differences of the kernels along ``w`` (four gradient and four
    energy launches, step 5e-4 Bohr on the largest component: truncation ~1e-11 relative,
    round-off ~1e-13), far inside the 1e-7 the reference's Hessian tests ask for
    (``test/test_grad/test_hessian.py:47``).  Everything stays on the CUDA kernels."""

    @staticmethod
    def forward(ctx, gout, positions, q, numbers, par, engine):
        gpos, gq = engine.gradient(par, numbers, positions.detach(), q.detach(), gout.detach().contiguous(), True, True)
        ctx.save_for_backward(gout, positions, q, numbers)
        ctx.par, ctx.engine = par, engine
        return gpos, gq

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, wpos, wq):
        gout, positions, q, numbers = ctx.saved_tensors
        par, engine = ctx.par, ctx.engine
        t = _fd_direction(wpos, wq)
        dpos = wpos * t[:, None, None] if wpos is not None else None
        dq = wq * t[:, None] if wq is not None else None
        g = gout.contiguous()
        hpos = torch.zeros_like(positions)
        hq = torch.zeros_like(q)
        de = torch.zeros_like(q)
        for mult, coef in _FD_STENCIL:
            x = positions + mult * dpos if dpos is not None else positions
            y = q + mult * dq if dq is not None else q
            gp, gqq = engine.gradient(par, numbers, x.contiguous(), y.contiguous(), g, True, True)
            hpos += coef * gp
            hq += coef * gqq
            if ctx.needs_input_grad[0]:
                de += coef * engine.energy(par, numbers, x.contiguous(), y.contiguous())[0]
        return de / t[:, None], hpos / t[:, None, None], hq / t[:, None], None, None, None


_PARAM_ORDER = ("s6", "s8", "s9", "s10", "a1", "a2", "alp")


def _param_tensors(param: Param) -> tuple:
    """The damping parameters the caller differentiates (tensors with ``requires_grad``), in the
    order of :data:`_PARAM_ORDER`; empty when there is none."""
    out = tuple(
        v if isinstance(v, Tensor) and v.requires_grad else None for v in (param.get(k) for k in _PARAM_ORDER)
    )
    return out if any(t is not None for t in out) else ()


def _param_vjp(engine: _Engine, par: _lib.Params, numbers: Tensor, positions: Tensor, q: Tensor,
               gout: Tensor) -> Tensor:  # fmt: skip
    """``d(sum g E)/d(s6, s8, s9, s10, a1, a2, alp)`` (float64 ``[7]``) on device: coordination
    numbers, reference weights and pair C6 of both flavours from the model-level kernels, then
    ``d4b200_param_vjp_*`` (csrc/d4b200_param.cu).  The reference differentiates its dense tape
    instead (test/test_grad/test_param.py:40-100)."""
    from .model import D4Model, D4SModel

    nbatch, nat = numbers.shape
    # the dense (N, N) intermediates (pair planes, C6 matrices, D4S weights) are processed in
    # slices of the batch: at most 32768 structures (grid limit of the kernels) and ~2 GB of planes
    per_structure = 8 * nat * nat * (6 + 2 + (14 if par.model == 1 else 0))
    chunk = max(1, min(32768, int(2e9 // max(per_structure, 1))))
    if nbatch > chunk:
        total = None
        for b0 in range(0, nbatch, chunk):
            sl = slice(b0, min(b0 + chunk, nbatch))
            part = _param_vjp(engine, par, numbers[sl].contiguous(), positions[sl].contiguous(),
                              q[sl].contiguous(), gout[sl].contiguous())  # fmt: skip
            total = part if total is None else total + part
        return total
    with torch.no_grad():
        pos = positions.detach()
        qd = q.detach()
        cn = get_properties(numbers, pos, q=torch.zeros_like(qd))[0]
        cls = D4SModel if par.model == 1 else D4Model
        model = cls(numbers, ga=engine.ga, gc=engine.gc, wf=par.wf, dtype=pos.dtype)
        c6q = model.get_atomic_c6(model.weight_references(cn, qd))
        c60 = model.get_atomic_c6(model.weight_references(cn, None))
        need = int(engine.lib.d4b200_param_vjp_workspace_bytes(nbatch, nat))
        ws = torch.empty(max(need, 8), dtype=torch.uint8, device=pos.device)
        out = torch.empty((nbatch, 7), dtype=torch.float64, device=pos.device)
        stream = torch.cuda.current_stream(pos.device).cuda_stream
        fn = engine.lib.d4b200_param_vjp_f64 if pos.dtype == torch.float64 else engine.lib.d4b200_param_vjp_f32
        _lib.check(
            fn(engine.handle, C.byref(par), nbatch, nat, numbers.data_ptr(), pos.data_ptr(), c6q.data_ptr(),
               c60.data_ptr(), gout.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), stream),
            "d4b200_param_vjp",
        )  # fmt: skip
        return out.sum(0)


# ---------------------------------------------------------------------------
# argument handling
# ---------------------------------------------------------------------------
_SCALAR_CACHE: dict[int, tuple[Any, int, float]] = {}


def _scalar(v: Any, name: str) -> float:
    """Host value of a damping parameter.  Device tensors (the reference's
    examples build them with ``positions.new_tensor``) cost one synchronisation
    the first time they are seen; the value is cached per tensor version."""
    if isinstance(v, Tensor):
        if v.requires_grad:  # differentiated parameter: value here, gradient through _param_vjp
            return float(v.detach())
        if v.device.type == "cpu":
            return float(v)
        hit = _SCALAR_CACHE.get(id(v))
        if hit is not None and hit[0]() is v and hit[1] == v._version:
            return hit[2]
        val = float(v)
        if len(_SCALAR_CACHE) > 256:
            _SCALAR_CACHE.clear()
        _SCALAR_CACHE[id(v)] = (weakref.ref(v), v._version, val)
        return val
    return float(v)


def _flatten_param(param: Param, cutoff: Cutoff | None, model_id: int, wf: float) -> _lib.Params:
    """``Param`` -> POD with the reference's defaults (twobody.py:177-178,
    threebody.py:233-234, functions.py:255-259)."""
    if param.get("a1") is None or param.get("a2") is None:
        missing = [k for k in ("a1", "a2") if param.get(k) is None]
        raise TypeError(f"RationalDamping (order 6) requires keyword(s): {', '.join(missing)}")
    par = _lib.Params()
    par.s6 = _scalar(param.get("s6", defaults.S6), "s6")
    par.s8 = _scalar(param.get("s8", defaults.S8), "s8")
    par.s9 = _scalar(param.get("s9", defaults.S9), "s9")
    par.has_s10 = 1 if "s10" in param and param["s10"] is not None else 0
    par.s10 = _scalar(param["s10"], "s10") if par.has_s10 else 0.0
    par.a1 = _scalar(param["a1"], "a1")
    par.a2 = _scalar(param["a2"], "a2")
    par.alp = _scalar(param.get("alp", defaults.ALP), "alp")
    if cutoff is None:
        par.disp2_cutoff, par.disp3_cutoff = defaults.D4_DISP2_CUTOFF, defaults.D4_DISP3_CUTOFF
    elif hasattr(cutoff, "as_float"):
        par.disp2_cutoff, par.disp3_cutoff = cutoff.as_float("disp2"), cutoff.as_float("disp3")
    else:  # e.g. the reference's own Cutoff object
        par.disp2_cutoff, par.disp3_cutoff = float(cutoff.disp2), float(cutoff.disp3)
    par.cn_cutoff = defaults.D4_CN_CUTOFF  # Cutoff.cn is not forwarded (dispersion/base.py:390)
    par.wf = wf
    par.model = model_id
    return par


class _ModelSpec(tuple):
    """(model id, ga, gc, wf) -- unpacks like the plain tuple -- plus ``ref_charges`` and ``c9_frequency``."""

    ref_charges = "eeq"
    c9_frequency: int | None = None


class _FrequencySlice:
    """A model restricted to one Casimir-Polder node of the exact C9 (``tables.ElementTables``); used by
    ``dispersion.D4ATMExact`` only."""

    def __init__(self, model: Any, frequency: int):
        self.model, self.frequency = model, int(frequency)


def _resolve_model(model: Any) -> "_ModelSpec":
    """-> (model id, ga, gc, wf), with ``.ref_charges`` ("eeq" | "gfn2") and ``.c9_frequency``"""
    if isinstance(model, _FrequencySlice):
        spec = _resolve_model(model.model)
        if spec[0] != 0:
            # the reference averages the pair-resolved D4S weights over the partners here
            # (model/d4s.py:292-316); not provided
            raise NotImplementedError("the exact C9 (D4ATMExact) is accelerated for model='d4' only")
        spec.c9_frequency = model.frequency
        return spec
    spec = _ModelSpec(_resolve_model_tuple(model))
    ref = getattr(model, "ref_charges", "eeq") if not isinstance(model, str) else "eeq"
    if ref not in ("eeq", "gfn2"):
        raise ValueError(f"Unknown reference charges: {ref}")
    spec.ref_charges = ref
    return spec


def _resolve_model_tuple(model: Any) -> tuple[int, float, float, float]:
    if isinstance(model, str):
        key = model.casefold()
        if key not in _ALLOWED_MODELS:
            raise ValueError(f"Unknown model '{key}'. Please use {', '.join(_ALLOWED_MODELS)}.")
        if key == "d4":
            return 0, defaults.GA_DEFAULT, defaults.GC_DEFAULT, defaults.WF_DEFAULT
        if key == "d4s":
            return 1, defaults.GA_DEFAULT, defaults.GC_DEFAULT, defaults.WF_DEFAULT
        raise NotImplementedError(f"model '{key}' is outside the accelerated D4 hot path")
    name = type(model).__name__
    if name in ("D4Model", "D4SModel"):
        wf = float(getattr(model, "wf", defaults.WF_DEFAULT))
        if name == "D4SModel" and wf != defaults.WF_DEFAULT:
            # the D4S kernels weight with the pair table wfpair (model/d4s.py:61-67); the reference
            # would apply the explicit uniform wf instead (model/base.py:150): not provided
            raise NotImplementedError("D4SModel with an explicit weighting factor wf is not accelerated")
        return (0 if name == "D4Model" else 1), float(model.ga), float(model.gc), wf
    raise NotImplementedError(f"model instance of type {name} is outside the accelerated D4 hot path")


def _small_limit(engine: "_Engine", dtype: torch.dtype, grad: bool, model_id: int) -> int:
    """Largest structure of the one-CTA-per-structure kernels (csrc/d4b200_flavour.cuh)."""
    key = (dtype == torch.float32, bool(grad), int(model_id))
    lim = engine.limits.get(key)
    if lim is None:
        lim = int(engine.lib.d4b200_small_limit(engine.handle, int(key[0]), int(key[1]), key[2]))
        if lim <= 0:
            raise _lib.D4B200Error(f"d4b200_small_limit failed with code {lim}")
        engine.limits[key] = lim
    return lim


def _compact_front(numbers: Tensor, positions: Tensor, q: Tensor, width: int):
    """Move the real atoms of every structure to the front and cut the atom axis to
    ``width`` (all structures here have at most ``width`` real atoms)."""
    order = torch.argsort((numbers == 0).to(torch.int8), dim=-1, stable=True)[:, :width]
    return (
        torch.gather(numbers, 1, order).contiguous(),
        torch.gather(positions, 1, order.unsqueeze(-1).expand(-1, -1, 3)).contiguous(),
        torch.gather(q, 1, order).contiguous(),
        order,
    )


def _scatter_back(values: Tensor, order: Tensor, nat: int) -> Tensor:
    return values.new_zeros((values.shape[0], nat)).scatter(1, order, values)


def _eeq_charges(numbers: Tensor, positions: Tensor, charge, cutoff: Cutoff | None) -> Tensor:
    """``get_eeq_charges(numbers, positions, charge, cutoff=cutoff.cn_eeq)`` of the reference
    (dispersion/base.py:401-407) on device: ``tad_dftd4_b200.eeq``."""
    from .eeq import get_eeq_charges

    if cutoff is None:
        cut = defaults.D4_CN_EEQ_CUTOFF
    elif hasattr(cutoff, "as_float"):
        cut = cutoff.as_float("cn_eeq")
    else:  # a foreign Cutoff object (e.g. the reference's own): tensor attributes
        cut = float(cutoff.cn_eeq)
    return get_eeq_charges(numbers, positions, charge, cutoff=cut)


def _check_arguments(numbers, positions, model, rcov=None, r4r2=None, rvdw=None, q=None, cn_function=None,
                     counting_function=None, damping_function=None) -> "_ModelSpec":  # fmt: skip
    """The argument checks of ``Disp.calculate`` (dispersion/base.py:354-399), in the reference's order: shapes
    first, then what lies outside the accelerated path.  Returns the resolved model."""
    if numbers.shape != positions.shape[:-1]:
        raise ValueError(
            f"Shape of positions ({positions.shape}) is not consistent "
            f"with atomic numbers ({numbers.shape}).",
        )
    spec = _resolve_model(model)
    for name, val in (("covalent radii", rcov), ("expectation values r4r2", r4r2)):
        if val is not None and numbers.shape != val.shape:
            raise ValueError(
                f"Shape of {name} ({val.shape}) is not consistent with atomic numbers ({numbers.shape})."
            )
    if rvdw is not None and numbers.shape != rvdw.shape[:-1]:
        raise ValueError(
            f"Shape of van der Waals radii ({rvdw.shape}) is not "
            f"consistent with atomic numbers ({numbers.shape}).",
        )
    if rcov is not None or r4r2 is not None or rvdw is not None:
        raise NotImplementedError("custom rcov/r4r2/rvdw are outside the accelerated hot path")
    for name, fn in (("cn_function", cn_function), ("counting_function", counting_function)):
        if fn is not None and getattr(fn, "__name__", "") not in ("cn_d4", "erf_count"):
            raise NotImplementedError(f"custom {name} is outside the accelerated hot path")
    if damping_function is not None and type(damping_function).__name__ != "RationalDamping":
        raise NotImplementedError("only RationalDamping is accelerated")
    if q is not None and numbers.shape != q.shape:
        raise ValueError(
            f"Shape of atomic charges ({q.shape}) is not consistent "
            f"with atomic numbers ({numbers.shape}).",
        )
    return spec


def dftd4(
    numbers: Tensor,
    positions: Tensor,
    charge: Tensor | float | int,
    param: Param,
    *,
    model: Any = "d4",
    rcov: Tensor | None = None,
    r4r2: Tensor | None = None,
    rvdw: Tensor | None = None,
    q: Tensor | None = None,
    cutoff: Cutoff | None = None,
    cn_function: Any = None,
    counting_function: Any = None,
    damping_function: Any = None,
) -> Tensor:
    """Atom-resolved DFT-D4 dispersion energy, shape ``(..., nat)``.

    Same arguments, padding convention (``numbers == 0``), return value and
    exception types as ``tad_dftd4.dftd4`` (``disp.py:44-146``).  Only the
    default method is accelerated (erf-count ``cn_d4``, rational damping, BJ
    radii, approximate C9, default element radii); any other plugin raises
    ``NotImplementedError`` instead of silently falling back.
    """
    spec = _check_arguments(numbers, positions, model, rcov, r4r2, rvdw, q, cn_function, counting_function,
                            damping_function)  # fmt: skip
    model_id, ga, gc, wf = spec
    if param.get("a1") is None or param.get("a2") is None:
        # raised by the reference's RationalDamping on any device (damping/functions.py:255-259)
        missing = [k for k in ("a1", "a2") if param.get(k) is None]
        raise TypeError(f"RationalDamping (order 6) requires keyword(s): {', '.join(missing)}")
    if positions.dtype not in (torch.float64, torch.float32):
        raise NotImplementedError(f"dtype {positions.dtype} is not supported (float64/float32)")
    if positions.device.type != "cuda":
        raise RuntimeError(
            "tad_dftd4_b200 runs on B200 GPUs only (no CPU fallback): move numbers/positions "
            f"to a CUDA device (got {positions.device})."
        )
    if q is None:
        q = _eeq_charges(numbers, positions, charge, cutoff)
    if numbers.shape != q.shape:
        raise ValueError(
            f"Shape of atomic charges ({q.shape}) is not consistent "
            f"with atomic numbers ({numbers.shape}).",
        )
    par = _flatten_param(param, cutoff, model_id, wf)
    ptens = _param_tensors(param)

    if spec.ref_charges == "gfn2" and _CHECKS and numbers.numel() and int(numbers.max()) > 86:
        # the reference indexes its (87, 7) GFN2 tables with the atomic numbers (model/d4.py:154)
        raise IndexError("ref_charges='gfn2' is tabulated for Z <= 86")
    engine = _Engine.get(positions.device, ga, gc, spec.ref_charges, spec.c9_frequency)
    nat = numbers.shape[-1]
    batch_shape = numbers.shape[:-1]
    num2 = numbers.reshape(-1, nat).to(torch.int64).contiguous()
    pos2 = positions.reshape(-1, nat, 3).contiguous()
    q2 = q.to(positions.dtype).reshape(-1, nat).contiguous()

    limit = _small_limit(engine, positions.dtype, positions.requires_grad or q.requires_grad, model_id)
    if nat > limit:
        # padded width beyond the one-CTA-per-structure kernels: structures that really
        # are that large go through the tiled kernels one by one (one host sync)
        counts = (num2 != 0).sum(-1)
        big = torch.nonzero(counts > limit).flatten().tolist()
        if big:
            if model_id != 0:
                raise NotImplementedError("the tiled large-system path supports model='d4' only")
            from .large import dftd4_large

            rows: list[Tensor | None] = [None] * num2.shape[0]
            small = [b for b in range(num2.shape[0]) if b not in set(big)]
            for b in big:
                rows[b] = dftd4_large(num2[b], pos2[b], param, q2[b], cutoff=cutoff,
                                      model=(ga, gc, wf, spec.ref_charges, spec.c9_frequency))
            if small:
                # compact the small structures to the front of the atom axis
                sel = torch.tensor(small, device=num2.device)
                ns, ps, qs, back = _compact_front(num2[sel], pos2[sel], q2[sel], limit)
                with torch.cuda.device(positions.device):
                    es = _scatter_back(_D4Function.apply(ps, qs, ns, par, engine, *ptens), back, nat)
                for n, b in enumerate(small):
                    rows[b] = es[n]
            return torch.stack(rows).reshape(*batch_shape, nat)
    with torch.cuda.device(positions.device):
        energy = _D4Function.apply(pos2, q2, num2, par, engine, *ptens)
    return energy.reshape(*batch_shape, nat)


def dftd4_host(
    numbers: Tensor,
    positions: Tensor,
    charge: Tensor | float | int,
    param: Param,
    *,
    q: Tensor,
    model: Any = "d4",
    cutoff: Cutoff | None = None,
    device: int | torch.device = 0,
    chunks: int = 0,
    out: Tensor | None = None,
    with_gradient: bool = False,
    out_gradient: Tensor | None = None,
):
    """Atom-resolved D4 energy for inputs that live in HOST memory, evaluated on the B200.

    Same arguments as :func:`dftd4` but ``numbers`` / ``positions`` / ``q`` are CPU tensors
    (pin them for full copy/compute overlap) and the result is a (pinned) CPU tensor.  The
    batch is pipelined in chunks through the C ABI (``d4b200_energy_host_*``): the
    host-to-device copy of one chunk overlaps the kernels of the previous one.

    ``with_gradient=True`` returns ``(energy, gradient)`` with ``gradient = d(energy.sum())/d
    positions`` (what ``torch.autograd.grad(energy.sum(), positions)`` gives in the reference's
    ``examples/forces.py``) from the fused energy+gradient kernels
    (``d4b200_energy_gradient_host_*``); for other upstream gradients use :func:`dftd4` with
    CUDA tensors."""
    if numbers.shape != positions.shape[:-1]:
        raise ValueError(
            f"Shape of positions ({positions.shape}) is not consistent "
            f"with atomic numbers ({numbers.shape}).",
        )
    if numbers.shape != q.shape:
        raise ValueError(
            f"Shape of atomic charges ({q.shape}) is not consistent with atomic numbers ({numbers.shape})."
        )
    if positions.device.type != "cpu" or positions.dtype not in (torch.float64, torch.float32):
        raise ValueError("dftd4_host expects float32/float64 CPU tensors")
    spec = _resolve_model(model)
    model_id, ga, gc, wf = spec
    par = _flatten_param(param, cutoff, model_id, wf)
    dev = torch.device("cuda", device) if isinstance(device, int) else device
    engine = _Engine.get(dev, ga, gc, spec.ref_charges, spec.c9_frequency)
    nat = numbers.shape[-1]
    if nat > _small_limit(engine, positions.dtype, bool(with_gradient), model_id):
        raise NotImplementedError("dftd4_host handles batches of small structures; use dftd4 for large ones")
    # atomic numbers travel as the caller holds them: uint8 / int32 tensors are uploaded narrow and
    # widened on the device (d4b200_energy_host_z_*); anything else goes as int64
    if numbers.dtype in (torch.uint8, torch.int32, torch.int64):
        num2 = numbers.reshape(-1, nat).contiguous()
    else:
        num2 = numbers.reshape(-1, nat).to(torch.int64).contiguous()
    pos2 = positions.reshape(-1, nat, 3).contiguous()
    q2 = q.to(positions.dtype).reshape(-1, nat).contiguous()

    def _buffer(t, shape, what):
        if t is None:
            return torch.empty(shape, dtype=positions.dtype, pin_memory=True)
        if (t.device.type != "cpu" or t.dtype != positions.dtype or t.numel() != int(np.prod(shape))
                or not t.is_contiguous()):  # fmt: skip
            raise ValueError(f"{what} must be a contiguous CPU tensor of dtype {positions.dtype} with "
                             f"{int(np.prod(shape))} elements (got {tuple(t.shape)}, {t.dtype}, {t.device})")  # fmt: skip
        return t

    out = _buffer(out, num2.shape, "out")
    if with_gradient:
        out_gradient = _buffer(out_gradient, pos2.shape, "out_gradient")
    f64 = positions.dtype == torch.float64
    fn = engine.lib.d4b200_energy_host_z_f64 if f64 else engine.lib.d4b200_energy_host_z_f32
    bits = C.c_int(0)
    code = fn(engine.handle, C.byref(par), num2.shape[0], nat, num2.data_ptr(), num2.element_size(),
              pos2.data_ptr(), q2.data_ptr(), out.data_ptr(),
              out_gradient.data_ptr() if with_gradient else None, None, int(chunks), C.byref(bits))  # fmt: skip
    if bits.value & 1:
        raise ValueError("numbers contains an atomic number outside 1..103 (0 = padding).")
    if bits.value & 2:
        raise NotImplementedError("a structure is larger than the one-CTA-per-structure kernels support")
    _lib.check(code, "d4b200_energy_host_z")
    if with_gradient:
        return out.reshape(numbers.shape), out_gradient.reshape(positions.shape)
    return out.reshape(numbers.shape)


def get_properties(
    numbers: Tensor,
    positions: Tensor,
    charge: Tensor | float | int | None = None,
    cutoff: Cutoff | None = None,
    *,
    q: Tensor | None = None,
):
    """``(cn, q, c6, alpha)`` as ``tad_dftd4.get_properties`` (``disp.py:149-197``):
    D4 coordination numbers, atomic charges, pair C6 ``(..., nat, nat)`` and static
    polarizabilities.  ``q`` (keyword, extension of the reference signature) passes
    explicit charges; without it EEQ charges are solved on device (``tad_dftd4_b200.eeq``)."""
    if numbers.shape != positions.shape[:-1]:
        raise ValueError(
            f"Shape of positions ({positions.shape}) is not consistent "
            f"with atomic numbers ({numbers.shape}).",
        )
    if positions.device.type != "cuda":
        raise RuntimeError("tad_dftd4_b200 runs on B200 GPUs only (no CPU fallback).")
    if torch.is_grad_enabled() and (positions.requires_grad or (q is not None and q.requires_grad)):
        # the reference's get_properties stays on the autograd tape (disp.py:149-197); these kernels
        # return plain values: refuse rather than hand back silently detached tensors
        raise NotImplementedError(
            "get_properties returns non-differentiable values (cn, C6, alpha); call it under "
            "torch.no_grad() or with detached inputs, and differentiate dftd4() instead"
        )
    if q is None:
        q = _eeq_charges(numbers, positions, 0.0 if charge is None else charge, cutoff)
    if numbers.shape != q.shape:
        raise ValueError(
            f"Shape of atomic charges ({q.shape}) is not consistent with atomic numbers ({numbers.shape})."
        )
    par = _flatten_param({"a1": defaults.A1, "a2": defaults.A2}, cutoff, 0, defaults.WF_DEFAULT)
    engine = _Engine.get(positions.device, defaults.GA_DEFAULT, defaults.GC_DEFAULT)
    nat = numbers.shape[-1]
    batch_shape = numbers.shape[:-1]
    num2 = numbers.reshape(-1, nat).to(torch.int64).contiguous()
    pos2 = positions.detach().reshape(-1, nat, 3).contiguous()
    q2 = q.detach().to(positions.dtype).reshape(-1, nat).contiguous()
    with torch.cuda.device(positions.device):
        cn, c6, alpha = engine.properties(par, num2, pos2, q2)
    return (
        cn.reshape(*batch_shape, nat),
        q,
        c6.reshape(*batch_shape, nat, nat),
        alpha.reshape(*batch_shape, nat),
    )
