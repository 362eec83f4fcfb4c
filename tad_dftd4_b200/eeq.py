"""
EEQ-2019 atomic partial charges on device -- stands in for
``tad_multicharge.get_eeq_charges`` (third-party ``tad-multicharge==0.5.0``, not
vendored by the reference; call sites ``/root/reference/src/tad_dftd4/dispersion/
base.py:401-407`` and ``disp.py:190``), the step immediately before the D4 hot
path.  With it ``dftd4(numbers, positions, charge, param)``
works without ``q=`` exactly like the reference, forces included (the charges
are on the autograd tape through ``d4b200_eeq_vjp_*``).

* padded batches with ``nat <= d4b200_eeq_limit()`` (160): one CUDA launch
  (``csrc/d4b200_eeq.cu``: one CTA per structure, bordered Coulomb system built and
  eliminated in shared memory), analytic vector-Jacobian product as a second kernel;
* larger structures: the same equations as dense torch operations on the caller's
  CUDA device (cuSOLVER through ``torch.linalg.solve``; a 20 000-atom system is a
  3.2 GB matrix) -- library code outside the hot path, differentiable by autograd.

No CPU fallback: CPU tensors raise like everywhere else in this package.
"""

from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib, defaults
from .data import eeq2019 as _p
from .data import elements as _el

__all__ = ["get_eeq_charges", "eeq_limit"]

Tensor = torch.Tensor

EEQ_KCN = 7.5
EEQ_CN_MAX = defaults.D4_CN_EEQ_MAX


def _param_blob() -> np.ndarray:
    n = _p.EEQ_MAX_ELEMENT + 1
    rcov = np.asarray(_el.COV_2009[:n], dtype=np.float64) * _el.AA2AU * 4.0 / 3.0
    blob = np.stack([
        np.asarray(_p.EEQ_CHI, dtype=np.float64), np.asarray(_p.EEQ_ETA, dtype=np.float64),
        np.asarray(_p.EEQ_KCN, dtype=np.float64), np.asarray(_p.EEQ_RAD, dtype=np.float64), rcov,
    ])  # fmt: skip
    return np.ascontiguousarray(blob)


class _EeqEngine:
    _cache: dict[int, "_EeqEngine"] = {}

    def __init__(self, index: int):
        self.lib = _lib.load()
        blob = _param_blob()
        handle = C.c_void_p()
        _lib.check(self.lib.d4b200_eeq_create(index, blob.ctypes.data, blob.size, C.byref(handle)),
                   "d4b200_eeq_create")  # fmt: skip
        self.handle = handle
        self.device = torch.device("cuda", index)
        self.limit = int(self.lib.d4b200_eeq_limit())
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)

    @classmethod
    def get(cls, device: torch.device) -> "_EeqEngine":
        index = device.index if device.index is not None else torch.cuda.current_device()
        eng = cls._cache.get(index)
        if eng is None:
            eng = cls._cache[index] = cls(index)
        return eng

    def _check(self) -> None:
        from . import disp

        if disp._CHECKS and int(self.status.item()) != 0:  # one stream sync, like Engine._status
            self.status.zero_()
            raise ValueError("numbers contains an atomic number outside 1..86 (EEQ-2019; 0 = padding).")

    #: the factor of the bordered matrix is kept for the backward pass up to this many bytes per call
    FACTOR_LIMIT = 4 << 30

    def charges(self, numbers: Tensor, positions: Tensor, charge: Tensor, cutoff: float, keep_factor: bool = False):
        """Charges; with ``keep_factor`` also the eliminated matrix of every structure (float64, opaque
        layout of ``d4b200_eeq_charges_factor_*``) for :meth:`vjp` -- or ``None`` when it would not fit."""
        nbatch, nat = numbers.shape
        q = torch.empty((nbatch, nat), dtype=positions.dtype, device=positions.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        f64 = positions.dtype == torch.float64
        args = (self.handle, nbatch, nat, numbers.data_ptr(), positions.data_ptr(), charge.data_ptr(), float(cutoff),
                q.data_ptr())  # fmt: skip
        if not keep_factor:
            fn = self.lib.d4b200_eeq_charges_f64 if f64 else self.lib.d4b200_eeq_charges_f32
            _lib.check(fn(*args, self.status.data_ptr(), stream), "d4b200_eeq_charges")
            self._check()
            return q
        per = int(self.lib.d4b200_eeq_factor_doubles(nat))
        factor = None
        if 0 < per * nbatch * 8 <= self.FACTOR_LIMIT:
            factor = torch.empty((nbatch, per), dtype=torch.float64, device=positions.device)
        fn = self.lib.d4b200_eeq_charges_factor_f64 if f64 else self.lib.d4b200_eeq_charges_factor_f32
        _lib.check(fn(*args, factor.data_ptr() if factor is not None else None, self.status.data_ptr(), stream),
                   "d4b200_eeq_charges_factor")  # fmt: skip
        self._check()
        return q, factor

    def vjp(self, numbers: Tensor, positions: Tensor, cutoff: float, q: Tensor, gq: Tensor,
            factor: Tensor | None = None) -> Tensor:  # fmt: skip
        """``J^T gq``; with the ``factor`` kept by :meth:`charges` for the same geometry the matrix is not
        eliminated again."""
        nbatch, nat = numbers.shape
        gpos = torch.empty_like(positions)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        f64 = positions.dtype == torch.float64
        head = (self.handle, nbatch, nat, numbers.data_ptr(), positions.data_ptr(), float(cutoff), q.data_ptr(),
                gq.data_ptr())  # fmt: skip
        if factor is None:
            fn = self.lib.d4b200_eeq_vjp_f64 if f64 else self.lib.d4b200_eeq_vjp_f32
            _lib.check(fn(*head, gpos.data_ptr(), None, stream), "d4b200_eeq_vjp")
        else:
            fn = self.lib.d4b200_eeq_vjp_factor_f64 if f64 else self.lib.d4b200_eeq_vjp_factor_f32
            _lib.check(fn(*head, factor.data_ptr(), gpos.data_ptr(), None, stream), "d4b200_eeq_vjp_factor")
        return gpos


class _EeqFunction(torch.autograd.Function):
    """q[b, i] = EEQ(numbers, positions, charge); backward = analytic VJP kernel."""

    @staticmethod
    def forward(ctx, positions: Tensor, numbers: Tensor, charge: Tensor, cutoff: float, engine: _EeqEngine):
        if ctx.needs_input_grad[0]:  # the backward pass reuses the eliminated matrix
            q, ctx.factor = engine.charges(numbers, positions, charge, cutoff, keep_factor=True)
        else:
            q, ctx.factor = engine.charges(numbers, positions, charge, cutoff), None
        ctx.save_for_backward(positions, numbers, q)
        ctx.cutoff = cutoff
        ctx.engine = engine
        ctx.charge = charge
        return q

    @staticmethod
    def backward(ctx, gq: Tensor):
        positions, numbers, q = ctx.saved_tensors
        if torch.is_grad_enabled():  # create_graph=True: keep the VJP on the tape (second derivatives)
            return _EeqVjp.apply(gq, positions, numbers, q, ctx.charge, ctx.cutoff, ctx.engine), None, None, None, None
        with torch.no_grad():
            gpos = ctx.engine.vjp(numbers, positions, ctx.cutoff, q, gq.contiguous(), ctx.factor)
        return gpos, None, None, None, None


class _EeqVjp(torch.autograd.Function):
    """``dL/dpositions = J^T gq`` of the EEQ charges (analytic VJP kernel) as a node that can be
    differentiated once more, semi-numerically like ``disp._D4Vjp``: along the upstream direction
    ``w``, ``d/dx (gq^T J w) = D_w (J^T gq)`` and ``d/dgq (gq^T J w) = J w = D_w q`` are fourth-order
    central differences of the VJP / charge kernels."""

    @staticmethod
    def forward(ctx, gq, positions, numbers, q, charge, cutoff, engine):
        ctx.save_for_backward(gq, positions, numbers, charge)
        ctx.cutoff, ctx.engine = cutoff, engine
        return engine.vjp(numbers, positions.detach(), cutoff, q.detach(), gq.detach().contiguous())

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, w):
        from .disp import _FD_STENCIL, _fd_direction

        gq, positions, numbers, charge = ctx.saved_tensors
        engine, cutoff = ctx.engine, ctx.cutoff
        t = _fd_direction(w)
        d = w * t[:, None, None]
        g = gq.contiguous()
        hpos = torch.zeros_like(positions)
        dq = torch.zeros_like(gq)
        for mult, coef in _FD_STENCIL:
            x = (positions + mult * d).contiguous()
            qx = engine.charges(numbers, x, charge, cutoff)
            dq += coef * qx
            hpos += coef * engine.vjp(numbers, x, cutoff, qx, g)
        return dq / t[:, None], hpos / t[:, None, None], None, None, None, None, None


_TABLES: dict[tuple[str, torch.dtype], Tensor] = {}


def _dense(numbers: Tensor, positions: Tensor, charge: Tensor, cutoff: float) -> Tensor:
    """The same equations with dense torch operations on the caller's device (structures beyond
    the shared-memory kernels); ``numbers`` [B, N], ``positions`` [B, N, 3], ``charge`` [B]."""
    dev, dtype = positions.device, positions.dtype
    key = (str(dev), dtype)
    tab = _TABLES.get(key)
    if tab is None:
        tab = _TABLES[key] = torch.from_numpy(_param_blob()).to(dev, dtype)
    if int(numbers.max()) > _p.EEQ_MAX_ELEMENT or int(numbers.min()) < 0:
        raise ValueError("numbers contains an atomic number outside 1..86 (EEQ-2019; 0 = padding).")
    chi, eta, kappa, rad, rcov = (t[numbers] for t in tab)
    real = numbers != 0
    n = numbers.shape[-1]
    eye = torch.eye(n, dtype=torch.bool, device=dev)
    mask = real.unsqueeze(-1) & real.unsqueeze(-2) & ~eye
    diff = positions.unsqueeze(-2) - positions.unsqueeze(-3)
    one = torch.ones((), dtype=dtype, device=dev)
    zero = torch.zeros((), dtype=dtype, device=dev)
    d = torch.sqrt(torch.where(mask, (diff * diff).sum(-1), one))
    del diff
    r0 = rcov.unsqueeze(-1) + rcov.unsqueeze(-2)
    count = torch.where(mask & (d <= cutoff), 0.5 * torch.erfc(EEQ_KCN * (d / r0 - 1.0)), zero)
    cn = count.sum(-1)
    del count, r0
    cn = math.log1p(math.exp(EEQ_CN_MAX)) - torch.log1p(torch.exp(EEQ_CN_MAX - cn))
    eps = torch.finfo(dtype).eps
    rhs = torch.where(real, -chi + kappa * torch.sqrt(torch.clamp(cn, min=eps)), zero)
    rhs = torch.cat((rhs, charge.unsqueeze(-1)), dim=-1)
    gam = torch.rsqrt(torch.where(mask, rad.unsqueeze(-1) ** 2 + rad.unsqueeze(-2) ** 2, one))
    diag = torch.where(real, eta + math.sqrt(2.0 / math.pi) / torch.where(real, rad, one), one)
    mat = torch.zeros((*numbers.shape[:-1], n + 1, n + 1), dtype=dtype, device=dev)
    mat[..., :n, :n] = torch.where(mask, torch.erf(d * gam) / d, zero) + torch.diag_embed(diag)
    del d, gam
    mat[..., :n, n] = real.to(dtype)
    mat[..., n, :n] = real.to(dtype)
    return torch.linalg.solve(mat, rhs)[..., :n]


def eeq_limit(device: torch.device | None = None) -> int:
    """Largest padded width the shared-memory EEQ kernels accept."""
    return int(_lib.load().d4b200_eeq_limit())


def get_eeq_charges(numbers: Tensor, positions: Tensor, chrg: Tensor | float | int, *,
                    cutoff: Tensor | float | None = None) -> Tensor:  # fmt: skip
    """EEQ-2019 charges ``(..., nat)`` of (batches of padded) structures with total charge
    ``chrg`` (``tad_multicharge.get_eeq_charges``); differentiable w.r.t. ``positions``."""
    if numbers.shape != positions.shape[:-1]:
        raise ValueError(
            f"Shape of positions ({positions.shape}) is not consistent with atomic numbers ({numbers.shape})."
        )
    if positions.device.type != "cuda":
        raise RuntimeError(
            "tad_dftd4_b200 runs on B200 GPUs only (no CPU fallback): move numbers/positions "
            f"to a CUDA device (got {positions.device})."
        )
    if positions.dtype not in (torch.float64, torch.float32):
        raise NotImplementedError(f"dtype {positions.dtype} is not supported (float64/float32)")
    cut = defaults.D4_CN_EEQ_CUTOFF if cutoff is None else float(cutoff)
    nat = numbers.shape[-1]
    batch_shape = numbers.shape[:-1]
    num2 = numbers.reshape(-1, nat).to(torch.int64).contiguous()
    pos2 = positions.reshape(-1, nat, 3).contiguous()
    nb = num2.shape[0]
    if isinstance(chrg, Tensor) and chrg.device.type == "cuda":
        chg = chrg.to(positions.device, positions.dtype).expand(batch_shape).reshape(-1).contiguous()
    elif isinstance(chrg, Tensor) and chrg.numel() > 1:
        chg = chrg.to(positions.device, positions.dtype, non_blocking=True).expand(batch_shape).reshape(-1).contiguous()
    else:  # host scalar: a device-side fill, no host-to-device copy and no synchronisation
        chg = torch.full((nb,), float(chrg), dtype=positions.dtype, device=positions.device)
    engine = _EeqEngine.get(positions.device)
    with torch.cuda.device(positions.device):
        if nat <= engine.limit:
            q = _EeqFunction.apply(pos2, num2, chg, cut, engine)
        else:
            q = _dense(num2, pos2, chg, cut)
    return q.reshape(*batch_shape, nat)
