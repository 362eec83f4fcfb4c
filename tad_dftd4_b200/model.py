"""
Model descriptors with the constructor signature of the reference's ``D4Model`` /
``D4SModel`` (``/root/reference/src/tad_dftd4/model/base.py:107-151``).

In the reference a model instance owns the ``(N, N, 7, 7)`` reference-C6 tensor of one
set of atoms.  Here that tensor never exists (C6 is a 23-term dot product of per-atom
vectors, see ``tables.py``), so a model is only a bundle of hyper-parameters
``(ga, gc, wf)`` that selects the device tables; it can be passed as ``model=`` to
:func:`tad_dftd4_b200.dftd4` exactly like the reference's instances.
"""

from __future__ import annotations

import torch

from . import defaults

__all__ = ["D4Model", "D4SModel"]


class _Model:
    __slots__ = ("numbers", "ga", "gc", "wf", "ref_charges", "_device", "_dtype")
    _key = "d4"

    def __init__(self, numbers=None, ga: float = defaults.GA_DEFAULT, gc: float = defaults.GC_DEFAULT,
                 wf=None, ref_charges: str = "eeq", rc6=None, device=None, dtype=None) -> None:  # fmt: skip
        if ref_charges not in ("eeq", "gfn2"):
            raise ValueError(f"Unknown reference charges: {ref_charges}")
        if rc6 is not None:
            raise NotImplementedError("user-supplied rc6 tensors are outside the accelerated hot path")
        self.numbers = numbers
        self.ga = float(ga)
        self.gc = float(gc)
        self.wf = defaults.WF_DEFAULT if wf is None else float(wf)
        self.ref_charges = ref_charges
        self._device = device if device is not None else (
            numbers.device if isinstance(numbers, torch.Tensor) else torch.device("cpu"))
        self._dtype = dtype if dtype is not None else torch.get_default_dtype()

    # read-only like the reference's TensorLike base (test/test_model/test_general.py:30-66)
    @property
    def device(self):
        return self._device

    @property
    def dtype(self):
        return self._dtype

    def _clone(self, **kw):
        args = dict(numbers=self.numbers, ga=self.ga, gc=self.gc, wf=self.wf, ref_charges=self.ref_charges,
                    device=self._device, dtype=self._dtype)  # fmt: skip
        args.update(kw)
        return type(self)(**args)

    def type(self, dtype: torch.dtype):
        """Copy of the model with another floating-point type."""
        if dtype not in (torch.float16, torch.float32, torch.float64):
            raise ValueError(f"Only float types are allowed (got {dtype}).")
        return self._clone(dtype=dtype)

    def to(self, device):
        """Copy of the model on another device."""
        device = torch.device(device)
        numbers = self.numbers.to(device) if isinstance(self.numbers, torch.Tensor) else self.numbers
        return self._clone(numbers=numbers, device=device)

    def __repr__(self) -> str:  # pragma: no cover
        return f"{type(self).__name__}(ga={self.ga}, gc={self.gc}, wf={self.wf}, ref_charges={self.ref_charges})"

    # -- device evaluation of the reference's model methods (csrc/d4b200_model.cu) ------------
    def _call_setup(self):
        import ctypes as C

        from . import _lib
        from .disp import _Engine

        numbers = self.numbers
        if not isinstance(numbers, torch.Tensor):
            raise ValueError("the model was constructed without atomic numbers")
        if numbers.device.type != "cuda":
            raise RuntimeError(
                "tad_dftd4_b200 runs on B200 GPUs only (no CPU fallback): construct the model with "
                f"numbers on a CUDA device (got {numbers.device})."
            )
        engine = _Engine.get(numbers.device, self.ga, self.gc, self.ref_charges)
        par = _lib.Params()
        par.wf = self.wf
        par.model = 1 if self._key == "d4s" else 0
        nat = numbers.shape[-1]
        num2 = numbers.reshape(-1, nat).to(torch.int64).contiguous()
        stream = torch.cuda.current_stream(numbers.device).cuda_stream
        return C, _lib, engine, par, num2, nat, stream

    def _real(self, t, name: str):
        if t is None:
            return None
        if t.shape != self.numbers.shape:
            raise ValueError(f"Shape of {name} ({t.shape}) is not consistent with atomic numbers ({self.numbers.shape}).")
        return t.to(self.numbers.device, self.dtype).reshape(-1, self.numbers.shape[-1]).contiguous()

    def weight_references(self, cn=None, q=None, *, with_dgwdq: bool = False, with_dgwdcn: bool = False):
        """``zeta(q) * gw`` of shape ``(..., nat, 7)`` -- D4S: ``(..., nat, nat, 7)``, the weights of
        atom n as seen by partner m -- and optionally the derivatives w.r.t. cn and q, in the
        reference's order ``(gw[, dgwdcn][, dgwdq])`` (model/d4.py:103-228, model/d4s.py:109-266)."""
        if self.ref_charges not in ("eeq", "gfn2"):
            raise ValueError(f"Unknown reference charges: {self.ref_charges}")
        C, _lib, engine, par, num2, nat, stream = self._call_setup()
        if self.dtype not in (torch.float64, torch.float32):
            raise NotImplementedError(f"dtype {self.dtype} is not supported (float64/float32)")
        cn2, q2 = self._real(cn, "cn"), self._real(q, "q")
        shape = (*self.numbers.shape, nat, 7) if self._key == "d4s" else (*self.numbers.shape, 7)
        with torch.cuda.device(self.numbers.device):
            gw = torch.empty(shape, dtype=self.dtype, device=self.numbers.device)
            dcn = torch.empty_like(gw) if with_dgwdcn else None
            dq = torch.empty_like(gw) if with_dgwdq else None
            fn = engine.lib.d4b200_weight_references_f64 if self.dtype == torch.float64 else engine.lib.d4b200_weight_references_f32
            _lib.check(
                fn(engine.handle, C.byref(par), num2.shape[0], nat, num2.data_ptr(),
                   cn2.data_ptr() if cn2 is not None else None, q2.data_ptr() if q2 is not None else None,
                   gw.data_ptr(), dcn.data_ptr() if dcn is not None else None,
                   dq.data_ptr() if dq is not None else None, None, stream),
                "d4b200_weight_references",
            )  # fmt: skip
        out = [gw]
        if with_dgwdcn:
            out.append(dcn)
        if with_dgwdq:
            out.append(dq)
        return gw if len(out) == 1 else tuple(out)

    def _gw(self, gw):
        nat = self.numbers.shape[-1]
        want = (*self.numbers.shape, nat, 7) if self._key == "d4s" else (*self.numbers.shape, 7)
        if tuple(gw.shape) != want:
            raise ValueError(f"Shape of the weights ({tuple(gw.shape)}) is not consistent with atomic numbers: expected {want}.")
        if gw.dtype not in (torch.float64, torch.float32):
            raise NotImplementedError(f"dtype {gw.dtype} is not supported (float64/float32)")
        return gw.to(self.numbers.device).contiguous()

    def get_atomic_c6(self, gw):
        """Pair C6 ``(..., nat, nat)`` from the reference weights (model/d4.py:268-289,
        model/d4s.py:268-290): contraction with the per-element-pair reference C6 table."""
        C, _lib, engine, par, num2, nat, stream = self._call_setup()
        gw = self._gw(gw)
        with torch.cuda.device(self.numbers.device):
            c6 = torch.empty((*self.numbers.shape, nat), dtype=gw.dtype, device=gw.device)
            fn = engine.lib.d4b200_atomic_c6_f64 if gw.dtype == torch.float64 else engine.lib.d4b200_atomic_c6_f32
            _lib.check(fn(engine.handle, par.model, num2.shape[0], nat, num2.data_ptr(), gw.data_ptr(),
                          c6.data_ptr(), stream), "d4b200_atomic_c6")  # fmt: skip
        return c6

    def _pols(self, gw, nfreq: int):
        if self._key == "d4s":
            raise NotImplementedError("polarizabilities are defined for atom-wise weights (D4Model)")
        C, _lib, engine, par, num2, nat, stream = self._call_setup()
        gw = self._gw(gw)
        with torch.cuda.device(self.numbers.device):
            out = torch.empty((*self.numbers.shape, nfreq), dtype=gw.dtype, device=gw.device)
            fn = engine.lib.d4b200_weighted_pols_f64 if gw.dtype == torch.float64 else engine.lib.d4b200_weighted_pols_f32
            _lib.check(fn(engine.handle, num2.shape[0], nat, nfreq, num2.data_ptr(), gw.data_ptr(),
                          out.data_ptr(), stream), "d4b200_weighted_pols")  # fmt: skip
        return out

    def get_weighted_pols(self, gw):
        """Weighted polarizabilities ``(..., nat, 23)`` (model/d4.py:291-307)."""
        return self._pols(gw, 23)

    def get_polarizabilities(self, weights):
        """Static polarizabilities ``(..., nat)`` (model/base.py:286-302)."""
        return self._pols(weights, 1)[..., 0]


class D4Model(_Model):
    """D4 model: one Gaussian weighting factor ``wf`` (default 6)."""


class D4SModel(_Model):
    """D4S model: element-pair specific weighting factors (``data/wfpair.py`` of the reference)."""

    _key = "d4s"
