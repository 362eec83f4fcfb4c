"""
Class-based interface mirroring the reference's plugin layer
(``/root/reference/src/tad_dftd4/dispersion/base.py``: ``Disp``, ``DispTerm``;
``dispersion/d4.py``: ``DispD4``, ``D4ATMApprox``; ``dispersion/twobody.py``:
``TwoBodyTerm``).

The reference sums ``term.calculate(...)`` over registered terms.  Here the registered
terms only *select* which parts of the fused kernel contribute: the default pair
``TwoBodyTerm(Rational, charge_dependent=True)`` + ``D4ATMApprox(Zero,
charge_dependent=False)`` runs in one launch; a single registered term maps onto the
same kernel with the other part switched off (``s9 = 0`` or ``s6 = s8 = 0``).  Any other
combination is outside the accelerated path and raises ``NotImplementedError``.
"""

from __future__ import annotations

from typing import Any

import torch

from .damping import RationalDamping, ZeroDamping
from .disp import dftd4

__all__ = ["Disp", "DispD4", "DispTerm", "TwoBodyTerm", "D4ATMApprox", "FusedD4Term", "ZeroDamping"]


class DispTerm:
    """Base class of the dispersion terms (``dispersion/base.py:38-101``)."""

    def __init__(self, damping_fn: Any, charge_dependent: bool):
        self.damping_fn = damping_fn
        self.charge_dependent = charge_dependent

    def __eq__(self, other: Any) -> bool:
        if self.__class__ is not other.__class__:
            return False
        return self.damping_fn == other.damping_fn and self.charge_dependent == other.charge_dependent

    def __hash__(self) -> int:
        return hash((type(self).__name__, self.charge_dependent))


class TwoBodyTerm(DispTerm):
    def __init__(self, *, damping_fn: Any = None, charge_dependent: bool = True):
        super().__init__(damping_fn if damping_fn is not None else RationalDamping(), charge_dependent)


class D4ATMApprox(DispTerm):
    def __init__(self, *, damping_fn: Any = None, charge_dependent: bool = False):
        super().__init__(damping_fn if damping_fn is not None else ZeroDamping(), charge_dependent)


class FusedD4Term(DispTerm):
    """Two-body + ATM in one launch (what ``DispD4`` registers as two terms)."""

    def __init__(self):
        super().__init__(RationalDamping(), True)


class Disp:
    """Dispersion calculator (``dispersion/base.py:140-431``)."""

    _ALLOWED_MODELS = ("d3", "d4", "d4s", "d5")
    TERMS: list[tuple[type, dict[str, Any] | None]] = []

    def __init__(self, model: Any = "d4", model_kwargs: dict[str, Any] | None = None, cn_fn: Any = None,
                 cn_fn_kwargs: dict[str, Any] | None = None, *, device=None, dtype=None):  # fmt: skip
        if isinstance(model, str):
            key = model.casefold()
            if key not in self._ALLOWED_MODELS:
                raise ValueError(f"Unknown model '{key}'. Please use {', '.join(self._ALLOWED_MODELS)}.")
            if model_kwargs:
                from .model import D4Model, D4SModel

                cls = {"d4": D4Model, "d4s": D4SModel}.get(key)
                if cls is None:
                    raise NotImplementedError(f"model '{key}' is outside the accelerated D4 hot path")
                model = cls(**model_kwargs)
        self.model = model
        self.cn_fn = cn_fn
        self.cn_fn_kwargs = cn_fn_kwargs or {}
        self.device, self.dtype = device, dtype
        self.terms: list[DispTerm] = []
        for term_cls, kw in self.TERMS:
            self.register(term_cls(**(kw or {})))

    def register(self, term: DispTerm) -> None:
        self.terms.append(term)

    def deregister(self, term: DispTerm) -> None:
        self.terms.remove(term)

    def calculate(self, numbers, positions, charge, param, *, cutoff=None, q=None, rcov=None,
                  r4r2=None, rvdw=None):  # fmt: skip
        is_c_dep = any(t.charge_dependent for t in self.terms)
        if q is not None and is_c_dep is False:
            raise RuntimeError(
                "Atomic charges are explicitly provided, but no term "
                "requires them. Please remove the `q` argument or "
                "provide a term that requires atomic charges.",
            )
        kinds = sorted(type(t).__name__ for t in self.terms)
        par = dict(param)
        if kinds == ["D4ATMApprox", "TwoBodyTerm"] or kinds == ["FusedD4Term"]:
            two = next((t for t in self.terms if isinstance(t, (TwoBodyTerm, FusedD4Term))))
            atm = next((t for t in self.terms if isinstance(t, D4ATMApprox)), None)
            if not two.charge_dependent or (atm is not None and atm.charge_dependent):
                raise NotImplementedError(
                    "only TwoBodyTerm(charge_dependent=True) + D4ATMApprox(charge_dependent=False) is fused"
                )
        elif kinds == ["TwoBodyTerm"]:
            if not self.terms[0].charge_dependent:
                raise NotImplementedError("charge-independent two-body term is outside the accelerated path")
            par["s9"] = 0.0
        elif kinds == ["D4ATMApprox"]:
            if self.terms[0].charge_dependent:
                raise NotImplementedError("charge-dependent ATM term is outside the accelerated path")
            par["s6"], par["s8"] = 0.0, 0.0
            par.pop("s10", None)
            if q is None:  # the ATM term does not use charges
                q = torch.zeros(numbers.shape, dtype=positions.dtype, device=positions.device)
        elif not kinds:
            return torch.zeros(numbers.shape, dtype=positions.dtype, device=positions.device)
        else:
            raise NotImplementedError(f"term combination {kinds} is outside the accelerated D4 hot path")
        for t in self.terms:
            if isinstance(t, (TwoBodyTerm, FusedD4Term)) and type(t.damping_fn).__name__ != "RationalDamping":
                raise NotImplementedError("only RationalDamping is accelerated for the two-body term")
            if isinstance(t, D4ATMApprox) and type(t.damping_fn).__name__ != "ZeroDamping":
                raise NotImplementedError("only ZeroDamping is accelerated for the ATM term")
        return dftd4(numbers, positions, charge, par, model=self.model, rcov=rcov, r4r2=r4r2, rvdw=rvdw,
                     q=q, cutoff=cutoff, cn_function=self.cn_fn)  # fmt: skip


class DispD4(Disp):
    """Standard DFT-D4 (``dispersion/d4.py:48-61``)."""

    TERMS = [
        (TwoBodyTerm, {"damping_fn": RationalDamping(), "charge_dependent": True}),
        (D4ATMApprox, {"damping_fn": ZeroDamping(), "charge_dependent": False}),
    ]
