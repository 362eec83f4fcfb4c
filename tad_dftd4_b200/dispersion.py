"""
Class-based interface mirroring the reference's plugin layer
(``/root/reference/src/tad_dftd4/dispersion/base.py``: ``Disp``, ``DispTerm``;
``dispersion/d4.py``: ``DispD4``, ``D4ATMApprox``; ``dispersion/twobody.py``:
``TwoBodyTerm``).

The reference sums ``term.calculate(...)`` over registered terms.  Here the registered
terms only *select* which parts of the fused kernel contribute: the default pair
``TwoBodyTerm(Rational, charge_dependent=True)`` + ``D4ATMApprox(Zero,
charge_dependent=False)`` runs in one launch; a single registered term maps onto the
same kernel with the other part switched off (``s9 = 0`` or ``s6 = s8 = 0``).  The exact
Casimir-Polder C9 (``D4ATMExact`` / ``DispD4Exact``, ``dispersion/d4.py:67-84``) is the sum of 23
ATM-only launches of the same kernels, one per integration node (``tables.ElementTables``).  Any
other combination is outside the accelerated path and raises ``NotImplementedError``.
"""

from __future__ import annotations

from typing import Any

import torch

from . import defaults
from .damping import RationalDamping, ZeroDamping
from .disp import _check_arguments, _FrequencySlice, dftd4
from .tables import NFREQ

__all__ = ["Disp", "DispD4", "DispD4Exact", "DispTerm", "TwoBodyTerm", "D4ATMApprox", "D4ATMExact", "FusedD4Term",
           "ZeroDamping"]  # fmt: skip


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

    # which part of the fused kernel this term stands for: parameters that are switched off
    _OFF: dict[str, float] = {}

    def _validate(self, param) -> None:
        """What the term cannot do, and the parameter check of its damping function
        (damping/functions.py:95-113), before anything is computed."""
        if type(self.damping_fn).__name__ != self._DAMPING:
            raise NotImplementedError(f"only {self._DAMPING} is accelerated for {type(self).__name__}")
        if self.charge_dependent not in self._CHARGE_DEPENDENT:
            raise NotImplementedError(
                f"{type(self).__name__}(charge_dependent={self.charge_dependent}) is outside the accelerated path"
            )
        if self._DAMPING == "RationalDamping":
            missing = [k for k in ("a1", "a2") if param.get(k) is None]
            if missing:
                raise TypeError(f"RationalDamping (order 6) requires keyword(s): {', '.join(missing)}")

    def calculate(self, numbers, positions, param, cn=None, model="d4", q=None, r4r2=None, rvdw=None,
                  cutoff=None):  # fmt: skip
        """Atom-resolved energy of this term: the plugin interface of the reference
        (``DispTerm.calculate``, ``dispersion/base.py:88-101``), so that an instance can be
        registered with the reference's own ``Disp`` (``base.py:215-219``) as well as with
        :class:`Disp` below.

        ``cn`` is accepted and ignored: the kernels evaluate the D4 coordination number of the
        structure themselves, fused with everything that consumes it (it is the same
        ``cn_d4`` / ``erf_count`` the reference computes at ``base.py:390``;
        :func:`tad_dftd4_b200.ncoord.cn_d4` returns it stand-alone).  ``r4r2`` / ``rvdw`` must
        be the defaults (the reference always passes the gathered default tables here, so they
        are not checked); ``model`` may be a model name, this package's ``D4Model`` /
        ``D4SModel`` or the reference's (read through ``ga``, ``gc``, ``wf``)."""
        self._validate(param)
        par = dict(param)
        for key, val in self._OFF.items():
            par[key] = val
        if self._DAMPING == "ZeroDamping":  # BJ radii of the ATM terms fall back to the defaults (threebody.py:244-256)
            par.setdefault("a1", defaults.A1)
            par.setdefault("a2", defaults.A2)
        if "s10" in self._OFF:
            par.pop("s10")
        if q is None:
            if self.charge_dependent:
                raise ValueError("a charge-dependent term needs the atomic charges q")
            q = torch.zeros(numbers.shape, dtype=positions.dtype, device=positions.device)
        elif not self.charge_dependent:
            q = torch.zeros_like(q)  # the reference evaluates such a term with zeta(q = 0)
        return dftd4(numbers, positions, 0.0, par, model=model, q=q, cutoff=cutoff)


class TwoBodyTerm(DispTerm):
    """Two-body term with rational damping (``dispersion/twobody.py:89-201``): the fused kernel
    with the ATM part switched off."""

    _DAMPING, _CHARGE_DEPENDENT, _OFF = "RationalDamping", (True, False), {"s9": 0.0}

    def __init__(self, *, damping_fn: Any = None, charge_dependent: bool = True):
        super().__init__(damping_fn if damping_fn is not None else RationalDamping(), charge_dependent)


class D4ATMApprox(DispTerm):
    """ATM term with the approximate C9 (``dispersion/d4.py:29-45``, ``threebody.py:210-256``):
    the fused kernel with the two-body part switched off."""

    _DAMPING, _CHARGE_DEPENDENT, _OFF = "ZeroDamping", (False,), {"s6": 0.0, "s8": 0.0, "s10": 0.0}

    def __init__(self, *, damping_fn: Any = None, charge_dependent: bool = False):
        super().__init__(damping_fn if damping_fn is not None else ZeroDamping(), charge_dependent)


class D4ATMExact(DispTerm):
    """ATM term with the exact C9 from the Casimir-Polder integral of the three weighted
    polarizabilities (``dispersion/d4.py:67-68``, ``threebody.py:276-302``, ``utils.py:155-212``):

        C9_ijk = (3/pi) sum_w t_w a_i(w) a_j(w) a_k(w).

    Every node ``w`` has the product form the kernels evaluate for the approximate C9, so the term is the
    sum of 23 ATM-only launches with single-node polarizability tables (see ``tables.ElementTables``);
    energies and gradients add up.  Costs 23 times the ATM part of the default term."""

    _DAMPING, _CHARGE_DEPENDENT, _OFF = "ZeroDamping", (False,), {"s6": 0.0, "s8": 0.0, "s10": 0.0}

    def __init__(self, *, damping_fn: Any = None, charge_dependent: bool = False):
        super().__init__(damping_fn if damping_fn is not None else ZeroDamping(), charge_dependent)

    def calculate(self, numbers, positions, param, cn=None, model="d4", q=None, r4r2=None, rvdw=None,
                  cutoff=None):  # fmt: skip
        total = None
        for w in range(NFREQ):
            e = super().calculate(numbers, positions, param, cn, _FrequencySlice(model, w), q, r4r2, rvdw, cutoff)
            total = e if total is None else total + e
        return total


class FusedD4Term(DispTerm):
    """Two-body + ATM in ONE launch (what ``DispD4`` registers as two terms).  Registered with
    the reference's ``Disp`` -- ``Disp(model="d4").register(FusedD4Term())`` -- it makes the
    reference's own class-based driver run on the kernels."""

    _DAMPING, _CHARGE_DEPENDENT, _OFF = "RationalDamping", (True,), {}

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
        if not self.terms:
            return torch.zeros(numbers.shape, dtype=positions.dtype, device=positions.device)
        kinds = sorted(type(t).__name__ for t in self.terms)
        fused = kinds == ["FusedD4Term"] or (
            kinds == ["D4ATMApprox", "TwoBodyTerm"]
            and all(t.charge_dependent == isinstance(t, TwoBodyTerm) for t in self.terms)
            and all(type(t.damping_fn).__name__ == t._DAMPING for t in self.terms)
        )
        if fused:  # the default combination (dispersion/d4.py:48-61): ONE launch
            return dftd4(numbers, positions, charge, param, model=self.model, rcov=rcov, r4r2=r4r2, rvdw=rvdw,
                         q=q, cutoff=cutoff, cn_function=self.cn_fn)  # fmt: skip
        # any other registered combination: the sum of the terms, as in the reference (base.py:409-431); every
        # term is the fused kernel with the other part switched off (D4ATMExact: 23 such launches)
        _check_arguments(numbers, positions, self.model, rcov, r4r2, rvdw, q, self.cn_fn)
        for t in self.terms:
            t._validate(param)
        if q is None and is_c_dep:  # one EEQ solve for all terms (base.py:401-407)
            from .disp import _eeq_charges

            q = _eeq_charges(numbers, positions, charge, cutoff)
        energy = None
        for t in self.terms:
            e = t.calculate(numbers, positions, param, None, self.model, q, r4r2, rvdw, cutoff)
            energy = e if energy is None else energy + e
        return energy


class DispD4(Disp):
    """Standard DFT-D4 (``dispersion/d4.py:48-61``)."""

    TERMS = [
        (TwoBodyTerm, {"damping_fn": RationalDamping(), "charge_dependent": True}),
        (D4ATMApprox, {"damping_fn": ZeroDamping(), "charge_dependent": False}),
    ]


class DispD4Exact(Disp):
    """DFT-D4 with the exact C9 coefficients via the Casimir-Polder formula (``dispersion/d4.py:71-84``)."""

    TERMS = [
        (TwoBodyTerm, {"damping_fn": RationalDamping(), "charge_dependent": True}),
        (D4ATMExact, {"damping_fn": ZeroDamping(), "charge_dependent": False}),
    ]
