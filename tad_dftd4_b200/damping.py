"""
Damping parameters: ``Param`` and ``get_params`` with the reference's
semantics (``/root/reference/src/tad_dftd4/damping/parameters/base.py:48-85``,
``loader.py:61-133``).  Only the rational (Becke-Johnson) damping of the
default D4 method is accelerated; the parameter table is the data of
``damping/parameters/d4.toml`` (exported by ``tools/export_reference_tables.py``).
"""

from __future__ import annotations

import json
from functools import lru_cache
from pathlib import Path
from typing import Any, Dict

__all__ = ["Param", "get_params", "Damping", "RationalDamping", "ZeroDamping", "MZeroDamping",
           "OptimisedPowerDamping"]

# ``Param`` is a plain dict with (all optional) keys
#   a1, a2, s6, s8, s9, s10, rs6, rs8, rs9, alp, bet, doi
Param = dict


class Damping:
    """Base of the damping-function markers.  Instances carry no state; two are equal iff they
    are of the same class, like the reference's ``Damping.__eq__``
    (``damping/functions.py:233-253``).  Only :class:`RationalDamping` (two-body) and
    :class:`ZeroDamping` (ATM) are evaluated by the kernels; the others exist so that code
    written against the reference imports and compares them, and raise
    ``NotImplementedError`` when a calculation is asked to use them."""

    radius_type = "r4r2"

    def __eq__(self, other: Any) -> bool:
        if not isinstance(other, Damping) and not hasattr(other, "radius_type"):
            return NotImplemented
        return type(other).__name__ == type(self).__name__

    def __ne__(self, other: Any) -> bool:
        res = self.__eq__(other)
        return True if res is NotImplemented else not res

    def __hash__(self) -> int:
        return hash(type(self).__name__)


class RationalDamping(Damping):
    """Becke-Johnson rational damping (``damping/functions.py:262-305``): fused into the kernels."""


class ZeroDamping(Damping):
    """Zero damping of the ATM term (``damping/functions.py:308-378``): fused into the kernels."""

    radius_type = "rvdw"


class MZeroDamping(Damping):
    """Modified zero damping (``damping/functions.py:381-431``): not accelerated."""

    radius_type = "rvdw"


class OptimisedPowerDamping(Damping):
    """Optimised-power damping (``damping/functions.py:434-484``): not accelerated."""


@lru_cache(maxsize=None)
def _load(method: str) -> Dict[str, Any]:
    path = Path(__file__).resolve().parent / "data" / f"{method}_damping.json"
    if not path.is_file():
        raise FileNotFoundError(f"Parameter file {path} missing.")
    with open(path, encoding="utf8") as fp:
        return json.load(fp)


def get_params(*, method: str, functional: str | None, variant: str | None = None,
               keep_doi: bool = False) -> Param:  # fmt: skip
    """Damping parameters of a functional, e.g.
    ``get_params(method="d4", functional="pbe0")``.  Like the reference this
    returns only the functional's own block (no merge of s6/s9/alp defaults)."""
    method = getattr(method, "value", method)
    if method not in ("d3", "d4", "d5"):  # the reference's DispersionMethod enum (parameters/base.py:88-93)
        raise ValueError(f"'{method}' is not a valid DispersionMethod")
    table = _load(method)  # only the D4 table ships with this package: d3/d5 -> FileNotFoundError
    if functional in (None, "default"):
        default_section = table["default"]
        if variant not in default_section[method]:
            raise KeyError(
                f"Variant '{variant}' not found in default parameters for method={method!r}."
            )
        section = default_section["parameter"]
    else:
        funcs = table["parameter"]
        if functional not in funcs:
            raise KeyError(f"Functional '{functional!r}' not found in damping parameters.")
        section = funcs[functional.casefold()]
        if method not in section:
            raise KeyError(f"Method '{method}' not found in damping parameters for '{functional!r}'.")
    if variant is None:
        variant = table["default"][method][0]
    variants = section[method]
    if variant not in variants:
        raise KeyError(
            f"Variant '{variant}' not found for functional={functional!r}, method={method!r}."
        )
    out = dict(variants[variant])
    if not keep_doi:
        out.pop("doi", None)
    return Param(**out)
