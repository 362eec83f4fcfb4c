"""
Host-side interface contract, CPU only (no kernels involved).  Each block checks the behaviour
that one file of the reference's own test-suite pins, so that code written against
``tad_dftd4`` meets the same exceptions, defaults and comparison semantics here:

* ``get_params``            -- test/test_param/test_read.py, test/test_param/test_fail.py
* ``Cutoff``                -- test/test_cutoff/test_general.py, test/test_cutoff/test_types.py
* ``dftd4`` argument checks -- test/test_d4/test_general.py (the checks that precede device work)
* ``Disp`` / terms          -- test/test_disp/test_general.py, test/test_disp/test_term.py
* damping markers           -- test/test_disp/test_damping.py
* model descriptors         -- test/test_model/test_general.py
"""
from __future__ import annotations

import itertools
from unittest.mock import patch

import pytest
import torch

import tad_dftd4_b200 as d4
from tad_dftd4_b200 import damping, defaults
from tad_dftd4_b200.cutoff import Cutoff
from tad_dftd4_b200.damping import (MZeroDamping, OptimisedPowerDamping, RationalDamping, ZeroDamping,
                                    get_params)  # fmt: skip
from tad_dftd4_b200.dispersion import Disp, DispD4, DispTerm, TwoBodyTerm

FLOATS = (torch.float16, torch.float32, torch.float64)
CUTOFF_FIELDS = ("disp2", "disp3", "cn", "cn_eeq")


# ------------------------------------------------------------------------------ get_params
@pytest.mark.parametrize("functional,needs", [(None, ("s6",)), ("pbe", ("a1", "a2")), ("b3lyp", ("a1", "a2")),
                                              ("revpbe", ("a1", "a2"))])  # fmt: skip
def test_get_params_blocks(functional, needs):
    block = get_params(method="d4", variant="bj-eeq-atm", functional=functional)
    assert isinstance(block, dict) and all(k in block for k in needs)


def test_get_params_doi_and_default_variant():
    assert "doi" in get_params(method="d4", variant="bj-eeq-atm", functional="pbe", keep_doi=True)
    assert "doi" not in get_params(method="d4", variant="bj-eeq-atm", functional="pbe")
    assert "a1" in get_params(method="d4", functional="pbe", variant=None)  # first default variant


@pytest.mark.parametrize("kwargs,exc,match", [
    (dict(method="d4", variant="d4-eeq-bj", functional="unknown"), KeyError, None),
    (dict(method="d4", functional="pbe", variant="unknown"), KeyError, "not found for functional"),
    (dict(method="d4", functional=None, variant="no-such-variant"), KeyError, "not found in default parameters"),
    (dict(method="d5", functional="pbe", variant="x"), FileNotFoundError, "missing"),  # no table for d5
    (dict(method="dx", functional="pbe"), ValueError, "not a valid DispersionMethod"),
])  # fmt: skip
def test_get_params_failures(kwargs, exc, match):
    with pytest.raises(exc, match=match):
        get_params(**kwargs)


def test_get_params_method_absent_for_functional():
    table = {"default": {"d4": ["bj-eeq-atm"]}, "parameter": {"pbe": {"reference": {}}}}
    with patch.object(damping, "_load", return_value=table), pytest.raises(KeyError, match="Method"):
        get_params(method="d4", functional="pbe", variant="bj-eeq-atm")


# ------------------------------------------------------------------------------ Cutoff
@pytest.mark.parametrize("dtype", FLOATS)
def test_cutoff_type_conversion(dtype):
    cut = Cutoff().type(dtype)
    assert cut.dtype == dtype and all(getattr(cut, f).dtype == dtype for f in CUTOFF_FIELDS)


def test_cutoff_is_read_only_and_float_only():
    cut = Cutoff()
    for attr, value in (("dtype", torch.float64), ("device", torch.device("cpu"))):
        with pytest.raises(AttributeError):
            setattr(cut, attr, value)
    with pytest.raises(ValueError):
        cut.type(torch.bool)
    moved = cut.to(torch.device("cpu"))
    assert moved.device == torch.device("cpu")
    assert all(getattr(moved, f).device == torch.device("cpu") for f in CUTOFF_FIELDS)


def test_cutoff_values():
    expect = (defaults.D4_DISP2_CUTOFF, defaults.D4_DISP3_CUTOFF, defaults.D4_CN_CUTOFF, defaults.D4_CN_EEQ_CUTOFF)
    cut = Cutoff()
    assert [float(getattr(cut, f)) for f in CUTOFF_FIELDS] == pytest.approx(list(expect))
    one = torch.tensor([1.0])
    mixed = Cutoff(disp2=one)
    assert all(isinstance(getattr(mixed, f), torch.Tensor) for f in CUTOFF_FIELDS)
    assert mixed.disp2 == pytest.approx(one)
    for vals in ((1, 2, -3, 4), (1.0, 2.0, 3.0, -4.0)):  # ints and floats, signs are not checked
        cut = Cutoff(*vals)
        assert [float(getattr(cut, f)) for f in CUTOFF_FIELDS] == pytest.approx(list(vals))
        assert all(isinstance(getattr(cut, f), torch.Tensor) for f in CUTOFF_FIELDS)


# ------------------------------------------------------------------------------ dftd4 / Disp argument checks
def _h2():
    return torch.tensor([1, 1]), torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 1.0]]), torch.tensor(0.0)


BAD_SHAPES = [dict(rcov=torch.tensor([1.0])), dict(r4r2=torch.tensor([1.0])), dict(rvdw=torch.tensor([1.0])),
              dict(q=torch.tensor([1.0]))]  # fmt: skip


@pytest.mark.parametrize("entry", ["function", "class"])
def test_shape_checks_come_first(entry):
    numbers, positions, charge = _h2()
    param = d4.Param(s6=torch.tensor(1.0))  # a1/a2 missing on purpose: shapes are checked before
    if entry == "function":
        call = lambda n, **kw: d4.dftd4(n, positions, charge, param, **kw)  # noqa: E731
    else:
        disp = Disp()
        disp.register(TwoBodyTerm(damping_fn=RationalDamping(), charge_dependent=True))
        call = lambda n, **kw: disp.calculate(n, positions, charge, param, **kw)  # noqa: E731
    for bad in BAD_SHAPES:
        with pytest.raises(ValueError):
            call(numbers, **bad)
    with pytest.raises(ValueError):  # numbers inconsistent with positions (charges given)
        call(torch.tensor([1]), q=torch.tensor([0.5, -0.5]))


def test_dftd4_model_parameters_device():
    numbers, positions, charge = _h2()
    good = {"a1": 0.4, "a2": 5.0}
    with pytest.raises(ValueError, match="Unknown model"):
        d4.dftd4(numbers, positions, charge, good, model="d6")
    with pytest.raises(TypeError):  # rational damping without a1/a2, on any device
        d4.dftd4(numbers, positions, charge, d4.Param(s6=torch.tensor(1.0)), q=torch.zeros(2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d4.dftd4(numbers, positions, charge, good, q=torch.zeros(2))


def test_disp_terms_and_models():
    numbers, positions, charge = _h2()
    param = d4.Param(s6=torch.tensor(1.0))
    disp = Disp()
    assert disp.terms == []
    disp.register(TwoBodyTerm())
    assert len(disp.terms) == 1
    disp.deregister(TwoBodyTerm())  # equal term, not the same object
    assert disp.terms == []

    no_charges = Disp()
    no_charges.register(TwoBodyTerm(damping_fn=RationalDamping(), charge_dependent=False))
    with pytest.raises(RuntimeError):  # q passed although no term uses charges
        no_charges.calculate(numbers=numbers, positions=positions, charge=charge, param=param,
                             q=torch.tensor([0.5, -0.5]))  # fmt: skip
    for fn, exc in ((RationalDamping(), TypeError), (OptimisedPowerDamping(), (TypeError, NotImplementedError))):
        incomplete = Disp()
        incomplete.register(TwoBodyTerm(damping_fn=fn))
        with pytest.raises(exc):
            incomplete.calculate(numbers=numbers, positions=positions, charge=charge, param=param)
    with pytest.raises(ValueError):
        DispD4(model="wrong")


def test_dispterm_comparison():
    class Dummy(DispTerm):
        def calculate(self, *args, **kwargs):
            return torch.tensor(0.0)

    base = Dummy(damping_fn=RationalDamping(), charge_dependent=True)
    assert base == Dummy(damping_fn=RationalDamping(), charge_dependent=True)
    assert base != Dummy(damping_fn=OptimisedPowerDamping(), charge_dependent=True)
    assert base != Dummy(damping_fn=RationalDamping(), charge_dependent=False)
    assert base != "not a disp term"


# ------------------------------------------------------------------------------ damping markers
def test_damping_markers_compare_by_class():
    kinds = (RationalDamping, ZeroDamping, MZeroDamping, OptimisedPowerDamping)
    for (i, a), (j, b) in itertools.product(enumerate(k() for k in kinds), enumerate(k() for k in kinds)):
        assert (a == b) is (i == j) and (a != b) is (i != j)
    one = RationalDamping()
    assert one == one
    for other in ("a string", 123, None, [RationalDamping()]):
        assert one != other

    class Shy:
        def __eq__(self, other):
            return NotImplemented

    assert (one == Shy()) is False


# ------------------------------------------------------------------------------ model descriptors
SIH4 = torch.tensor([14, 1, 1, 1, 1])


@pytest.mark.parametrize("cls", [d4.D4Model, d4.D4SModel])
def test_model_descriptor_contract(cls):
    model = cls(SIH4)
    for dtype in FLOATS:
        assert model.type(dtype).dtype == dtype
    with pytest.raises(ValueError):
        model.type(torch.bool)
    assert model.to(torch.device("cpu")).device == torch.device("cpu")
    for attr, value in (("dtype", torch.float64), ("device", torch.device("cpu"))):
        with pytest.raises(AttributeError):
            setattr(model, attr, value)
    with pytest.raises(ValueError):
        cls(SIH4, ref_charges="wrong")
    model.ref_charges = "wrong"  # changed after construction: caught when the weights are asked for
    with pytest.raises(ValueError):
        model.weight_references()
    assert d4.D4Model(SIH4, wf=6).wf == 6
