"""
Host-side interface tests written after the reference's own test-suite (same names, same
assertions), run against ``tad_dftd4_b200`` on the CPU -- no kernels are involved:

* ``test/test_param/test_read.py:29-51`` and ``test/test_param/test_fail.py:31-85``
  (``get_params``: defaults, functionals, DOI, unknown functional / variant / method file),
* ``test/test_cutoff/test_general.py:29-72`` and ``test/test_cutoff/test_types.py:30-65``
  (``Cutoff``: dtype / device handling, defaults, tensors and numbers),
* ``test/test_disp/test_general.py`` / ``test/test_d4/test_general.py`` style argument
  checks of ``dftd4`` that fail before any device work (shape mismatches, unknown model,
  missing damping parameters).
"""
from __future__ import annotations

from unittest.mock import patch

import pytest
import torch

import tad_dftd4_b200 as d4
from tad_dftd4_b200 import damping, defaults
from tad_dftd4_b200.cutoff import Cutoff
from tad_dftd4_b200.damping import get_params


# ---- test/test_param/test_read.py ------------------------------------------------------------
def test_default() -> None:
    params = get_params(method="d4", variant="bj-eeq-atm", functional=None)
    assert isinstance(params, dict)
    assert "s6" in params


@pytest.mark.parametrize("func", ["pbe", "b3lyp", "revpbe"])
def test_func(func: str) -> None:
    params = get_params(method="d4", variant="bj-eeq-atm", functional=func)
    assert isinstance(params, dict)
    assert "a1" in params
    assert "a2" in params


def test_with_doi() -> None:
    params = get_params(method="d4", variant="bj-eeq-atm", functional="pbe", keep_doi=True)
    assert isinstance(params, dict)
    assert "doi" in params
    assert "doi" not in get_params(method="d4", variant="bj-eeq-atm", functional="pbe")


# ---- test/test_param/test_fail.py ------------------------------------------------------------
def test_unknown_func() -> None:
    with pytest.raises(KeyError):
        get_params(method="d4", variant="d4-eeq-bj", functional="unknown")


def test_unknown_variant() -> None:
    with pytest.raises(KeyError):
        get_params(method="d4", functional="pbe", variant="unknown")


def test_unknown_variant_default() -> None:
    with pytest.raises(KeyError, match="not found in default parameters"):
        get_params(method="d4", functional=None, variant="no-such-variant")


def test_unknown_variant_functional() -> None:
    with pytest.raises(KeyError, match="not found for functional"):
        get_params(method="d4", functional="pbe", variant="no-such-variant")


def test_missing_toml() -> None:
    """Parameter file does not exist (d5 has none; this package ships the D4 table only)."""
    with pytest.raises(FileNotFoundError, match="missing"):
        get_params(method="d5", functional="pbe", variant="x")


def test_invalid_method() -> None:
    with pytest.raises(ValueError, match="not a valid DispersionMethod"):
        get_params(method="dx", functional="pbe")


def test_default_variant_for_functional() -> None:
    params = get_params(method="d4", functional="pbe", variant=None)
    assert isinstance(params, dict)
    assert "a1" in params


def test_method_missing_in_functional() -> None:
    fake_table = {"default": {"d4": ["bj-eeq-atm"]}, "parameter": {"pbe": {"reference": {}}}}
    with patch.object(damping, "_load", return_value=fake_table):
        with pytest.raises(KeyError, match="Method"):
            get_params(method="d4", functional="pbe", variant="bj-eeq-atm")


# ---- test/test_cutoff/test_general.py --------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.float64])
def test_change_type(dtype: torch.dtype) -> None:
    cutoff = Cutoff().type(dtype)
    assert cutoff.dtype == dtype
    assert cutoff.disp2.dtype == dtype
    assert cutoff.disp3.dtype == dtype
    assert cutoff.cn.dtype == dtype
    assert cutoff.cn_eeq.dtype == dtype


def test_change_type_fail() -> None:
    cutoff = Cutoff()
    with pytest.raises(AttributeError):
        cutoff.dtype = torch.float64
    with pytest.raises(ValueError):
        cutoff.type(torch.bool)


def test_change_device_cpu() -> None:
    device = torch.device("cpu")
    cutoff = Cutoff().to(device)
    assert cutoff.device == device
    assert cutoff.disp2.device == device
    assert cutoff.disp3.device == device
    assert cutoff.cn.device == device
    assert cutoff.cn_eeq.device == device


def test_change_device_fail() -> None:
    cutoff = Cutoff()
    with pytest.raises(AttributeError):
        cutoff.device = torch.device("cpu")


# ---- test/test_cutoff/test_types.py ----------------------------------------------------------
def test_defaults() -> None:
    cutoff = Cutoff()
    assert pytest.approx(defaults.D4_DISP2_CUTOFF) == cutoff.disp2.cpu()
    assert pytest.approx(defaults.D4_DISP3_CUTOFF) == cutoff.disp3.cpu()
    assert pytest.approx(defaults.D4_CN_CUTOFF) == cutoff.cn.cpu()
    assert pytest.approx(defaults.D4_CN_EEQ_CUTOFF) == cutoff.cn_eeq.cpu()


def test_tensor() -> None:
    tmp = torch.tensor([1.0])
    cutoff = Cutoff(disp2=tmp)
    assert isinstance(cutoff.disp2, torch.Tensor)
    assert isinstance(cutoff.disp3, torch.Tensor)
    assert isinstance(cutoff.cn, torch.Tensor)
    assert isinstance(cutoff.cn_eeq, torch.Tensor)
    assert pytest.approx(tmp.cpu()) == cutoff.disp2.cpu()


@pytest.mark.parametrize("vals", [(1, 2, -3, 4), (1.0, 2.0, 3.0, -4.0)])
def test_int_float(vals) -> None:
    disp2, disp3, cn, cn_eeq = vals
    cutoff = Cutoff(disp2, disp3, cn, cn_eeq)
    for name in ("disp2", "disp3", "cn", "cn_eeq"):
        assert isinstance(getattr(cutoff, name), torch.Tensor)
    assert pytest.approx(vals[0]) == cutoff.disp2.cpu()
    assert pytest.approx(vals[1]) == cutoff.disp3.cpu()
    assert pytest.approx(vals[2]) == cutoff.cn.cpu()
    assert pytest.approx(vals[3]) == cutoff.cn_eeq.cpu()


# ---- argument checks of dftd4 that precede any device work -------------------------------------
PARAM = {"a1": 0.4, "a2": 5.0}


def test_fail_shape_positions() -> None:
    numbers = torch.tensor([1, 1])
    positions = torch.zeros(3, 3)
    with pytest.raises(ValueError, match="positions"):
        d4.dftd4(numbers, positions, 0.0, PARAM)


def test_fail_shape_q_and_radii() -> None:
    numbers = torch.tensor([1, 1])
    positions = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 1.4]])
    with pytest.raises(ValueError, match="charges"):
        d4.dftd4(numbers, positions, 0.0, PARAM, q=torch.zeros(3))
    with pytest.raises(ValueError, match="covalent radii"):
        d4.dftd4(numbers, positions, 0.0, PARAM, rcov=torch.zeros(3))
    with pytest.raises(ValueError, match="r4r2"):
        d4.dftd4(numbers, positions, 0.0, PARAM, r4r2=torch.zeros(3))


def test_fail_unknown_model() -> None:
    numbers = torch.tensor([1, 1])
    positions = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 1.4]])
    with pytest.raises(ValueError, match="Unknown model"):
        d4.dftd4(numbers, positions, 0.0, PARAM, model="d6")


def test_no_cpu_fallback() -> None:
    numbers = torch.tensor([1, 1])
    positions = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 1.4]])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d4.dftd4(numbers, positions, 0.0, PARAM, q=torch.zeros(2))


# ---- test/test_disp/test_general.py:31-145 (class interface) ------------------------------------
def _h2():
    return (torch.tensor([1, 1]), torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, 1.0]]), torch.tensor(0.0),
            d4.Param(s6=torch.tensor(1.0)))  # fmt: skip


def test_fail_class() -> None:
    from tad_dftd4_b200.damping import RationalDamping
    from tad_dftd4_b200.dispersion import Disp, TwoBodyTerm

    numbers, positions, charge, param = _h2()
    disp = Disp()
    disp.register(TwoBodyTerm(damping_fn=RationalDamping(), charge_dependent=True))
    with pytest.raises(ValueError):  # rcov wrong shape
        disp.calculate(numbers, positions, charge, param, rcov=torch.tensor([1.0]))
    with pytest.raises(ValueError):  # r4r2 wrong shape
        disp.calculate(numbers, positions, charge, param, r4r2=torch.tensor([1.0]))
    with pytest.raises(ValueError):  # rvdw wrong shape
        disp.calculate(numbers, positions, charge, param, rvdw=torch.tensor([1.0]))
    with pytest.raises(ValueError):  # atomic partial charges wrong shape
        disp.calculate(numbers, positions, charge, param, q=torch.tensor([1.0]))
    with pytest.raises(ValueError):  # wrong numbers
        disp.calculate(torch.tensor([1]), positions, charge, param, q=torch.tensor([0.5, -0.5]))


def test_fail_charges_not_required() -> None:
    from tad_dftd4_b200.damping import RationalDamping
    from tad_dftd4_b200.dispersion import Disp, TwoBodyTerm

    numbers, positions, charge, param = _h2()
    disp = Disp()
    disp.register(TwoBodyTerm(damping_fn=RationalDamping(), charge_dependent=False))
    with pytest.raises(RuntimeError):
        disp.calculate(numbers=numbers, positions=positions, charge=charge, param=param,
                       q=torch.tensor([0.5, -0.5]))  # fmt: skip


def test_fail_damping_param() -> None:
    from tad_dftd4_b200.damping import RationalDamping
    from tad_dftd4_b200.dispersion import Disp, TwoBodyTerm

    numbers, positions, charge, param = _h2()
    disp = Disp()
    disp.register(TwoBodyTerm(damping_fn=RationalDamping()))
    with pytest.raises(TypeError):  # a1 / a2 missing
        disp.calculate(numbers=numbers, positions=positions, charge=charge, param=param)


def test_fail_model() -> None:
    from tad_dftd4_b200.dispersion import DispD4

    with pytest.raises(ValueError):
        DispD4(model="wrong")


def test_terms() -> None:
    from tad_dftd4_b200.dispersion import Disp, TwoBodyTerm

    disp = Disp()
    assert len(disp.terms) == 0
    disp.register(TwoBodyTerm())
    assert len(disp.terms) == 1
    disp.deregister(TwoBodyTerm())
    assert len(disp.terms) == 0


# ---- test/test_model/test_general.py:30-103 (model descriptors) ---------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.float64])
def test_model_change_type(dtype: torch.dtype) -> None:
    numbers = torch.tensor([14, 1, 1, 1, 1])
    model = d4.D4Model(numbers).type(dtype)
    assert model.dtype == dtype


def test_model_change_type_fail() -> None:
    model = d4.D4Model(torch.tensor([14, 1, 1, 1, 1]))
    with pytest.raises(AttributeError):
        model.dtype = torch.float64
    with pytest.raises(ValueError):
        model.type(torch.bool)


def test_model_change_device() -> None:
    device = torch.device("cpu")
    model = d4.D4Model(torch.tensor([14, 1, 1, 1, 1])).to(device)
    assert model.device == device
    with pytest.raises(AttributeError):
        model.device = torch.device("cpu")


@pytest.mark.parametrize("model", ["d4", "d4s"])
def test_ref_charges_fail(model: str) -> None:
    numbers = torch.tensor([14, 1, 1, 1, 1])
    cls = d4.D4Model if model == "d4" else d4.D4SModel
    with pytest.raises(ValueError):
        cls(numbers, ref_charges="wrong")


@pytest.mark.parametrize("model", ["d4", "d4s"])
def test_ref_charges_fail_2(model: str) -> None:
    numbers = torch.tensor([14, 1, 1, 1, 1])
    cls = d4.D4Model if model == "d4" else d4.D4SModel
    m = cls(numbers, ref_charges="eeq")
    m.ref_charges = "wrong"
    with pytest.raises(ValueError):
        m.weight_references()


def test_model_args() -> None:
    numbers = torch.tensor([14, 1, 1, 1, 1])
    model = d4.D4Model(numbers, wf=6)
    assert model.wf == 6


# ---- test/test_disp/test_damping.py:29-98 (damping-function markers) ----------------------------
def test_damping_equality() -> None:
    from tad_dftd4_b200.damping import MZeroDamping, OptimisedPowerDamping, RationalDamping, ZeroDamping

    damp = RationalDamping()
    assert damp == damp
    assert RationalDamping() == RationalDamping()
    assert not (RationalDamping() != RationalDamping())
    assert RationalDamping() != ZeroDamping()
    assert not (RationalDamping() == ZeroDamping())
    instances = [cls() for cls in (RationalDamping, ZeroDamping, MZeroDamping, OptimisedPowerDamping)]
    for i, a in enumerate(instances):
        for j, b in enumerate(instances):
            assert (a == b) is (i == j)
            assert (a != b) is (i != j)
    assert damp != "a string"
    assert damp != 123
    assert damp != None  # noqa: E711
    assert damp != [RationalDamping()]

    class Unrelated:
        def __eq__(self, other):
            return NotImplemented

    assert (damp == Unrelated()) is False


def test_fail_damping_param_other_schemes() -> None:
    """test_disp/test_general.py:95-112: missing parameters raise TypeError for every damping
    function of the reference; the schemes the kernels do not evaluate raise NotImplementedError
    once the parameters are complete."""
    from tad_dftd4_b200.damping import OptimisedPowerDamping
    from tad_dftd4_b200.dispersion import Disp, TwoBodyTerm

    numbers, positions, charge, param = _h2()
    disp = Disp()
    disp.register(TwoBodyTerm(damping_fn=OptimisedPowerDamping()))
    with pytest.raises((TypeError, NotImplementedError)):
        disp.calculate(numbers=numbers, positions=positions, charge=charge, param=param)


# ---- test/test_disp/test_term.py:31-80 (DispTerm equality) --------------------------------------
def test_dispterm_equality() -> None:
    from tad_dftd4_b200.damping import OptimisedPowerDamping, RationalDamping
    from tad_dftd4_b200.dispersion import DispTerm

    class DummyDispTerm(DispTerm):
        def calculate(self, numbers, positions, param, cn, model, q, r4r2, rvdw, cutoff):
            return torch.tensor(0.0)

    t1 = DummyDispTerm(damping_fn=RationalDamping(), charge_dependent=True)
    t2 = DummyDispTerm(damping_fn=RationalDamping(), charge_dependent=True)
    assert t1 == t2
    assert t1 != DummyDispTerm(damping_fn=OptimisedPowerDamping(), charge_dependent=True)
    assert t1 != DummyDispTerm(damping_fn=RationalDamping(), charge_dependent=False)
    assert t1 != "not a disp term"
