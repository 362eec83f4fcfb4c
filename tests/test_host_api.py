"""Host-side contract of the drop-in (no GPU needed): the C-ABI library loads and exports
every symbol include/d4b200.h declares; argument validation raises the reference's
exception types (reference tests test/test_d4/test_general.py:55-117,
test/test_disp/test_general.py:30-153, test/test_param/test_fail.py)."""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

import pytest
import torch

import tad_dftd4_b200 as d4
from tad_dftd4_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
PARAM = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "d4b200.h").read_text()
    declared = set(re.findall(r"\b(d4b200_[a-z0-9_]+)\s*\(", header))
    declared -= {"d4b200_params", "d4b200_tables"}
    assert len(declared) >= 20
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert lib.d4b200_version() >= 100
    assert set(_lib.EXPORTED_SYMBOLS) <= declared


def test_params_struct_matches_header():
    # 11 doubles + 2 int32 (struct d4b200_params)
    assert ctypes.sizeof(_lib.Params) == 11 * 8 + 2 * 4


def test_error_string_and_invalid_arguments():
    lib = _lib.load()
    assert lib.d4b200_error_string(0) == b"ok"
    assert b"workspace" in lib.d4b200_error_string(-2)
    par = _lib.Params()
    # null handle / pointers are rejected without touching a device
    assert lib.d4b200_energy_f64(None, ctypes.byref(par), 1, 1, None, None, None, None, None, None, 0, None) == -1
    assert lib.d4b200_tables_create(0, None, 0, None, 0, 3.0, 2.0, None) == -1


def test_shape_mismatch_raises_value_error():
    numbers = torch.tensor([6, 1, 1])
    positions = torch.zeros(2, 3)
    with pytest.raises(ValueError, match="Shape of positions"):
        d4.dftd4(numbers, positions, 0.0, PARAM, q=torch.zeros(3))
    with pytest.raises(ValueError, match="atomic charges"):
        d4.dftd4(numbers, torch.zeros(3, 3), 0.0, PARAM, q=torch.zeros(2))
    for kw in ("rcov", "r4r2"):
        with pytest.raises(ValueError):
            d4.dftd4(numbers, torch.zeros(3, 3), 0.0, PARAM, q=torch.zeros(3), **{kw: torch.zeros(2)})
    with pytest.raises(ValueError, match="van der Waals"):
        d4.dftd4(numbers, torch.zeros(3, 3), 0.0, PARAM, q=torch.zeros(3), rvdw=torch.zeros(2, 2))


def test_unknown_model_and_unsupported_plugins():
    numbers, positions, q = torch.tensor([6, 1]), torch.zeros(2, 3), torch.zeros(2)
    with pytest.raises(ValueError, match="Unknown model"):
        d4.dftd4(numbers, positions, 0.0, PARAM, q=q, model="xyz")
    with pytest.raises(NotImplementedError):
        d4.dftd4(numbers, positions, 0.0, PARAM, q=q, model="d3")
    with pytest.raises(NotImplementedError):
        d4.dftd4(numbers, positions, 0.0, PARAM, q=q, cn_function=lambda *a, **k: None)
    with pytest.raises(NotImplementedError):
        d4.dftd4(numbers, positions, 0.0, PARAM, q=q, rcov=torch.ones(2))


def test_cpu_tensors_fail_loudly():
    """No CPU fallback: the product path refuses CPU tensors instead of computing."""
    numbers, positions, q = torch.tensor([6, 1]), torch.tensor([[0.0, 0, 0], [0, 0, 2.0]]), torch.zeros(2)
    with pytest.raises(RuntimeError, match="B200 GPUs only"):
        d4.dftd4(numbers, positions, 0.0, PARAM, q=q)
    with pytest.raises(RuntimeError, match="B200 GPUs only"):
        d4.get_properties(numbers, positions, q=q)


def test_missing_damping_parameters_raise_type_error():
    from tad_dftd4_b200.disp import _flatten_param

    with pytest.raises(TypeError, match="a1"):
        _flatten_param({"s8": 1.0, "a2": 5.0}, None, 0, 6.0)
    par = _flatten_param({"a1": 0.4, "a2": 5.0}, None, 0, 6.0)
    assert (par.s6, par.s8, par.s9, par.alp, par.has_s10) == (1.0, 1.0, 1.0, 16.0, 0)
    assert (par.disp2_cutoff, par.disp3_cutoff, par.cn_cutoff) == (60.0, 40.0, 30.0)
    par = _flatten_param({"a1": 0.4, "a2": 5.0, "s10": 0.0}, d4.Cutoff(disp2=50, cn=20.0), 0, 6.0)
    assert par.has_s10 == 1 and par.disp2_cutoff == 50.0
    assert par.cn_cutoff == 30.0  # Cutoff.cn is not forwarded (reference dispersion/base.py:390)
    # differentiated parameters (test/test_grad/test_param.py of the reference): value flattened,
    # the tensors are handed to the autograd function
    from tad_dftd4_b200.disp import _param_tensors

    a1 = torch.tensor(0.4, dtype=torch.float64, requires_grad=True)
    par = _flatten_param({"a1": a1, "a2": 5.0}, None, 0, 6.0)
    assert par.a1 == 0.4
    pt = _param_tensors({"a1": a1, "a2": torch.tensor(5.0), "s8": 1.0})
    assert len(pt) == 7 and pt[4] is a1 and all(t is None for k, t in enumerate(pt) if k != 4)
    assert _param_tensors({"a1": 0.4, "a2": torch.tensor(5.0)}) == ()


def test_get_params():
    p = d4.get_params(method="d4", functional="pbe0")
    assert p == {"s8": 1.20065498, "a1": 0.40085597, "a2": 5.02928789}  # d4.toml:269
    assert "doi" in d4.get_params(method="d4", functional="b3lyp", keep_doi=True)
    assert d4.get_params(method="d4", functional=None, variant="bj-eeq-atm")["s9"] == 1.0
    with pytest.raises(KeyError):
        d4.get_params(method="d4", functional="not-a-functional")
    with pytest.raises(KeyError):
        d4.get_params(method="d4", functional="pbe0", variant="nope")


def test_cutoff_defaults_and_conversions():
    c = d4.Cutoff()
    assert (float(c.disp2), float(c.disp3), float(c.cn), float(c.cn_eeq)) == (60.0, 40.0, 30.0, 25.0)
    assert d4.Cutoff(disp2=torch.tensor(33.0)).as_float("disp2") == 33.0
    assert c.type(torch.float64).disp2.dtype == torch.float64
    with pytest.raises(ValueError):
        c.type(torch.bool)


def test_class_interface_term_selection():
    disp = d4.dispersion.DispD4()
    assert [type(t).__name__ for t in disp.terms] == ["TwoBodyTerm", "D4ATMApprox"]
    assert d4.dispersion.TwoBodyTerm() == d4.dispersion.TwoBodyTerm()
    only_atm = d4.dispersion.Disp()
    only_atm.register(d4.dispersion.D4ATMApprox())
    with pytest.raises(RuntimeError, match="no term"):
        only_atm.calculate(torch.tensor([6, 1]), torch.zeros(2, 3), 0.0, PARAM, q=torch.zeros(2))
    with pytest.raises(ValueError):
        d4.dispersion.Disp(model="foo")


def test_pack_pads_with_zeros():
    a, b = torch.tensor([1, 2, 3]), torch.tensor([4])
    assert d4.pack([a, b]).tolist() == [[1, 2, 3], [4, 0, 0]]
