"""
Plugin boundary, CPU contract (SURVEY.md 8b secondary; no kernels involved).

* The reference's four ``examples/*.py`` run UNCHANGED (read from ``/root/reference/examples``
  in the build container, skipped where that tree does not exist) against
  ``tad_dftd4_b200.install()``: ``import tad_dftd4 as d4`` resolves to the alias module and every
  name the scripts use exists with the reference's call signature.  The arithmetic behind the
  entry point is the CPU oracle here (patched in below -- the product path has no CPU route);
  ``tests/test_gpu_install.py`` runs the same call sequence on the B200 kernels.
* ``FusedD4Term.calculate`` has the reference's ``DispTerm.calculate`` signature and can be
  driven by a ``Disp`` that follows the reference's ``Disp.calculate`` loop.
"""
from __future__ import annotations

import inspect
import runpy
import sys
from pathlib import Path
from unittest.mock import patch

import pytest
import torch

import d4_oracle as orc
import tad_dftd4_b200 as d4
from tad_dftd4_b200 import dispersion

REF_EXAMPLES = Path("/root/reference/examples")
SHIM = Path(__file__).resolve().parent.parent / "oracle" / "mctc_shim"


def _oracle_dftd4(numbers, positions, charge, param, *, model="d4", q=None, cutoff=None, **kw):
    """Stand-in for the C call: same signature as ``tad_dftd4_b200.dftd4``."""
    import eeq_oracle

    name = model if isinstance(model, str) else {"D4Model": "d4", "D4SModel": "d4s"}[type(model).__name__]
    par = {k: (float(v) if v is not None else v) for k, v in param.items() if k != "doi"}
    if q is None:
        q = eeq_oracle.get_eeq_charges(numbers, positions.double(), charge)
    return orc.dftd4(numbers, positions.double(), par, q.double(), model=name).to(positions.dtype)


@pytest.fixture()
def installed():
    sys.path.insert(0, str(SHIM))
    with patch.object(d4, "dftd4", _oracle_dftd4), patch.object(dispersion, "dftd4", _oracle_dftd4):
        kind = d4.install()
        try:
            yield kind
        finally:
            d4.uninstall()
            sys.path.remove(str(SHIM))
            for key in [k for k in sys.modules if k == "tad_mctc" or k.startswith("tad_mctc.")]:
                del sys.modules[key]


@pytest.mark.skipif(not REF_EXAMPLES.is_dir(), reason="reference tree not present on this box")
@pytest.mark.parametrize("script", ["single.py", "batch.py", "d4s.py", "forces.py"])
def test_reference_example_runs_unchanged(installed, script, capsys):
    assert installed == "alias"  # the reference itself is not importable here (tad-mctc absent)
    runpy.run_path(str(REF_EXAMPLES / script), run_name="__main__")  # single.py / forces.py carry their own asserts
    out = capsys.readouterr().out
    if script == "batch.py":  # the energies the script documents in its comments (float32 run)
        assert "-0.00883414" in out and "-0.00270136" in out
    if script == "d4s.py":
        assert "tensor([" in out


def test_install_is_reversible():
    assert d4.install() == "alias"
    import tad_dftd4

    assert tad_dftd4.dftd4 is d4.dftd4 and tad_dftd4.Param is d4.Param
    assert tad_dftd4.dispersion.DispD4 is dispersion.DispD4 and tad_dftd4.disp.dftd4 is d4.dftd4
    d4.uninstall()
    assert "tad_dftd4" not in sys.modules and "tad_dftd4.disp" not in sys.modules
    with pytest.raises(ImportError):
        import tad_dftd4  # noqa: F401,F811


def test_term_calculate_signature_is_the_reference_plugin_interface():
    # dispersion/base.py:88-101 of the reference
    want = ["self", "numbers", "positions", "param", "cn", "model", "q", "r4r2", "rvdw", "cutoff"]
    for cls in (dispersion.FusedD4Term, dispersion.TwoBodyTerm, dispersion.D4ATMApprox):
        assert list(inspect.signature(cls.calculate).parameters) == want


def test_fused_term_driven_by_a_reference_style_disp_loop():
    """The delegate loop of the reference's Disp.calculate (dispersion/base.py:413-431) over a
    registered FusedD4Term reproduces dftd4; the two single terms add up to the same energy."""
    numbers, positions, q = orc.organic_batch([9, 14], seed=23)
    param = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)
    want = orc.dftd4(numbers, positions, param, q)
    cn = orc.cn_d4(numbers, positions)  # what the reference hands to every term (ignored by the fused kernel)
    with patch.object(dispersion, "dftd4", _oracle_dftd4):
        for terms in ([dispersion.FusedD4Term()], [dispersion.TwoBodyTerm(), dispersion.D4ATMApprox()]):
            energy = torch.zeros_like(positions[..., 0])
            for term in terms:
                energy = energy + term.calculate(numbers=numbers, positions=positions, param=param, cn=cn, model="d4",
                                                 q=q, r4r2=None, rvdw=None, cutoff=None)  # fmt: skip
            assert torch.allclose(energy, want, rtol=1e-12, atol=1e-16)
        # a charge-independent two-body term is the same kernel with zeta(q = 0) (twobody.py:76-80)
        e0 = dispersion.TwoBodyTerm(charge_dependent=False).calculate(numbers, positions, param, cn, "d4", q)
        want0, *_ = orc.dftd4(numbers, positions, param, torch.zeros_like(q), parts=True)
        assert torch.allclose(e0, want0, rtol=1e-12, atol=1e-16)
        with pytest.raises(NotImplementedError):
            dispersion.D4ATMApprox(charge_dependent=True).calculate(numbers, positions, param, cn, "d4", q)
        with pytest.raises(ValueError):
            dispersion.FusedD4Term().calculate(numbers, positions, param, cn, "d4", None)


REF_SRC = Path("/root/reference/src")


@pytest.mark.skipif(not REF_SRC.is_dir(), reason="reference tree not present on this box")
def test_install_rebinds_an_importable_reference_and_its_disp_drives_the_terms():
    """With the reference importable (here: its unmodified sources on top of oracle/mctc_shim) ``install()`` rebinds its
    entry points instead of registering an alias, a reference ``DispD4`` instance keeps its model, and the reference's
    OWN ``Disp.calculate`` loop (dispersion/base.py:409-431) drives a registered ``FusedD4Term``."""
    numbers, positions, q = orc.organic_batch([9, 14], seed=29)
    param = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)
    charge = torch.zeros(2, dtype=torch.float64)
    sys.path[:0] = [str(SHIM), str(REF_SRC)]
    try:
        with patch.object(d4, "dftd4", _oracle_dftd4), patch.object(dispersion, "dftd4", _oracle_dftd4):
            assert d4.install() == "rebound"
            import tad_dftd4 as ref
            from tad_dftd4.dispersion.base import Disp as RefDisp

            assert ref.__file__.startswith(str(REF_SRC)) and ref.dftd4 is _oracle_dftd4 and ref.disp.dftd4 is _oracle_dftd4
            tpar = {k: torch.tensor(v, dtype=torch.float64) for k, v in param.items()}
            for model in ("d4", "d4s"):
                want = orc.dftd4(numbers, positions, param, q, model=model)
                got = ref.dispersion.DispD4(model=model, dtype=torch.float64).calculate(numbers, positions, charge, tpar, q=q)
                assert torch.allclose(got, want, rtol=1e-12, atol=1e-16)
                driver = RefDisp(model=model, dtype=torch.float64)  # the reference's class, the reference's loop
                driver.register(dispersion.FusedD4Term())
                got = driver.calculate(numbers, positions, charge, tpar, q=q)
                assert torch.allclose(got, want, rtol=1e-12, atol=1e-16)
            d4.uninstall()
            assert ref.dftd4 is not _oracle_dftd4  # the reference's own function is back
    finally:
        d4.uninstall()
        sys.path.remove(str(SHIM))
        sys.path.remove(str(REF_SRC))
        for key in [k for k in sys.modules if k.split(".")[0] in ("tad_mctc", "tad_dftd4")]:
            del sys.modules[key]
