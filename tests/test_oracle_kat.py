"""
Pins ``oracle/d4_oracle.py`` against the known-answer vectors that the
reference keeps in its own test-suite.  Geometries of
LiH / SiH4 follow from symmetry + bond length; the golden
numbers are the Fortran ``dftd4`` values quoted in
``/root/reference/test/test_model/samples.py:63-124`` and
``/root/reference/test/test_d4/samples.py:51-140`` with the reference's own
tolerances (``test_twobody.py:72``, ``test_c6.py:41``, ``test_weights.py:45``).
"""
from __future__ import annotations

import pytest
import torch

import d4_oracle as orc

F64 = torch.float64
TPSSH = dict(s6=1.0, s8=1.85897750, s9=1.0, s10=0.0, alp=16.0, a1=0.44286966, a2=4.60230534)
TOL = torch.finfo(F64).eps ** 0.5 * 10  # test_twobody.py:72

A = 1.61768389755830
SIH4_Z = torch.tensor([14, 1, 1, 1, 1])
SIH4_XYZ = torch.tensor([[0, 0, 0], [A, A, -A], [-A, -A, -A], [A, -A, A], [-A, A, A]], dtype=F64)
SIH4_Q = torch.tensor(
    [-8.412842390895063e-02, 2.103210597723753e-02, 2.103210597723774e-02,
     2.103210597723764e-02, 2.103210597723773e-02], dtype=F64)  # fmt: skip
Z0 = 1.50796743897235
LIH_Z = torch.tensor([3, 1])
LIH_XYZ = torch.tensor([[0, 0, -Z0], [0, 0, Z0]], dtype=F64)
LIH_Q = torch.tensor([3.708714958301688e-01, -3.708714958301688e-01], dtype=F64)


def test_k1_lih_c6_from_golden_weights():
    # test_model/samples.py:72-85 (gw, Fortran order (3,2)) and :104-112 (c6)
    gw = torch.zeros(2, 7, dtype=F64)
    gw[0, :3] = torch.tensor([1.8699287753787968e-02, 9.7889292523075033e-01, 1.8718044551687104e-37])
    gw[1, :3] = torch.tensor([7.9608926855620182e-02, 3.5225968112617356e00, 0.0])
    c6 = orc.atomic_c6_d4(orc.reference_c6(LIH_Z), gw)
    ref = torch.tensor([[4.1059628873073926e01, 2.9129176877403175e01],
                        [2.9129176877403175e01, 4.0408036338319796e01]], dtype=F64)  # fmt: skip
    assert pytest.approx(ref, rel=1e-7) == c6  # reference tolerance is 1e-5 (test_c6.py:41)


def test_k3_lih_weights():
    cn = orc.cn_d4(LIH_Z, LIH_XYZ)
    assert pytest.approx(0.80226843, abs=1e-7) == cn[0].item()
    gw = orc.weight_references_d4(LIH_Z, cn, LIH_Q)
    ref0 = torch.tensor([1.8699287753787968e-02, 9.7889292523075033e-01, 1.8718044551687104e-37])
    ref1 = torch.tensor([7.9608926855620182e-02, 3.5225968112617356e00, 0.0])
    assert pytest.approx(ref0, abs=1e-6) == gw[0, :3]  # test_weights.py:45
    assert pytest.approx(ref1, abs=1e-6) == gw[1, :3]


def test_k4_sih4_twobody_threebody_total():
    e2, e3, cn, _, _ = orc.dftd4(SIH4_Z, SIH4_XYZ, TPSSH, SIH4_Q, parts=True)
    assert pytest.approx([3.64990496, 0.91247624], abs=1e-7) == cn[:2].tolist()
    disp2 = torch.tensor([-7.3576887054011008e-04] + [-2.9019661352652499e-04] * 4, dtype=F64)
    disp3 = torch.tensor([+1.0737848175727815e-09] + [+1.3076135979507488e-08] * 4, dtype=F64)
    disp = torch.tensor([-7.3576779675529251e-04] + [-2.9018353739054548e-04] * 4, dtype=F64)
    assert pytest.approx(disp2, abs=TOL) == e2
    assert pytest.approx(disp3, rel=1e-9) == e3  # geometry-only term: 1e-10 level
    assert pytest.approx(disp, abs=TOL) == e2 + e3


def test_k4_lih_twobody_d4_and_d4s():
    e = orc.dftd4(LIH_Z, LIH_XYZ, dict(TPSSH, s9=0.0), LIH_Q)
    assert pytest.approx([-2.5509356117699210e-04] * 2, abs=TOL) == e.tolist()
    e = orc.dftd4(LIH_Z, LIH_XYZ, dict(TPSSH, s9=0.0), LIH_Q, model="d4s")
    assert pytest.approx([-8.7076436586852103e-04] * 2, abs=TOL) == e.tolist()


def test_sih4_d4s_twobody():
    e = orc.dftd4(SIH4_Z, SIH4_XYZ, dict(TPSSH, s9=0.0), SIH4_Q, model="d4s")
    ref = [-8.5424954549356738e-04] + [-3.2471197868320798e-04] * 4
    assert pytest.approx(ref, abs=TOL) == e.tolist()


def test_atm_mask_tests_two_distances_only():
    """open triple (two short edges) contributes to the two end
    atoms only, e/6 each, through the centre form."""
    z = torch.tensor([6, 6, 6])
    xyz = torch.tensor([[0.0, 0, 0], [30.0, 0, 0], [60.0, 0.5, 0]], dtype=F64)
    q = torch.zeros(3, dtype=F64)
    e2, e3, *_ = orc.dftd4(z, xyz, dict(TPSSH), q, parts=True)
    assert e3[1].item() == 0.0
    assert e3[0].item() != 0.0 and pytest.approx(e3[0].item(), rel=1e-12) == e3[2].item()


def test_errors():
    with pytest.raises(TypeError):
        orc.dftd4(LIH_Z, LIH_XYZ, dict(s8=1.0), LIH_Q)
    with pytest.raises(ValueError):
        orc.dftd4(LIH_Z, LIH_XYZ[:1], TPSSH, LIH_Q)
