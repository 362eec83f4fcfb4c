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


# --------------------------------------------------------------------------
# K5-K7: full-path vectors of the reference that run through the EEQ charges
# (q=None).  They pin oracle/eeq_oracle.py (restated tad-multicharge) together
# with the D4 restatement; K6 was produced by the reference's own autograd.
# --------------------------------------------------------------------------
import eeq_oracle as eeq  # noqa: E402

SINGLE_Z = torch.tensor([6, 6, 6, 6, 7, 6, 16, 1, 1, 1, 1, 1])  # examples/single.py:7-9
SINGLE_XYZ = torch.tensor([
    [-2.56745685564671, -0.02509985979910, 0.0], [-1.39177582455797, +2.27696188880014, 0.0],
    [+1.27784995624894, +2.45107479759386, 0.0], [+2.62801937615793, +0.25927727028120, 0.0],
    [+1.41097033661123, -1.99890996077412, 0.0], [-1.17186102298849, -2.34220576284180, 0.0],
    [-2.39505990368378, -5.22635838332362, 0.0], [+2.41961980455457, -3.62158019253045, 0.0],
    [-2.51744374846065, +3.98181713686746, 0.0], [+2.24269048384775, +4.24389473203647, 0.0],
    [+4.66488984573956, +0.17907568006409, 0.0], [-4.60044244782237, -0.17794734637413, 0.0],
], dtype=F64)  # fmt: skip
FORMAMIDE2_Z = torch.tensor([6, 6, 7, 7, 1, 1, 1, 1, 1, 1, 8, 8])  # examples/batch.py:8-13
FORMAMIDE2_XYZ = torch.tensor([
    [-3.81469488143921, +0.09993441402912, 0.0], [+3.81469488143921, -0.09993441402912, 0.0],
    [-2.66030049324036, -2.15898251533508, 0.0], [+2.66030049324036, +2.15898251533508, 0.0],
    [-0.73178529739380, -2.28237795829773, 0.0], [-5.89039325714111, -0.02589114569128, 0.0],
    [-3.71254944801331, -3.73605775833130, 0.0], [+3.71254944801331, +3.73605775833130, 0.0],
    [+0.73178529739380, +2.28237795829773, 0.0], [+5.89039325714111, +0.02589114569128, 0.0],
    [-2.74426102638245, +2.16115570068359, 0.0], [+2.74426102638245, -2.16115570068359, 0.0],
], dtype=F64)  # fmt: skip
FORMAMIDE_Z = torch.tensor([6, 8, 7, 1, 1, 1, 0, 0, 0, 0, 0, 0])
FORMAMIDE_XYZ = torch.tensor([
    [-0.55569743203406, +1.09030425468557, 0.0], [+0.51473634678469, +3.15152550263611, 0.0],
    [+0.59869690244446, -1.16861263789477, 0.0], [-0.45355203669134, -2.74568780438064, 0.0],
    [+2.52721209544999, -1.29200800956867, 0.0], [-2.63139587595376, +0.96447869452240, 0.0],
] + [[0.0, 0.0, 0.0]] * 6, dtype=F64)  # fmt: skip
TPSS0 = dict(s6=1.0, s8=1.62438102, s9=1.0, a1=0.40329022, a2=4.80537871)  # test_grad/test_pos.py:51-57


def test_eeq_charges_of_the_reference_samples():
    # q of test/test_d4/samples.py (Fortran dftd4 values): agreement at the 1e-7 level like K3
    q = eeq.get_eeq_charges(SIH4_Z, SIH4_XYZ, 0.0)
    assert pytest.approx(SIH4_Q, abs=5e-7) == q
    assert abs(q.sum().item()) < 1e-14
    q = eeq.get_eeq_charges(LIH_Z, LIH_XYZ, 0.0)
    assert pytest.approx(LIH_Q, abs=5e-7) == q
    q = eeq.get_eeq_charges(LIH_Z, LIH_XYZ, 1.0)
    assert pytest.approx(1.0, abs=1e-14) == q.sum().item()


def test_k5_sih4_s10_full_path():
    # test/test_d4/test_twobody.py:172-205 (q=None -> EEQ), produced by the reference itself
    q = eeq.get_eeq_charges(SIH4_Z, SIH4_XYZ, 0.0)
    e = orc.dftd4(SIH4_Z, SIH4_XYZ, dict(s8=1.85897750, s9=0.0, s10=1.0, a1=0.44286966, a2=4.60230534), q)
    ref = torch.tensor([-8.8928018057670788e-04] + [-3.3765541880036940e-04] * 4, dtype=F64)
    assert pytest.approx(ref, abs=1e-14) == e


@pytest.mark.parametrize("name", ["LiH", "SiH4"])
def test_k6_gradients_through_eeq(name):
    # test/test_grad/samples_grad.py:40-92, test_pos.py:151-190: autograd of the reference, EEQ on the tape
    z, xyz = (LIH_Z, LIH_XYZ) if name == "LiH" else (SIH4_Z, SIH4_XYZ)
    pos = xyz.clone().requires_grad_(True)
    e = orc.dftd4(z, pos, TPSS0, eeq.get_eeq_charges(z, pos, 0.0))
    (g,) = torch.autograd.grad(e.sum(), pos)
    if name == "LiH":
        ref = torch.tensor([[0, 0, -6.8677584018156501e-05], [0, 0, +6.8677584018156501e-05]], dtype=F64)
    else:
        s = 3.5863777807514914e-06
        ref = s * torch.tensor([[0, 0, 0], [-1, -1, 1], [1, 1, 1], [-1, 1, -1], [1, -1, -1]], dtype=F64)
    assert pytest.approx(ref, abs=1e-14) == g  # agreement is 2e-15 / 9e-16


def test_k7_single_py_energies():
    # examples/single.py:52-71 (asserted there with atol 1e-8; printed to 10 decimals, float32 run)
    q = eeq.get_eeq_charges(SINGLE_Z, SINGLE_XYZ, 0.0)
    e = orc.dftd4(SINGLE_Z, SINGLE_XYZ, TPSSH, q)
    ref = torch.tensor([-0.0020841344, -0.0018971195, -0.0018107513, -0.0018305695, -0.0021737693, -0.0019484236,
                        -0.0022788253, -0.0004080658, -0.0004261866, -0.0004199839, -0.0004280768, -0.0005108935],
                       dtype=F64)  # fmt: skip
    assert pytest.approx(ref, abs=1e-9) == e


def test_k7_formamide_dimer_batch():
    # README.md:302, __init__.py:82-87 doctest, examples/batch.py:60-65 (padded batch of two)
    z = torch.stack([FORMAMIDE2_Z, FORMAMIDE_Z])
    xyz = torch.stack([FORMAMIDE2_XYZ, FORMAMIDE_XYZ])
    q = eeq.get_eeq_charges(z, xyz, torch.zeros(2, dtype=F64))
    assert (q[1, 6:] == 0).all()
    e = orc.dftd4(z, xyz, TPSSH, q).sum(-1)
    # the published digits come from a float32 run (rounding noise ~2e-9: the reference itself gives
    # -0.0088341413 / -0.0027013614 in float32 and -0.0088341428 / -0.0027013615 in float64 here)
    assert pytest.approx([-0.0088341432, -0.0027013607], abs=2e-9) == e.tolist()
    assert pytest.approx(-0.0034314217, abs=4e-9) == (e[0] - 2 * e[1]).item()
