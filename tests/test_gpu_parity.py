"""
Parity of the CUDA path (through the public API -> C ABI) with the committed
golden vectors of the reference and with the oracle.

Tolerances are the ones BASELINE.json's north_star states: FP64 energy
<= 1e-10 relative (to the largest atomic energy of the structure / to the total
energy), FP64 gradient <= 1e-9 absolute; FP32 <= 1e-5 relative.
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

import d4_oracle as orc
from helpers import GOLDEN_CASES, as_torch, load_golden

pytestmark = pytest.mark.gpu

E_RTOL64 = 1e-10
G_ATOL64 = 1e-9
RTOL32 = 1e-5


def _d4():
    import tad_dftd4_b200 as d4

    return d4


def _run(case, dtype, model="d4", grad=True):
    d4 = _d4()
    dev = torch.device("cuda:0")
    numbers, positions, q = as_torch(case, dev, dtype)
    pos = positions.clone().requires_grad_(grad)
    cut = d4.Cutoff(**case["cutoff"], device=dev, dtype=dtype) if case["cutoff"] else None
    e = d4.dftd4(numbers, pos, 0.0, dict(case["param"]), q=q, model=model, cutoff=cut)
    g = None
    if grad:
        (g,) = torch.autograd.grad(e.sum(), pos)
        g = g.cpu().numpy().astype(np.float64)
    return e.detach().cpu().numpy().astype(np.float64), g


def _small(name):
    n = load_golden(name)["numbers"]
    return (n != 0).sum(-1).max()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_energy_f64(name):
    case = load_golden(name)
    if _small(name) > 128:
        pytest.skip("beyond the FP64 small-family limit")
    e, _ = _run(case, torch.float64, grad=False)
    ref = case["energy_d4"]
    scale = np.abs(ref).max()
    assert e.shape == ref.shape
    assert np.abs(e - ref).max() / scale < E_RTOL64
    tot, rtot = e.sum(-1), ref.sum(-1)
    assert np.all(np.abs(tot - rtot) <= E_RTOL64 * np.abs(rtot) + 1e-300)
    assert np.all(e[case["numbers"] == 0] == 0.0)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_gradient_f64(name):
    case = load_golden(name)
    if _small(name) > 100:
        pytest.skip("beyond the FP64 gradient small-family limit")
    _, g = _run(case, torch.float64)
    ref = case["grad_d4"]
    assert g.shape == ref.shape
    assert np.abs(g - ref).max() < G_ATOL64
    assert np.all(g[case["numbers"] == 0] == 0.0)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_energy_and_gradient_f32(name):
    case = load_golden(name)
    e, g = _run(case, torch.float32)
    ref, gref = case["energy_d4"], case["grad_d4"]
    tot, rtot = e.sum(-1), ref.sum(-1)
    assert np.all(np.abs(tot - rtot) <= RTOL32 * np.abs(rtot) + 1e-12)
    assert np.abs(e - ref).max() / np.abs(ref).max() < 5 * RTOL32
    assert np.abs(g - gref).max() < 5 * RTOL32 * max(np.abs(gref).max(), 1e-3)


@pytest.mark.parametrize("name", ["sih4_tpssh", "organic_33", "ragged_batch", "holes", "all_elements"])
def test_properties_f64(name):
    """get_properties (cn, C6 matrix, static polarizabilities) vs the reference's model classes."""
    d4 = _d4()
    case = load_golden(name)
    dev = torch.device("cuda:0")
    numbers, positions, q = as_torch(case, dev, torch.float64)
    cn, qout, c6, alpha = d4.get_properties(numbers, positions, q=q)
    assert qout is q
    assert np.allclose(cn.cpu().numpy(), case["cn"], rtol=1e-12, atol=1e-14)
    assert np.allclose(c6.cpu().numpy(), case["c6"], rtol=1e-11, atol=1e-13)
    assert np.allclose(alpha.cpu().numpy(), case["alpha"], rtol=1e-11, atol=1e-13)


D4S_LIMIT64 = 120  # largest structure of the FP64 D4S kernels


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_d4s_energy_and_gradient_f64(name):
    """model="d4s": pair-dependent Gaussian weights (reference model/d4s.py)."""
    case = load_golden(name)
    if _small(name) > D4S_LIMIT64:
        pytest.skip("beyond the FP64 D4S small-family limit")
    e, g = _run(case, torch.float64, model="d4s")
    ref, gref = case["energy_d4s"], case["grad_d4s"]
    assert np.abs(e - ref).max() / np.abs(ref).max() < E_RTOL64
    tot, rtot = e.sum(-1), ref.sum(-1)
    assert np.all(np.abs(tot - rtot) <= E_RTOL64 * np.abs(rtot) + 1e-300)
    assert np.abs(g - gref).max() < G_ATOL64


@pytest.mark.parametrize("name", ["single_pbe0", "organic_64", "ragged_batch", "tight_cutoffs"])
def test_d4s_f32(name):
    case = load_golden(name)
    e, g = _run(case, torch.float32, model="d4s")
    ref, gref = case["energy_d4s"], case["grad_d4s"]
    tot, rtot = e.sum(-1), ref.sum(-1)
    assert np.all(np.abs(tot - rtot) <= RTOL32 * np.abs(rtot) + 1e-12)
    assert np.abs(g - gref).max() < 5 * RTOL32 * max(np.abs(gref).max(), 1e-3)


def test_class_interface_and_term_selection_gpu():
    """DispD4 == dftd4; single registered terms select two-body / ATM only (reference:
    test/test_d4/test_class.py, test_twobody.py, test_atm.py)."""
    d4 = _d4()
    import d4_oracle as orc

    case = load_golden("organic_33")
    dev = torch.device("cuda:0")
    numbers, positions, q = as_torch(case, dev, torch.float64)
    param = dict(case["param"])
    e = d4.dftd4(numbers, positions, 0.0, param, q=q)
    assert torch.equal(d4.dispersion.DispD4().calculate(numbers, positions, 0.0, param, q=q), e)
    n, p, qq = as_torch(case)
    e2, e3, *_ = orc.dftd4(n, p, param, qq, parts=True)
    two = d4.dispersion.Disp()
    two.register(d4.dispersion.TwoBodyTerm())
    assert (two.calculate(numbers, positions, 0.0, param, q=q).cpu() - e2).abs().max() < 1e-12 * e2.abs().max()
    atm = d4.dispersion.Disp()
    atm.register(d4.dispersion.D4ATMApprox())
    assert (atm.calculate(numbers, positions, 0.0, param).cpu() - e3).abs().max() < 1e-10 * e3.abs().max()
    cn = d4.ncoord.cn_d4(numbers, positions)
    assert np.allclose(cn.cpu().numpy(), case["cn"], rtol=1e-12)


def test_device_status_reports_bad_atomic_number():
    d4 = _d4()
    dev = torch.device("cuda:0")
    numbers = torch.tensor([[6, 1, 120]], device=dev)
    positions = torch.tensor([[[0.0, 0, 0], [0, 0, 2.0], [0, 3.0, 0]]], dtype=torch.float64, device=dev)
    with pytest.raises(ValueError, match="atomic number"):
        d4.dftd4(numbers, positions, 0.0, {"a1": 0.4, "a2": 5.0}, q=torch.zeros(1, 3, dtype=torch.float64, device=dev))


def test_idempotent_and_permutation_invariant():
    """Size-independent properties: same result on repeated calls; permuting the structures
    of a batch permutes the energies; permuting atoms permutes atomic energies (<=1e-13)."""
    d4 = _d4()
    case = load_golden("ragged_batch")
    dev = torch.device("cuda:0")
    numbers, positions, q = as_torch(case, dev, torch.float64)
    param = dict(case["param"])
    e1 = d4.dftd4(numbers, positions, 0.0, param, q=q)
    e2 = d4.dftd4(numbers, positions, 0.0, param, q=q)
    assert torch.equal(e1, e2)
    perm = torch.randperm(numbers.shape[0], device=dev)
    assert torch.equal(d4.dftd4(numbers[perm], positions[perm], 0.0, param, q=q[perm]), e1[perm])  # bitwise
    ap = torch.randperm(numbers.shape[1], device=dev)
    e3 = d4.dftd4(numbers[:, ap], positions[:, ap], 0.0, param, q=q[:, ap])
    assert (e3 - e1[:, ap]).abs().max() < 1e-13 * e1.abs().max()


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("model", ["d4", "d4s"])
def test_weighted_upstream_and_charge_gradient(fused, model):
    """General VJP: arbitrary upstream weights g_i and dL/dq vs oracle autograd, with the
    fused forward (cached sum-gradient unused because g is not a broadcast scalar) and
    with the energy-only forward."""
    d4 = _d4()
    import d4_oracle as orc

    d4.set_fused_forward(fused)
    try:
        case = load_golden("ragged_batch")
        n, p, q = as_torch(case)
        g = torch.from_numpy(np.random.default_rng(4).normal(size=tuple(n.shape)))
        pos = p.clone().requires_grad_(True)
        qq = q.clone().requires_grad_(True)
        e_ref = orc.dftd4(n, pos, case["param"], qq, model=model)
        gp_ref, gq_ref = torch.autograd.grad((e_ref * g).sum(), (pos, qq))
        dev = torch.device("cuda:0")
        pd, qd = p.to(dev).requires_grad_(True), q.to(dev).requires_grad_(True)
        e = d4.dftd4(n.to(dev), pd, 0.0, dict(case["param"]), q=qd, model=model)
        gp, gq = torch.autograd.grad((e * g.to(dev)).sum(), (pd, qd))
        assert (e.detach().cpu() - e_ref.detach()).abs().max() / e_ref.detach().abs().max() < E_RTOL64
        assert (gp.cpu() - gp_ref).abs().max() < G_ATOL64
        assert (gq.cpu() - gq_ref).abs().max() < G_ATOL64
        # scaled sum: broadcast upstream gradient -> cached gradient times the scalar
        pd2 = p.to(dev).requires_grad_(True)
        e2 = d4.dftd4(n.to(dev), pd2, 0.0, dict(case["param"]), q=q.to(dev), model=model)
        (g2,) = torch.autograd.grad(2.5 * e2.sum(), pd2)
        (g2_ref,) = torch.autograd.grad(2.5 * orc.dftd4(n, pos, case["param"], q, model=model).sum(), pos)
        assert (g2.cpu() - g2_ref).abs().max() < G_ATOL64
    finally:
        d4.set_fused_forward(True)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("chunks", [0, 1, 3, 7])
def test_host_buffer_entry_matches_device_entry(dtype, chunks):
    """d4b200_energy_host_* (chunked H2D / kernels / D2H pipeline) == the device-pointer call,
    bitwise, for every chunking (structures are independent, chunk assignment is static)."""
    d4 = _d4()
    case = load_golden("ragged_batch")
    numbers, positions, q = as_torch(case, torch.device("cpu"), dtype)
    rep = 9  # make the batch large enough for several chunks per pipeline slot
    numbers, positions, q = numbers.repeat(rep, 1), positions.repeat(rep, 1, 1), q.repeat(rep, 1)
    param = dict(case["param"])
    e_host = d4.dftd4_host(numbers.pin_memory(), positions.pin_memory(), 0.0, param, q=q.pin_memory(), chunks=chunks)
    assert e_host.device.type == "cpu" and e_host.shape == numbers.shape
    e_dev = d4.dftd4(numbers.cuda(), positions.cuda(), 0.0, param, q=q.cuda()).cpu()
    assert torch.equal(e_host, e_dev)
    if dtype == torch.float64:
        ref = np.tile(case["energy_d4"], (rep, 1))
        assert np.abs(e_host.numpy() - ref).max() / np.abs(ref).max() < E_RTOL64
    # pageable host memory works too (the driver stages it)
    e_page = d4.dftd4_host(numbers, positions, 0.0, param, q=q, chunks=2)
    assert torch.equal(e_page, e_dev)


@pytest.mark.parametrize("model", ["d4", "d4s"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("chunks", [0, 1, 5])
def test_host_buffer_forces_entry_matches_device_entry(dtype, chunks, model):
    """d4b200_energy_gradient_host_* == device-pointer call + autograd, bitwise, for every chunking;
    FP64 results within the north_star tolerances of the golden vectors."""
    d4 = _d4()
    case = load_golden("ragged_batch")
    numbers, positions, q = as_torch(case, torch.device("cpu"), dtype)
    rep = 5
    numbers, positions, q = numbers.repeat(rep, 1), positions.repeat(rep, 1, 1), q.repeat(rep, 1)
    param = dict(case["param"])
    e_host, g_host = d4.dftd4_host(numbers.pin_memory(), positions.pin_memory(), 0.0, param, q=q.pin_memory(),
                                   chunks=chunks, model=model, with_gradient=True)
    assert g_host.device.type == "cpu" and g_host.shape == positions.shape and e_host.shape == numbers.shape
    pos = positions.cuda().requires_grad_(True)
    e_dev = d4.dftd4(numbers.cuda(), pos, 0.0, param, q=q.cuda(), model=model)
    (g_dev,) = torch.autograd.grad(e_dev.sum(), pos)
    assert torch.equal(e_host, e_dev.detach().cpu())
    assert torch.equal(g_host, g_dev.cpu())
    if dtype == torch.float64:
        key = "d4" if model == "d4" else "d4s"
        ref, gref = np.tile(case[f"energy_{key}"], (rep, 1)), np.tile(case[f"grad_{key}"], (rep, 1, 1))
        assert np.abs(e_host.numpy() - ref).max() / np.abs(ref).max() < E_RTOL64
        assert np.abs(g_host.numpy() - gref).max() < G_ATOL64


def test_host_entry_narrow_numbers_and_status():
    """dftd4_host: uint8 / int32 atomic numbers are uploaded narrow and give bitwise the int64 result;
    a bad atomic number or a wrong output buffer raises (the host entries read the kernels' status)."""
    import tad_dftd4_b200 as d4

    numbers, positions, q = orc.organic_batch([12, 33, 7, 50, 20], seed=5)
    param = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)
    ref_e, ref_g = d4.dftd4_host(numbers, positions, 0.0, param, q=q, with_gradient=True)
    for dt in (torch.uint8, torch.int32):
        e, g = d4.dftd4_host(numbers.to(dt), positions, 0.0, param, q=q, with_gradient=True)
        assert torch.equal(e, ref_e) and torch.equal(g, ref_g)
        assert torch.equal(d4.dftd4_host(numbers.to(dt), positions, 0.0, param, q=q), d4.dftd4_host(numbers, positions, 0.0, param, q=q))
    dev_e = d4.dftd4(numbers.cuda(), positions.cuda(), 0.0, param, q=q.cuda()).cpu()
    assert torch.allclose(ref_e, dev_e, rtol=1e-13, atol=0)
    bad = numbers.clone()
    bad[1, 3] = 110
    with pytest.raises(ValueError):
        d4.dftd4_host(bad, positions, 0.0, param, q=q)
    with pytest.raises(ValueError):
        d4.dftd4_host(bad.to(torch.uint8), positions, 0.0, param, q=q, with_gradient=True)
    with pytest.raises(ValueError):
        d4.dftd4_host(numbers, positions, 0.0, param, q=q, out=torch.empty(3, dtype=torch.float64))
    with pytest.raises(ValueError):
        d4.dftd4_host(numbers, positions, 0.0, param, q=q, out=torch.empty(numbers.shape, dtype=torch.float32))


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_unaligned_views_and_odd_width_are_bitwise_consistent(dtype):
    """Odd padded width (rows of odd structures are only 8-byte aligned) and views that start inside an
    allocation: the bulk-async staging windows of the persistent kernels must not change a single bit,
    whichever structures a CTA happens to stage or to load directly."""
    import tad_dftd4_b200 as d4

    dev = torch.device("cuda:0")
    numbers, positions, q = orc.organic_batch([12, 33, 7, 21, 5, 30, 33, 2, 19] * 30, seed=6)
    n, p, qq = numbers.to(dev), positions.to(dev, dtype), q.to(dev, dtype)
    param = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)

    def run(sl):
        pos = p[sl].clone().requires_grad_(True)
        e = d4.dftd4(n[sl], pos, 0.0, param, q=qq[sl])
        (g,) = torch.autograd.grad(e.sum(), pos)
        return d4.dftd4(n[sl], p[sl], 0.0, param, q=qq[sl]), e.detach(), g

    full = run(slice(None))
    if dtype == torch.float64:
        ref = orc.dftd4(numbers[:9], positions[:9], param, q[:9])
        assert ((full[0][:9].cpu() - ref).abs().max() / ref.abs().max()) < 1e-10
    for sl in (slice(1, None), slice(3, -2), slice(0, 1), slice(-1, None)):
        part = run(sl)
        for a, b in zip(part, full):
            assert torch.equal(a, b[sl])


def test_calls_from_two_host_threads_on_their_own_streams():
    """ctypes releases the GIL during a library call; the calls on one tables handle are serialised (_lib._serialised)
    and workspaces are kept per stream, so two threads get the results of a single-threaded run."""
    import threading

    import tad_dftd4_b200 as d4

    dev = torch.device("cuda:0")
    cases = []
    for seed, sizes in ((1, [12, 33, 60, 7] * 8), (2, [100, 80, 64] * 6)):
        n, p, q = orc.organic_batch(sizes, seed=seed)
        cases.append((n.to(dev), p.to(dev), q.to(dev)))
    par = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)

    def run(case):
        n, p, q = case
        pos = p.clone().requires_grad_(True)
        e = d4.dftd4(n, pos, 0.0, par, q=q)
        (g,) = torch.autograd.grad(e.sum(), pos)
        return e.detach(), g

    want = [run(c) for c in cases]
    torch.cuda.synchronize()
    out: dict = {}

    def worker(k):
        with torch.cuda.stream(torch.cuda.Stream(dev)):
            res = [run(cases[k]) for _ in range(25)]
            torch.cuda.current_stream().synchronize()
        out[k] = res

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for k in range(2):
        for e, g in out[k]:
            assert torch.equal(e, want[k][0]) and torch.equal(g, want[k][1])
