#!/usr/bin/env python
"""
bench.py -- headline measurement of the B200-native DFT-D4 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload c3|c2|c5|c1|c4|c4g] [--dtype f64|f32] [--no-subs]

A *step* is one pass of the hot path over one batch of synthetic structures.  The
headline workload is BASELINE.json's metric, "D4 energy+forces":

* ``c3`` (default, configs[2]): 1024 synthetic 100-atom molecules, atom-resolved D4
  energy + analytic gradient (``autograd.grad(E.sum(), positions)``), PBE0-D4, FP64;
* ``c2`` (configs[1]): 4096 synthetic 20-60-atom organics, energy only;
* ``c5`` (configs[4]): the C3 inputs with the D4S model (``--dtype f32|f64``);
* ``c4`` / ``c4g`` (configs[3]): one 20 001-atom water cluster, energy / energy+gradient,
  row-block split over the ranks + NCCL all-reduces;
* ``c1`` (configs[0]): the 12-atom molecule of ``examples/single.py`` (latency).

With ``--gpus N`` (launched under torchrun, one rank per GPU) the headline ``value`` is
WEAK scaling: every rank works on its own batch of the workload's shape (structures are
independent, no data-path collective), value = all structures of all ranks / slowest
rank's device time.  The same JSON line carries, unless ``--no-subs``:

* ``strong``: the FIXED global batch of C3 and C2 sharded over the ranks through
  ``tad_dftd4_b200.parallel.dftd4_sharded`` (north_star: "4096 ... sharded 1/2/4/8");
* ``sub``: C2, C5-FP64, C5-FP32 measured the same way as the headline (fewer steps);
* ``c4``: the single large system with forces, STRONG scaling over the ranks, with the
  time of the all-reduces split out and the work counted in SURVEY 8(d)'s units.

``value`` is measured with inputs resident in HBM, ``e2e`` through the public
host-buffer API from pinned HOST tensors (H2D + kernels + D2H inside the timed region).

``--impl reference`` times the CPU implementation of the same path on the host cores:
the reference itself cannot be installed here (its dependencies tad-mctc /
tad-multicharge are not in the image), so this arm runs the oracle restatement
(``oracle/d4_oracle.py``: the reference's dense torch formulation, bit-identical to
the unmodified reference on every golden case) and says so (``cpu_baseline.kind =
"port"``); each of its steps is a bounded sample of the workload.
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

# torchrun exports OMP_NUM_THREADS=1 to every rank; rank 0 also times the CPU baseline on all host
# cores (the other ranks only wait for it), so give its OpenMP / MKL pools the whole box back
# BEFORE torch initialises them
if os.environ.get("RANK", "0") == "0" and os.environ.get("OMP_NUM_THREADS") == "1":
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import bench_inputs  # noqa: E402  (synthetic input generators; numpy only)

PBE0 = dict(s8=1.20065498, a1=0.40085597, a2=5.02928789)  # d4.toml:269 of the reference

# algorithmic flop per unit of work (contract figures)
F_PCN, F_P2, F_T, F_W = 18, 248, 30, 120
F_PCN_G, F_P2_G, F_T_G = 30, 320, 90
NCLS = 5  # D4B200_NCLASS
METRIC = "D4 energy+forces throughput (batched molecules/s)"

WORKLOADS = {
    "c2": dict(nbatch=4096, lo=20, hi=60, seed=2, grad=False,
               name="C2: 4096 synthetic 20-60-atom organics, D4 energy (two-body + ATM), padded to 60"),
    "c3": dict(nbatch=1024, lo=100, hi=100, seed=3, grad=True,
               name="C3: 1024 synthetic 100-atom molecules, D4 energy + analytic gradient incl. ATM"),
    "c5": dict(nbatch=1024, lo=100, hi=100, seed=3, grad=True, model="d4s",
               name="C5: 1024 synthetic 100-atom molecules (the C3 inputs), D4S model, energy + analytic "
                    "gradient incl. ATM"),
    "c1": dict(single=True, nbatch=1, grad=False, seed=0,
               name="C1: examples/single.py 12-atom molecule, D4 energy, PBE0 parameters (one structure: latency)"),
    "c4": dict(nmol=6667, seed=4, grad=False, large=True,
               name="C4: single 20001-atom water cluster (6667 H2O), D4 energy, default 60/40/30 Bohr cutoffs, "
                    "row-block split over the GPUs + all-reduce"),
    "c4g": dict(nmol=6667, seed=4, grad=True, large=True,
                name="C4: single 20001-atom water cluster (6667 H2O), D4 energy + analytic gradient, default "
                     "60/40/30 Bohr cutoffs, row-block split over the GPUs + all-reduces"),
}  # fmt: skip


_OUT = None  # the process's real stdout (see main): exactly one JSON line goes there


def emit(line: dict) -> None:
    print(json.dumps(line), file=_OUT or sys.stdout, flush=True)


def oracle():
    """The CPU oracle: ONLY the cpu_baseline leg and the reference arm call this."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import d4_oracle  # noqa: E402

    return d4_oracle


def make_batch(wl: dict, rank: int):
    if wl.get("single"):
        return bench_inputs.single_molecule()
    rng = np.random.default_rng(wl["seed"] + 1000 * rank)
    sizes = rng.integers(wl["lo"], wl["hi"] + 1, size=wl["nbatch"])
    world = int(os.environ.get("WORLD_SIZE", "1"))  # the ranks of one box share its host cores
    numbers, positions, q = bench_inputs.organic_batch_parallel(
        sizes, seed=wl["seed"] + 1000 * rank, workers=max(2, min(32, (os.cpu_count() or 8) // world)))
    pad = wl["hi"] - numbers.shape[1]  # every rank's batch has the workload's padded width
    if pad > 0:
        numbers = torch.nn.functional.pad(numbers, (0, pad))
        positions = torch.nn.functional.pad(positions, (0, 0, 0, pad))
        q = torch.nn.functional.pad(q, (0, pad))
    return numbers, positions, q


def config_dict(wl: dict, world: int, dtype: str, eeq: bool) -> dict:
    """``config`` of the JSON line: identical for the B200 arm and the reference arm."""
    nb = wl.get("nbatch", 1)
    return {
        "workload": wl["name"],
        "structures_per_gpu": nb,
        "global_batch": nb * world,
        "parallelism": f"structure-sharded x{world}, no data-path collective",
        "l2": "flushed (256 MB write) between timed iterations",
        "param": "PBE0-D4 (s8 1.20065498, a1 0.40085597, a2 5.02928789), "
                 + ("q=None: EEQ-2019 charges solved inside every step" if eeq else "explicit charges q"),
        "dispersion_model": wl.get("model", "d4"),
        "precision": dtype,
    }  # fmt: skip


def work_counts(numbers: torch.Tensor, grad: bool):
    n = (numbers != 0).sum(-1).to(torch.float64)
    pairs = n * (n - 1) / 2
    triples = n * (n - 1) * (n - 2) / 6
    flop = (F_PCN + F_P2) * pairs + F_T * triples + F_W * n
    if grad:
        flop = flop + (F_PCN_G + F_P2_G) * pairs + F_T_G * triples
    return n, pairs, triples, flop


def measured_traffic(workload: str, dtype: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full
    capture of the same command (profiles/traffic.json), or None."""
    try:
        with open(ROOT / "profiles" / "traffic.json", encoding="utf8") as fp:
            return json.load(fp).get(f"{workload}_{dtype}", {}).get("dram_bytes_per_launch")
    except OSError:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")  # fmt: skip

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
            self.thread = threading.Thread(target=self._read)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
                self.proc.wait()
            self.thread.join()  # the pipe is closed: the reader has seen EOF
            self.proc.stdout.close()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}  # fmt: skip


# --------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# --------------------------------------------------------------------------
def cpu_sample_shape(wl: dict, nbatch: int):
    chunk = 8 if wl["grad"] else 64
    return min(chunk * 2, nbatch), chunk


def cpu_time_sample(wl: dict, numbers, positions, q, nsample: int, chunk: int, eeq: bool = False, keep=None):
    """Seconds the dense CPU formulation needs for ``nsample`` structures of the
    workload (model rebuilt per call like the reference, dispersion/base.py:363).
    ``keep`` (a list) receives the (energy, gradient) of every chunk for the parity report."""
    orc = oracle()
    model = wl.get("model", "d4")
    torch.set_num_threads(os.cpu_count() or 1)
    if eeq:
        import eeq_oracle
    t0 = time.perf_counter()
    for s in range(0, nsample, chunk):
        sl = slice(s, min(s + chunk, nsample))
        if eeq:  # default q=None path of the reference: EEQ charges on the tape
            pos = positions[sl].clone().requires_grad_(wl["grad"])
            e = orc.dftd4(numbers[sl], pos, PBE0, eeq_oracle.get_eeq_charges(numbers[sl], pos, 0.0), model=model)
            g = torch.autograd.grad(e.sum(), pos)[0] if wl["grad"] else None
            e = e.detach()
        elif wl["grad"]:
            e, g = orc.energy_and_gradient(numbers[sl], positions[sl], PBE0, q[sl], model=model)
        else:
            e, g = orc.dftd4(numbers[sl], positions[sl], PBE0, q[sl], model=model), None
        if keep is not None:
            keep.append((e, g))
    return time.perf_counter() - t0


def cpu_sample_text(nsample: int, chunk: int) -> str:
    return (f"first {nsample} structures of the workload per step in chunks of {chunk} (dense N^3 temporaries), "
            f"float64, torch threads = {torch.get_num_threads()}; oracle/d4_oracle.py = the reference's dense torch "
            "formulation (the reference itself is not installable here: tad-mctc / tad-multicharge absent)")  # fmt: skip


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    if wl.get("large"):
        emit({"impl": "reference", "unavailable": "the dense reference formulation needs "
              "157 GB for rc6 and 6.4e13 B per N^3 temporary at 20k atoms (BASELINE.md)"})
        return
    numbers, positions, q = make_batch(wl, 0)
    nsample, chunk = cpu_sample_shape(wl, numbers.shape[0])
    times = []
    for it in range(args.warmup + args.steps):
        dt = cpu_time_sample(wl, numbers, positions, q, nsample, chunk, eeq=args.eeq)
        if it >= args.warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    value = nsample / sec
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value, "unit": "molecules/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(wl, world, "f64", args.eeq),
        "cpu_baseline": {"value": value, "unit": "molecules/s", "cores": os.cpu_count() or 1, "kind": "port",
                         "sample": cpu_sample_text(nsample, chunk)},
        "e2e": {"value": value, "unit": "molecules/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": "CPU arm: the host cores of this box run the oracle port of tad-dftd4 0.8.0 (kind = port); "
                "it does not scale with --gpus",
    }  # fmt: skip
    emit(line)


# --------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------
class Ctx:
    """Process-wide state of the B200 arm."""

    def __init__(self, args, rank, world, local_rank):
        import tad_dftd4_b200 as d4
        from tad_dftd4_b200 import _lib

        self.args, self.rank, self.world = args, rank, world
        self.d4 = d4
        self.dev = torch.device("cuda", local_rank)
        self.local_rank = local_rank
        torch.cuda.set_device(self.dev)
        self.dist = None
        if world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
            # the other ranks WAIT (blocking sockets, no spinning on the host cores) while rank 0 times the CPU baseline
            self.wait_group = dist.new_group(backend="gloo")
        d4.set_checks(False)  # fully asynchronous steps; parity is the tests' job (and cpu_baseline.parity)
        self.lib = _lib.load()
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)  # > 126 MB L2

    def launches(self) -> int:
        return int(self.lib.d4b200_total_launch_count()) + int(self.lib.d4b200_eeq_launch_count())

    def barrier(self):
        torch.cuda.synchronize(self.dev)
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def timed(self, fn, steps, warmup, host_call=False):
        """Sum of the per-step device times (ms) of ``steps`` calls, L2 flushed before each."""
        for _ in range(warmup):
            fn()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        n0 = self.launches()
        for s in range(steps):
            self.flush.fill_(float(s))  # evict inputs/tables from L2 between timed iterations
            if host_call:  # the host-buffer API runs on its own streams: let the flush finish first
                torch.cuda.current_stream(self.dev).synchronize()
            ev[s][0].record()
            fn()
            ev[s][1].record()
        self.barrier()
        return sum(a.elapsed_time(b) for a, b in ev), self.launches() - n0

    def reduce(self, values, op="max"):
        t = torch.tensor(values, dtype=torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return t.tolist()


class BatchBench:
    """One padded-batch workload on this rank: resident step, host-buffer step, work counts."""

    def __init__(self, ctx: Ctx, key: str, dtype: str, batch, eeq: bool = False):
        self.ctx, self.key, self.wl, self.dtype_name = ctx, key, WORKLOADS[key], dtype
        wl, dev = self.wl, ctx.dev
        self.dtype = torch.float64 if dtype == "f64" else torch.float32
        numbers_h, positions_h, q_h = batch
        positions_h, q_h = positions_h.to(self.dtype), q_h.to(self.dtype)
        self.nat, self.pairs, self.triples, self.flop = work_counts(numbers_h, wl["grad"])
        self.numbers_h, self.positions_h, self.q_h = numbers_h.pin_memory(), positions_h.pin_memory(), q_h.pin_memory()
        # the host-buffer API takes the atomic numbers as the caller holds them; one byte per atom is enough
        self.numbers_h8 = numbers_h.to(torch.uint8).pin_memory()
        self.numbers, self.positions, self.q = self.numbers_h.to(dev), self.positions_h.to(dev), self.q_h.to(dev)
        self.model = wl.get("model", "d4")
        self.eeq = eeq
        self.out_e = torch.empty(numbers_h.shape, dtype=self.dtype).pin_memory()
        self.out_g = torch.empty(positions_h.shape, dtype=self.dtype).pin_memory() if wl["grad"] else None
        es = positions_h.element_size()
        self.h2d = numbers_h.numel() * (8 if eeq else 1) + (positions_h.numel() + (0 if eeq else q_h.numel())) * es
        self.d2h = self.out_e.numel() * es + (self.out_g.numel() * es if wl["grad"] else 0)

    def call(self, n, p, qq):
        return self.ctx.d4.dftd4(n, p, 0.0, PBE0, q=None if self.eeq else qq, model=self.model)

    def step_resident(self):
        if self.wl["grad"]:
            pos = self.positions.detach().requires_grad_(True)
            e = self.call(self.numbers, pos, self.q)
            (g,) = torch.autograd.grad(e.sum(), pos)
            return e, g
        return self.call(self.numbers, self.positions, self.q), None

    def step_e2e(self):
        dev, wl = self.ctx.dev, self.wl
        if not self.eeq:
            # public host-buffer API: pinned host tensors in, pinned host tensors out; the C ABI pipelines
            # H2D copy / kernels / D2H copy in chunks (d4b200_energy_host_*, d4b200_energy_gradient_host_*)
            self.ctx.d4.dftd4_host(self.numbers_h8, self.positions_h, 0.0, PBE0, q=self.q_h, model=self.model,
                                   device=dev, out=self.out_e, with_gradient=wl["grad"], out_gradient=self.out_g)
            return
        n = self.numbers_h.to(dev, non_blocking=True)
        p = self.positions_h.to(dev, non_blocking=True)
        if wl["grad"]:
            p.requires_grad_(True)
            e = self.call(n, p, None)
            (g,) = torch.autograd.grad(e.sum(), p)
            self.out_g.copy_(g, non_blocking=True)
        else:
            e = self.call(n, p, None)
        self.out_e.copy_(e.detach(), non_blocking=True)

    # ---- measurements ---------------------------------------------------
    def measure(self, steps, warmup):
        """value / e2e of this workload, aggregated over the ranks (weak scaling)."""
        ctx = self.ctx
        total_ms, launches = ctx.timed(self.step_resident, steps, warmup)
        e2e_ms = 1.0
        if not ctx.args.no_e2e:
            e2e_ms, _ = ctx.timed(self.step_e2e, steps, max(warmup, 3), host_call=not self.eeq)
        total_ms, e2e_ms = ctx.reduce([total_ms, e2e_ms], "max")
        nmol, npair, ntrip, nflop = ctx.reduce([float(self.numbers.shape[0]), float(self.pairs.sum()),
                                                float(self.triples.sum()), float(self.flop.sum())], "sum")  # fmt: skip
        sec, e2e_sec = total_ms / steps * 1e-3, e2e_ms / steps * 1e-3
        return {
            "value": nmol / sec, "unit": "molecules/s", "ms_per_step": sec * 1e3, "steps": steps,
            "pair_terms_per_s": npair / sec, "triple_terms_per_s": ntrip / sec,
            "algorithmic_tflops": nflop / sec / 1e12,
            "e2e": None if ctx.args.no_e2e else {"value": nmol / e2e_sec, "unit": "molecules/s", "h2d_bytes_per_step": int(self.h2d),
                    "d2h_bytes_per_step": int(self.d2h), "ms_per_step": e2e_sec * 1e3,
                    "api": "q=None: dftd4() on tensors copied from pinned host memory" if self.eeq else
                           "tad_dftd4_b200.dftd4_host (C ABI d4b200_energy_host_z_*): pinned host tensors in and out, "
                           "atomic numbers as uint8"},
            "gpu_launches": int(launches),
        }  # fmt: skip

    def roofline(self, step_ms, nsteps=6):
        """Dominant kernel of this workload: per-launch duration measured live with CUDA events
        (on the streams the class kernels are launched on, inside the C library)."""
        from tad_dftd4_b200.disp import _Engine

        ctx, lib, wl = self.ctx, self.ctx.lib, self.wl
        engine = _Engine.get(ctx.dev, 3.0, 2.0)
        f32, d4s = int(self.dtype == torch.float32), int(self.model == "d4s")
        lib.d4b200_profile_enable(engine.handle, 1)
        caps = (C.c_int * NCLS)()
        lib.d4b200_class_caps_model(engine.handle, f32, int(wl["grad"]), d4s, caps)
        per_class = [[] for _ in range(NCLS)]
        prep_ms, call_ms = [], []
        for _ in range(nsteps):
            ctx.flush.fill_(1.0)
            self.step_resident()
            ms = (C.c_float * (NCLS + 2))()
            lib.d4b200_profile_read(engine.handle, ms)
            for c in range(NCLS):
                if ms[c] >= 0:
                    per_class[c].append(ms[c])
            prep_ms.append(ms[NCLS])
            call_ms.append(ms[NCLS + 1])
        lib.d4b200_profile_enable(engine.handle, 0)
        class_ms = [statistics.mean(v[1:] if len(v) > 1 else v) if v else 0.0 for v in per_class]
        lo, class_flop = 0, []
        for c in range(NCLS):
            sel = (self.nat >= lo) & (self.nat <= caps[c]) if caps[c] >= lo else torch.zeros_like(self.nat, dtype=torch.bool)
            class_flop.append(float(self.flop[sel].sum()))
            lo = caps[c] + 1
        dom = max(range(NCLS), key=lambda c: class_ms[c])
        peak_tf = C.c_double(0.0)
        scratch = torch.empty(64 * 1024 * 1024, dtype=torch.uint8, device=ctx.dev)
        lib.d4b200_measure_fp64_peak(engine.handle, scratch.data_ptr(), scratch.numel(),
                                     torch.cuda.current_stream(ctx.dev).cuda_stream, C.byref(peak_tf))  # fmt: skip
        fp64 = self.dtype == torch.float64
        nominal_tf = 37.2 if fp64 else 74.4
        peak = peak_tf.value if fp64 else 2 * peak_tf.value
        achieved = class_flop[dom] / (class_ms[dom] * 1e-3) / 1e12 if class_ms[dom] > 0 else 0.0
        whole = float(self.flop.sum()) / (step_ms * 1e-3) / 1e12 if step_ms > 0 else 0.0
        return {
            "bound": "fp64" if fp64 else "fp32",
            "kernel": f"small_kernel<{'double' if fp64 else 'float'},{'grad' if wl['grad'] else 'energy'},"
                      f"{self.model}> size class <= {caps[dom]} atoms",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak if peak else None,
            "traffic": measured_traffic(self.key, self.dtype_name),
            "peak_source": "measured in this run: DFMA chain microbenchmark (d4b200_measure_fp64_peak"
                           + ("" if fp64 else ", x2 for FP32") + f"); nominal 148 SM x 64 lanes x 2 x 1.965 GHz = "
                           f"{nominal_tf} TFLOP/s; MEASURED_PEAKS.json has no FP64 entry",
            "nominal_peak": nominal_tf, "frac_of_nominal": achieved / nominal_tf,
            "kernel_ms": class_ms[dom], "kernel_algorithmic_flop": class_flop[dom],
            "all_class_ms": class_ms, "prep_ms": statistics.mean(prep_ms[1:]),
            "serialised_call_ms": statistics.mean(call_ms[1:]),
            "step_share": class_ms[dom] / step_ms if step_ms else None,
            "whole_step_frac": whole / peak if peak else None,
            "hbm_gbs_algorithmic": (self.h2d + self.d2h) / (step_ms * 1e-3) / 1e9 if step_ms else None,
        }  # fmt: skip

    def cpu_baseline(self):
        """Oracle port on the host cores (bounded sample) + parity of this run's GPU results against it."""
        wl = self.wl
        nsample, chunk = cpu_sample_shape(wl, self.numbers_h.shape[0])
        keep: list = []
        sec = cpu_time_sample(wl, self.numbers_h, self.positions_h.double(), self.q_h.double(), nsample, chunk,
                              eeq=self.eeq, keep=keep)  # fmt: skip
        e_gpu, g_gpu = self.step_resident()
        e_ref = torch.cat([k[0] for k in keep])
        scale = e_ref.abs().amax(-1, keepdim=True)
        parity = {"against": "float64 oracle, same sample", "structures": nsample,
                  "max_rel_energy": float(((e_gpu[:nsample].detach().double().cpu() - e_ref).abs() / scale).max())}  # fmt: skip
        if wl["grad"]:
            g_ref = torch.cat([k[1] for k in keep])
            parity["max_abs_gradient"] = float((g_gpu[:nsample].double().cpu() - g_ref).abs().max())
        return {"value": nsample / sec, "unit": "molecules/s", "cores": os.cpu_count(), "kind": "port",
                "sample": cpu_sample_text(nsample, chunk), "parity": parity}  # fmt: skip


def strong_record(ctx: Ctx, key: str, batch, steps, warmup):
    """The FIXED global batch of the workload, sharded over the ranks through the public
    ``parallel.dftd4_sharded`` (contiguous shards, no data-path collective)."""
    from tad_dftd4_b200.parallel import dftd4_sharded

    wl, dev = WORKLOADS[key], ctx.dev
    numbers, positions, q = (t.to(dev) for t in batch)
    if ctx.dist is not None:  # every rank evaluates shards of rank 0's batch
        for t in (numbers, positions, q):
            ctx.dist.broadcast(t, src=0)

    def step():
        if wl["grad"]:
            pos = positions.detach().requires_grad_(True)
            e = dftd4_sharded(numbers, pos, 0.0, PBE0, q=q, gather=False)
            (g,) = torch.autograd.grad(e.sum(), pos)
            return e, g
        return dftd4_sharded(numbers, positions, 0.0, PBE0, q=q, gather=False), None

    ms, _ = ctx.timed(step, steps, warmup)
    (ms,) = ctx.reduce([ms], "max")
    sec = ms / steps * 1e-3
    _, pairs, triples, flop = work_counts(batch[0], wl["grad"])
    return {"workload": wl["name"], "global_batch": int(numbers.shape[0]), "scaling": "strong",
            "api": "tad_dftd4_b200.parallel.dftd4_sharded(gather=False)",
            "value": numbers.shape[0] / sec, "unit": "molecules/s", "ms_per_step": sec * 1e3, "steps": steps,
            "pair_terms_per_s": float(pairs.sum()) / sec, "triple_terms_per_s": float(triples.sum()) / sec,
            "algorithmic_tflops": float(flop.sum()) / sec / 1e12}  # fmt: skip


def large_record(ctx: Ctx, key: str, cluster, steps, warmup, dtype_name="f64"):
    """C4: one large structure, STRONG scaling over the ranks (row-block + all-reduces)."""
    from tad_dftd4_b200 import large

    wl, dev = WORKLOADS[key], ctx.dev
    dtype = torch.float64 if dtype_name == "f64" else torch.float32
    numbers_h, positions_h, q_h = cluster
    nat = numbers_h.shape[0]
    numbers, positions, q = numbers_h.to(dev), positions_h.to(dev, dtype), q_h.to(dev, dtype)

    def step():
        if wl["grad"]:
            pos = positions.detach().requires_grad_(True)
            e = large.dftd4_large(numbers, pos, PBE0, q)
            (g,) = torch.autograd.grad(e.sum(), pos)
            return e.detach()
        return large.dftd4_large(numbers, positions, PBE0, q)

    for _ in range(max(1, warmup)):
        e = step()
    ctx.barrier()
    prof = large.profile_begin()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s in range(steps):
        ev[s][0].record()
        e = step()
        ev[s][1].record()
    ctx.barrier()
    large.profile_end()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    parts = prof.totals_ms()  # per-section device time on this rank, summed over the steps
    ms, ar_ms, plan_ms = ctx.reduce([ms, parts.get("all_reduce", 0.0), parts.get("plan", 0.0)], "max")
    sec = ms / steps * 1e-3
    # exact work counts in SURVEY 8(d)'s units: P_2 (r <= 60), T = sum_j C(n_j, 2) - 2 #closed
    counts = large.work_counts(numbers, positions.double())
    p2, trip, ctrip = counts["pairs_disp2"], counts["triples"], counts["centre_triples"]
    flop = (F_PCN * counts["pairs_cn"] + (F_P2 + (F_P2_G if wl["grad"] else 0)) * p2
            + (F_PCN_G * counts["pairs_cn"] if wl["grad"] else 0)
            + (F_T + (F_T_G if wl["grad"] else 0)) * trip + F_W * nat)  # fmt: skip
    peak_tf = C.c_double(0.0)
    from tad_dftd4_b200.disp import _Engine

    engine = _Engine.get(dev, 3.0, 2.0)
    scratch = torch.empty(64 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ctx.lib.d4b200_measure_fp64_peak(engine.handle, scratch.data_ptr(), scratch.numel(),
                                     torch.cuda.current_stream(dev).cuda_stream, C.byref(peak_tf))  # fmt: skip
    peak = peak_tf.value * (1 if dtype == torch.float64 else 2)
    tf = flop / sec / 1e12
    return {
        "workload": wl["name"], "atoms": nat, "scaling": "strong", "n_gpus": ctx.world,
        "value": nat / sec, "unit": "atoms/s", "ms_per_step": sec * 1e3, "steps": steps,
        "all_reduce_ms_per_step": ar_ms / steps, "plan_ms_per_step": plan_ms / steps,
        "collectives_per_step": 3 if wl["grad"] else 1,
        "pair_terms": p2, "triple_terms": trip, "centre_triple_terms": ctrip,
        "pair_terms_per_s": p2 / sec, "triple_terms_per_s": trip / sec,
        "algorithmic_tflops": tf, "per_gpu_frac_of_fp64_peak": tf / ctx.world / peak if peak else None,
        "peak_tflops": peak, "energy_sum": float(e.sum().item()),
        "reference": "N/A: the dense reference formulation cannot hold a 20k-atom system (BASELINE.md)",
    }  # fmt: skip


def run_b200(args, rank, world, local_rank):
    key = args.workload
    wl = WORKLOADS[key]
    subs = not args.no_subs and not wl.get("large") and not wl.get("single")
    # ---- all synthetic inputs first: the generator's worker pool runs before CUDA / NCCL exist
    batches = {}
    need = {key} | ({"c2", "c3"} if subs else set())
    for k in sorted(need - {"c4", "c4g"}):
        batches[k] = make_batch(WORKLOADS[k], rank)
    strong_batches = {}
    if subs:
        for k in ("c3", "c2"):
            strong_batches[k] = batches[k] if rank == 0 else tuple(torch.empty_like(t) for t in batches[k])
    if "c5" in need:
        batches["c5"] = batches.get("c3") or make_batch(WORKLOADS["c5"], rank)
    cluster = None
    if subs or wl.get("large"):
        cluster = bench_inputs.water_cluster(args.nmol or WORKLOADS["c4g"]["nmol"], WORKLOADS["c4g"]["seed"])

    ctx = Ctx(args, rank, world, local_rank)
    if wl.get("large"):
        rec = large_record(ctx, key, cluster, args.steps, min(args.warmup, 2), args.dtype)
        if rank == 0:
            line = {"metric": "D4 dispersion throughput, single large system (atoms/s)", "n_gpus": world,
                    "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "dtype": args.dtype,
                    "data": "synthetic", "config": {"workload": wl["name"], "atoms": rec["atoms"],
                                                     "parallelism": f"row-block x{world}, NCCL all-reduces"},
                    "e2e": None, "gpu_launches": None, "cpu_baseline": None, **rec}  # fmt: skip
            emit(line)
        return ctx

    head = BatchBench(ctx, key, args.dtype, batches[key], eeq=args.eeq)
    head.step_resident()  # builds tables / workspace
    with ClockSampler(local_rank) as clocks:
        res = head.measure(args.steps, args.warmup)
        # keep the same load running (untimed) until nvidia-smi has had time to sample it
        t_end = time.perf_counter() + 1.5
        while time.perf_counter() < t_end:
            for _ in range(20):
                head.step_resident()
            torch.cuda.synchronize(ctx.dev)
    clk = clocks.summary()
    roof = head.roofline(res["ms_per_step"] / 1.0)

    sub, strong, c4 = {}, {}, None
    if subs:
        ksteps = max(3, min(args.steps, 10))
        for name, k, dt in (("c2_f64", "c2", "f64"), ("c3_f64", "c3", "f64"), ("c5_f64", "c5", "f64"), ("c5_f32", "c5", "f32")):
            if k == key and dt == args.dtype and not args.eeq:
                continue  # that is the headline
            bb = BatchBench(ctx, k, dt, batches.get(k) or batches["c3"])
            bb.step_resident()
            r = bb.measure(ksteps, 3)
            rf = bb.roofline(r["ms_per_step"], nsteps=4)
            r["workload"] = WORKLOADS[k]["name"]
            r["dtype"] = dt
            r["roofline"] = {kk: rf[kk] for kk in ("bound", "kernel", "achieved", "peak", "unit", "frac", "kernel_ms",
                                                   "all_class_ms", "whole_step_frac")}  # fmt: skip
            sub[name] = r
            del bb
        for k in ("c3", "c2"):
            strong[k] = strong_record(ctx, k, strong_batches[k], ksteps, 3)
        c4 = large_record(ctx, "c4g", cluster, 2, 1)

    cpu = None
    if not args.no_cpu:
        torch.cuda.synchronize(ctx.dev)
        if rank == 0:  # rank 0 has the host cores to itself; the other ranks sleep in the gloo barrier below
            cpu = head.cpu_baseline()
        if ctx.dist is not None:
            ctx.dist.barrier(group=ctx.wait_group)

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": res["value"], "unit": "molecules/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": config_dict(wl, world, args.dtype, args.eeq),
            "pair_terms_per_s": res["pair_terms_per_s"], "triple_terms_per_s": res["triple_terms_per_s"],
            "algorithmic_tflops": res["algorithmic_tflops"],
            "e2e": res["e2e"],
            "gpu_launches": res["gpu_launches"],
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": clk,
            "strong": strong or None,
            "sub": sub or None,
            "c4": c4,
            "native_library": str(Path(ctx.lib._name).resolve().relative_to(ROOT)) if hasattr(ctx.lib, "_name") else None,
        }  # fmt: skip
        emit(line)
    return ctx


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="development: skip the host-buffer leg (A/B of older library builds)")
    ap.add_argument("--no-subs", action="store_true",
                    help="headline workload only (no strong-scaling / C2 / C5 / C4 sub-records)")
    ap.add_argument("--eeq", action="store_true",
                    help="default q=None path: EEQ charges on device inside the step (and on the autograd tape)")
    ap.add_argument("--nmol", type=int, default=0, help="c4: number of water molecules (default 6667)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: whatever libraries write to file descriptor 1 (NCCL's
    # version banner at communicator creation, ...) is routed to stderr for the whole run
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, WORKLOADS[args.workload], rank, world)
        return
    ctx = run_b200(args, rank, world, local_rank)
    # orderly teardown: device idle, process group destroyed, then normal interpreter exit (the
    # driver's exit-time hook records which shared libraries this process mapped)
    torch.cuda.synchronize(ctx.dev)
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


def leave() -> None:
    """Let the autograd worker threads go idle, then exit the interpreter normally.

    A plain exit right after the last backward pass aborted now and then ("terminate called without an
    active exception", exit code 134 after the JSON line was out; 2 of 22 runs in
    ``profiles/r02_exit_probe.txt``).  The C++ stack of the aborting thread
    (``tools/probes/abort_trace.c``) names the cause: an autograd ENGINE worker thread
    (``torch::autograd::Engine::thread_main``) is still releasing the buffers of the last graph task --
    tensors created in a Python ``autograd.Function.backward`` carry a Python object, so dropping them
    needs the GIL (``TensorImpl::decref_pyobject`` -> ``PyEval_RestoreThread``) -- while the main thread,
    which was woken as soon as the last task completed, runs on to ``Py_Finalize`` holding the GIL.  A
    thread that asks for the GIL of a finalizing interpreter is ended with ``pthread_exit``; its forced
    unwind crosses a ``noexcept`` C++ frame, hence ``std::terminate``.  Giving the GIL away for a moment
    after the last backward lets the worker finish and park on its (C++) ready queue, where finalization
    does not touch it.  No ``os._exit``: every exit-time hook, static destructor and library finalizer
    runs (``D4_BENCH_EXIT=hard`` keeps the old behaviour for diagnosis)."""
    import gc
    import time

    sys.stdout.flush()
    sys.stderr.flush()
    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_available():
        torch.cuda.synchronize()
    gc.collect()
    time.sleep(0.25)  # sleeping releases the GIL
    if os.environ.get("D4_BENCH_EXIT", "normal") != "hard":
        return
    import atexit

    try:
        threading._shutdown()
    except Exception:  # noqa: BLE001
        pass
    atexit._run_exitfuncs()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    import faulthandler

    faulthandler.enable()  # a fatal signal during teardown leaves the Python stacks on stderr
    if os.environ.get("D4_BENCH_ABORT_TRACE"):  # diagnosis: tools/probes/abort_trace.c (C++ stack on SIGABRT)
        import ctypes

        ctypes.CDLL(os.environ["D4_BENCH_ABORT_TRACE"]).abort_trace_install()
    main()
    leave()
