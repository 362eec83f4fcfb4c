"""ctypes binding of ``libd4b200.so`` (C ABI: ``include/d4b200.h``).

There is no CPU or PyTorch fallback: if the library is missing or the device
is not a B200 the import / first call fails loudly.
"""

from __future__ import annotations

import ctypes as C
import atexit
import os
import threading
import time
from pathlib import Path

# D4B200_LIBRARY points at an alternative build of the same C ABI (A/B timing of kernel variants)
LIB_PATH = Path(os.environ.get("D4B200_LIBRARY") or Path(__file__).resolve().parent / "libd4b200.so").resolve()


class D4B200Error(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [
        ("s6", C.c_double), ("s8", C.c_double), ("s9", C.c_double), ("s10", C.c_double),
        ("a1", C.c_double), ("a2", C.c_double), ("alp", C.c_double),
        ("disp2_cutoff", C.c_double), ("disp3_cutoff", C.c_double), ("cn_cutoff", C.c_double),
        ("wf", C.c_double), ("has_s10", C.c_int32), ("model", C.c_int32),
    ]  # fmt: skip


_lib = None

_VP = C.c_void_p
_SIGS = {
    "d4b200_version": (C.c_int, []),
    "d4b200_error_string": (C.c_char_p, [C.c_int]),
    "d4b200_tables_create": (C.c_int, [C.c_int, _VP, C.c_size_t, _VP, C.c_size_t, C.c_double, C.c_double, C.POINTER(_VP)]),
    "d4b200_tables_destroy": (C.c_int, [_VP]),
    "d4b200_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "d4b200_energy_f64": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_energy_f32": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_energy_host_f64": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, C.c_int]),
    "d4b200_energy_host_f32": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, C.c_int]),
    "d4b200_energy_gradient_host_f64": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP, C.c_int]),
    "d4b200_energy_gradient_host_f32": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP, C.c_int]),
    "d4b200_energy_host_z_f64": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, C.c_int, _VP, _VP, _VP, _VP, _VP, C.c_int, C.POINTER(C.c_int)]),
    "d4b200_energy_host_z_f32": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, C.c_int, _VP, _VP, _VP, _VP, _VP, C.c_int, C.POINTER(C.c_int)]),
    "d4b200_gradient_f64": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_gradient_f32": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_energy_gradient_f64": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_energy_gradient_f32": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_properties_f64": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_properties_f32": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_large_group_size": (C.c_int, []),
    "d4b200_large_workspace_bytes": (C.c_size_t, [_VP, C.c_int, C.c_int]),
    "d4b200_large_energy_f64": (C.c_int, [_VP, C.POINTER(Params), C.c_int, _VP, _VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_large_energy_f32": (C.c_int, [_VP, C.POINTER(Params), C.c_int, _VP, _VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_large_gradient_f64": (C.c_int, [_VP, C.POINTER(Params), C.c_int, _VP, _VP, _VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_large_gradient_f32": (C.c_int, [_VP, C.POINTER(Params), C.c_int, _VP, _VP, _VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_large_energy_gradient_f64": (C.c_int, [_VP, C.POINTER(Params), C.c_int, _VP, _VP, _VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_large_energy_gradient_f32": (C.c_int, [_VP, C.POINTER(Params), C.c_int, _VP, _VP, _VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_large_cn_chain_f64": (C.c_int, [_VP, C.POINTER(Params), C.c_int, _VP, _VP, _VP, C.c_int, C.c_int, _VP, _VP]),
    "d4b200_large_cn_chain_f32": (C.c_int, [_VP, C.POINTER(Params), C.c_int, _VP, _VP, _VP, C.c_int, C.c_int, _VP, _VP]),
    "d4b200_weight_references_f64": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "d4b200_weight_references_f32": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "d4b200_atomic_c6_f64": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP, _VP]),
    "d4b200_atomic_c6_f32": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP, _VP]),
    "d4b200_weighted_pols_f64": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP, _VP]),
    "d4b200_weighted_pols_f32": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, _VP, _VP, _VP, _VP]),
    "d4b200_param_vjp_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "d4b200_param_vjp_f64": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_param_vjp_f32": (C.c_int, [_VP, C.POINTER(Params), C.c_int, C.c_int, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_size_t, _VP]),
    "d4b200_eeq_create": (C.c_int, [C.c_int, _VP, C.c_size_t, C.POINTER(_VP)]),
    "d4b200_eeq_destroy": (C.c_int, [_VP]),
    "d4b200_eeq_limit": (C.c_int, []),
    "d4b200_eeq_launch_count": (C.c_longlong, []),
    "d4b200_eeq_charges_f64": (C.c_int, [_VP, C.c_int, C.c_int, _VP, _VP, _VP, C.c_double, _VP, _VP, _VP]),
    "d4b200_eeq_charges_f32": (C.c_int, [_VP, C.c_int, C.c_int, _VP, _VP, _VP, C.c_double, _VP, _VP, _VP]),
    "d4b200_eeq_vjp_f64": (C.c_int, [_VP, C.c_int, C.c_int, _VP, _VP, C.c_double, _VP, _VP, _VP, _VP, _VP]),
    "d4b200_eeq_vjp_f32": (C.c_int, [_VP, C.c_int, C.c_int, _VP, _VP, C.c_double, _VP, _VP, _VP, _VP, _VP]),
    "d4b200_eeq_factor_doubles": (C.c_size_t, [C.c_int]),
    "d4b200_eeq_charges_factor_f64": (C.c_int, [_VP, C.c_int, C.c_int, _VP, _VP, _VP, C.c_double, _VP, _VP, _VP, _VP]),
    "d4b200_eeq_charges_factor_f32": (C.c_int, [_VP, C.c_int, C.c_int, _VP, _VP, _VP, C.c_double, _VP, _VP, _VP, _VP]),
    "d4b200_eeq_vjp_factor_f64": (C.c_int, [_VP, C.c_int, C.c_int, _VP, _VP, C.c_double, _VP, _VP, _VP, _VP, _VP, _VP]),
    "d4b200_eeq_vjp_factor_f32": (C.c_int, [_VP, C.c_int, C.c_int, _VP, _VP, C.c_double, _VP, _VP, _VP, _VP, _VP, _VP]),
    "d4b200_status": (C.c_int, [_VP, _VP, C.POINTER(C.c_int)]),
    "d4b200_last_launch_count": (C.c_int, []),
    "d4b200_total_launch_count": (C.c_longlong, []),
    "d4b200_profile_enable": (C.c_int, [_VP, C.c_int]),
    "d4b200_profile_read": (C.c_int, [_VP, C.POINTER(C.c_float)]),
    "d4b200_class_caps": (C.c_int, [_VP, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "d4b200_class_caps_model": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "d4b200_small_limit": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int]),
    "d4b200_phase_profile": (C.c_int, [_VP, C.c_int, C.POINTER(C.c_ulonglong)]),
    "d4b200_measure_fp64_peak": (C.c_int, [_VP, _VP, C.c_size_t, _VP, C.POINTER(C.c_double)]),
}  # fmt: skip

EXPORTED_SYMBOLS = tuple(_SIGS)


def load() -> C.CDLL:
    """Load the shared library (once) and declare the signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.is_file():
        raise D4B200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  tad_dftd4_b200 has no CPU / PyTorch fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name, None)
        if fn is None:
            if os.environ.get("D4B200_LIBRARY"):  # an older build of the ABI, loaded on purpose for A/B timing
                continue
            raise D4B200Error(f"{LIB_PATH} does not export {name}: rebuild it (python -c 'import __graft_entry__ as g; g.build()')")
        fn.restype = res
        fn.argtypes = args
        if name not in _LOCK_FREE:
            setattr(lib, name, _serialised(fn))
    _lib = lib
    atexit.register(_let_autograd_workers_finish)
    return lib


# Entry points that only read constants: no lock.
_LOCK_FREE = frozenset({"d4b200_version", "d4b200_error_string", "d4b200_eeq_limit", "d4b200_large_group_size",
                        "d4b200_workspace_bytes", "d4b200_eeq_factor_doubles"})  # fmt: skip
_CALL_LOCK = threading.RLock()


def _serialised(fn):
    """One library call at a time per process.  A tables handle owns ONE set of fork / join events and class streams
    (csrc/d4b200_handle.cuh); ctypes releases the GIL during a call, so two Python threads could otherwise interleave
    their event records and stream waits on the same handle.  Calls only ENQUEUE work (a few microseconds), so the lock
    costs nothing measurable; calls from one thread on different CUDA streams were never a problem (workspaces are kept
    per stream, `disp._Engine`).  C callers: serialise the calls on one handle, or use one handle per thread."""

    def call(*args):
        with _CALL_LOCK:
            return fn(*args)

    call.__name__ = getattr(fn, "__name__", "d4b200_call")
    call.__wrapped__ = fn
    return call


def _let_autograd_workers_finish() -> None:
    """Exit hook: give the GIL away for a moment before the interpreter finalizes.

    The gradients come out of Python ``autograd.Function.backward`` methods, so the tensors the autograd
    engine's worker thread drops after the last task of a backward pass carry Python objects, and
    dropping them needs the GIL.  The caller of ``backward`` is woken before that; if it goes straight
    on to interpreter exit, the worker asks for the GIL of a finalizing interpreter, is ended with
    ``pthread_exit`` and the forced unwind through a ``noexcept`` frame calls ``std::terminate``
    (exit code 134 after all work is done; stack in ``profiles/r02_exit_probe.txt``).  ``atexit`` hooks run
    before the interpreter is marked as finalizing, and sleeping releases the GIL."""
    time.sleep(0.05)


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().d4b200_error_string(code).decode()
        raise D4B200Error(f"{what} failed with code {code}: {msg}")
