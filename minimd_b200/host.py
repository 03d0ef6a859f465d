"""ctypes binding of the host layer (include/minimd_host.h): a whole miniMD run as an object.

`Simulation(args, precision)` is the reference's main() up to the step-0 thermo record;
`run()` is Integrate::run; `finish()` prints PERF_SUMMARY.  `Simulation.plan(...)` runs only the
host-side setup (no GPU needed) so that decomposition, bins, stencil, tables and the synthetic
atoms can be inspected anywhere.  This module holds no numerics.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import BinGeometry, RunParams, SwapTable

PKG = os.path.dirname(os.path.abspath(__file__))
INPUTS = os.path.join(os.path.dirname(PKG), "inputs")

REDUCE_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_void_p)

HOST_SIGNATURES = {
    "mmd_sim_precision_bytes": (C.c_int, []),
    "mmd_sim_last_error": (C.c_char_p, []),
    "mmd_sim_create": (C.c_int, [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                 C.POINTER(C.c_void_p)]),
    "mmd_sim_plan": (C.c_int, [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int, REDUCE_FN, C.c_void_p,
                               C.POINTER(C.c_void_p)]),
    "mmd_sim_host_array": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]),
    "mmd_sim_bin_geometry": (C.c_int, [C.c_void_p, C.POINTER(BinGeometry)]),
    "mmd_sim_swap_table": (C.c_int, [C.c_void_p, C.POINTER(SwapTable)]),
    "mmd_sim_run": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "mmd_sim_finish": (C.c_int, [C.c_void_p]),
    "mmd_sim_destroy": (C.c_int, [C.c_void_p]),
    "mmd_sim_thermo": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mmd_sim_get_int": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_longlong)]),
    "mmd_sim_get_real": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_double)]),
    "mmd_sim_ctx": (C.c_void_p, [C.c_void_p]),
    "mmd_sim_run_params": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(RunParams)]),
}

_host: dict[str, C.CDLL] = {}


def host_lib_path(precision: str) -> str:
    return os.path.join(PKG, "lib", f"libminimd_host_{precision}.so")


def load_host(precision: str = "f64") -> C.CDLL:
    if precision not in _host:
        _lib.load()  # the device library first (RTLD_GLOBAL not needed: the host lib links it by rpath)
        path = host_lib_path(precision)
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `python -m minimd_b200.build`")
        lib = C.CDLL(path)
        for name, (res, args) in HOST_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        assert lib.mmd_sim_precision_bytes() == (8 if precision == "f64" else 4)
        _host[precision] = lib
    return _host[precision]


class HostError(RuntimeError):
    pass


def input_file(name: str) -> str:
    """Path of a shipped input deck (inputs/in.lj.miniMD, inputs/in.eam.miniMD)."""
    return os.path.join(INPUTS, name)


class Simulation:
    """args: the reference's command line as a list of strings (without argv[0])."""

    def __init__(self, args, precision: str = "f64", rank: int = 0, nranks: int = 1, device: int = -1,
                 nccl_id: bytes | None = None, _plan=False, reduce=None):
        self.lib = load_host(precision)
        self.precision = precision
        self.real = np.float64 if precision == "f64" else np.float32
        self.args = [str(a) for a in args]
        argv = (C.c_char_p * len(self.args))(*[a.encode() for a in self.args])
        h = C.c_void_p()
        if _plan:
            self._cb = REDUCE_FN(reduce) if reduce is not None else C.cast(None, REDUCE_FN)
            rc = self.lib.mmd_sim_plan(len(self.args), argv, rank, nranks, self._cb, None, C.byref(h))
        else:
            idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id else None
            rc = self.lib.mmd_sim_create(len(self.args), argv, rank, nranks, device, idbuf, C.byref(h))
        if rc:
            raise HostError(self.lib.mmd_sim_last_error().decode(errors="replace"))
        self.h = h
        self.planned = _plan

    @classmethod
    def plan(cls, args, precision: str = "f64", rank: int = 0, nranks: int = 1, reduce=None) -> "Simulation":
        """Host-only setup (no GPU).  reduce(values: np.ndarray, op: int) must reduce IN PLACE over ranks."""
        cb = None
        if reduce is not None:
            def cb(ptr, n, op, _user):  # noqa: ANN001
                reduce(np.ctypeslib.as_array(ptr, shape=(n,)), op)
        return cls(args, precision, rank, nranks, _plan=True, reduce=cb)

    def close(self):
        if getattr(self, "h", None):
            self.lib.mmd_sim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc:
            raise HostError(self.lib.mmd_sim_last_error().decode(errors="replace"))

    # ---- scalars / arrays ---------------------------------------------------------------
    def geti(self, key: str) -> int:
        v = C.c_longlong()
        self._check(self.lib.mmd_sim_get_int(self.h, key.encode(), C.byref(v)))
        return int(v.value)

    def getr(self, key: str) -> float:
        v = C.c_double()
        self._check(self.lib.mmd_sim_get_real(self.h, key.encode(), C.byref(v)))
        return float(v.value)

    def host_array(self, name: str) -> np.ndarray:
        p, n = C.c_void_p(), C.c_longlong()
        self._check(self.lib.mmd_sim_host_array(self.h, name.encode(), C.byref(p), C.byref(n)))
        dtype = np.int32 if name in ("type", "stencil") else self.real
        if not p.value or n.value == 0:
            return np.zeros(0, dtype=dtype)
        buf = (C.c_char * (np.dtype(dtype).itemsize * n.value)).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype, count=n.value).copy()

    def bin_geometry(self) -> BinGeometry:
        g = BinGeometry()
        self._check(self.lib.mmd_sim_bin_geometry(self.h, C.byref(g)))
        return g

    def swap_table(self) -> SwapTable:
        t = SwapTable()
        self._check(self.lib.mmd_sim_swap_table(self.h, C.byref(t)))
        return t

    # ---- running ------------------------------------------------------------------------
    def run(self, nsteps: int = -1) -> float:
        """Integrate::run; returns the CUDA-event time of the loop in ms."""
        ms = C.c_double()
        self._check(self.lib.mmd_sim_run(self.h, nsteps, C.byref(ms)))
        return ms.value

    def finish(self):
        self._check(self.lib.mmd_sim_finish(self.h))

    def thermo(self):
        n = self.lib.mmd_sim_thermo(self.h, 0, None, None, None, None)
        st = (C.c_int * max(n, 1))()
        T, U, P = ((C.c_double * max(n, 1))() for _ in range(3))
        self.lib.mmd_sim_thermo(self.h, n, st, T, U, P)
        return list(st[:n]), list(T[:n]), list(U[:n]), list(P[:n])

    def run_params(self, nsteps: int) -> RunParams:
        p = RunParams()
        self._check(self.lib.mmd_sim_run_params(self.h, nsteps, C.byref(p)))
        return p

    def context(self):
        """The simulation's device context as an api.Context view (not owning)."""
        from .api import Context
        c = Context.__new__(Context)
        c.lib = _lib.load()
        c.real = self.real
        c.precision = self.precision
        c.ntypes = self.geti("ntypes")
        c.h = C.c_void_p(self.lib.mmd_sim_ctx(self.h))
        c.nswap = self.geti("nswap")
        c.mbins = self.geti("mbins")
        c._borrowed = True
        return c
