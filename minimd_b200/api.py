"""numpy-facing wrapper over the C ABI: one `Context` = one `mmd_ctx` (one rank on one GPU).

This is a binding, not an implementation: every method is one C-ABI call.  Host arrays use the
reference's layout (AoS with stride PAD, MMD_float = float32|float64).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BinGeometry, RunParams, SwapTable, ThermoSample, check


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    def __init__(self, precision: str = "f64", ntypes: int = 4, device: int = 0, stream: int | None = None):
        self.lib = _lib.load()
        self.real = np.float64 if precision == "f64" else np.float32
        self.precision = precision
        self.ntypes = ntypes
        h = C.c_void_p()
        check(self.lib.mmd_ctx_create(device, np.dtype(self.real).itemsize, ntypes, C.c_void_p(stream or 0), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            if not getattr(self, "_borrowed", False):
                self.lib.mmd_ctx_destroy(self.h)
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

    # ---- misc -------------------------------------------------------------------------
    def sync(self):
        check(self.lib.mmd_ctx_sync(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.mmd_ctx_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(self.lib.mmd_ctx_launches(self.h))

    def query(self, key: str) -> int:
        v = C.c_longlong()
        check(self.lib.mmd_query_int(self.h, key.encode(), C.byref(v)))
        return int(v.value)

    def set_option(self, key: str, value: int):
        check(self.lib.mmd_set_option(self.h, key.encode(), int(value)))

    def _real(self, a):
        return np.ascontiguousarray(a, dtype=self.real)

    # ---- Atom -------------------------------------------------------------------------
    def set_box(self, prd, lo=None, hi=None):
        prd = np.asarray(prd, dtype=np.float64)
        lo = np.zeros(3) if lo is None else np.asarray(lo, dtype=np.float64)
        hi = prd.copy() if hi is None else np.asarray(hi, dtype=np.float64)
        dp = C.POINTER(C.c_double)
        check(self.lib.mmd_atom_set_box(self.h, prd.ctypes.data_as(dp), lo.ctypes.data_as(dp), hi.ctypes.data_as(dp)))

    def upload(self, x, v, type_=None):
        x = self._real(x)
        v = self._real(v)
        n, pad = x.shape
        t = None if type_ is None else np.ascontiguousarray(type_, dtype=np.int32)
        check(self.lib.mmd_atom_upload(self.h, _ptr(x), _ptr(v), _ptr(t), n, pad))

    def split(self, nlocal: int):
        """The last (uploaded - nlocal) atoms of the preceding upload are ghosts (host-side Comm::borders)."""
        check(self.lib.mmd_atom_split(self.h, int(nlocal)))

    def update(self, x=None, v=None, first=0):
        x = None if x is None else self._real(x)
        v = None if v is None else self._real(v)
        ref = x if x is not None else v
        n, pad = ref.shape
        check(self.lib.mmd_atom_update(self.h, _ptr(x), _ptr(v), first, n, pad))

    def counts(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        check(self.lib.mmd_atom_counts(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    @property
    def nlocal(self):
        return self.counts()[0]

    @property
    def nghost(self):
        return self.counts()[1]

    def download(self, what: str = "xvft", first: int = 0, count: int | None = None, pad: int = 3):
        """Returns a dict with any of x, v, f (count x pad) and type (count)."""
        nl, ng, _ = self.counts()
        if count is None:
            count = (nl if "v" in what else nl + ng) - first
        out = {}
        for k in "xvf":
            if k in what:
                out[k] = np.zeros((count, pad), dtype=self.real)
        if "t" in what:
            out["type"] = np.zeros(count, dtype=np.int32)
        check(self.lib.mmd_atom_download(self.h, _ptr(out.get("x")), _ptr(out.get("v")), _ptr(out.get("f")),
                                         _ptr(out.get("type")), first, count, pad))
        return out

    def pbc(self):
        check(self.lib.mmd_atom_pbc(self.h))

    def sort(self):
        check(self.lib.mmd_atom_sort(self.h))

    # ---- Neighbor ---------------------------------------------------------------------
    def neigh_setup(self, geo: dict, stencil, cutneighsq):
        g = BinGeometry(**{k: (int(v) if not k.startswith("bininv") else float(v)) for k, v in geo.items()})
        st = np.ascontiguousarray(stencil, dtype=np.int32)
        cs = self._real(cutneighsq)
        assert cs.size == self.ntypes * self.ntypes
        check(self.lib.mmd_neigh_setup(self.h, C.byref(g), _ptr(st), st.size, _ptr(cs)))
        self.mbins = g.mbinx * g.mbiny * g.mbinz

    def binatoms(self, count: int = -1, atoms_per_bin: int = 8):
        apb, mx = C.c_int(atoms_per_bin), C.c_int()
        check(self.lib.mmd_neigh_binatoms(self.h, count, C.byref(apb), C.byref(mx)))
        return apb.value, mx.value

    def build(self, halfneigh: int, ghost_newton: int, maxneighs: int = 100):
        mx, tot = C.c_int(maxneighs), C.c_longlong()
        check(self.lib.mmd_neigh_build(self.h, halfneigh, ghost_newton, C.byref(mx), C.byref(tot)))
        return mx.value, tot.value

    def neigh_download(self, nrows: int | None = None, maxneighs: int | None = None, lists: bool = True):
        nrows = self.nlocal if nrows is None else nrows
        maxneighs = self.query("maxneighs") if maxneighs is None else maxneighs
        num = np.zeros(nrows, dtype=np.int32)
        nb = np.zeros((nrows, maxneighs), dtype=np.int32) if lists else None
        check(self.lib.mmd_neigh_download(self.h, _ptr(num), _ptr(nb), nrows, maxneighs))
        return num, nb

    def neigh_upload(self, numneigh, neighbors):
        num = np.ascontiguousarray(numneigh, dtype=np.int32)
        nb = np.ascontiguousarray(neighbors, dtype=np.int32)
        check(self.lib.mmd_neigh_upload(self.h, _ptr(num), _ptr(nb), nb.shape[0], nb.shape[1]))

    def bins_download(self, atoms_per_bin: int):
        cnt = np.zeros(self.mbins, dtype=np.int32)
        rows = np.zeros((self.mbins, atoms_per_bin), dtype=np.int32)
        check(self.lib.mmd_neigh_bins_download(self.h, _ptr(cnt), _ptr(rows), atoms_per_bin))
        return cnt, rows

    def atom_bins(self, count: int):
        b = np.zeros(count, dtype=np.int32)
        check(self.lib.mmd_neigh_atom_bins_download(self.h, _ptr(b), count))
        return b

    # ---- Force ------------------------------------------------------------------------
    def lj_setup(self, cutforcesq, sigma6, epsilon):
        a, b, c = self._real(cutforcesq), self._real(sigma6), self._real(epsilon)
        check(self.lib.mmd_force_lj_setup(self.h, _ptr(a), _ptr(b), _ptr(c)))

    def lj_compute(self, halfneigh: int, ghost_newton: int, evflag: int):
        e, v = np.zeros(1, dtype=self.real), np.zeros(1, dtype=self.real)
        check(self.lib.mmd_force_lj_compute(self.h, halfneigh, ghost_newton, evflag, _ptr(e), _ptr(v)))
        return float(e[0]), float(v[0])

    def eam_setup(self, rhor_spline, z2r_spline, frho_spline, nr, nrho, nr_tot, nrho_tot, rdr, rdrho, cutforcesq):
        a, b, c, d = (self._real(t) for t in (rhor_spline, z2r_spline, frho_spline, cutforcesq))
        check(self.lib.mmd_force_eam_setup(self.h, _ptr(a), _ptr(b), _ptr(c), nr, nrho, nr_tot, nrho_tot,
                                           float(rdr), float(rdrho), _ptr(d)))

    def eam_compute(self, halfneigh: int, evflag: int):
        e, v = np.zeros(1, dtype=self.real), np.zeros(1, dtype=self.real)
        check(self.lib.mmd_force_eam_compute(self.h, halfneigh, evflag, _ptr(e), _ptr(v)))
        return float(e[0]), float(v[0])

    # ---- Integrate / Thermo -----------------------------------------------------------
    def initial_integrate(self, dt: float, dtforce: float):
        check(self.lib.mmd_integrate_initial(self.h, dt, dtforce))

    def final_integrate(self, dtforce: float):
        check(self.lib.mmd_integrate_final(self.h, dtforce))

    def sum_mv2(self, mass: float) -> float:
        v = C.c_double()
        check(self.lib.mmd_thermo_sum_mv2(self.h, mass, C.byref(v)))
        return v.value

    # ---- Comm -------------------------------------------------------------------------
    def comm_setup(self, table: SwapTable):
        check(self.lib.mmd_comm_setup(self.h, C.byref(table)))
        self.nswap = table.nswap

    def exchange(self):
        check(self.lib.mmd_comm_exchange(self.h))

    def borders(self):
        check(self.lib.mmd_comm_borders(self.h))

    def communicate(self):
        check(self.lib.mmd_comm_communicate(self.h))

    def reverse_communicate(self):
        check(self.lib.mmd_comm_reverse_communicate(self.h))

    def swap_counts(self):
        n = self.nswap
        a, b, c = (np.zeros(n, dtype=np.int32) for _ in range(3))
        ip = C.POINTER(C.c_int)
        check(self.lib.mmd_comm_swap_counts(self.h, a.ctypes.data_as(ip), b.ctypes.data_as(ip), c.ctypes.data_as(ip)))
        return a, b, c

    def sendlist(self, iswap: int):
        n = int(self.swap_counts()[0][iswap])
        out = np.zeros(n, dtype=np.int32)
        if n:
            check(self.lib.mmd_comm_sendlist_download(self.h, iswap, _ptr(out), n))
        return out

    def nccl_init(self, id128: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(id128, 128)
        check(self.lib.mmd_comm_nccl_init(self.h, buf, rank, nranks))

    # ---- time loop ----------------------------------------------------------------------
    def run(self, params: RunParams, max_samples: int = 64, timed: bool = True):
        samples = (ThermoSample * max_samples)()
        ns = C.c_int()
        ms = C.c_float()
        check(self.lib.mmd_run(self.h, C.byref(params), samples, max_samples, C.byref(ns),
                               C.byref(ms) if timed else None))
        out = [(s.step, s.sum_mv2, s.eng_vdwl, s.virial) for s in samples[:min(ns.value, max_samples)]]
        return out, (ms.value if timed else None)


    def phase_times(self, reset: bool = False):
        """{phase: (ms, intervals)} accumulated by run() while option phase_timing is on."""
        ms = (C.c_double * 5)()
        calls = (C.c_longlong * 5)()
        check(self.lib.mmd_run_phase_times(self.h, ms, calls, int(reset)))
        names = ("integrate", "comm", "neigh", "force", "other")
        return {n: (ms[i], calls[i]) for i, n in enumerate(names)}


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(_lib.load().mmd_comm_nccl_unique_id(buf))
    return buf.raw
