"""ctypes wrapper around the plain-C oracle (oracle/minimd_oracle.c) and a runner for the
unmodified reference binary in oracle/_ref/.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under minimd_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
import tempfile
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def build(force: bool = False) -> None:
    """Compile the oracle shared objects (and oracle/_ref when /root/reference exists)."""
    need = force or not all(
        os.path.exists(os.path.join(HERE, f"libminimd_oracle_{p}.so")) for p in ("f64", "f32")
    )
    src = os.path.join(HERE, "minimd_oracle.c")
    if not need:
        need = any(
            os.path.getmtime(src) > os.path.getmtime(os.path.join(HERE, f"libminimd_oracle_{p}.so"))
            for p in ("f64", "f32")
        )
    if need:
        subprocess.check_call(["make", "-C", HERE, "libminimd_oracle_f64.so", "libminimd_oracle_f32.so"],
                              stdout=subprocess.DEVNULL)
    if os.path.isdir(os.environ.get("MINIMD_REFERENCE", "/root/reference")):
        subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
        # the drop-in proof: the reference program with one translation unit replaced by a C-ABI caller (tests/dropin/);
        # rebuilt when the library or the replacement units are newer than the binaries
        lib = os.path.join(os.path.dirname(HERE), "minimd_b200", "lib", "libminimd_b200.so")
        outs = [os.path.join(REF_DIR, n) for n in ("miniMD_dropin_run_f64", "miniMD_dropin_force_f64")]
        srcs = [lib] + [os.path.join(os.path.dirname(HERE), "tests", "dropin", f) for f in ("integrate_b200.cpp", "force_lj_b200.cpp")]
        if os.path.exists(lib) and (force or not all(os.path.exists(o) for o in outs) or
                                    max(os.path.getmtime(s_) for s_ in srcs[1:]) > min(os.path.getmtime(o) for o in outs)):
            subprocess.check_call(["bash", os.path.join(HERE, "build_dropin.sh")], stdout=subprocess.DEVNULL)


class OrcConfig(C.Structure):
    _fields_ = [
        ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
        ("ntimes", C.c_int),
        ("forcetype", C.c_int), ("units", C.c_int),
        ("halfneigh", C.c_int), ("ghost_newton", C.c_int),
        ("neigh_every", C.c_int), ("sort", C.c_int),
        ("thermo_nstat", C.c_int), ("nbin_override", C.c_int),
        ("epsilon", C.c_double), ("sigma", C.c_double),
        ("dt", C.c_double), ("t_request", C.c_double), ("rho", C.c_double),
        ("force_cut", C.c_double), ("neigh_cut", C.c_double),
    ]


@dataclass
class Config:
    """The 14-line input file + CLI overrides of ref/ljs.cpp as one record."""
    nx: int = 8
    ny: int = 8
    nz: int = 8
    ntimes: int = 100
    force: str = "lj"          # "lj" | "eam"
    halfneigh: int = 1
    ghost_newton: int = 1
    neigh_every: int = 20
    sort: int = -1
    thermo_nstat: int = 100
    nbins: int = -1
    epsilon: float = 1.0
    sigma: float = 1.0
    dt: float | None = None
    t_request: float | None = None
    rho: float | None = None
    force_cut: float | None = None
    skin: float | None = None
    ntypes: int = 4

    def resolved(self) -> "Config":
        d = dict(self.__dict__)
        lj = self.force == "lj"
        dflt = dict(dt=0.005 if lj else 0.001, t_request=1.44 if lj else 600.0,
                    rho=0.8442 if lj else 0.07041125, force_cut=2.5 if lj else 4.95,
                    skin=0.30 if lj else 1.00)
        for k, v in dflt.items():
            if d[k] is None:
                d[k] = v
        return Config(**d)

    @property
    def units(self) -> str:
        return "lj" if self.force == "lj" else "metal"

    def input_text(self) -> str:
        """Render as the positional in.*.miniMD format (ref/input.cpp:124-176)."""
        c = self.resolved()
        return "\n".join([
            "generated miniMD input", "",
            c.units, "none", c.force, f"{c.epsilon!r} {c.sigma!r}",
            f"{c.nx} {c.ny} {c.nz}", f"{c.ntimes}", f"{c.dt!r}", f"{c.t_request!r}", f"{c.rho!r}",
            f"{c.neigh_every}", f"{c.force_cut!r} {c.skin!r}", f"{c.thermo_nstat}", ""])

    def as_struct(self, real=np.float64) -> OrcConfig:
        c = self.resolved()
        # in.neigh_cut += in.force_cut is evaluated in MMD_float (ref/input.cpp:183)
        neigh_cut = float(real(c.skin) + real(c.force_cut))
        # struct In holds MMD_float members (ref/ljs.h:37-51): every value is rounded to the
        # build's precision by sscanf before main() widens it again for create_box & co.
        r = lambda v: float(real(v))
        return OrcConfig(c.nx, c.ny, c.nz, c.ntimes, 0 if c.force == "lj" else 1,
                         0 if c.force == "lj" else 1, c.halfneigh, c.ghost_newton, c.neigh_every,
                         c.sort, c.thermo_nstat, c.nbins, r(c.epsilon), r(c.sigma), r(c.dt),
                         r(c.t_request), r(c.rho), r(c.force_cut), neigh_cut)


_libs: dict[str, C.CDLL] = {}
_libc = C.CDLL(None)


def _lib(precision: str) -> C.CDLL:
    if precision not in _libs:
        build()
        lib = C.CDLL(os.path.join(HERE, f"libminimd_oracle_{precision}.so"))
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.c_int]
        lib.orc_destroy.argtypes = [C.c_void_p]
        lib.orc_init.argtypes = [C.c_void_p, C.POINTER(OrcConfig), C.c_char_p]
        lib.orc_run.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.orc_get_int.argtypes = [C.c_void_p, C.c_char_p]
        lib.orc_get_real.argtypes = [C.c_void_p, C.c_char_p]
        lib.orc_get_real.restype = C.c_double
        lib.orc_set_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.orc_set_real.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        lib.orc_get_ptr.argtypes = [C.c_void_p, C.c_char_p]
        lib.orc_get_ptr.restype = C.c_void_p
        lib.orc_get_sendlist.argtypes = [C.c_void_p, C.c_int]
        lib.orc_get_sendlist.restype = C.c_void_p
        lib.orc_reserve_atoms.argtypes = [C.c_void_p, C.c_int]
        for fn in ("orc_binatoms",):
            getattr(lib, fn).argtypes = [C.c_void_p, C.c_int]
        for fn in ("orc_neighbor_build", "orc_sort", "orc_pbc", "orc_borders", "orc_communicate",
                   "orc_reverse_communicate", "orc_force_compute", "orc_initial_integrate",
                   "orc_final_integrate", "orc_neighbor_setup", "orc_comm_setup", "orc_force_lj_setup"):
            getattr(lib, fn).argtypes = [C.c_void_p]
        lib.orc_thermo_record.argtypes = [C.c_void_p, C.c_int]
        _libs[precision] = lib
    return _libs[precision]


def default_eam_file() -> str | None:
    """Cu_u6.eam location: oracle/_ref/ copy (made by build_ref.sh) or the regenerated fixture."""
    p = os.path.join(REF_DIR, "Cu_u6.eam")
    if os.path.exists(p):
        return p
    return None


class Oracle:
    """One single-rank oracle simulation (FP64 by default)."""

    def __init__(self, cfg: Config, precision: str = "f64", eam_file: str | None = None, seed_types: bool = True):
        self.cfg = cfg.resolved()
        self.precision = precision
        self.real = np.float64 if precision == "f64" else np.float32
        self.lib = _lib(precision)
        self.h = self.lib.orc_create(self.cfg.ntypes)
        if seed_types:
            _libc.srand(5413)  # ref/ljs.cpp:110
        st = self.cfg.as_struct(self.real)
        ef = eam_file or default_eam_file()
        rc = self.lib.orc_init(self.h, C.byref(st), (ef or "Cu_u6.eam").encode())
        if rc:
            raise RuntimeError(f"orc_init failed rc={rc}")
        self.step = 0

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- scalars -------------------------------------------------------------------------
    def geti(self, k: str) -> int:
        return self.lib.orc_get_int(self.h, k.encode())

    def getr(self, k: str) -> float:
        return self.lib.orc_get_real(self.h, k.encode())

    def seti(self, k: str, v: int) -> None:
        self.lib.orc_set_int(self.h, k.encode(), int(v))

    def setr(self, k: str, v: float) -> None:
        self.lib.orc_set_real(self.h, k.encode(), float(v))

    # -- arrays (views into oracle memory; copy if you need them to survive a call) -------
    def _arr(self, k: str, dtype, n: int) -> np.ndarray:
        p = self.lib.orc_get_ptr(self.h, k.encode())
        if not p or n == 0:
            return np.zeros(0, dtype=dtype)
        buf = (C.c_char * (np.dtype(dtype).itemsize * n)).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=n)

    @property
    def nlocal(self) -> int:
        return self.geti("nlocal")

    @property
    def nghost(self) -> int:
        return self.geti("nghost")

    @property
    def nall(self) -> int:
        return self.nlocal + self.nghost

    def x(self, n=None):
        n = self.nall if n is None else n
        return self._arr("x", self.real, 3 * n).reshape(n, 3)

    def v(self, n=None):
        n = self.nlocal if n is None else n
        return self._arr("v", self.real, 3 * n).reshape(n, 3)

    def f(self, n=None):
        n = self.nall if n is None else n
        return self._arr("f", self.real, 3 * n).reshape(n, 3)

    def type(self, n=None):
        n = self.nall if n is None else n
        return self._arr("type", np.int32, n)

    def numneigh(self):
        return self._arr("numneigh", np.int32, self.nlocal)

    def neighbors(self):
        m = self.geti("maxneighs")
        return self._arr("neighbors", np.int32, self.nlocal * m).reshape(self.nlocal, m)

    def bincount(self):
        return self._arr("bincount", np.int32, self.geti("mbins"))

    def bins(self):
        return self._arr("bins", np.int32, self.geti("mbins") * self.geti("atoms_per_bin")).reshape(
            self.geti("mbins"), self.geti("atoms_per_bin"))

    def stencil(self):
        return self._arr("stencil", np.int32, self.geti("nstencil"))

    def ivec(self, k: str, n: int):
        return self._arr(k, np.int32, n)

    def rvec(self, k: str, n: int):
        return self._arr(k, self.real, n)

    def sendlist(self, iswap: int) -> np.ndarray:
        n = int(self.ivec("sendnum", self.geti("nswap"))[iswap])
        p = self.lib.orc_get_sendlist(self.h, iswap)
        if not p or n == 0:
            return np.zeros(0, dtype=np.int32)
        buf = (C.c_char * (4 * n)).from_address(p)
        return np.frombuffer(buf, dtype=np.int32, count=n)

    def thermo_log(self):
        n = self.geti("nlog")
        return (self._arr("log_step", np.int32, n).copy(), self._arr("log_t", np.float64, n).copy(),
                self._arr("log_e", np.float64, n).copy(), self._arr("log_p", np.float64, n).copy())

    # -- stepping ------------------------------------------------------------------------
    def run(self, nsteps: int) -> None:
        self.lib.orc_run(self.h, self.step, nsteps)
        self.step += nsteps

    def call(self, fn: str, *args) -> None:
        getattr(self.lib, "orc_" + fn)(self.h, *args)


# --------------------------------------------------------------------------------------
# unmodified reference binary (oracle/_ref/)
# --------------------------------------------------------------------------------------

@dataclass
class RefResult:
    steps: list = field(default_factory=list)
    T: list = field(default_factory=list)
    U: list = field(default_factory=list)
    P: list = field(default_factory=list)
    nlocal: int = -1
    nghost: int = -1
    neighs: int = -1
    perf: float = float("nan")      # atom-steps/s from PERF_SUMMARY
    t_total: float = float("nan")
    t_force: float = float("nan")
    t_neigh: float = float("nan")
    t_comm: float = float("nan")
    natoms: int = -1
    stdout: str = ""


def ref_binary(precision: str = "f64") -> str | None:
    p = os.path.join(REF_DIR, f"miniMD_ref_{precision}")
    return p if os.path.exists(p) else None


def run_reference(cfg: Config, precision: str = "f64", threads: int = 1, yaml: bool = True,
                  timeout: float | None = None) -> RefResult:
    """Run oracle/_ref/miniMD_ref_<precision> on `cfg`; parse thermo, counts and PERF_SUMMARY."""
    exe = ref_binary(precision)
    if exe is None:
        raise FileNotFoundError("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    c = cfg.resolved()
    with tempfile.TemporaryDirectory() as td:
        inp = os.path.join(td, "in.miniMD")
        with open(inp, "w") as fh:
            fh.write(c.input_text())
        cmd = [exe, "-i", inp, "-t", str(threads), "--half_neigh", str(c.halfneigh), "-gn", str(c.ghost_newton),
               "--sort", str(c.sort), "--ntypes", str(c.ntypes)]
        if c.nbins > 0:
            cmd += ["-b", str(c.nbins)]
        if yaml:
            cmd += ["-o", "1", "--yaml_screen"]
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))
        out = subprocess.run(cmd, cwd=REF_DIR, capture_output=True, text=True, timeout=timeout, env=env).stdout
        for fn in os.listdir(REF_DIR):
            if fn.startswith("miniMD-") and fn.endswith(".yaml"):
                os.remove(os.path.join(REF_DIR, fn))
    return parse_reference_output(out, yaml)


def parse_reference_output(out: str, yaml: bool = True) -> RefResult:
    r = RefResult(stdout=out)
    if yaml and "thermodynamic_output:" in out:
        blk = out.split("thermodynamic_output:")[1].split("time:")[0]
        for m in re.finditer(r"timestep:\s*(-?\d+)\s*\n\s*T\*:\s*(\S+)\s*\n\s*U\*:\s*(\S+)\s*\n\s*P\*:\s*(\S+)", blk):
            r.steps.append(int(m.group(1)))
            r.T.append(float(m.group(2)))
            r.U.append(float(m.group(3)))
            r.P.append(float(m.group(4)))
    else:
        body = out.split("# Timestep T U P Time")[1].split("# Performance Summary")[0] if "# Timestep T U P Time" in out else ""
        for ln in body.strip().splitlines():
            p = ln.split()
            if len(p) == 5:
                r.steps.append(int(p[0])); r.T.append(float(p[1])); r.U.append(float(p[2])); r.P.append(float(p[3]))
    for key, attr in (("Nlocal", "nlocal"), ("Nghost", "nghost"), ("Neighs", "neighs")):
        m = re.search(rf"# {key}:\s+(\S+) ave", out)
        if m:
            setattr(r, attr, int(float(m.group(1))))
    m = re.search(r"^(\d+) (\d+) (\d+) (\d+) (\S+) (\S+) (\S+) (\S+) (\S+) (\S+) (\S+) PERF_SUMMARY", out, re.M)
    if m:
        r.natoms = int(m.group(4))
        r.t_total, r.t_force, r.t_neigh, r.t_comm = (float(m.group(i)) for i in (5, 6, 7, 8))
        r.perf = float(m.group(10))
    return r
