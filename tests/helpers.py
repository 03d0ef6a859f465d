"""Glue between the oracle (checker) and the CUDA path (thing under test)."""
import ctypes as C
import json
import os

import numpy as np

from minimd_b200 import Context
from minimd_b200._lib import RunParams, SwapTable
from oracle.oracle import Config, Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def golden(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def eam_file(tmpdir=None):
    """A funcfl file regenerated from the committed table fixture (17 significant digits)."""
    from oracle.oracle import default_eam_file
    p = default_eam_file()
    if p:
        return p
    d = np.load(os.path.join(GOLDEN, "cu_u6_funcfl.npz"))
    path = os.path.join(tmpdir or "/tmp", "Cu_u6.eam")
    with open(path, "w") as fh:
        fh.write("Cu funcfl table regenerated from tests/golden/cu_u6_funcfl.npz\n")
        fh.write(f"{int(d['atomic_number'])} {float(d['mass'])!r} {float(d['lattice'])!r} FCC\n")
        fh.write(f"{int(d['nrho'])} {float(d['drho'])!r} {int(d['nr'])} {float(d['dr'])!r} {float(d['cut'])!r}\n")
        for arr in (d["frho"], d["zr"], d["rhor"]):
            for i in range(0, arr.size, 5):
                fh.write(" ".join(f"{v:.16e}" for v in arr[i:i + 5]) + "\n")
    return path


def geometry_of(o: Oracle) -> dict:
    g = {k: o.geti(k) for k in ("nbinx", "nbiny", "nbinz", "mbinx", "mbiny", "mbinz", "mbinxlo", "mbinylo", "mbinzlo")}
    g.update({k: o.getr(k) for k in ("bininvx", "bininvy", "bininvz")})
    return g


def swap_table_of(o: Oracle) -> SwapTable:
    t = SwapTable()
    t.me, t.nprocs, t.nswap = 0, 1, o.geti("nswap")
    need = o.ivec("need", 3)
    for d in range(3):
        t.need[d] = int(need[d])
        t.procgrid[d] = 1
        t.procneigh[d][0] = t.procneigh[d][1] = 0
    n = t.nswap
    for name in ("pbc_any", "pbc_flagx", "pbc_flagy", "pbc_flagz"):
        a = o.ivec(name, n)
        for w in range(n):
            getattr(t, name)[w] = int(a[w])
    lo, hi = o.rvec("slablo", n), o.rvec("slabhi", n)
    for w in range(n):
        t.slablo[w], t.slabhi[w] = float(lo[w]), float(hi[w])
        t.sendproc[w] = t.recvproc[w] = 0
    return t


def context_from_oracle(o: Oracle, upload_atoms=True) -> Context:
    """Configure a device context exactly as the oracle is configured (box, bins, stencil, swap
    table, force parameters) and hand it the oracle's LOCAL atoms."""
    c = Context(o.precision, ntypes=o.cfg.ntypes)
    prd = [o.getr("box.xprd"), o.getr("box.yprd"), o.getr("box.zprd")]
    c.set_box(prd, [o.getr("box.xlo"), o.getr("box.ylo"), o.getr("box.zlo")],
              [o.getr("box.xhi"), o.getr("box.yhi"), o.getr("box.zhi")])
    nn = o.cfg.ntypes ** 2
    c.neigh_setup(geometry_of(o), o.stencil(), o.rvec("cutneighsq", nn))
    c.comm_setup(swap_table_of(o))
    if o.cfg.force == "lj":
        c.lj_setup(o.rvec("cutforcesq", nn), o.rvec("sigma6", nn), o.rvec("epsilon", nn))
    else:
        nr_tot, nrho_tot = o.geti("nr_tot"), o.geti("nrho_tot")
        c.eam_setup(o.rvec("rhor_spline", nn * nr_tot), o.rvec("z2r_spline", nn * nr_tot),
                    o.rvec("frho_spline", nn * nrho_tot), o.geti("nr"), o.geti("nrho"), nr_tot, nrho_tot,
                    o.getr("rdr"), o.getr("rdrho"), o.rvec("cutforcesq", nn))
    if upload_atoms:
        n = o.nlocal
        c.upload(o.x(n), o.v(n), o.type(n))
    return c


def run_params(o: Oracle, ntimes, first=0) -> RunParams:
    cfg = o.cfg
    return RunParams(ntimes=ntimes, first_step=first, total_steps=cfg.ntimes, neigh_every=cfg.neigh_every,
                     sort_every=o.geti("sort_every"), thermo_nstat=cfg.thermo_nstat, halfneigh=o.geti("halfneigh"),
                     ghost_newton=o.geti("ghost_newton"), force_style=0 if cfg.force == "lj" else 1,
                     dt=o.getr("dt"), dtforce=o.getr("dtforce"), mass=o.getr("mass"))


def thermo_from_samples(o: Oracle, samples):
    """Apply Thermo's scalings (ref/thermo.cpp:119-194) to raw device reductions."""
    out = []
    natoms = o.geti("natoms")
    for step, mv2, eng, vir in samples:
        t = mv2 * o.getr("t_scale")
        e = eng * (2.0 if o.geti("halfneigh") else 1.0) * o.getr("e_scale") / natoms
        p = (t * o.getr("dof_boltz") + vir) * o.getr("p_scale")
        out.append((step, t, e, p))
    return out


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def g6(v):
    """A count as the reference's YAML/histogram report prints it (%g, ref/output.cpp:402-481)."""
    return int(float(f"{float(v):.6g}"))


def write_lammps_data(path, x, v, prd, mass=1.0):
    """A LAMMPS data file in the layout the reference's reader expects (ref/setup.cpp:54-301):
    header counts + box bounds, then Masses / Atoms ('id type x y z') / Velocities ('id vx vy vz')."""
    with open(path, "w") as fh:
        fh.write("LAMMPS data file written by tests/helpers.py\n\n")
        fh.write(f"{len(x)} atoms\n1 atom types\n\n")
        for d, name in enumerate("xyz"):
            fh.write(f"0.0 {float(prd[d])!r} {name}lo {name}hi\n")
        fh.write("\nMasses\n\n")
        fh.write(f"1 {float(mass)!r}\n")
        fh.write("\nAtoms\n\n")
        for i, r in enumerate(x):
            fh.write(f"{i + 1} 1 {float(r[0])!r} {float(r[1])!r} {float(r[2])!r}\n")
        fh.write("\nVelocities\n\n")
        for i, r in enumerate(v):
            fh.write(f"{i + 1} {float(r[0])!r} {float(r[1])!r} {float(r[2])!r}\n")
    return path


def lattice_for_datafile(cells=6, jitter=0.05, seed=11):
    """Deterministic non-trivial configuration: the synthetic FCC lattice of the host layer, every atom
    displaced by a small seeded offset (so the data-file run differs from the built-in lattice run)."""
    from minimd_b200 import Simulation, input_file
    s = Simulation.plan(["-i", input_file("in.lj.miniMD"), "-s", str(cells)])
    x = s.host_array("x").reshape(-1, 3).copy()
    v = s.host_array("v").reshape(-1, 3).copy()
    prd = np.array([s.getr("xprd"), s.getr("yprd"), s.getr("zprd")])
    rng = np.random.default_rng(seed)
    x = np.mod(x + rng.uniform(-jitter, jitter, x.shape), prd)
    x = np.minimum(x, np.nextafter(prd, 0))
    return x, v, prd
