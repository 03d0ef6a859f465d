"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs.  Integer / index work must be bit-exact; floating point within the stated
tolerances (FP64: 1e-11 relative per kernel, 1e-9 on T/U/P over 100 steps; the reference's
own acceptance is 1e-5).
"""
import numpy as np
import pytest

from helpers import context_from_oracle, rel, run_params, thermo_from_samples
from oracle.oracle import Config, Oracle

pytestmark = pytest.mark.gpu

F64_KERNEL_RTOL = 1e-11
F32_KERNEL_RTOL = 2e-4


def ktol(prec):
    return F64_KERNEL_RTOL if prec == "f64" else F32_KERNEL_RTOL


def maxabs(a):
    return float(np.max(np.abs(a))) if np.size(a) else 0.0


def assert_close(a, b, rtol, what=""):
    scale = max(maxabs(b), 1e-300)
    err = maxabs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)) / scale
    assert err <= rtol, f"{what}: max err {err:.3e} > {rtol:.1e} (scale {scale:.3e})"


def melted(cfg_kwargs, steps, prec="f64"):
    """Oracle advanced to a re-neighboring step, so x / ghosts / lists are mutually consistent."""
    cfg = Config(**cfg_kwargs)
    o = Oracle(cfg, prec)
    if steps:
        assert steps % cfg.resolved().neigh_every == 0
        o.run(steps)
    return o


CASES = [
    dict(nx=8, ny=8, nz=8, halfneigh=1, ghost_newton=1),
    dict(nx=8, ny=8, nz=8, halfneigh=1, ghost_newton=0),
    dict(nx=8, ny=8, nz=8, halfneigh=0, ghost_newton=0),
    dict(nx=6, ny=8, nz=10, halfneigh=1, ghost_newton=1, sort=0),
    dict(nx=3, ny=3, nz=3, halfneigh=0, ghost_newton=1),
]


# ------------------------------------------------------------------------------------------
# binning / borders / neighbor build / sort : bit-exact
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tile", [1, 3, 4, 2, 0])
@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("steps", [0, 40])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_borders_bins_and_lists_are_bit_exact(case, steps, prec, tile):
    """tile=1: tile-resident 16-bit rows built on x-sorted windows by the interval build, two atoms of a bin per sweep
    (the default), and exported back to the reference's format; tile=3: the same, one atom per sweep; tile=4: the same
    rows from the one-lane-per-atom build; tile=2: from the per-bin candidate-table build (windows in CSR order);
    tile=0: classic rows."""
    o = melted(CASES[case], steps, prec)
    c = context_from_oracle(o)
    c.set_option("tile_lists", 1 if tile else 0)
    c.set_option("tile_xsort", 1 if tile in (1, 3, 4) else 0)
    c.set_option("tile_pair_build", 0 if tile == 3 else 1)
    c.set_option("tile_lane_build", 1 if tile == 4 else 0)
    c.exchange()
    c.borders()
    # ghosts: counts, send lists, positions (+- prd shifts) and types
    assert c.counts()[:2] == (o.nlocal, o.nghost)
    sn, rn, fr = c.swap_counts()
    ns = o.geti("nswap")
    assert list(sn) == list(o.ivec("sendnum", ns)) and list(fr) == list(o.ivec("firstrecv", ns))
    for w in range(ns):
        assert np.array_equal(c.sendlist(w), o.sendlist(w)), f"sendlist {w}"
    d = c.download("xt")
    assert np.array_equal(d["x"], o.x()), "ghost positions differ"
    assert np.array_equal(d["type"], o.type())
    # bins
    apb, mx = c.binatoms(-1, 8)
    o.call("binatoms", -1)
    assert apb == o.geti("atoms_per_bin")
    cnt, rows = c.bins_download(apb)
    assert np.array_equal(cnt, o.bincount())
    ob = o.bins()
    for b in np.nonzero(cnt)[0]:
        assert np.array_equal(rows[b, :cnt[b]], ob[b, :cnt[b]])
    # neighbor lists
    half, gn = o.geti("halfneigh"), o.geti("ghost_newton")
    mxn, total = c.build(half, gn, 100)
    assert c.query("list_tile") == (1 if tile else 0)
    assert c.query("list_xsorted") == (1 if tile in (1, 3, 4) else 0)
    num, nb = c.neigh_download()
    onum, onb = o.numneigh(), o.neighbors()
    assert mxn == o.geti("maxneighs")
    assert total == int(onum.sum())
    assert np.array_equal(num, onum)
    for i in range(o.nlocal):
        assert np.array_equal(nb[i, :num[i]], onb[i, :num[i]]), f"row {i}"


@pytest.mark.parametrize("tile", [1, 0])
def test_neighbor_resize_protocol_matches_reference(tile):
    o = Oracle(Config(nx=6, ny=6, nz=6, halfneigh=0), "f64")
    c = context_from_oracle(o)
    c.set_option("tile_lists", tile)
    c.exchange()
    c.borders()
    o.seti("maxneighs", 20)
    o.call("neighbor_build")
    mxn, total = c.build(0, 1, 20)
    assert mxn == o.geti("maxneighs") and mxn > 20
    assert c.query("neigh_resizes") >= 1
    num, nb = c.neigh_download()
    assert np.array_equal(num, o.numneigh())
    onb = o.neighbors()
    for i in range(0, o.nlocal, 7):
        assert np.array_equal(nb[i, :num[i]], onb[i, :num[i]])


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_sort_permutation_is_the_reference_order(prec):
    o = melted(dict(nx=8, ny=8, nz=8, sort=0), 40, prec)   # never sorted so far
    c = context_from_oracle(o)
    c.sort()
    o.call("sort")
    d = c.download("xvt", count=o.nlocal)
    assert np.array_equal(d["x"], o.x(o.nlocal))
    assert np.array_equal(d["v"], o.v())
    assert np.array_equal(d["type"], o.type(o.nlocal))


def test_coord2bin_edge_cases_bit_exact():
    """Atoms exactly on bin / box boundaries and ghosts outside [0,prd)."""
    o = Oracle(Config(nx=4, ny=4, nz=4), "f64")
    prd = o.getr("box.xprd")
    bs = 1.0 / o.getr("bininvx")
    rng = np.random.default_rng(7)
    n = 4096
    x = rng.uniform(-2.7, prd + 2.7, size=(n, 3))
    edges = np.array([0.0, prd, bs, 2 * bs, prd - bs, np.nextafter(prd, 0), np.nextafter(0.0, -1), -bs, prd + bs,
                      np.nextafter(bs, 0), np.nextafter(bs, 10)])
    x[:edges.size * 3] = np.repeat(edges, 3)[:, None]
    x[:edges.size, 1] = rng.uniform(0, prd, edges.size)
    c = context_from_oracle(o, upload_atoms=False)
    c.upload(x, np.zeros_like(x), np.zeros(n, dtype=np.int32))
    c.binatoms(n, 8)
    got = c.atom_bins(n)
    lib = o.lib
    import ctypes as C
    lib.orc_coord2bin.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    want = np.array([lib.orc_coord2bin(o.h, *map(float, p)) for p in x], dtype=np.int32)
    assert np.array_equal(got, want)


# ------------------------------------------------------------------------------------------
# forces
# ------------------------------------------------------------------------------------------
def setup_lists(o, tile=1):
    c = context_from_oracle(o)
    c.set_option("tile_lists", tile)
    c.set_option("tile_eam", tile)      # EAM uses classic rows unless asked (measured faster); the tests cover both
    c.exchange()
    c.borders()
    c.build(o.geti("halfneigh"), o.geti("ghost_newton"), 100)
    return c


@pytest.mark.parametrize("dealt", [1, 0])
@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("half,gn", [(1, 1), (1, 0), (0, 0)])
@pytest.mark.parametrize("size", [(8, 8, 8), (5, 7, 9)])
def test_lj_tile_force_energy_virial(half, gn, size, prec, dealt):
    """Tile-resident lists: every local atom's force is complete after the kernel (no scatter, no reverse
    halo), so half+ghost_newton is compared with the oracle AFTER its reverse_communicate.
    dealt=1: bank-dealt rows walked by a quarter warp per atom (the default); dealt=0: one lane pair per row."""
    o = melted(dict(nx=size[0], ny=size[1], nz=size[2], halfneigh=half, ghost_newton=gn), 40, prec)
    c = context_from_oracle(o)
    c.set_option("tile_xsort", 1 if size[0] == 8 else 0)   # one size per build flavour
    c.set_option("tile_dealt", dealt)
    c.exchange()
    c.borders()
    c.build(half, gn, 100)
    assert c.query("list_tile") == 1
    assert c.query("list_dealt") == dealt
    o.seti("evflag", 1)
    o.call("force_compute")
    if half and gn:
        o.call("reverse_communicate")
    eng, vir = c.lj_compute(half, gn, 1)
    nl = o.nlocal
    f = c.download("f", count=nl)["f"]
    assert_close(f, o.f(nl), ktol(prec), "f")
    etol = 1e-11 if prec == "f64" else 2e-3
    assert abs(eng - o.getr("eng_vdwl")) <= etol * abs(o.getr("eng_vdwl"))
    assert abs(vir - o.getr("virial")) <= etol * max(abs(o.getr("virial")), abs(o.getr("eng_vdwl")))
    # a reverse halo after the kernel must be a no-op (ghost forces are kept at zero)
    c.reverse_communicate()
    assert np.array_equal(c.download("f", count=nl)["f"], f)
    c.lj_compute(half, gn, 0)
    assert_close(c.download("f", count=nl)["f"], f, 1e-13 if prec == "f64" else 1e-5, "f(ev=0) vs f(ev=1)")


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("tpa", [1, 4, 8, 32])
@pytest.mark.parametrize("half,gn", [(1, 1), (1, 0), (0, 0)])
def test_lj_force_energy_virial(half, gn, tpa, prec):
    o = melted(dict(nx=8, ny=8, nz=8, halfneigh=half, ghost_newton=gn), 40, prec)
    c = setup_lists(o, tile=0)
    c.set_option("lj_threads_per_atom", tpa)
    o.seti("evflag", 1)
    o.call("force_compute")
    eng, vir = c.lj_compute(half, gn, 1)
    n = o.nall if half else o.nlocal
    f = c.download("f", count=n)["f"]
    assert_close(f, o.f(n), ktol(prec), "f")
    # the FP32 reference accumulates energy in one float; the device reduces in FP64
    etol = 1e-11 if prec == "f64" else 2e-3
    assert abs(eng - o.getr("eng_vdwl")) <= etol * abs(o.getr("eng_vdwl"))
    assert abs(vir - o.getr("virial")) <= etol * max(abs(o.getr("virial")), abs(o.getr("eng_vdwl")))
    # evflag=0 path gives the same forces
    c.lj_compute(half, gn, 0)
    f0 = c.download("f", count=n)["f"]
    assert_close(f0, f, 1e-13 if prec == "f64" else 1e-5, "f(ev=0) vs f(ev=1)")


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("half,gn", [(1, 1), (1, 0), (0, 0)])
def test_lj_compute_seam_with_host_lists(half, gn, prec):
    """The mpi-spec compute_lj seam (mpi-spec/force_lj_custom.cpp:17-30): the HOST owns ghosts and neighbor lists and
    hands over x of nlocal + nghost atoms and its rows on every call -- mmd_atom_upload + mmd_atom_split +
    mmd_neigh_upload + mmd_force_lj_compute; forces (ghost contributions of half lists included) come back."""
    o = melted(dict(nx=7, ny=6, nz=8, halfneigh=half, ghost_newton=gn), 40, prec)
    c = context_from_oracle(o, upload_atoms=False)
    nl, na = o.nlocal, o.nall
    v = np.zeros((na, 3), dtype=c.real)
    v[:nl] = o.v(nl)
    c.upload(o.x(na), v, o.type(na))
    c.split(nl)
    assert c.counts()[:2] == (nl, na - nl)
    num, nb = o.numneigh(), o.neighbors()
    c.neigh_upload(num, nb)
    assert c.query("list_tile") == 0
    o.seti("evflag", 1)
    o.call("force_compute")
    eng, vir = c.lj_compute(half, gn, 1)
    n = na if half else nl
    assert_close(c.download("f", count=n)["f"], o.f(n), ktol(prec), "f")
    etol = 1e-11 if prec == "f64" else 2e-3
    assert abs(eng - o.getr("eng_vdwl")) <= etol * abs(o.getr("eng_vdwl"))
    # positions change, lists stay: the per-step path of the seam
    o.call("initial_integrate")
    o.call("communicate")
    c.update(x=o.x(na))
    o.seti("evflag", 0)
    o.call("force_compute")
    c.lj_compute(half, gn, 0)
    assert_close(c.download("f", count=n)["f"], o.f(n), ktol(prec), "f after update")


@pytest.mark.parametrize("tile", [1, 0])
def test_lj_per_type_tables_path(tile):
    """Distinct epsilon/sigma per type pair exercises the non-uniform table kernels."""
    o = melted(dict(nx=6, ny=6, nz=6, halfneigh=0, ghost_newton=0, ntypes=3), 20, "f64")
    nn = 9
    rng = np.random.default_rng(3)
    eps = o.rvec("epsilon", nn); s6 = o.rvec("sigma6", nn); cut = o.rvec("cutforcesq", nn)
    sym = lambda a: (a + a.T) / 2
    eps[:] = sym(rng.uniform(0.8, 1.2, (3, 3))).ravel()
    s6[:] = sym(rng.uniform(0.9, 1.1, (3, 3))).ravel()
    cut[:] = sym(rng.uniform(5.0, 6.25, (3, 3))).ravel()
    c = setup_lists(o, tile)
    assert c.query("list_tile") == tile
    assert c.query("lj_uniform") == 0
    o.seti("evflag", 1)
    o.call("force_compute")
    eng, vir = c.lj_compute(0, 0, 1)
    assert_close(c.download("f", count=o.nlocal)["f"], o.f(o.nlocal), 1e-11, "f")
    assert abs(eng - o.getr("eng_vdwl")) <= 1e-11 * abs(o.getr("eng_vdwl"))
    assert abs(vir - o.getr("virial")) <= 1e-10 * abs(o.getr("eng_vdwl"))


@pytest.mark.parametrize("tile", [1, 0])
@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("half", [1, 0])
@pytest.mark.parametrize("uniform", [1, 0])
def test_eam_force_energy_virial(half, uniform, prec, tile):
    """tile=1: both EAM passes as owner-computes shared-memory kernels on tile-resident rows; tile=0: classic rows."""
    o = melted(dict(nx=6, ny=6, nz=6, force="eam", halfneigh=half, ghost_newton=0), 20, prec)
    c = setup_lists(o, tile)
    assert c.query("list_tile") == tile
    if not uniform:
        c.set_option("force_nonuniform", 1)
    o.seti("evflag", 1)
    o.call("force_compute")
    eng, vir = c.eam_compute(half, 1)
    n = o.nall if half else o.nlocal
    f = c.download("f", count=n)["f"]
    assert_close(f, o.f(n), 1e-10 if prec == "f64" else 5e-4, "f")
    # FP32: the reference (and so the oracle) accumulates energy/virial in ONE float over ~1e5 pairs
    # (ref/force_eam.cpp:96,259); the device reduces in FP64 and is closer to the FP64 truth.
    etol = 1e-11 if prec == "f64" else 2e-3
    assert abs(eng - o.getr("eng_vdwl")) <= etol * abs(o.getr("eng_vdwl"))
    assert abs(vir - o.getr("virial")) <= max(etol, 1e-10) * max(abs(o.getr("virial")), 1.0) * 10


# ------------------------------------------------------------------------------------------
# integrate / halo / thermo
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tile", [1, 0])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_integrate_pbc_halo_and_thermo_kernels(prec, tile):
    o = melted(dict(nx=8, ny=8, nz=8), 40, prec)
    c = setup_lists(o, tile)
    nl, na = o.nlocal, o.nall
    tol = 1e-14 if prec == "f64" else 1e-6
    # forces first (both sides), then reverse halo
    o.seti("evflag", 0)
    o.call("force_compute")
    c.lj_compute(1, 1, 0)
    o.call("reverse_communicate")
    c.reverse_communicate()
    assert_close(c.download("f", count=nl)["f"], o.f(nl), ktol(prec), "f after reverse")
    # make device f identical to the oracle's so the integrator is tested in isolation
    dt, dtf = o.getr("dt"), o.getr("dtforce")
    # (velocity-Verlet halves: FMA contraction on the device => 1 ulp)
    o.call("final_integrate")
    c.final_integrate(dtf)
    assert_close(c.download("v")["v"], o.v(), max(tol, ktol(prec)), "v final")
    mv2 = c.sum_mv2(o.getr("mass"))
    want = float(np.sum(o.v().astype(np.float64) ** 2) * o.getr("mass"))
    assert abs(mv2 - want) <= 1e-6 * want
    o.call("initial_integrate")
    c.initial_integrate(dt, dtf)
    d = c.download("xv", count=nl)
    assert_close(d["x"], o.x(nl), max(tol, ktol(prec)), "x initial")
    assert_close(d["v"], o.v(), max(tol, ktol(prec)), "v initial")
    # forward halo: ghosts follow their owners exactly (same inputs => compare device with itself)
    o.call("communicate")
    c.communicate()
    xs = c.download("x")["x"]
    prd = np.array([o.getr("box.xprd"), o.getr("box.yprd"), o.getr("box.zprd")])
    ns = o.geti("nswap")
    sn, fr = o.ivec("sendnum", ns), o.ivec("firstrecv", ns)
    flags = np.stack([o.ivec("pbc_flagx", ns), o.ivec("pbc_flagy", ns), o.ivec("pbc_flagz", ns)], axis=1)
    ref = xs.copy()
    for w in range(ns):
        ref[fr[w]:fr[w] + sn[w]] = ref[o.sendlist(w)] + (flags[w] * prd).astype(xs.dtype)
    assert np.array_equal(xs, ref), "forward halo is not an exact copy+shift"
    assert_close(xs, o.x(), max(tol, ktol(prec)), "x after halo")
    # pbc wrap: push some atoms out of the box and wrap on both sides
    x = o.x(nl).copy()
    x[::5] += prd.astype(x.dtype)
    x[1::5] -= prd.astype(x.dtype)
    o.x(nl)[:] = x
    c.update(x=x)
    o.call("pbc")
    c.pbc()
    assert np.array_equal(c.download("x", count=nl)["x"], o.x(nl))


# ------------------------------------------------------------------------------------------
# whole time loop
# ------------------------------------------------------------------------------------------
RUN_CASES = [
    ("lj", 1, 1, "f64", 1e-9), ("lj", 1, 0, "f64", 1e-9), ("lj", 0, 0, "f64", 1e-9),
    ("eam", 1, 0, "f64", 1e-9), ("eam", 0, 0, "f64", 1e-9),
    ("lj", 0, 0, "f32", 1e-4), ("lj", 1, 1, "f32", 1e-4), ("eam", 0, 0, "f32", 1e-4),
]


@pytest.mark.parametrize("tile", [1, 0])
@pytest.mark.parametrize("force,half,gn,prec,tol", RUN_CASES)
def test_time_loop_matches_oracle(force, half, gn, prec, tol, tile):
    cfg = Config(nx=8, ny=8, nz=8, ntimes=100, force=force, halfneigh=half, ghost_newton=gn, thermo_nstat=10)
    o64 = Oracle(cfg, "f64")          # FP32 runs are judged against the FP64 oracle
    o = Oracle(cfg, prec)
    c = context_from_oracle(o)
    c.set_option("tile_lists", tile)
    c.set_option("tile_eam", tile)
    c.exchange()
    c.borders()
    c.build(o.geti("halfneigh"), o.geti("ghost_newton"), 100)
    assert c.query("list_tile") == tile
    samples, ms = c.run(run_params(o, 100))
    assert c.query("list_tile") == tile
    got = thermo_from_samples(o, samples)
    o64.run(100)
    st, T, U, P = o64.thermo_log()
    assert [g[0] for g in got] == list(st[1:])
    pscale = max(1.0, float(np.max(np.abs(P))))
    for (step, t, e, p), tw, ew, pw in zip(got, T[1:], U[1:], P[1:]):
        assert abs(t - tw) <= tol * abs(tw), (step, t, tw)
        assert abs(e - ew) <= tol * abs(ew), (step, e, ew)
        assert abs(p - pw) <= 10 * tol * pscale, (step, p, pw)
    # integer pins at the end of the run
    o.run(100)
    assert c.counts()[:2] == (o.nlocal, o.nghost) or prec == "f32"
    if prec == "f64":
        assert c.query("total_neigh") == int(o.numneigh().sum())
        assert c.query("maxneighs") == o.geti("maxneighs")


@pytest.mark.parametrize("force,half,prec", [("lj", 1, "f64"), ("lj", 0, "f32"), ("eam", 0, "f64")])
def test_graph_replay_is_the_eager_loop(force, half, prec):
    """Option graph_steps: pairs of plain steps replayed from a CUDA graph (captured once per neighbor list) must leave
    exactly the state of the step-by-step launches -- same kernels, same arguments, same count."""
    def run(graph):
        cfg = Config(nx=8, ny=8, nz=8, ntimes=70, force=force, halfneigh=half, ghost_newton=half, thermo_nstat=30)
        o = Oracle(cfg, prec)
        c = context_from_oracle(o)
        c.set_option("graph_steps", graph)
        c.exchange()
        c.borders()
        c.build(half, half, 100)
        samples, _ = c.run(run_params(o, 70))
        d = c.download("xv", count=c.counts()[0])
        return samples, d["x"], d["v"], c.query("launches"), c.query("graph_replays"), c.query("graph_captures")
    eager, graph = run(0), run(1)
    assert eager[4] == 0 and graph[4] > 20 and 0 < graph[5] <= 6, (eager[4:], graph[4:])
    assert np.array_equal(eager[1], graph[1]) and np.array_equal(eager[2], graph[2])
    # (energy / virial are summed over CTAs with floating-point atomics: equal up to the order of that sum)
    for a, b in zip(eager[0], graph[0]):
        assert a[0] == b[0] and all(abs(u - w) <= 1e-13 * abs(u) for u, w in zip(a[1:], b[1:])), (a, b)
    assert eager[3] == graph[3], "a replay accounts for the launches it stands for"


@pytest.mark.parametrize("opts", [dict(tile_lists=1, tile_dealt=1), dict(tile_lists=1, tile_dealt=0),
                                  dict(tile_lists=1, tile_dealt=1, fuse_force=0), dict(tile_lists=0)])
def test_isolated_atoms_are_still_integrated(opts):
    """A gas so thin that no atom has a neighbor inside cutneigh (every row is empty): the atoms must still be owned by
    their rows -- forces stored as zero, velocities and positions advanced -- in every list format and fusion mode."""
    cfg = Config(nx=6, ny=6, nz=6, ntimes=40, rho=0.02, halfneigh=1, ghost_newton=1, thermo_nstat=10)
    o = Oracle(cfg, "f64")
    c = context_from_oracle(o)
    for k, v in opts.items():
        c.set_option(k, v)
    c.exchange()
    c.borders()
    c.build(1, 1, 100)
    assert c.query("total_neigh") == 0
    samples, _ = c.run(run_params(o, 40))
    o.run(40)
    nl = o.nlocal
    assert c.counts()[0] == nl
    d = c.download("xvf", count=nl)
    assert_close(d["x"], o.x(nl), 1e-13, "x")
    assert_close(d["v"], o.v(nl), 1e-13, "v")
    assert maxabs(d["f"]) == 0.0
    st, T, U, P = o.thermo_log()
    got = thermo_from_samples(o, samples)
    for (step, t, e, p), tw in zip(got, T[1:]):
        assert abs(t - tw) <= 1e-12 * abs(tw) and e == 0.0


# ------------------------------------------------------------------------------------------
# run-time switches: every optional fusion must reproduce the unfused path
# ------------------------------------------------------------------------------------------
def _run_with(options, prec="f64", steps=60):
    cfg = Config(nx=8, ny=8, nz=8, ntimes=steps, halfneigh=1, ghost_newton=1, thermo_nstat=20)
    o = Oracle(cfg, prec)
    c = context_from_oracle(o)
    for k, v in options.items():
        c.set_option(k, v)
    c.exchange()
    c.borders()
    c.build(1, 1, 100)
    samples, _ = c.run(run_params(o, steps))
    d = c.download("xv", count=c.counts()[0])
    return samples, d["x"], d["v"], c.query("launches")


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_fused_paths_reproduce_the_unfused_ones(prec):
    base = _run_with(dict(fuse_halo=0, fuse_force=0, fuse_integrate=0), prec)
    # one-launch forward halo: an exact copy+shift either way
    halo = _run_with(dict(fuse_halo=1, fuse_force=0, fuse_integrate=0), prec)
    assert np.array_equal(halo[1], base[1]) and np.array_equal(halo[2], base[2])
    assert halo[3] < base[3]
    # Verlet halves in one kernel / in the force kernel's epilogue: same operations in the same order
    for opts in (dict(fuse_halo=1, fuse_force=0, fuse_integrate=1), dict(fuse_halo=1, fuse_force=1, fuse_integrate=1),
                 dict(fuse_halo=1, fuse_force=1, fuse_integrate=1, tile_dealt=0),
                 dict(fuse_halo=1, fuse_force=1, fuse_integrate=1, fuse_ghosts=1)):
        got = _run_with(opts, prec)
        tol = 1e-12 if prec == "f64" else 5e-5
        assert_close(got[1], base[1], tol, f"x {opts}")
        assert_close(got[2], base[2], tol, f"v {opts}")
        for (s0, m0, e0, v0), (s1, m1, e1, v1) in zip(base[0], got[0]):
            assert s0 == s1 and abs(m0 - m1) <= tol * abs(m0) and abs(e0 - e1) <= tol * abs(e0)
        assert got[3] < base[3]


# ------------------------------------------------------------------------------------------
# geometries off the beaten path: wider stencils, fat and thin bins -- whichever list format the
# library picks (tile rows from either build, or the classic fallback) must stay exact
# ------------------------------------------------------------------------------------------
ODD_CASES = [
    dict(nx=8, ny=8, nz=8, force_cut=4.0),                 # stencil 7x7x7: 49 runs -> per-bin table build, large windows
    dict(nx=8, ny=8, nz=8, nbins=3),                       # fat bins (~75 atoms): no x-sort, 3x3x3 stencil
    dict(nx=8, ny=8, nz=8, nbins=4, halfneigh=0),          # ~32 atoms per bin
    dict(nx=8, ny=8, nz=8, nbins=13),                      # thin bins: stencil wider than the tiling supports
    dict(nx=6, ny=9, nz=7, nbins=9, halfneigh=1, ghost_newton=0),
]


@pytest.mark.parametrize("case", range(len(ODD_CASES)))
def test_unusual_geometries_lists_and_time_loop(case):
    kw = dict(ODD_CASES[case])
    cfg = Config(ntimes=40, thermo_nstat=10, **kw)
    o = Oracle(cfg, "f64")
    c = context_from_oracle(o)
    c.exchange()
    c.borders()
    half, gn = o.geti("halfneigh"), o.geti("ghost_newton")
    mxn, total = c.build(half, gn, 100)
    num, nb = c.neigh_download()
    onum, onb = o.numneigh(), o.neighbors()
    assert mxn == o.geti("maxneighs") and total == int(onum.sum())
    assert np.array_equal(num, onum)
    for i in range(o.nlocal):
        assert np.array_equal(nb[i, :num[i]], onb[i, :num[i]]), f"row {i}"
    samples, _ = c.run(run_params(o, 40))
    got = thermo_from_samples(o, samples)
    o.run(40)
    st, T, U, P = o.thermo_log()
    pscale = max(1.0, float(np.max(np.abs(P))))
    for (step, t, e, p), tw, ew, pw in zip(got, T[1:], U[1:], P[1:]):
        assert abs(t - tw) <= 1e-9 * abs(tw) and abs(e - ew) <= 1e-9 * abs(ew) and abs(p - pw) <= 1e-8 * pscale, (step, case)
    assert c.query("total_neigh") == int(o.numneigh().sum())
    print(case, "list_tile", c.query("list_tile"), "xsorted", c.query("list_xsorted"), "fallbacks", c.query("tile_fallbacks"),
          "max halo", c.query("tile_max_halo"))
