"""Host layer (minimd_b200/csrc/host) without a GPU: input parsing, CLI overrides, the synthetic
FCC/Park-Miller atoms, bin geometry + stencil, the swap table, force tables and thermo scalings
must equal the oracle's (which is pinned to the reference) bit for bit -- they define the inputs
and the bins every atom lands in.  Multi-rank decomposition is checked with 2 gloo processes."""
import os
import sys

import numpy as np
import pytest

from helpers import eam_file, geometry_of, swap_table_of
from minimd_b200 import HostError, Simulation, input_file
from oracle.oracle import Config, Oracle

GEO_KEYS = ("nbinx", "nbiny", "nbinz", "mbinx", "mbiny", "mbinz", "mbinxlo", "mbinylo", "mbinzlo",
            "bininvx", "bininvy", "bininvz")


def plan_for(cfg: Config, prec: str, extra=()):
    c = cfg.resolved()
    deck = "in.lj.miniMD" if c.force == "lj" else "in.eam.miniMD"
    args = ["-i", input_file(deck), "-nx", c.nx, "-ny", c.ny, "-nz", c.nz, "--half_neigh", c.halfneigh,
            "-gn", c.ghost_newton, "--sort", c.sort, "--ntypes", c.ntypes, "-n", c.ntimes]
    if c.nbins > 0:
        args += ["-b", c.nbins]
    if c.force == "eam":
        args += ["--eam_file", eam_file()]
    return Simulation.plan(list(args) + list(extra), prec)


CASES = [
    dict(nx=8, ny=8, nz=8), dict(nx=8, ny=8, nz=8, halfneigh=0, ghost_newton=0), dict(nx=8, ny=8, nz=8, halfneigh=1, ghost_newton=0),
    dict(nx=6, ny=8, nz=10, sort=0), dict(nx=3, ny=3, nz=3, halfneigh=0), dict(nx=12, ny=5, nz=7, nbins=4),
    dict(nx=6, ny=6, nz=6, force="eam", halfneigh=1), dict(nx=6, ny=6, nz=6, force="eam", halfneigh=0),
    dict(nx=5, ny=5, nz=5, ntypes=2), dict(nx=20, ny=20, nz=20),
]


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_single_rank_plan_equals_oracle(case, prec):
    cfg = Config(**CASES[case])
    o = Oracle(cfg, prec)          # oracle state after init: same atoms (no initial sort unless --sort > 0)
    s = plan_for(cfg, prec)
    n = o.nlocal
    assert s.geti("natoms") == o.geti("natoms") and s.geti("nlocal") == n
    assert np.array_equal(s.host_array("x").reshape(-1, 3), o.x(n))
    assert np.array_equal(s.host_array("v").reshape(-1, 3), o.v())
    assert np.array_equal(s.host_array("type"), o.type(n))
    # bins + stencil
    g, want = s.bin_geometry(), geometry_of(o)
    for k in GEO_KEYS:
        assert getattr(g, k) == want[k], k
    assert np.array_equal(s.host_array("stencil"), o.stencil())
    assert s.geti("mbins") == o.geti("mbins")
    nn = cfg.ntypes ** 2
    assert np.array_equal(s.host_array("cutneighsq"), o.rvec("cutneighsq", nn))
    assert np.array_equal(s.host_array("cutforcesq"), o.rvec("cutforcesq", nn))
    # swap table
    t, tw = s.swap_table(), swap_table_of(o)
    assert t.nswap == tw.nswap and list(t.need) == list(tw.need)
    for w in range(t.nswap):
        for f in ("pbc_any", "pbc_flagx", "pbc_flagy", "pbc_flagz", "slablo", "slabhi", "sendproc", "recvproc"):
            assert getattr(t, f)[w] == getattr(tw, f)[w], (f, w)
    # force parameters / tables
    if cfg.force == "lj":
        assert np.array_equal(s.host_array("epsilon"), o.rvec("epsilon", nn))
        assert np.array_equal(s.host_array("sigma6"), o.rvec("sigma6", nn))
    else:
        for k in ("nr", "nrho", "nr_tot", "nrho_tot"):
            assert s.geti(k) == o.geti(k)
        assert s.getr("rdr") == o.getr("rdr") and s.getr("rdrho") == o.getr("rdrho")
        assert np.array_equal(s.host_array("rhor_spline"), o.rvec("rhor_spline", nn * o.geti("nr_tot")))
        assert np.array_equal(s.host_array("z2r_spline"), o.rvec("z2r_spline", nn * o.geti("nr_tot")))
        assert np.array_equal(s.host_array("frho_spline"), o.rvec("frho_spline", nn * o.geti("nrho_tot")))
    # scalings and step sizes (dtforce: the oracle has already folded in 1/mass, ref/integrate.cpp:81)
    for k in ("t_scale", "e_scale", "p_scale", "dof_boltz", "mass", "dt"):
        assert s.getr(k) == o.getr(k), k
    assert np.float64(s.real(s.getr("dtforce")) / s.real(s.getr("mass"))) == o.getr("dtforce")
    for k in ("halfneigh", "ghost_newton", "sort_every"):
        assert s.geti(k) == o.geti(k), k


def test_cli_overrides_follow_the_reference():
    # -s wins over the deck; -nx alone sets all three; -nx with -s keeps ny,nz from -s (ref/ljs.cpp:332-349)
    s = Simulation.plan(["-i", input_file("in.lj.miniMD"), "-s", "4"])
    assert s.geti("natoms") == 4 * 4 ** 3 and (s.geti("nbinx"), s.geti("nbiny"), s.geti("nbinz")) == (3, 3, 3)
    s = Simulation.plan(["-i", input_file("in.lj.miniMD"), "-nx", "5"])
    assert s.geti("natoms") == 4 * 5 ** 3
    s = Simulation.plan(["-i", input_file("in.lj.miniMD"), "-s", "4", "-nx", "6"])
    assert s.geti("natoms") == 4 * 6 * 4 * 4
    # --sort: >0 value, <0 reneigh frequency, 0 never (ref/ljs.cpp:375)
    assert Simulation.plan(["-i", input_file("in.lj.miniMD"), "-s", "4", "--sort", "7"]).geti("sort_every") == 7
    assert Simulation.plan(["-i", input_file("in.lj.miniMD"), "-s", "4", "--sort", "0"]).geti("sort_every") == 0
    assert Simulation.plan(["-i", input_file("in.lj.miniMD"), "-s", "4"]).geti("sort_every") == 20
    # EAM forces ghost_newton off (ref/ljs.cpp:277-282); options the reference only lists in its help are tolerated
    s = Simulation.plan(["-i", input_file("in.eam.miniMD"), "-s", "4", "--eam_file", eam_file(), "-dm", "-ng", "2", "-t", "4"])
    assert s.geti("ghost_newton") == 0 and s.geti("force_style") == 1
    # -n overrides the step count
    assert Simulation.plan(["-i", input_file("in.lj.miniMD"), "-s", "4", "-n", "37"]).geti("ntimes") == 37


def test_errors_are_reported_not_fatal(tmp_path):
    with pytest.raises(HostError):
        Simulation.plan(["-i", str(tmp_path / "missing.in")])
    bad = tmp_path / "bad.in"
    bad.write_text("title\n\nparsecs\nnone\nlj\n1 1\n4 4 4\n10\n0.005\n1.44\n0.8442\n20\n2.5 0.3\n100\n")
    with pytest.raises(HostError):
        Simulation.plan(["-i", str(bad)])
    with pytest.raises(HostError):
        Simulation.plan(["-i", input_file("in.lj.miniMD"), "--no_such_flag"])
    with pytest.raises(HostError):   # running needs a device context
        Simulation.plan(["-i", input_file("in.lj.miniMD"), "-s", "4"]).run(1)


def test_input_deck_format(tmp_path):
    """Values are positional; anything after them on a line is a comment; skin is added to the cutoff."""
    deck = tmp_path / "in.custom"
    deck.write_text("my deck\n\nlj   units\nnone data\nlj  style\n0.5 1.25 eps sigma\n3 4 5 cells\n60 steps\n0.002 dt\n"
                    "1.1 T\n0.9 rho\n10 every\n2.0 0.4 cut skin\n20 thermo\n")
    s = Simulation.plan(["-i", str(deck)])
    assert s.geti("natoms") == 4 * 3 * 4 * 5 and s.geti("ntimes") == 60 and s.geti("neigh_every") == 10
    assert s.geti("thermo_nstat") == 20 and s.getr("dt") == 0.002 and s.getr("cutforce") == 2.0
    assert s.getr("cutneigh") == 2.0 + 0.4
    assert np.all(s.host_array("epsilon") == 0.5) and np.all(s.host_array("sigma6") == 1.25 ** 6)


# ------------------------------------------------------------------------------------------------
# multi-rank decomposition over gloo (2 processes)
# ------------------------------------------------------------------------------------------------
def _rank_main(rank, world, port, shape, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def reduce(values, op):
        t = torch.from_numpy(values)
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == 0 else dist.ReduceOp.MAX)

    nx, ny, nz = shape
    s = Simulation.plan(["-i", input_file("in.lj.miniMD"), "-nx", nx, "-ny", ny, "-nz", nz], "f64", rank, world, reduce)
    t = s.swap_table()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=s.host_array("x").reshape(-1, 3), v=s.host_array("v").reshape(-1, 3),
             box=np.array([s.getr(k) for k in ("xlo", "xhi", "ylo", "yhi", "zlo", "zhi")]),
             grid=np.array([s.geti(f"procgrid{d}") for d in range(3)]), loc=np.array([s.geti(f"myloc{d}") for d in range(3)]),
             sendproc=np.array(list(t.sendproc)[:t.nswap]), recvproc=np.array(list(t.recvproc)[:t.nswap]),
             pbc_any=np.array(list(t.pbc_any)[:t.nswap]), slablo=np.array(list(t.slablo)[:t.nswap]),
             slabhi=np.array(list(t.slabhi)[:t.nswap]), natoms=s.geti("natoms"))
    dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(8, 8, 16), (6, 6, 6)])
def test_two_rank_decomposition_over_gloo(shape, tmp_path):
    import torch.multiprocessing as mp
    port = 29000 + (os.getpid() % 500) + shape[2]
    mp.spawn(_rank_main, args=(2, port, shape, str(tmp_path)), nprocs=2, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(2)]
    one = Simulation.plan(["-i", input_file("in.lj.miniMD"), "-nx", shape[0], "-ny", shape[1], "-nz", shape[2]])
    x1, v1 = one.host_array("x").reshape(-1, 3), one.host_array("v").reshape(-1, 3)
    # the ranks partition the lattice: every atom exactly once, inside its owner's sub-box
    assert len(r[0]["x"]) + len(r[1]["x"]) == int(r[0]["natoms"]) == len(x1)
    xs = np.concatenate([r[0]["x"], r[1]["x"]])
    vs = np.concatenate([r[0]["v"], r[1]["v"]])
    key = lambda a: np.lexsort((a[:, 0], a[:, 1], a[:, 2]))
    assert np.array_equal(xs[key(xs)], x1[key(x1)])
    # velocities: same streams, normalised with cross-rank sums (summation order differs => 1e-13)
    assert np.allclose(vs[key(xs)], v1[key(x1)], rtol=1e-12, atol=1e-13)
    for k in range(2):
        b = r[k]["box"]
        for d in range(3):
            assert np.all(r[k]["x"][:, d] >= b[2 * d]) and np.all(r[k]["x"][:, d] < b[2 * d + 1])
    # the split goes along the longest axis (smallest cut surface, ref/comm.cpp:86-120) and
    # neighbours are mutual: what rank a sends to, receives from it in the same swap
    grid = tuple(r[0]["grid"])
    assert int(np.prod(grid)) == 2 and (shape[2] <= shape[0] or grid == (1, 1, 2))
    for w in range(len(r[0]["sendproc"])):
        for a in range(2):
            b = int(r[a]["sendproc"][w])
            assert int(r[b]["recvproc"][w]) == a
    # periodic shifts only on swaps that cross the box boundary; slabs lie inside the sender's reach
    split = int(np.argmax(grid))
    for k in range(2):
        lo, hi = r[k]["box"][2 * split], r[k]["box"][2 * split + 1]
        w0 = 2 * split
        assert r[k]["slablo"][w0] == pytest.approx(lo) and r[k]["slabhi"][w0 + 1] == pytest.approx(hi)
        assert bool(r[k]["pbc_any"][w0]) == (r[k]["loc"][split] == 0)
        assert bool(r[k]["pbc_any"][w0 + 1]) == (r[k]["loc"][split] == 1)


def test_lammps_data_file_setup(tmp_path):
    """-f <data file> (ref/setup.cpp:215-301): atoms and velocities come from the file in id order, the bin
    grid from the density (about 16 atoms per 2x2x2 bins); golden bins from the reference binary."""
    from helpers import golden, lattice_for_datafile, write_lammps_data
    g = golden("reference_datafile.json")["lj_data_half"]
    x, v, prd = lattice_for_datafile()
    data = write_lammps_data(str(tmp_path / "atoms.data"), x, v, prd)
    s = Simulation.plan(["-i", input_file("in.lj.miniMD"), "-f", data])
    assert s.geti("natoms") == len(x) == s.geti("nlocal")
    assert [s.geti("nbinx"), s.geti("nbiny"), s.geti("nbinz")] == g["bins"]
    assert np.array_equal(s.host_array("x").reshape(-1, 3), x) and np.array_equal(s.host_array("v").reshape(-1, 3), v)
    assert s.getr("xprd") == prd[0]
    # -b overrides the density rule; a deck that names the file on line 4 is equivalent to -f
    assert Simulation.plan(["-i", input_file("in.lj.miniMD"), "-f", data, "-b", "3"]).geti("nbinx") == 3
    deck = tmp_path / "in.data.miniMD"
    deck.write_text(open(input_file("in.lj.miniMD")).read().replace("none ", data + " ", 1))
    assert Simulation.plan(["-i", str(deck)]).geti("nlocal") == len(x)
    with pytest.raises(HostError):
        Simulation.plan(["-i", input_file("in.lj.miniMD"), "-f", str(tmp_path / "nope.data")])
